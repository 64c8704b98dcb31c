// Dependent-issue latency of the leaf's rounding chain on B200 (development aid):
//   FSETP -> FSEL -> FSETP -> FSEL -> FSETP -> FSEL -> FADD -> FMUL -> FFMA, repeated.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/lat_chain tools/lat_chain.cu && /tmp/lat_chain
#include <cstdio>
#include <cuda_runtime.h>
__global__ void chain(float* out, const float* X, const float* V, float w, float uy, float u, int iters, long long* cyc) {
  float x[8], v[8];
  for (int k = 0; k < 8; ++k) { x[k] = X[k]; v[k] = V[k]; }
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    const bool b2 = w >= x[4];
    const float t1 = b2 ? x[6] : x[2];
    const float xa = b2 ? x[7] : x[3], xb = b2 ? x[5] : x[1];
    const float va = b2 ? v[7] : v[3], vb = b2 ? v[5] : v[1], vc = b2 ? v[6] : v[2], vd = b2 ? v[4] : v[0];
    const bool b1 = w >= t1;
    const float tt = b1 ? xa : xb;
    const float c1 = b1 ? va : vb, c0 = b1 ? vc : vd;
    const float qq = (w >= tt) ? c1 : c0;
    const float res = __fmul_rn(__fsub_rn(w, qq), uy);
    w = __fmaf_rn(-res, u, w + 0.37f);
  }
  long long t1 = clock64();
  out[threadIdx.x] = w;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  float hx[8] = {0, -0.857f, -0.571f, -0.286f, 0.0f, 0.286f, 0.571f, 0.857f}, hv[8] = {-1, -0.714f, -0.429f, -0.143f, 0.143f, 0.429f, 0.714f, 1};
  float *X, *V, *out; long long* cyc;
  cudaMalloc(&X, 32); cudaMalloc(&V, 32); cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
  cudaMemcpy(X, hx, 32, cudaMemcpyHostToDevice); cudaMemcpy(V, hv, 32, cudaMemcpyHostToDevice);
  for (int threads : {32, 128}) {
    chain<<<1, threads>>>(out, X, V, 0.3f, 1.7f, 0.9f, 4096, cyc);
    chain<<<1, threads>>>(out, X, V, 0.3f, 1.7f, 0.9f, 4096, cyc);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("threads %d: %.1f cycles per column-chain iteration\n", threads, (double)h / 4096);
  }
  return 0;
}
