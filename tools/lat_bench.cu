// Dependent-issue latencies of the primitives the serial chains are made of (development aid).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bin/lat_bench tools/lat_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 256
__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}
__global__ void k(long long* out, double* sink, double x0, float f0) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = x0 + i;
  __syncthreads();
  long long t0, t1; double x = x0, y = x0 * 0.5; float f = f0, g = f0 * 0.5f; double c[2] = {x0, x0};
  int idx = threadIdx.x & 31;
  // DFMA
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) x = __fma_rn(x, y, y);
  t1 = clock64(); if (threadIdx.x == 0) out[0] = t1 - t0;
  // rsqrt(double)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = rsqrt(x) + 1.0;
  t1 = clock64(); if (threadIdx.x == 0) out[1] = t1 - t0;
  // DMMA dependent
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) dmma(c, y, y);
  t1 = clock64(); if (threadIdx.x == 0) out[2] = t1 - t0;
  // FFMA
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) f = __fmaf_rn(f, g, g);
  t1 = clock64(); if (threadIdx.x == 0) out[3] = t1 - t0;
  // SHFL f32
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) f = __shfl_sync(0xffffffffu, f, (idx + 1) & 31);
  t1 = clock64(); if (threadIdx.x == 0) out[4] = t1 - t0;
  // SHFL f64
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (idx + 1) & 31);
  t1 = clock64(); if (threadIdx.x == 0) out[5] = t1 - t0;
  // LDS.64 dependent (pointer chase)
  int p = idx;
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) p = ((int)sm[p & 1023]) & 1023;
  t1 = clock64(); if (threadIdx.x == 0) out[6] = t1 - t0;
  // __syncthreads
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) __syncthreads();
  t1 = clock64(); if (threadIdx.x == 0) out[7] = t1 - t0;
  // FMNMX / fminf
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) f = fminf(f + 1.0f, g);
  t1 = clock64(); if (threadIdx.x == 0) out[8] = t1 - t0;
  // double rcp
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __drcp_rn(x) + 1.0;
  t1 = clock64(); if (threadIdx.x == 0) out[9] = t1 - t0;
  // DMUL
  t0 = clock64();
#pragma unroll
  for (int i = 0; i < N; ++i) x = __dmul_rn(x, y);
  t1 = clock64(); if (threadIdx.x == 0) out[10] = t1 - t0;
  // rsqrtf via float + newton (candidate)
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    double y0 = (double)rsqrtf((float)x);
    double e = __fma_rn(-__dmul_rn(x, y0), y0, 1.0); y0 = __fma_rn(__dmul_rn(y0, 0.5), e, y0);
    e = __fma_rn(-__dmul_rn(x, y0), y0, 1.0); y0 = __fma_rn(__dmul_rn(y0, 0.5), e, y0);
    x = y0 + 1.0;
  }
  t1 = clock64(); if (threadIdx.x == 0) out[11] = t1 - t0;
  sink[threadIdx.x] = x + f + c[0] + c[1] + p;
}
int main() {
  long long* out; double* sink; cudaMalloc(&out, 128); cudaMalloc(&sink, 8 * 1024);
  const char* names[] = {"DFMA", "rsqrt(double)+DADD", "DMMA m8n8k4 (dependent)", "FFMA", "SHFL f32", "SHFL f64", "LDS.64 chase (+cvt)", "__syncthreads", "FADD+FMNMX", "__drcp_rn+DADD", "DMUL", "rsqrtf+2 Newton (double)+DADD"};
  for (int threads : {32, 256}) {
    k<<<1, threads>>>(out, sink, 1.37, 0.73f); cudaDeviceSynchronize();
    k<<<1, threads>>>(out, sink, 1.37, 0.73f); cudaDeviceSynchronize();
    long long h[16]; cudaMemcpy(h, out, 128, cudaMemcpyDeviceToHost);
    printf("threads=%d\n", threads);
    for (int i = 0; i < 12; ++i) printf("  %-32s %7.1f cycles/op\n", names[i], (double)h[i] / N);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
