// Peer-visible device memory for the multi-GPU factorisation (chol_dag.cu, slk_chol_dist_*).
// One process per GPU: a rank allocates its workspace here, exports a CUDA IPC handle, the ranks
// exchange the handles over torch.distributed and map each other's workspaces; kernels then store
// into the peers' memory through NVLink.  These four calls are the only ones in the library that own
// device memory (everything else works on caller-provided buffers), because an IPC handle refers to a
// whole cudaMalloc allocation and a framework's caching allocator hands out sub-blocks.
#include "common.cuh"

#include <string.h>

using namespace slk;

namespace slk {

struct PeerPtrs { float* p[8]; };

// In-place sum all-reduce over buffers that every rank can address (NVLink peer memory): rank r owns the
// r-th slice of the buffer, reads that slice from all ranks (peer loads), adds in rank order and stores the
// sum back into every rank's copy (peer stores).  One rank computes each element, so all copies end up
// bit-identical, and no rank touches another rank's slice: a barrier before (inputs complete) and one after
// (stores landed) are all the synchronisation there is.  16-byte accesses, two elements in flight per thread.
__global__ void __launch_bounds__(512) peer_allreduce_kernel(const __grid_constant__ PeerPtrs P, int nranks, int rank,
                                                             int64_t n4) {
  const int64_t lo = n4 * rank / nranks, hi = n4 * (rank + 1) / nranks;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += 2 * stride) {
    const int64_t i2 = i + stride;
    const bool two = i2 < hi;
    float4 a[8], b[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (q < nranks) {
        a[q] = __ldcg(reinterpret_cast<const float4*>(P.p[q]) + i);
        if (two) b[q] = __ldcg(reinterpret_cast<const float4*>(P.p[q]) + i2);
      }
    }
    float4 sa = a[0], sb = b[0];
#pragma unroll
    for (int q = 1; q < 8; ++q) {
      if (q < nranks) {
        sa.x = __fadd_rn(sa.x, a[q].x); sa.y = __fadd_rn(sa.y, a[q].y); sa.z = __fadd_rn(sa.z, a[q].z); sa.w = __fadd_rn(sa.w, a[q].w);
        if (two) { sb.x = __fadd_rn(sb.x, b[q].x); sb.y = __fadd_rn(sb.y, b[q].y); sb.z = __fadd_rn(sb.z, b[q].z); sb.w = __fadd_rn(sb.w, b[q].w); }
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      if (q < nranks) {
        __stcg(reinterpret_cast<float4*>(P.p[q]) + i, sa);
        if (two) __stcg(reinterpret_cast<float4*>(P.p[q]) + i2, sb);
      }
    }
  }
}

__global__ void timestamp_kernel(unsigned long long* slot) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  *slot = t;
}
}  // namespace slk

extern "C" {

/* A CUDA stream of the given priority on the current device (0 = default, negative = higher; clamped to
   the device's range).  The layer-set driver needs one real stream per independent layer -- a framework's
   stream pool may hand out the same few streams again and again -- and high-priority streams for the
   longest chains.  *stream_host receives the cudaStream_t. */
int slk_stream_create(int priority, void** stream_host) {
  SLK_REQUIRE(stream_host, "NULL output");
  int lo = 0, hi = 0;
  SLK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));   // lo = least priority (largest number)
  if (priority < hi) priority = hi;
  if (priority > lo) priority = lo;
  cudaStream_t s;
  SLK_CUDA(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, priority));
  *stream_host = (void*)s;
  return SLK_OK;
}

int slk_stream_destroy(void* stream) {
  SLK_REQUIRE(stream, "NULL stream");
  SLK_CUDA(cudaStreamDestroy((cudaStream_t)stream));
  return SLK_OK;
}

/* In-place sum all-reduce of `count` floats (count % 4 == 0) over nranks peer-visible buffers (peers_host:
   HOST array of the nranks device pointers, entry `rank` this rank's own; slk_peer_alloc / slk_peer_open).
   The caller brackets it with barriers over the ranks on the same stream (see slk_chol_dist_*).  Replaces the
   NCCL all-reduce of the packed statistics (statistics.py:76-87 summed over ranks) with NVLink peer loads /
   stores issued by our own kernel. */
int slk_peer_allreduce_f32(void* const* peers_host, int32_t nranks, int32_t rank, int64_t count, void* stream) {
  SLK_REQUIRE(peers_host && nranks >= 1 && nranks <= 8 && rank >= 0 && rank < nranks && count >= 0 && count % 4 == 0,
              "bad arguments");
  if (count == 0 || nranks == 1) return SLK_OK;
  PeerPtrs P;
  for (int q = 0; q < 8; ++q) P.p[q] = q < nranks ? (float*)peers_host[q] : nullptr;
  for (int q = 0; q < nranks; ++q) SLK_REQUIRE(P.p[q] && ((uintptr_t)P.p[q] % 16) == 0, "peer buffer %d missing or misaligned", q);
  const int64_t n4 = count / 4;
  int64_t blocks = ceil_div(ceil_div(n4, nranks), 2 * 512);
  const int64_t cap = (int64_t)sm_count() * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  peer_allreduce_kernel<<<(unsigned)blocks, 512, 0, (cudaStream_t)stream>>>(P, nranks, rank, n4);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

/* development aid (tools/timeline.py): a one-thread kernel that stores %globaltimer (ns) into *slot when the
   stream reaches it -- stream-ordered phase marks inside a multi-stream CUDA graph */
int slk_debug_timestamp(void* slot, void* stream) {
  SLK_REQUIRE(slot, "NULL slot");
  timestamp_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((unsigned long long*)slot);
  SLK_LAUNCH_CHECK();
  return SLK_OK;
}

/* bytes of device memory on the current device; *dptr_host receives the pointer, handle_host (64
   bytes, host) the cudaIpcMemHandle_t to send to the peers */
int slk_peer_alloc(size_t bytes, void** dptr_host, void* handle_host) {
  SLK_REQUIRE(bytes > 0 && dptr_host && handle_host, "bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  void* p = nullptr;
  SLK_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t hd;
  cudaError_t e = cudaIpcGetMemHandle(&hd, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    set_error("cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    return SLK_ERR_CUDA;
  }
  memcpy(handle_host, &hd, sizeof(hd));
  *dptr_host = p;
  return SLK_OK;
}

/* map a peer's allocation (its 64-byte handle) into this process; enables peer access lazily */
int slk_peer_open(const void* handle_host, void** dptr_host) {
  SLK_REQUIRE(handle_host && dptr_host, "bad arguments");
  cudaIpcMemHandle_t hd;
  memcpy(&hd, handle_host, sizeof(hd));
  void* p = nullptr;
  SLK_CUDA(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
  *dptr_host = p;
  return SLK_OK;
}

int slk_peer_close(void* dptr) {
  SLK_REQUIRE(dptr, "NULL pointer");
  SLK_CUDA(cudaIpcCloseMemHandle(dptr));
  return SLK_OK;
}

int slk_peer_free(void* dptr) {
  SLK_REQUIRE(dptr, "NULL pointer");
  SLK_CUDA(cudaFree(dptr));
  return SLK_OK;
}

}  // extern "C"
