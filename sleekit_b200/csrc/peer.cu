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
