// Peer-visible device memory for the multi-GPU factorisation (chol_dag.cu, slk_chol_dist_*).
// One process per GPU: a rank allocates its workspace here, exports a CUDA IPC handle, the ranks
// exchange the handles over torch.distributed and map each other's workspaces; kernels then store
// into the peers' memory through NVLink.  These four calls are the only ones in the library that own
// device memory (everything else works on caller-provided buffers), because an IPC handle refers to a
// whole cudaMalloc allocation and a framework's caching allocator hands out sub-blocks.
#include "common.cuh"

#include <string.h>

using namespace slk;

extern "C" {

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
