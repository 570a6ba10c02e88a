// peer.cu -- CUDA-IPC plumbing of the data-parallel update: the gradient blocks and arrival flags that k_adam_peer (update.cu)
// reads across GPUs live in one cudaMalloc region per rank, exported to the other ranks of the box (one process per GPU) as an
// IPC handle and mapped there with peer access over NVLink.  torch's caching allocator sub-allocates and cannot hand out an
// exportable base pointer, hence the four calls below; the region is wrapped as a torch tensor on the Python side
// (cacto_b200/parallel.py: PeerRegion).  The reference has no multi-GPU support (main.py:59).
#include <string.h>
#include "common.cuh"

static_assert(sizeof(cudaIpcMemHandle_t) == CACTO_IPC_HANDLE_BYTES, "IPC handle size");

extern "C" int cacto_peer_alloc(int64_t bytes, void** out) {
  if (bytes <= 0 || !out) return CACTO_E_ARG;
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(p, 0, (size_t)bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    cudaFree(p);
    return (int)e;
  }
  *out = p;
  return 0;
}

extern "C" int cacto_peer_free(void* p) {
  if (!p) return CACTO_E_ARG;
  return (int)cudaFree(p);
}

extern "C" int cacto_peer_export(void* p, void* handle) {
  if (!p || !handle) return CACTO_E_ARG;
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) return (int)e;
  memcpy(handle, &h, sizeof(h));
  return 0;
}

extern "C" int cacto_peer_open(const void* handle, void** out) {
  if (!handle || !out) return CACTO_E_ARG;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return (int)e;
  *out = p;
  return 0;
}

extern "C" int cacto_peer_close(void* p) {
  if (!p) return CACTO_E_ARG;
  return (int)cudaIpcCloseMemHandle(p);
}
