// common.cuh -- launch helpers and shared-memory tile staging used by the element-wise kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "cacto_b200.h"

namespace cacto {

#define CACTO_LAUNCH_CHECK()                         \
  do {                                               \
    cudaError_t e__ = cudaGetLastError();            \
    if (e__ != cudaSuccess) return (int)e__;         \
  } while (0)

// Programmatic dependent launch (the kernels of one update form a chain of dependent launches: schedule -> critic gradient ->
// Adam -> schedule -> actor gradient -> Adam).  A kernel launched with launch_pdl may be scheduled while its predecessor in
// the stream is still running; it must call pdl_wait() before it touches global memory (blocks until the predecessor grid
// has completed and its writes are visible).  pdl_wait() also releases the NEXT kernel of the chain for scheduling, so the
// launch latency between two dependent kernels (~2 us each, a tenth of the small-batch update) overlaps the running kernel.
// EVERY kernel of such a chain has to call pdl_wait(): completion of a grid implies completion of its predecessors only if
// the grid itself waited.  Inside stream capture the attribute becomes a programmatic edge of the CUDA graph.
__device__ __forceinline__ void pdl_wait() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... P, typename... A>
inline cudaError_t launch_pdl(void (*kernel)(P...), unsigned grid, unsigned block, size_t smem, cudaStream_t st, A&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, P(args)...);
}

// Padded row stride (odd number of elements) so that "thread r touches row r" is bank-conflict free.
__host__ __device__ constexpr int pad_odd(int w) { return (w & 1) ? w : w + 1; }

// Rows [row0, row0+rows) of a row-major [B][W] array  <->  a shared tile with stride pad_odd(W).
// Global accesses are fully coalesced (consecutive threads touch consecutive elements).  With an odd W the tile is the
// row block itself (stride W): it moves as 16-byte vectors without any index arithmetic -- the div/mod per element of the
// general path was most of the instruction count of the dynamics-step kernel (profiles/README.md).
template <typename T>
__device__ __forceinline__ bool vec16_ok(const void* g, int count) {
  return (reinterpret_cast<uintptr_t>(g) & 15) == 0 && (count * (int)sizeof(T)) % 16 == 0;
}
template <int W, int NT, typename T>
__device__ __forceinline__ void tile_load_rows(const T* __restrict__ g, T* __restrict__ s, int rows) {
  constexpr int WP = pad_odd(W);
  if (WP == W) {
    const int count = rows * W;
    if (vec16_ok<T>(g, count)) {
      const int4* g4 = reinterpret_cast<const int4*>(g);
      int4* s4 = reinterpret_cast<int4*>(s);
      for (int i = threadIdx.x; i < count * (int)sizeof(T) / 16; i += NT) s4[i] = g4[i];
    } else {
      for (int i = threadIdx.x; i < count; i += NT) s[i] = g[i];
    }
    return;
  }
  for (int i = threadIdx.x; i < rows * W; i += NT) {
    int r = i / W, c = i - r * W;
    s[r * WP + c] = g[i];
  }
}
template <int W, int NT, typename T>
__device__ __forceinline__ void tile_store_rows(T* __restrict__ g, const T* __restrict__ s, int rows) {
  constexpr int WP = pad_odd(W);
  if (WP == W) {
    const int count = rows * W;
    if (vec16_ok<T>(g, count)) {
      int4* g4 = reinterpret_cast<int4*>(g);
      const int4* s4 = reinterpret_cast<const int4*>(s);
      for (int i = threadIdx.x; i < count * (int)sizeof(T) / 16; i += NT) g4[i] = s4[i];
    } else {
      for (int i = threadIdx.x; i < count; i += NT) g[i] = s[i];
    }
    return;
  }
  for (int i = threadIdx.x; i < rows * W; i += NT) {
    int r = i / W, c = i - r * W;
    g[i] = s[r * WP + c];
  }
}

// Thread `threadIdx.x` publishes its W register values as row threadIdx.x of the tile, then the CTA
// writes the tile out coalesced.  All threads of the CTA must call it.
template <int W, int NT, typename T>
__device__ __forceinline__ void stage_out_rows(T* __restrict__ g_tile, const T* vals, T* smem, int rows) {
  constexpr int WP = pad_odd(W);
  __syncthreads();
#pragma unroll
  for (int c = 0; c < W; ++c) smem[threadIdx.x * WP + c] = vals[c];
  __syncthreads();
  tile_store_rows<W, NT, T>(g_tile, smem, rows);
}

}  // namespace cacto
