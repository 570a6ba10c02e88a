// peak.cu -- in-run measurement of the FP32 CUDA-core FMA peak, the roofline denominator of the
// fp32-FMA kernels ("CUDA-core FP32 FMA peak must be measured in-run").
// Measurement infrastructure only: called by bench.py, never on the product path.
#include "common.cuh"

namespace cacto {

constexpr int PEAK_ILP = 16;

__global__ void __launch_bounds__(256) k_fma_peak(float* __restrict__ out, int iters, float a, float b) {
  float acc[PEAK_ILP];
#pragma unroll
  for (int i = 0; i < PEAK_ILP; ++i) acc[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < PEAK_ILP; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PEAK_ILP; ++i) s += acc[i];
  if (s == 12345.678f) out[0] = s;     // never true in practice; keeps the chain alive
}

}  // namespace cacto

// Launches blocks x 256 threads, each doing iters x 16 dependent-chain FMAs: flops = blocks*256*iters*16*2.
extern "C" int cacto_peak_fma_fp32(float* out, int32_t iters, int32_t blocks, void* stream) {
  if (!out || iters <= 0 || blocks <= 0) return CACTO_E_ARG;
  cacto::k_fma_peak<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, iters, 0.999f, 0.001f);
  CACTO_LAUNCH_CHECK();
  return 0;
}
