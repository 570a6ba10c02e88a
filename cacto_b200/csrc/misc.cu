// misc.cu -- ABI version, parameter-block sizes, strided device-to-host copy used by the host feeder.
#include "common.cuh"
#include "mlp_layout.cuh"

extern "C" int32_t cacto_abi_version(void) { return CACTO_ABI_VERSION; }
extern "C" int64_t cacto_actor_param_count(int32_t ns, int32_t na) { return cacto::ActorLayout(ns, na).total; }
extern "C" int64_t cacto_critic_param_count(int32_t ns) { return cacto::CriticLayout(ns).total; }

// Column block [rows x width_bytes] of a device array with row pitch src_pitch -> host array with row pitch dst_pitch
// (one cudaMemcpy2DAsync on the copy engine): how RL_AC.rollout_to_host streams the SoA trajectories of a sub-batch into
// its columns of the pinned host buffers while the next sub-batch is being rolled out.
extern "C" int cacto_copy2d_to_host(void* dst_host, int64_t dst_pitch, const void* src_dev, int64_t src_pitch, int64_t width_bytes,
                                    int64_t rows, void* stream) {
  if (!dst_host || !src_dev) return CACTO_E_ARG;
  if (width_bytes < 0 || rows < 0 || dst_pitch < width_bytes || src_pitch < width_bytes) return CACTO_E_SIZE;
  if (width_bytes == 0 || rows == 0) return 0;
  return (int)cudaMemcpy2DAsync(dst_host, (size_t)dst_pitch, src_dev, (size_t)src_pitch, (size_t)width_bytes, (size_t)rows,
                                cudaMemcpyDeviceToHost, (cudaStream_t)stream);
}

// The same for a stack of column blocks: slab k of the source holds src_rows rows (pitch src_pitch), of which the first `rows`
// go to slab k of the destination (dst_rows rows of pitch dst_pitch) -- one cudaMemcpy3DAsync.  The compact transfer format of
// RL_AC.rollout_to_host drops the time row of every knot this way (rows = ns - 1 of src_rows = ns) without a staging pass.
extern "C" int cacto_copy3d_to_host(void* dst_host, int64_t dst_pitch, int64_t dst_rows, const void* src_dev, int64_t src_pitch,
                                    int64_t src_rows, int64_t width_bytes, int64_t rows, int64_t slabs, void* stream) {
  if (!dst_host || !src_dev) return CACTO_E_ARG;
  if (width_bytes < 0 || rows < 0 || slabs < 0 || dst_pitch < width_bytes || src_pitch < width_bytes || dst_rows < rows || src_rows < rows)
    return CACTO_E_SIZE;
  if (width_bytes == 0 || rows == 0 || slabs == 0) return 0;
  cudaMemcpy3DParms c = {};
  c.srcPtr = make_cudaPitchedPtr(const_cast<void*>(src_dev), (size_t)src_pitch, (size_t)width_bytes, (size_t)src_rows);
  c.dstPtr = make_cudaPitchedPtr(dst_host, (size_t)dst_pitch, (size_t)width_bytes, (size_t)dst_rows);
  c.extent = make_cudaExtent((size_t)width_bytes, (size_t)rows, (size_t)slabs);
  c.kind = cudaMemcpyDeviceToHost;
  return (int)cudaMemcpy3DAsync(&c, (cudaStream_t)stream);
}

// fp64 -> fp32 of values that ARE fp32 numbers (the controls: fp32 actor outputs widened for the fp64 dynamics, RL.py:223): exact.
// HBM-bound: 12 bytes per element, 32-byte loads / 16-byte stores.
namespace cacto {
__global__ void __launch_bounds__(256) k_narrow_f64(const double* __restrict__ src, float* __restrict__ dst, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = n >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const double2 a = reinterpret_cast<const double2*>(src)[2 * i], b = reinterpret_cast<const double2*>(src)[2 * i + 1];
    reinterpret_cast<float4*>(dst)[i] = make_float4((float)a.x, (float)a.y, (float)b.x, (float)b.y);
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = (float)src[i];
}
}  // namespace cacto

extern "C" int cacto_narrow_f64_to_f32(const double* src, float* dst, int64_t n, void* stream) {
  if (n < 0) return CACTO_E_SIZE;
  if (n == 0) return 0;
  if (!src || !dst || (reinterpret_cast<uintptr_t>(src) & 15) || (reinterpret_cast<uintptr_t>(dst) & 15)) return CACTO_E_ARG;
  int64_t ctas = (n / 4 + 255) / 256;
  if (ctas < 1) ctas = 1;
  if (ctas > 148 * 8) ctas = 148 * 8;
  cacto::k_narrow_f64<<<(unsigned)ctas, 256, 0, (cudaStream_t)stream>>>(src, dst, n);
  CACTO_LAUNCH_CHECK();
  return 0;
}

// n draws of Python's random.random() on the HOST from a copy of the interpreter's generator state (random.getstate(): 624 words
// of MT19937 + position; CPython _randommodule.c: a = next32 >> 5, b = next32 >> 6, (a * 2^26 + b) / 2^53), state advanced in
// place: the stratified PER sampler (replay_buffer.py:139-157) consumes B draws per round from that stream, and a Python-level
// loop over 4096 calls costs more than the rest of the round.  Same stream, same bits (tests/test_host_random.py).
static inline uint32_t mt_next(uint32_t* mt, uint32_t* pos) {
  if (*pos >= 624u) {
    for (int k = 0; k < 624; ++k) {
      const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
      mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    *pos = 0;
  }
  uint32_t y = mt[(*pos)++];
  y ^= y >> 11;
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= y >> 18;
  return y;
}
extern "C" int cacto_host_mt19937_random(uint32_t* state625, double* out, int64_t n) {
  if (n < 0) return CACTO_E_SIZE;
  if (!state625 || (n > 0 && !out)) return CACTO_E_ARG;
  if (state625[624] > 624u) return CACTO_E_ARG;
  for (int64_t i = 0; i < n; ++i) {
    const uint32_t a = mt_next(state625, state625 + 624) >> 5, b = mt_next(state625, state625 + 624) >> 6;
    out[i] = (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
  }
  return 0;
}

// p_i ** alpha on the HOST with the C library's pow -- the function CPython's float ** float calls, so the priorities written
// into the trees carry the bits of the reference's `priority ** self._alpha` (replay_buffer.py:210-216); CUDA's pow is not
// correctly rounded and would change them.  One call per batch instead of a Python loop over B floats (0.4 ms at B = 4096).
#include <math.h>
extern "C" int cacto_host_pow(const double* x, double exponent, double* out, int64_t n) {
  if (n < 0) return CACTO_E_SIZE;
  if (n > 0 && (!x || !out)) return CACTO_E_ARG;
  for (int64_t i = 0; i < n; ++i) out[i] = pow(x[i], exponent);
  return 0;
}
