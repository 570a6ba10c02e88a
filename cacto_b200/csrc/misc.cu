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
