// rtg.cu -- K5: n-step reward-to-go windows of RL_AC.RL_Solve (RL.py:173-187) for a ragged batch of
// TO trajectories.  One CTA per trajectory, one thread per knot; rewards staged in shared memory; the s_next block is
// written as one coalesced shifted copy of the state block.
// Bit-exact with the reference: each window is summed left-to-right in fp64 starting from 0
// (Python's builtin sum), then rounded to float32 and stored back as fp64 (quirk Q12).  There is no
// discount factor in the reference (gamma == 1).
#include "common.cuh"

namespace cacto {

constexpr int RTG_THREADS = 128;

__global__ void __launch_bounds__(RTG_THREADS) k_rtg_window(const int64_t* __restrict__ offsets, const double* __restrict__ rwrd,
                                                            const double* __restrict__ states, int ns, int nsteps_td, int mc,
                                                            double* __restrict__ partial, double* __restrict__ total,
                                                            double* __restrict__ s_next, double* __restrict__ done,
                                                            double* __restrict__ term, double* __restrict__ ep_return,
                                                            int smem_knots) {
  extern __shared__ double s_r[];
  const int e = blockIdx.x;
  const int64_t o = offsets[e];
  const int K = (int)(offsets[e + 1] - o);        // knots = T + 1
  if (K <= 0) return;
  const int T = K - 1;
  const bool staged = K <= smem_knots;
  if (staged) {
    for (int i = threadIdx.x; i < K; i += RTG_THREADS) s_r[i] = rwrd[o + i];
    __syncthreads();
  }
  const double* r = staged ? s_r : (rwrd + o);
  for (int i = threadIdx.x; i < K; i += RTG_THREADS) {
    int final_step;
    double d = 0.0;
    if (mc) {
      final_step = T;
      d = 1.0;
    } else {
      final_step = min(i + nsteps_td, T);
      if (final_step == T) d = 1.0;
    }
    double acc = 0.0;
    for (int k = i; k <= final_step; ++k) acc = __dadd_rn(acc, r[k]);
    const double part = (double)(float)acc;
    for (int k = final_step + 1; k <= T; ++k) acc = __dadd_rn(acc, r[k]);
    partial[o + i] = part;
    total[o + i] = (double)(float)acc;
    if (i == 0) ep_return[e] = acc;                 // the same left-to-right sum over all knots, unrounded
    done[o + i] = d;
    term[o + i] = (i == T) ? 1.0 : 0.0;
  }
  // s_next[i] = state[final_step + 1] where the window ends before the trajectory does (done == 0), else 0 (RL.py:180-184):
  // rows i < T - n are a copy of the state block shifted by n + 1 knots, so the CTA moves it as one contiguous, fully
  // coalesced stream instead of ns strided stores per knot
  const int n_copy = mc ? 0 : max(0, T - nsteps_td);        // knots with i + n < T
  const int64_t shift = (int64_t)(nsteps_td + 1) * ns;
  const double* src = states + o * ns;
  double* dst = s_next + o * ns;
  for (int j = threadIdx.x; j < K * ns; j += RTG_THREADS) dst[j] = (j < n_copy * ns) ? src[j + shift] : 0.0;
}

}  // namespace cacto

using namespace cacto;

extern "C" int cacto_rtg_window(const int64_t* offsets, int32_t E, const double* rwrd, const double* states, int32_t ns,
                                int32_t nsteps_td, int32_t mc, double* partial, double* total_rtg, double* s_next, double* done,
                                double* term, double* ep_return, void* stream) {
  if (!offsets || !rwrd || !states || !partial || !total_rtg || !s_next || !done || !term || !ep_return) return CACTO_E_ARG;
  if (E < 0 || ns < 1 || nsteps_td < 0) return CACTO_E_SIZE;
  if (E == 0) return 0;
  const int smem_knots = 512;              // 4 KB of rewards (every conf has <= 501 knots) so that 16 CTAs fit an SM; longer trajectories read global memory
  k_rtg_window<<<E, RTG_THREADS, smem_knots * sizeof(double), (cudaStream_t)stream>>>(
      offsets, rwrd, states, ns, nsteps_td, mc, partial, total_rtg, s_next, done, term, ep_return, smem_knots);
  CACTO_LAUNCH_CHECK();
  return 0;
}
