// rollout.cu -- K1: batched policy rollouts, RL_AC.create_TO_init (RL.py:197-233) for B initial
// conditions at once, fusing per time step
//     NN.eval(actor, x)  (utils.py:17-24 normalisation + Dense/LeakyReLU x2 + Dense, NeuralNetwork.py:51-63,130-138)
//     Env.simulate(x, u) (environment.py:80-91 / :235 / :437 / :584; robot_utils.py:399-405)
// in one persistent kernel: a CTA owns S rollouts for their whole horizon, the hidden activations
// never leave shared memory, the 256x256 layer is computed in place, and the state / control
// trajectories are written structure-of-arrays ([t][component][rollout]) so that every store is
// coalesced.  As in the reference the actor runs in float32 on the float32-rounded state and the
// dynamics run in float64 on the float64 state (quirk Q13).
#include "common.cuh"
#include "mlp.cuh"
#include "systems.cuh"

namespace cacto {

constexpr int RO_S = 64;      // rollouts per CTA
constexpr int RO_NT = 256;    // threads per CTA
constexpr int RO_MIN_CTAS = 2; // resident CTAs per SM the register budget is capped for (<= 128 registers)
constexpr int NSP = 16;       // padded input width (>= CACTO_MAX_NS, multiple of 4)
constexpr int NAP = 8;        // padded action width

struct RolloutSmem {
  float H[RO_S][ACTOR_H];
  float X[RO_S][NSP];
  float ACT[RO_S][NAP];
  int tmax;
};

template <int SYS>
__global__ void __launch_bounds__(RO_NT, RO_MIN_CTAS) k_rollout(const __grid_constant__ cacto_sys_params P, const float* __restrict__ actor,
                                                   int use_actor, const double* __restrict__ ics, const int32_t* __restrict__ horizon,
                                                   int T_max, double* __restrict__ states, double* __restrict__ controls,
                                                   int32_t* __restrict__ flags, double* __restrict__ rewards, int64_t B) {
  constexpr int NX = SysDims<SYS>::NX, NA = SysDims<SYS>::NA, NS = NX + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  RolloutSmem& sm = *reinterpret_cast<RolloutSmem*>(smem_raw);
  const ActorLayout L(NS, NA);
  const int tid = threadIdx.x;
  const int64_t b = (int64_t)blockIdx.x * RO_S + tid;
  const bool owner = tid < RO_S && b < B;

  double x[NS];
  int h = 0, ok = 1;
  if (tid == 0) sm.tmax = 0;
  __syncthreads();
  if (owner) {
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      x[j] = ics[b * NS + j];
      states[(int64_t)j * B + b] = x[j];
    }
    h = min(max(horizon[b], 0), T_max);
    atomicMax(&sm.tmax, h);
  }
  if (tid < RO_S) {
#pragma unroll
    for (int j = 0; j < NSP; ++j) sm.X[tid][j] = 0.f;
  }
  __syncthreads();
  const int tmax = sm.tmax;

  for (int t = 0; t < tmax; ++t) {
    const bool live = owner && ok && t < h;
    if (use_actor) {
      if (tid < RO_S) {
#pragma unroll
        for (int j = 0; j < NS; ++j) sm.X[tid][j] = live ? normalize_component(P, j, (float)x[j]) : 0.f;
      }
      __syncthreads();
      tile_gemm<RO_S, ACTOR_H, RO_NT, false>(&sm.X[0][0], NSP, NS, actor + L.W1, ACTOR_H, [&](int row, int col, const float4& acc) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(actor + L.b1 + col));
        *reinterpret_cast<float4*>(&sm.H[row][col]) =
            make_float4(leaky(acc.x + bb.x), leaky(acc.y + bb.y), leaky(acc.z + bb.z), leaky(acc.w + bb.w));
      });
      __syncthreads();
      tile_gemm<RO_S, ACTOR_H, RO_NT, true>(&sm.H[0][0], ACTOR_H, ACTOR_H, actor + L.W2, ACTOR_H, [&](int row, int col, const float4& acc) {
        const float4 bb = __ldg(reinterpret_cast<const float4*>(actor + L.b2 + col));
        *reinterpret_cast<float4*>(&sm.H[row][col]) =
            make_float4(leaky(acc.x + bb.x), leaky(acc.y + bb.y), leaky(acc.z + bb.z), leaky(acc.w + bb.w));
      });
      __syncthreads();
      tile_gemm_small<RO_S, RO_NT, 8, true>(&sm.H[0][0], ACTOR_H, ACTOR_H, actor + L.W3, NA, 0,
                                            [&](int s, int j, float v) { sm.ACT[s][j] = v + __ldg(actor + L.b3 + j); });
      __syncthreads();
    }
    if (live) {
      double u[NA], xn[NS];
#pragma unroll
      for (int j = 0; j < NA; ++j) {
        u[j] = use_actor ? (double)sm.ACT[tid][j] : 0.0;
        controls[((int64_t)t * NA + j) * B + b] = u[j];
      }
      if (rewards != nullptr) rewards[(int64_t)t * B + b] = sys_reward<SYS, double>(P, P.w_running, x, u, false);
      sys_step<SYS, double>(P, x, u, xn);
      xn[NX] = x[NX] + P.dt;
      bool nan = false;
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        x[j] = xn[j];
        nan |= (xn[j] != xn[j]);
        states[((int64_t)(t + 1) * NS + j) * B + b] = xn[j];
      }
      if (nan) ok = 0;                          // RL.py:229-231
      if (rewards != nullptr && t + 1 == h && !nan)
        rewards[(int64_t)(t + 1) * B + b] = sys_reward<SYS, double>(P, P.w_terminal, x, (const double*)nullptr, false);
    }
  }
  if (owner) flags[b] = ok;
}

template <int SYS>
static int launch_rollout(const cacto_sys_params& P, const float* actor, int use_actor, const double* ics, const int32_t* horizon,
                          int T_max, double* states, double* controls, int32_t* flags, double* rewards, int64_t B, cudaStream_t st) {
  auto k = k_rollout<SYS>;
  const size_t sm = sizeof(RolloutSmem);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return (int)e;
  k<<<(unsigned)((B + RO_S - 1) / RO_S), RO_NT, sm, st>>>(P, actor, use_actor, ics, horizon, T_max, states, controls, flags, rewards, B);
  CACTO_LAUNCH_CHECK();
  return 0;
}

}  // namespace cacto

using namespace cacto;

extern "C" int cacto_rollout(const cacto_sys_params* p, const float* actor_params, int use_actor, const double* ics,
                             const int32_t* horizon, int32_t T_max, double* states, double* controls, int32_t* flags,
                             double* rewards, int64_t B, void* stream) {
  if (!p) return CACTO_E_ARG;
  if (B < 0 || T_max < 0) return CACTO_E_SIZE;
  if (B == 0) return 0;
  if (!ics || !horizon || !states || !flags || (T_max > 0 && !controls) || (use_actor && !actor_params)) return CACTO_E_ARG;
  if (use_actor && (reinterpret_cast<uintptr_t>(actor_params) & 15)) return CACTO_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  switch (p->system) {
    case CACTO_SINGLE_INTEGRATOR: return launch_rollout<CACTO_SINGLE_INTEGRATOR>(*p, actor_params, use_actor, ics, horizon, T_max, states, controls, flags, rewards, B, st);
    case CACTO_DOUBLE_INTEGRATOR: return launch_rollout<CACTO_DOUBLE_INTEGRATOR>(*p, actor_params, use_actor, ics, horizon, T_max, states, controls, flags, rewards, B, st);
    case CACTO_CAR: return launch_rollout<CACTO_CAR>(*p, actor_params, use_actor, ics, horizon, T_max, states, controls, flags, rewards, B, st);
    case CACTO_CAR_PARK: return launch_rollout<CACTO_CAR_PARK>(*p, actor_params, use_actor, ics, horizon, T_max, states, controls, flags, rewards, B, st);
    case CACTO_MANIPULATOR: return launch_rollout<CACTO_MANIPULATOR>(*p, actor_params, use_actor, ics, horizon, T_max, states, controls, flags, rewards, B, st);
    case CACTO_UR5: return launch_rollout<CACTO_UR5>(*p, actor_params, use_actor, ics, horizon, T_max, states, controls, flags, rewards, B, st);
    default: return CACTO_E_SYSTEM;
  }
}
