// dyn_kernels.cu -- batched element-wise kernels over (state, action) samples:
//   K1' simulate_batch, K2 derivative_batch / augmented_derivative, EE position, reward (+ dr/da).
// One thread per sample.  All of them are HBM-bound: with the reference's row
// layout [B][width] every global access goes through a shared-memory tile so that it is coalesced;
// with the structure-of-arrays layout [width][B] threads read/write global memory directly.
#include "common.cuh"
#include "systems.cuh"

namespace cacto {

constexpr int LAYOUT_ROWS = 0, LAYOUT_SOA = 1;

template <int SYS> struct Tpb { static constexpr int V = (SYS == CACTO_UR5) ? 64 : 128; };

// ---- per-thread sample I/O --------------------------------------------------------------------
template <int W, int NT, int LAYOUT, typename T>
__device__ __forceinline__ void read_sample(const T* __restrict__ g, int64_t B, int64_t row0, int rows, T* smem, T* out) {
  if (LAYOUT == LAYOUT_ROWS) {
    constexpr int WP = pad_odd(W);
    __syncthreads();
    tile_load_rows<W, NT, T>(g + row0 * W, smem, rows);
    __syncthreads();
    if ((int)threadIdx.x < rows) {
#pragma unroll
      for (int c = 0; c < W; ++c) out[c] = smem[threadIdx.x * WP + c];
    }
  } else {
    if ((int)threadIdx.x < rows) {
#pragma unroll
      for (int c = 0; c < W; ++c) out[c] = g[(int64_t)c * B + row0 + threadIdx.x];
    }
  }
}
// Two inputs of a sample (state and action) with ONE barrier pair: both tiles are requested back to back into disjoint parts of
// the shared buffer (the step kernel moves 68 bytes per sample: a second load that only starts after the first has been waited
// for and consumed doubles the exposed memory latency of a block).
template <int W1, int W2, int NT, int LAYOUT, typename T>
__device__ __forceinline__ void read_sample2(const T* __restrict__ g1, const T* __restrict__ g2, int64_t B, int64_t row0, int rows, T* smem,
                                             T* out1, T* out2) {
  if (LAYOUT == LAYOUT_ROWS) {
    constexpr int WP1 = pad_odd(W1), WP2 = pad_odd(W2);
    T* smem2 = smem + ((NT * WP1 * (int)sizeof(T) + 15) / 16) * 16 / (int)sizeof(T);
    __syncthreads();
    tile_load_rows<W1, NT, T>(g1 + row0 * W1, smem, rows);
    tile_load_rows<W2, NT, T>(g2 + row0 * W2, smem2, rows);
    __syncthreads();
    if ((int)threadIdx.x < rows) {
#pragma unroll
      for (int c = 0; c < W1; ++c) out1[c] = smem[threadIdx.x * WP1 + c];
#pragma unroll
      for (int c = 0; c < W2; ++c) out2[c] = smem2[threadIdx.x * WP2 + c];
    }
  } else {
    read_sample<W1, NT, LAYOUT, T>(g1, B, row0, rows, smem, out1);
    read_sample<W2, NT, LAYOUT, T>(g2, B, row0, rows, smem, out2);
  }
}
template <int W, int NT, int LAYOUT, typename T>
__device__ __forceinline__ void write_sample(T* __restrict__ g, int64_t B, int64_t row0, int rows, T* smem, const T* vals) {
  if (LAYOUT == LAYOUT_ROWS) {
    stage_out_rows<W, NT, T>(g + row0 * W, vals, smem, rows);
  } else {
    if ((int)threadIdx.x < rows) {
#pragma unroll
      for (int c = 0; c < W; ++c) g[(int64_t)c * B + row0 + threadIdx.x] = vals[c];
    }
  }
}

// ---- kernels -------------------------------------------------------------------------------------
template <int SYS, typename T, int LAYOUT>
__global__ void __launch_bounds__(Tpb<SYS>::V) k_dyn_step(const __grid_constant__ cacto_sys_params P, const T* __restrict__ state,
                                                          const T* __restrict__ action, T* __restrict__ next, int64_t B) {
  constexpr int NX = SysDims<SYS>::NX, NA = SysDims<SYS>::NA, NS = NX + 1, NT = Tpb<SYS>::V;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  const int64_t row0 = (int64_t)blockIdx.x * NT;
  const int rows = (int)min((int64_t)NT, B - row0);
  T x[NS], u[NA], xn[NS];
  read_sample2<NS, NA, NT, LAYOUT, T>(state, action, B, row0, rows, smem, x, u);
  if ((int)threadIdx.x < rows) {
    sys_step<SYS, T>(P, x, u, xn);
    xn[NX] = x[NX] + T(P.dt);
  }
  write_sample<NS, NT, LAYOUT, T>(next, B, row0, rows, smem, xn);
}

template <int SYS, typename T, int LAYOUT>
__global__ void __launch_bounds__(Tpb<SYS>::V) k_dyn_derivative(const __grid_constant__ cacto_sys_params P, const T* __restrict__ state,
                                                                T* __restrict__ Fu_out, int64_t B) {
  constexpr int NX = SysDims<SYS>::NX, NA = SysDims<SYS>::NA, NS = NX + 1, NT = Tpb<SYS>::V;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  const int64_t row0 = (int64_t)blockIdx.x * NT;
  const int rows = (int)min((int64_t)NT, B - row0);
  T x[NS], Fu[NS * NA];
  read_sample<NS, NT, LAYOUT, T>(state, B, row0, rows, smem, x);
  if ((int)threadIdx.x < rows) {
    sys_Fu<SYS, T>(P, x, Fu);
    if (P.normalize) {                                  // environment.py:106-107
#pragma unroll
      for (int i = 0; i < NX; ++i) {
        const T inv = T(1) / T(P.state_norm[i]);
#pragma unroll
        for (int j = 0; j < NA; ++j) Fu[i * NA + j] *= inv;
      }
    }
#pragma unroll
    for (int j = 0; j < NA; ++j) Fu[NX * NA + j] = T(0);
  }
  write_sample<NS * NA, NT, LAYOUT, T>(Fu_out, B, row0, rows, smem, Fu);
}

template <int SYS, typename T, int LAYOUT>
__global__ void __launch_bounds__(Tpb<SYS>::V) k_dyn_augmented(const __grid_constant__ cacto_sys_params P, const T* __restrict__ state,
                                                               const T* __restrict__ action, T* __restrict__ Fx_out,
                                                               T* __restrict__ Fu_out, int64_t B) {
  constexpr int NX = SysDims<SYS>::NX, NA = SysDims<SYS>::NA, NS = NX + 1, NT = Tpb<SYS>::V;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  const int64_t row0 = (int64_t)blockIdx.x * NT;
  const int rows = (int)min((int64_t)NT, B - row0);
  T x[NS], u[NA], Fx[NX * NX], Fu[NX * NA];
  read_sample2<NS, NA, NT, LAYOUT, T>(state, action, B, row0, rows, smem, x, u);
  if ((int)threadIdx.x < rows) sys_jac<SYS, T>(P, x, u, Fx, Fu);
  write_sample<NX * NX, NT, LAYOUT, T>(Fx_out, B, row0, rows, smem, Fx);
  write_sample<NX * NA, NT, LAYOUT, T>(Fu_out, B, row0, rows, smem, Fu);
}

template <int SYS, typename T, int LAYOUT>
__global__ void __launch_bounds__(Tpb<SYS>::V) k_ee(const __grid_constant__ cacto_sys_params P, const T* __restrict__ state,
                                                    T* __restrict__ ee, int64_t B) {
  constexpr int NX = SysDims<SYS>::NX, NS = NX + 1, NT = Tpb<SYS>::V;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  const int64_t row0 = (int64_t)blockIdx.x * NT;
  const int rows = (int)min((int64_t)NT, B - row0);
  T x[NS], p[3];
  read_sample<NS, NT, LAYOUT, T>(state, B, row0, rows, smem, x);
  if ((int)threadIdx.x < rows) sys_ee<SYS, T>(P, x, p);
  write_sample<3, NT, LAYOUT, T>(ee, B, row0, rows, smem, p);
}

template <int SYS, typename T, int LAYOUT>
__global__ void __launch_bounds__(Tpb<SYS>::V) k_reward(const __grid_constant__ cacto_sys_params P, const double* __restrict__ weights,
                                                        const T* __restrict__ state, const T* __restrict__ action, int plain_ucost,
                                                        T* __restrict__ reward, T* __restrict__ dr_da, int64_t B) {
  constexpr int NX = SysDims<SYS>::NX, NA = SysDims<SYS>::NA, NS = NX + 1, NT = Tpb<SYS>::V;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* smem = reinterpret_cast<T*>(smem_raw);
  const int64_t row0 = (int64_t)blockIdx.x * NT;
  const int rows = (int)min((int64_t)NT, B - row0);
  T x[NS], u[NA], g[NA];
  double w[8];
  read_sample<NS, NT, LAYOUT, T>(state, B, row0, rows, smem, x);
  if (action != nullptr) read_sample<NA, NT, LAYOUT, T>(action, B, row0, rows, smem, u);
  {
    double* sw = reinterpret_cast<double*>(smem_raw);
    read_sample<8, NT, LAYOUT_ROWS, double>(weights, B, row0, rows, sw, w);
  }
  T r = T(0);
  if ((int)threadIdx.x < rows) {
    r = sys_reward<SYS, T>(P, w, x, action != nullptr ? u : nullptr, plain_ucost != 0);
    if (dr_da != nullptr) sys_dr_da<SYS, T>(P, T(w[6]), u, g);
  }
  if ((int)threadIdx.x < rows) reward[row0 + threadIdx.x] = r;
  if (dr_da != nullptr) write_sample<NA, NT, LAYOUT, T>(dr_da, B, row0, rows, smem, g);
}

// ---- dispatch ------------------------------------------------------------------------------------
template <int SYS, typename T>
constexpr size_t tile_bytes() {
  constexpr int NX = SysDims<SYS>::NX, NA = SysDims<SYS>::NA;
  constexpr int W1 = pad_odd(NX * NX), W2 = pad_odd((NX + 1) * NA);
  constexpr size_t w = (W1 > W2 ? W1 : W2);
  constexpr size_t a = w * Tpb<SYS>::V * sizeof(T), b = (size_t)9 * Tpb<SYS>::V * sizeof(double);
  constexpr size_t c = (size_t)(pad_odd(NX + 1) + pad_odd(NA)) * Tpb<SYS>::V * sizeof(T) + 16;       // state and action tiles side by side
  return (a > b ? a : b) > c ? (a > b ? a : b) : c;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  if (bytes > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
  }
  return 0;
}

#define CACTO_FOR_SYSTEM(sys, MACRO)                                  \
  switch (sys) {                                                      \
    case CACTO_SINGLE_INTEGRATOR: MACRO(CACTO_SINGLE_INTEGRATOR); break; \
    case CACTO_DOUBLE_INTEGRATOR: MACRO(CACTO_DOUBLE_INTEGRATOR); break; \
    case CACTO_CAR: MACRO(CACTO_CAR); break;                          \
    case CACTO_CAR_PARK: MACRO(CACTO_CAR_PARK); break;                \
    case CACTO_MANIPULATOR: MACRO(CACTO_MANIPULATOR); break;          \
    case CACTO_UR5: MACRO(CACTO_UR5); break;                          \
    default: return CACTO_E_SYSTEM;                                   \
  }

static int check_common(const cacto_sys_params* p, int dtype, int layout, int64_t B) {
  if (p == nullptr) return CACTO_E_ARG;
  if (dtype != 0 && dtype != 1) return CACTO_E_DTYPE;
  if (layout != 0 && layout != 1) return CACTO_E_ARG;
  if (B < 0) return CACTO_E_SIZE;
  if (p->system < 0 || p->system > CACTO_UR5) return CACTO_E_SYSTEM;
  return 0;
}

template <int SYS, typename T, int LAYOUT>
static int launch_step(const cacto_sys_params& P, const void* s, const void* a, void* n, int64_t B, cudaStream_t st) {
  constexpr int NT = Tpb<SYS>::V;
  auto k = k_dyn_step<SYS, T, LAYOUT>;
  size_t sm = tile_bytes<SYS, T>();
  if (int e = set_smem(k, sm)) return e;
  k<<<(unsigned)((B + NT - 1) / NT), NT, sm, st>>>(P, (const T*)s, (const T*)a, (T*)n, B);
  CACTO_LAUNCH_CHECK();
  return 0;
}
template <int SYS, typename T, int LAYOUT>
static int launch_derivative(const cacto_sys_params& P, const void* s, void* Fu, int64_t B, cudaStream_t st) {
  constexpr int NT = Tpb<SYS>::V;
  auto k = k_dyn_derivative<SYS, T, LAYOUT>;
  size_t sm = tile_bytes<SYS, T>();
  if (int e = set_smem(k, sm)) return e;
  k<<<(unsigned)((B + NT - 1) / NT), NT, sm, st>>>(P, (const T*)s, (T*)Fu, B);
  CACTO_LAUNCH_CHECK();
  return 0;
}
template <int SYS, typename T, int LAYOUT>
static int launch_augmented(const cacto_sys_params& P, const void* s, const void* a, void* Fx, void* Fu, int64_t B, cudaStream_t st) {
  constexpr int NT = Tpb<SYS>::V;
  auto k = k_dyn_augmented<SYS, T, LAYOUT>;
  size_t sm = tile_bytes<SYS, T>();
  if (int e = set_smem(k, sm)) return e;
  k<<<(unsigned)((B + NT - 1) / NT), NT, sm, st>>>(P, (const T*)s, (const T*)a, (T*)Fx, (T*)Fu, B);
  CACTO_LAUNCH_CHECK();
  return 0;
}
template <int SYS, typename T, int LAYOUT>
static int launch_ee(const cacto_sys_params& P, const void* s, void* ee, int64_t B, cudaStream_t st) {
  constexpr int NT = Tpb<SYS>::V;
  auto k = k_ee<SYS, T, LAYOUT>;
  size_t sm = tile_bytes<SYS, T>();
  if (int e = set_smem(k, sm)) return e;
  k<<<(unsigned)((B + NT - 1) / NT), NT, sm, st>>>(P, (const T*)s, (T*)ee, B);
  CACTO_LAUNCH_CHECK();
  return 0;
}
template <int SYS, typename T, int LAYOUT>
static int launch_reward(const cacto_sys_params& P, const double* w, const void* s, const void* a, int plain, void* r, void* g,
                         int64_t B, cudaStream_t st) {
  constexpr int NT = Tpb<SYS>::V;
  auto k = k_reward<SYS, T, LAYOUT>;
  size_t sm = tile_bytes<SYS, T>();
  if (int e = set_smem(k, sm)) return e;
  k<<<(unsigned)((B + NT - 1) / NT), NT, sm, st>>>(P, w, (const T*)s, (const T*)a, plain, (T*)r, (T*)g, B);
  CACTO_LAUNCH_CHECK();
  return 0;
}

}  // namespace cacto

using namespace cacto;

#define DISPATCH_TL(SYS, FN, ...)                                                         \
  if (dtype == 0) {                                                                       \
    if (layout == 0) return FN<SYS, float, LAYOUT_ROWS>(__VA_ARGS__);                     \
    return FN<SYS, float, LAYOUT_SOA>(__VA_ARGS__);                                       \
  } else {                                                                                \
    if (layout == 0) return FN<SYS, double, LAYOUT_ROWS>(__VA_ARGS__);                    \
    return FN<SYS, double, LAYOUT_SOA>(__VA_ARGS__);                                      \
  }

extern "C" int cacto_dyn_step(const cacto_sys_params* p, int dtype, int layout, const void* state, const void* action,
                              void* state_next, int64_t B, void* stream) {
  if (int e = check_common(p, dtype, layout, B)) return e;
  if (B == 0) return 0;
  if (!state || !action || !state_next) return CACTO_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
#define M_(SYS) DISPATCH_TL(SYS, launch_step, *p, state, action, state_next, B, st)
  CACTO_FOR_SYSTEM(p->system, M_)
#undef M_
  return 0;
}

extern "C" int cacto_dyn_derivative(const cacto_sys_params* p, int dtype, int layout, const void* state, const void* action,
                                    void* Fu, int64_t B, void* stream) {
  (void)action;  // ds'/da does not depend on a for any of the six systems (environment.py:93-109)
  if (int e = check_common(p, dtype, layout, B)) return e;
  if (B == 0) return 0;
  if (!state || !Fu) return CACTO_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
#define M_(SYS) DISPATCH_TL(SYS, launch_derivative, *p, state, Fu, B, st)
  CACTO_FOR_SYSTEM(p->system, M_)
#undef M_
  return 0;
}

extern "C" int cacto_dyn_augmented(const cacto_sys_params* p, int dtype, int layout, const void* state, const void* action,
                                   void* Fx, void* Fu, int64_t B, void* stream) {
  if (int e = check_common(p, dtype, layout, B)) return e;
  if (B == 0) return 0;
  if (!state || !action || !Fx || !Fu) return CACTO_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
#define M_(SYS) DISPATCH_TL(SYS, launch_augmented, *p, state, action, Fx, Fu, B, st)
  CACTO_FOR_SYSTEM(p->system, M_)
#undef M_
  return 0;
}

extern "C" int cacto_ee_position(const cacto_sys_params* p, int dtype, int layout, const void* state, void* ee, int64_t B,
                                 void* stream) {
  if (int e = check_common(p, dtype, layout, B)) return e;
  if (B == 0) return 0;
  if (!state || !ee) return CACTO_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
#define M_(SYS) DISPATCH_TL(SYS, launch_ee, *p, state, ee, B, st)
  CACTO_FOR_SYSTEM(p->system, M_)
#undef M_
  return 0;
}

extern "C" int cacto_reward(const cacto_sys_params* p, int dtype, int layout, const double* weights, const void* state,
                            const void* action, int ur5_plain_ucost, void* reward, void* dr_da, int64_t B, void* stream) {
  if (int e = check_common(p, dtype, layout, B)) return e;
  if (B == 0) return 0;
  if (!state || !weights || !reward) return CACTO_E_ARG;
  if (dr_da && !action) return CACTO_E_ARG;
  cudaStream_t st = (cudaStream_t)stream;
#define M_(SYS) DISPATCH_TL(SYS, launch_reward, *p, weights, state, action, ur5_plain_ucost, reward, dr_da, B, st)
  CACTO_FOR_SYSTEM(p->system, M_)
#undef M_
  return 0;
}
