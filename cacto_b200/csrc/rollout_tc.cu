// rollout_tc.cu -- K1 on the 5th-generation tensor cores: the same fused actor + dynamics rollout as rollout.cu,
// with the 256x256 hidden layer on tcgen05.mma (kind::tf32, accumulator in TMEM) using 3xTF32 operand splitting
//     a = a_hi + a_lo,  w = w_hi + w_lo,  a*w ~= a_hi*w_hi + a_hi*w_lo + a_lo*w_hi   (error ~2^-21 relative)
// so that the result keeps fp32-class accuracy (the rollout parity gate is 1e-5; single-pass tf32/bf16 gives 1e-3).
//
// One CTA owns 256 rollouts = two M=128 tiles.  Warp roles (320 threads):
//   warps 0-7  workers: thread t <-> rollout t <-> TMEM lane t%128 of tile t/128.  Per K-chunk of 32 hidden units they
//              compute layer 1 (K = ns, CUDA cores) for their row, split it into hi/lo and store it into the tile's
//              A-operand buffer in the UMMA K-major no-swizzle layout; at the end of a step they read their accumulator
//              row with tcgen05.ld, apply bias + LeakyReLU, contract with W3 (layer 3) in registers and advance the
//              fp64 dynamics of their rollout.
//   warp 8     MMA issuer (one lane): for every chunk waits for the W2 chunk (TMA) and the two A chunks, issues
//              3 products x 4 k-steps of tcgen05.mma M128 N256 K8 per tile, commits completion to the mbarriers that free
//              the A / W2 buffers and publish the accumulators.
//   warp 9     TMA producer (one lane): streams the pre-split, pre-laid-out W2 image (hi and lo, 64 KB per chunk) from
//              L2 with cp.async.bulk into a 2-slot ring.
// W2 is re-laid out once per policy version by k_actor_tc_prepare (cacto_actor_tc_prepare).
#include "common.cuh"
#include "mlp.cuh"
#include "systems.cuh"

namespace cacto {

constexpr int TC_TILE = 128;                 // rollouts per UMMA tile (M)
constexpr int TC_TILES = 2;                  // tiles per CTA.  Measured: 2 tiles x KC 32, 1 CTA/SM = 7.9 ms; 1 tile x KC 16, 2 CTAs/SM = 10.2 ms
constexpr int TC_CTAS_PER_SM = TC_TILES == 1 ? 2 : 1;
constexpr int TC_ROLLOUTS = TC_TILE * TC_TILES;
constexpr int TC_WORKERS = TC_ROLLOUTS;      // worker threads
constexpr int TC_THREADS = TC_WORKERS + 64;  // + MMA warp + TMA warp
constexpr int TC_MMA_WARP = TC_WORKERS / 32;
constexpr int TC_TMEM_COLS = TC_TILES * ACTOR_H;
constexpr int TC_KC = TC_TILES == 1 ? 16 : 32;   // hidden units per K-chunk (sized so that the CTA's shared memory allows TC_CTAS_PER_SM)
constexpr int TC_NCHUNK = ACTOR_H / TC_KC;   // 8
constexpr int TC_A_IMG = TC_TILE * TC_KC;    // floats of one A image (hi or lo) of a chunk: 4096 (16 KB)
constexpr int TC_B_IMG = ACTOR_H * TC_KC;    // floats of one W2 image (hi or lo) of a chunk: 8192 (32 KB)
constexpr int TC_W2IMG_FLOATS = TC_NCHUNK * 2 * TC_B_IMG;   // 131072 floats (512 KB)
constexpr int NSP_TC = 8;

// UMMA K-major, no-swizzle canonical layout: 16-byte units of 4 consecutive k; 8 rows x 16 B = one 128-byte core
// matrix; core matrices of consecutive 8-row groups are SBO = 128 B apart, the next k-unit is LBO = (rows/8)*128 B away.
__host__ __device__ constexpr int umma_offset(int rows, int r, int k) {      // float index of element (r, k)
  return (k >> 2) * (rows / 8) * 32 + (r >> 3) * 32 + (r & 7) * 4 + (k & 3);
}

__device__ __forceinline__ float tf32_round(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// W2 (Keras [in k][out n]) -> per chunk kc: hi image then lo image of B[n][k] = W2[kc*32 + k][n] in the UMMA layout.
__global__ void __launch_bounds__(256) k_actor_tc_prepare(const float* __restrict__ actor, int ns, int na, float* __restrict__ img) {
  const ActorLayout L(ns, na);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ACTOR_H * ACTOR_H) return;
  const int k = i / ACTOR_H, n = i - k * ACTOR_H;          // coalesced read of W2[k][n]
  const float w = actor[L.W2 + i];
  const float hi = tf32_round(w);
  const int kc = k / TC_KC, kk = k - kc * TC_KC;
  float* base = img + (size_t)kc * 2 * TC_B_IMG;
  const int o = umma_offset(ACTOR_H, n, kk);
  base[o] = hi;
  base[TC_B_IMG + o] = w - hi;
}

struct TcSmem {
  alignas(1024) float B[2][2 * TC_B_IMG];            // W2 chunk ring: [slot][hi | lo]               128 KB
  alignas(1024) float A[TC_TILES][2 * TC_A_IMG];     // A chunk per tile: [hi | lo]                  64 KB
  alignas(16) float W1[CACTO_MAX_NS][ACTOR_H];       // layer-1 weights                              13 KB
  alignas(16) float b1[ACTOR_H];
  alignas(16) float4 head[ACTOR_H];                  // (b2[c], W3[c][0..2]) ... see HEAD_W below
  alignas(16) float W3x[ACTOR_H][4];                 // W3[c][3..5] for na > 3 (UR5)
  float b3[8];
  uint64_t b_full[2], b_empty[2], a_full[TC_TILES], a_empty[TC_TILES], d_full[TC_TILES];
  uint32_t tmem_base;
  int tmax;
};

__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* b) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  const uint32_t mb = smem_u32(b);
  uint32_t done = 0, spins = 0;
  while (!done) {
    if (++spins > (1u << 26)) __trap();      // fail loudly instead of hanging the GPU
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(mb), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void umma_commit(uint64_t* b) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) |
         ((uint64_t)1 << 46);                // version 1 (Blackwell), base offset 0, SWIZZLE_NONE
}
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int SYS>
__global__ void __launch_bounds__(TC_THREADS, TC_CTAS_PER_SM) k_rollout_tc(const __grid_constant__ cacto_sys_params P, const float* __restrict__ actor,
                                                              const float* __restrict__ w2img, const double* __restrict__ ics,
                                                              const int32_t* __restrict__ horizon, int T_max, double* __restrict__ states,
                                                              double* __restrict__ controls, int32_t* __restrict__ flags,
                                                              double* __restrict__ rewards, int64_t B, int num_sms) {
  constexpr int NX = SysDims<SYS>::NX, NA = SysDims<SYS>::NA, NS = NX + 1;
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  TcSmem& sm = *reinterpret_cast<TcSmem*>(smem_raw);
  const ActorLayout L(NS, NA);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool worker = tid < TC_WORKERS;
  const int tile = (tid / TC_TILE) & (TC_TILES - 1), row = tid % TC_TILE;
  const int64_t b = (int64_t)blockIdx.x * TC_ROLLOUTS + tid;
  const bool owner = worker && b < B;

  // ---- one-time setup: small weights to shared memory, barriers, TMEM
  for (int i = tid; i < NS * ACTOR_H; i += TC_THREADS) sm.W1[i / ACTOR_H][i % ACTOR_H] = actor[L.W1 + i];
  for (int c = tid; c < ACTOR_H; c += TC_THREADS) {
    sm.b1[c] = actor[L.b1 + c];
    sm.head[c] = make_float4(actor[L.b2 + c], actor[L.W3 + c * NA + 0], NA > 1 ? actor[L.W3 + c * NA + 1] : 0.f,
                             NA > 2 ? actor[L.W3 + c * NA + 2] : 0.f);
    for (int j = 3; j < 7; ++j) sm.W3x[c][j - 3] = (j < NA) ? actor[L.W3 + c * NA + j] : 0.f;
  }
  if (tid < 8) sm.b3[tid] = tid < NA ? actor[L.b3 + tid] : 0.f;
  if (tid == 0) {
    sm.tmax = 0;
    for (int s = 0; s < 2; ++s) { mbar_init(&sm.b_full[s], 1); mbar_init(&sm.b_empty[s], 1); }
    for (int m = 0; m < TC_TILES; ++m) { mbar_init(&sm.a_full[m], TC_TILE); mbar_init(&sm.a_empty[m], 1); mbar_init(&sm.d_full[m], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == TC_MMA_WARP) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(TC_TMEM_COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  __syncthreads();

  double x[NS];
  int h = 0, ok = 1;
  if (owner) {
#pragma unroll
    for (int j = 0; j < NS; ++j) {
      x[j] = ics[b * NS + j];
      states[(int64_t)j * B + b] = x[j];
    }
    h = min(max(horizon[b], 0), T_max);
    atomicMax(&sm.tmax, h);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int tmax = sm.tmax;
  const uint32_t tmem = sm.tmem_base;
  if (TC_CTAS_PER_SM == 2 && ((blockIdx.x / (unsigned)num_sms) & 1u)) {
    // CTAs i and i + num_sms share an SM in the first wave: start the second one about half a step later so that the
    // MMAs of one CTA fill the epilogue / dynamics bubble of the other (later waves are staggered by completion order).
    const long long t0 = clock64();
    while (clock64() - t0 < 9000) {}
  }

  if (worker) {
    // =================================================================== workers
    float* Ah = &sm.A[tile][0];
    float* Al = Ah + TC_A_IMG;
    const uint32_t d_taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(tile * ACTOR_H);
    for (int t = 0; t < tmax; ++t) {
      const bool live = owner && ok && t < h;
      float xn[NS];
#pragma unroll
      for (int j = 0; j < NS; ++j) xn[j] = live ? normalize_component(P, j, (float)x[j]) : 0.f;
      // ---- layer 1 chunk by chunk into the A operand buffer
      for (int kc = 0; kc < TC_NCHUNK; ++kc) {
        const int q = t * TC_NCHUNK + kc;
        if (q > 0) mbar_wait(&sm.a_empty[tile], (uint32_t)((q - 1) & 1));
#pragma unroll
        for (int ku = 0; ku < TC_KC / 4; ++ku) {
          const int c = kc * TC_KC + ku * 4;
          float4 z = *reinterpret_cast<const float4*>(&sm.b1[c]);
#pragma unroll
          for (int j = 0; j < NS; ++j) {
            const float4 w = *reinterpret_cast<const float4*>(&sm.W1[j][c]);
            z.x = fmaf(xn[j], w.x, z.x); z.y = fmaf(xn[j], w.y, z.y); z.z = fmaf(xn[j], w.z, z.z); z.w = fmaf(xn[j], w.w, z.w);
          }
          z.x = leaky(z.x); z.y = leaky(z.y); z.z = leaky(z.z); z.w = leaky(z.w);
          const float4 hi = make_float4(tf32_round(z.x), tf32_round(z.y), tf32_round(z.z), tf32_round(z.w));
          const int o = ku * (TC_TILE / 8) * 32 + (row >> 3) * 32 + (row & 7) * 4;
          *reinterpret_cast<float4*>(Ah + o) = hi;
          *reinterpret_cast<float4*>(Al + o) = make_float4(z.x - hi.x, z.y - hi.y, z.z - hi.z, z.w - hi.w);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // my generic-proxy stores -> visible to the UMMA (async proxy)
        mbar_arrive(&sm.a_full[tile]);
      }
      // ---- epilogue: accumulator row -> bias, LeakyReLU, layer 3 in registers
      mbar_wait(&sm.d_full[tile], (uint32_t)(t & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float act[NA];
#pragma unroll
      for (int j = 0; j < NA; ++j) act[j] = sm.b3[j];
#pragma unroll 1
      for (int c0 = 0; c0 < ACTOR_H; c0 += 16) {
        uint32_t v[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(d_taddr + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float4 hw = sm.head[c0 + i];
          const float hval = leaky(__uint_as_float(v[i]) + hw.x);
          act[0] = fmaf(hval, hw.y, act[0]);
          if (NA > 1) act[1] = fmaf(hval, hw.z, act[1]);
          if (NA > 2) act[2] = fmaf(hval, hw.w, act[2]);
          if (NA > 3) {
            const float4 wx = *reinterpret_cast<const float4*>(&sm.W3x[c0 + i][0]);
            act[3] = fmaf(hval, wx.x, act[3]);
            if (NA > 4) act[4] = fmaf(hval, wx.y, act[4]);
            if (NA > 5) act[5] = fmaf(hval, wx.z, act[5]);
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");    // my TMEM reads precede the next step's MMAs (ordered via a_full)
      // ---- dynamics
      if (live) {
        double u[NA], xnext[NS];
#pragma unroll
        for (int j = 0; j < NA; ++j) {
          u[j] = (double)act[j];
          controls[((int64_t)t * NA + j) * B + b] = u[j];
        }
        if (rewards != nullptr) rewards[(int64_t)t * B + b] = sys_reward<SYS, double>(P, P.w_running, x, u, false);
        sys_step<SYS, double>(P, x, u, xnext);
        xnext[NX] = x[NX] + P.dt;
        bool nan = false;
#pragma unroll
        for (int j = 0; j < NS; ++j) {
          x[j] = xnext[j];
          nan |= (xnext[j] != xnext[j]);
          states[((int64_t)(t + 1) * NS + j) * B + b] = xnext[j];
        }
        if (nan) ok = 0;
        if (rewards != nullptr && t + 1 == h && !nan)
          rewards[(int64_t)(t + 1) * B + b] = sys_reward<SYS, double>(P, P.w_terminal, x, (const double*)nullptr, false);
      }
    }
    if (owner) flags[b] = ok;
  } else if (warp == TC_MMA_WARP) {
    // =================================================================== MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(ACTOR_H >> 3) << 17) | ((uint32_t)(TC_TILE >> 4) << 24);
      constexpr uint32_t lboA = (TC_TILE / 8) * 128, lboB = (ACTOR_H / 8) * 128, sbo = 128;
      const int total = tmax * TC_NCHUNK;
      for (int g = 0; g < total; ++g) {
        const int s = g & 1, kc = g % TC_NCHUNK;
        mbar_wait(&sm.b_full[s], (uint32_t)((g >> 1) & 1));
        const uint32_t bh = smem_u32(&sm.B[s][0]), bl = bh + TC_B_IMG * 4;
#pragma unroll
        for (int m = 0; m < TC_TILES; ++m) {
          mbar_wait(&sm.a_full[m], (uint32_t)(g & 1));
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t ah = smem_u32(&sm.A[m][0]), al = ah + TC_A_IMG * 4;
          const uint32_t d = tmem + (uint32_t)(m * ACTOR_H);
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            const uint32_t a0 = (p == 2) ? al : ah, b0 = (p == 1) ? bl : bh;
#pragma unroll
            for (int ks = 0; ks < TC_KC / 8; ++ks) {
              umma_tf32(d, umma_desc(a0 + ks * 2 * lboA, lboA, sbo), umma_desc(b0 + ks * 2 * lboB, lboB, sbo), idesc,
                        (kc > 0 || p > 0 || ks > 0) ? 1u : 0u);
            }
          }
          umma_commit(&sm.a_empty[m]);                     // A chunk of tile m may be overwritten once these MMAs are done
          if (kc == TC_NCHUNK - 1) umma_commit(&sm.d_full[m]);
        }
        umma_commit(&sm.b_empty[s]);                       // W2 ring slot may be refilled
      }
    }
    __syncwarp();
  } else {
    // =================================================================== TMA producer
    if (lane == 0) {
      const int total = tmax * TC_NCHUNK;
      constexpr uint32_t bytes = 2 * TC_B_IMG * 4;         // hi + lo image of one chunk: 64 KB
      for (int g = 0; g < total; ++g) {
        const int s = g & 1, kc = g % TC_NCHUNK;
        if (g >= 2) mbar_wait(&sm.b_empty[s], (uint32_t)(((g - 2) >> 1) & 1));
        const uint32_t mb = smem_u32(&sm.b_full[s]);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(&sm.B[s][0])),
                     "l"(w2img + (size_t)kc * 2 * TC_B_IMG), "r"(bytes), "r"(mb)
                     : "memory");
      }
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == TC_MMA_WARP) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(TC_TMEM_COLS));
}

template <int SYS>
static int launch_rollout_tc(const cacto_sys_params& P, const float* actor, const float* w2img, const double* ics, const int32_t* horizon,
                             int T_max, double* states, double* controls, int32_t* flags, double* rewards, int64_t B, cudaStream_t st) {
  auto k = k_rollout_tc<SYS>;
  const size_t sm = sizeof(TcSmem);
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (e != cudaSuccess) return (int)e;
  int dev = 0, num_sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
  k<<<(unsigned)((B + TC_ROLLOUTS - 1) / TC_ROLLOUTS), TC_THREADS, sm, st>>>(P, actor, w2img, ics, horizon, T_max, states, controls, flags,
                                                                              rewards, B, num_sms);
  CACTO_LAUNCH_CHECK();
  return 0;
}

}  // namespace cacto

using namespace cacto;

extern "C" int64_t cacto_actor_tc_image_floats(void) { return TC_W2IMG_FLOATS; }

extern "C" int cacto_actor_tc_prepare(const float* actor_params, int32_t ns, int32_t na, float* w2img, void* stream) {
  if (!actor_params || !w2img) return CACTO_E_ARG;
  if (ns < 2 || ns > CACTO_MAX_NS || na < 1 || na > CACTO_MAX_NA) return CACTO_E_SIZE;
  k_actor_tc_prepare<<<(ACTOR_H * ACTOR_H + 255) / 256, 256, 0, (cudaStream_t)stream>>>(actor_params, ns, na, w2img);
  CACTO_LAUNCH_CHECK();
  return 0;
}

extern "C" int cacto_rollout_tc(const cacto_sys_params* p, const float* actor_params, const float* w2img, const double* ics,
                                const int32_t* horizon, int32_t T_max, double* states, double* controls, int32_t* flags,
                                double* rewards, int64_t B, void* stream) {
  if (!p) return CACTO_E_ARG;
  if (B < 0 || T_max < 0) return CACTO_E_SIZE;
  if (B == 0) return 0;
  if (!actor_params || !w2img || !ics || !horizon || !states || !flags || (T_max > 0 && !controls)) return CACTO_E_ARG;
  if ((reinterpret_cast<uintptr_t>(w2img) & 127) || (reinterpret_cast<uintptr_t>(actor_params) & 15)) return CACTO_E_ALIGN;
  cudaStream_t st = (cudaStream_t)stream;
  switch (p->system) {
    case CACTO_SINGLE_INTEGRATOR: return launch_rollout_tc<CACTO_SINGLE_INTEGRATOR>(*p, actor_params, w2img, ics, horizon, T_max, states, controls, flags, rewards, B, st);
    case CACTO_DOUBLE_INTEGRATOR: return launch_rollout_tc<CACTO_DOUBLE_INTEGRATOR>(*p, actor_params, w2img, ics, horizon, T_max, states, controls, flags, rewards, B, st);
    case CACTO_CAR: return launch_rollout_tc<CACTO_CAR>(*p, actor_params, w2img, ics, horizon, T_max, states, controls, flags, rewards, B, st);
    case CACTO_CAR_PARK: return launch_rollout_tc<CACTO_CAR_PARK>(*p, actor_params, w2img, ics, horizon, T_max, states, controls, flags, rewards, B, st);
    case CACTO_MANIPULATOR: return launch_rollout_tc<CACTO_MANIPULATOR>(*p, actor_params, w2img, ics, horizon, T_max, states, controls, flags, rewards, B, st);
    case CACTO_UR5: return launch_rollout_tc<CACTO_UR5>(*p, actor_params, w2img, ics, horizon, T_max, states, controls, flags, rewards, B, st);
    default: return CACTO_E_SYSTEM;
  }
}
