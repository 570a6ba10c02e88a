// gemm16_proto.cu -- PROTOTYPE for the round-2 large-batch update ("Plan for the large-batch update on tcgen05").
// NOT part of the library.  k_gemm16 (forward / input-gradient form) has run on a B200 (profiles/r1_gemm16_proto.txt): all cases OK,
// 16384 x 256 x 256 in 54 us.  k_wgrad16 (weight-gradient form, reduction over the batch) was added afterwards: compiled, not yet run.
// Self-checking standalone program:
//     nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o gemm16_proto profiles/scripts/gemm16_proto.cu && ./gemm16_proto
//
// C[M x N] = A[M x K] * W[K x N], fp32 in global memory, computed on tcgen05.mma kind::f16 with fp16 hi/lo operand splitting
// (a w ~ a_hi w_hi + a_hi w_lo + a_lo w_hi, fp32 accumulation in TMEM) -- the arithmetic of rollout_tc16.cu, whose descriptor /
// instruction-descriptor / commit / TMEM code this file reuses, generalised from "A produced by layer 1" to "A read from HBM".
// One CTA (128 threads) per 128-row tile of A; K is walked in chunks of 64: every thread converts the chunk of its own row into
// the hi / lo images (UMMA K-major, no swizzle: 16-byte units of 8 consecutive k, 8 rows x 16 B = one 128-byte core matrix,
// 8-row groups SBO = 128 B apart, k-units LBO = (rows / 8) * 128 B apart), the CTA converts the W chunk cooperatively, thread 0
// issues 4 k-steps x 3 UMMAs and commits to an mbarrier.  Single-buffered on purpose (correctness first); the production
// kernel needs the A / W rings and the two tile pipelines of rollout_tc16.cu.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

constexpr int TILE_M = 128, KC = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) | ((uint64_t)(128u >> 4) << 32) | ((uint64_t)1 << 46);
}
// byte offset of element (row, kk) inside an image of `rows` rows x KC k-values of fp16
__host__ __device__ constexpr int img_offset(int rows, int row, int kk) {
  return (kk >> 3) * (rows / 8) * 128 + (row >> 3) * 128 + (row & 7) * 16 + (kk & 7) * 2;
}
__device__ __forceinline__ void split_store(unsigned char* hi_img, unsigned char* lo_img, int off, float x) {
  const __half h = __float2half_rn(x);
  *reinterpret_cast<__half*>(hi_img + off) = h;
  *reinterpret_cast<__half*>(lo_img + off) = __float2half_rn(x - __half2float(h));
}

template <int N>
__global__ void __launch_bounds__(128) k_gemm16(const float* __restrict__ A, const float* __restrict__ W, float* __restrict__ C, int M, int K,
                                                float scale_a, float scale_w) {
  static_assert(N % 16 == 0 && N >= 16 && N <= 256, "UMMA N");
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int A_IMG = TILE_M * KC * 2, B_IMG = N * KC * 2;
  unsigned char* Ah = smem;
  unsigned char* Al = Ah + A_IMG;
  unsigned char* Bh = Al + A_IMG;
  unsigned char* Bl = Bh + B_IMG;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t row = (int64_t)blockIdx.x * TILE_M + tid;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base;
  constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);   // f16 x f16 -> f32, K-major
  constexpr uint32_t lboA = (TILE_M / 8) * 128, lboB = (N / 8) * 128;
  uint32_t parity = 0;
  for (int k0 = 0; k0 < K; k0 += KC) {
    // ---- operand images of this K-chunk
    if (row < M) {
      const float4* a4 = reinterpret_cast<const float4*>(A + row * K + k0);
#pragma unroll 4
      for (int q = 0; q < KC / 4; ++q) {
        const float4 v = a4[q];
        split_store(Ah, Al, img_offset(TILE_M, tid, 4 * q + 0), v.x * scale_a);
        split_store(Ah, Al, img_offset(TILE_M, tid, 4 * q + 1), v.y * scale_a);
        split_store(Ah, Al, img_offset(TILE_M, tid, 4 * q + 2), v.z * scale_a);
        split_store(Ah, Al, img_offset(TILE_M, tid, 4 * q + 3), v.w * scale_a);
      }
    } else {
      for (int kk = 0; kk < KC; ++kk) split_store(Ah, Al, img_offset(TILE_M, tid, kk), 0.f);
    }
    for (int i = tid; i < KC * N; i += 128) {            // W[k0 + kk][n], coalesced over n
      const int kk = i / N, n = i - kk * N;
      split_store(Bh, Bl, img_offset(N, n, kk), W[(int64_t)(k0 + kk) * N + n] * scale_w);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy smem writes -> visible to the UMMA (async proxy)
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
      for (int ks = 0; ks < KC / 16; ++ks) {
        const uint64_t ah = make_desc(smem_u32(Ah) + ks * 2 * lboA, lboA), al = make_desc(smem_u32(Al) + ks * 2 * lboA, lboA);
        const uint64_t bh = make_desc(smem_u32(Bh) + ks * 2 * lboB, lboB), bl = make_desc(smem_u32(Bl) + ks * 2 * lboB, lboB);
        const uint64_t da[3] = {ah, ah, al}, db[3] = {bh, bl, bh};
        for (int p = 0; p < 3; ++p) {
          const uint32_t acc = (k0 == 0 && ks == 0 && p == 0) ? 0u : 1u;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase),
              "l"(da[p]), "l"(db[p]), "r"(idesc), "r"(acc)
              : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    // ---- the images may be overwritten once the MMAs that read them have completed
    uint32_t done = 0, spins = 0;
    while (!done) {
      if (++spins > (1u << 24)) __trap();
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done)
                   : "r"(smem_u32(&bar)), "r"(parity)
                   : "memory");
    }
    parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  // ---- epilogue: thread tid <-> TMEM lane tid, 8 columns at a time
  const float unscale = 1.f / (scale_a * scale_w);
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (row < M) {
      float4* c4 = reinterpret_cast<float4*>(C + row * N + c0);
      c4[0] = make_float4(__uint_as_float(v[0]) * unscale, __uint_as_float(v[1]) * unscale, __uint_as_float(v[2]) * unscale,
                          __uint_as_float(v[3]) * unscale);
      c4[1] = make_float4(__uint_as_float(v[4]) * unscale, __uint_as_float(v[5]) * unscale, __uint_as_float(v[6]) * unscale,
                          __uint_as_float(v[7]) * unscale);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256));
}

// Weight gradient dW[F x N] += X^T D with X [B x F] (layer input) and D [B x N] (upstream gradient), both row-major fp32: the
// reduction runs over the BATCH.  NOT YET RUN on a GPU (added after the round's GPU budget was spent; it differs from k_gemm16,
// which has run, only in the index mapping of the image builds and in the epilogue).  UMMA view: M = 128 features of a feature
// tile, N = N, K = samples.  Both operands are batch-major in memory, so both image builds transpose: a thread reads X / D
// coalesced along the feature / column index and scatters fp16 pairs into the K-major images (k = sample).  A CTA owns
// (feature tile, batch slice), accumulates its slice in TMEM and adds the 128 x N result to dW with 16-byte vector reductions
// (gridDim.y batch slices).  Rows of X / D beyond B are zero-filled.
template <int N>
__global__ void __launch_bounds__(128) k_wgrad16(const float* __restrict__ X, const float* __restrict__ D, float* __restrict__ dW, int B, int F,
                                                 int slice, float scale_x, float scale_d) {
  static_assert(N % 16 == 0 && N >= 16 && N <= 256, "UMMA N");
  extern __shared__ __align__(1024) unsigned char smem[];
  constexpr int A_IMG = TILE_M * KC * 2, B_IMG = N * KC * 2;
  unsigned char* Ah = smem;
  unsigned char* Al = Ah + A_IMG;
  unsigned char* Bh = Al + A_IMG;
  unsigned char* Bl = Bh + B_IMG;
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int f0 = blockIdx.x * TILE_M;                       // feature tile (F % 128 == 0)
  const int b_begin = blockIdx.y * slice, b_end = min(B, b_begin + slice);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(256));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tbase = tmem_base;
  constexpr uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24);
  constexpr uint32_t lboA = (TILE_M / 8) * 128, lboB = (N / 8) * 128;
  uint32_t parity = 0;
  bool first = true;
  for (int b0 = b_begin; b0 < b_end; b0 += KC) {
    for (int i = tid; i < KC * TILE_M; i += 128) {          // A[m][kk] = X[b0 + kk][f0 + m], coalesced over m
      const int kk = i / TILE_M, m = i - kk * TILE_M;
      const float v = (b0 + kk < b_end) ? X[(int64_t)(b0 + kk) * F + f0 + m] * scale_x : 0.f;
      split_store(Ah, Al, img_offset(TILE_M, m, kk), v);
    }
    for (int i = tid; i < KC * N; i += 128) {               // B[n][kk] = D[b0 + kk][n], coalesced over n
      const int kk = i / N, n = i - kk * N;
      const float v = (b0 + kk < b_end) ? D[(int64_t)(b0 + kk) * N + n] * scale_d : 0.f;
      split_store(Bh, Bl, img_offset(N, n, kk), v);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
      for (int ks = 0; ks < KC / 16; ++ks) {
        const uint64_t ah = make_desc(smem_u32(Ah) + ks * 2 * lboA, lboA), al = make_desc(smem_u32(Al) + ks * 2 * lboA, lboA);
        const uint64_t bh = make_desc(smem_u32(Bh) + ks * 2 * lboB, lboB), bl = make_desc(smem_u32(Bl) + ks * 2 * lboB, lboB);
        const uint64_t da[3] = {ah, ah, al}, db[3] = {bh, bl, bh};
        for (int p = 0; p < 3; ++p) {
          const uint32_t acc = (first && ks == 0 && p == 0) ? 0u : 1u;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tbase),
              "l"(da[p]), "l"(db[p]), "r"(idesc), "r"(acc)
              : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    first = false;
    uint32_t done = 0, spins = 0;
    while (!done) {
      if (++spins > (1u << 24)) __trap();
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(done)
                   : "r"(smem_u32(&bar)), "r"(parity)
                   : "memory");
    }
    parity ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }
  if (b_begin < b_end) {                                    // thread tid <-> feature f0 + tid
    const float unscale = 1.f / (scale_x * scale_d);
    for (int c0 = 0; c0 < N; c0 += 8) {
      uint32_t v[8];
      const uint32_t taddr = tbase + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                   : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      float4* d4 = reinterpret_cast<float4*>(dW + (int64_t)(f0 + tid) * N + c0);
      atomicAdd(d4, make_float4(__uint_as_float(v[0]) * unscale, __uint_as_float(v[1]) * unscale, __uint_as_float(v[2]) * unscale,
                                __uint_as_float(v[3]) * unscale));
      atomicAdd(d4 + 1, make_float4(__uint_as_float(v[4]) * unscale, __uint_as_float(v[5]) * unscale, __uint_as_float(v[6]) * unscale,
                                    __uint_as_float(v[7]) * unscale));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(256));
}

static float pow2_scale_for(const std::vector<float>& x, int target_exp) {      // 2^s with max|x| 2^s ~ 2^target_exp
  float m = 0.f;
  for (float v : x) m = fmaxf(m, fabsf(v));
  if (m == 0.f) return 1.f;
  int e;
  frexpf(m, &e);                                    // m = f 2^e, f in [0.5, 1)
  return ldexpf(1.f, target_exp - e);
}

template <int N>
static int run(int M, int K) {
  std::vector<float> A((size_t)M * K), W((size_t)K * N), C((size_t)M * N);
  srand(1);
  for (auto& x : A) x = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& x : W) x = ((float)rand() / RAND_MAX * 2 - 1) * 0.15f;
  const float sa = pow2_scale_for(A, 11), sw = pow2_scale_for(W, 13);      // |a'| < 2^11 as in K1 (guard +-2047), |w'| < 2^13
  float *dA, *dW, *dC;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dC, C.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  const size_t sm = 2 * (size_t)TILE_M * KC * 2 + 2 * (size_t)N * KC * 2;
  cudaFuncSetAttribute(k_gemm16<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  const int grid = (M + TILE_M - 1) / TILE_M;
  k_gemm16<N><<<grid, 128, sm>>>(dA, dW, dC, M, K, sa, sw);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d: CUDA error: %s\n", N, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(C.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int m = 0; m < M; m += 37)                                           // a sample of rows, all columns
    for (int n = 0; n < N; ++n) {
      double r = 0;
      for (int k = 0; k < K; ++k) r += (double)A[(size_t)m * K + k] * W[(size_t)k * N + n];
      maxerr = fmax(maxerr, fabs(r - C[(size_t)m * N + n]));
      maxref = fmax(maxref, fabs(r));
    }
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0); cudaEventCreate(&t1);
  cudaEventRecord(t0);
  for (int i = 0; i < 20; ++i) k_gemm16<N><<<grid, 128, sm>>>(dA, dW, dC, M, K, sa, sw);
  cudaEventRecord(t1); cudaEventSynchronize(t1);
  float ms = 0;
  cudaEventElapsedTime(&ms, t0, t1);
  ms /= 20;
  const bool ok = maxerr <= 2e-6 * fmax(1.0, maxref) * sqrt((double)K);
  printf("M=%d N=%d K=%d: max abs err %.3e (max |ref| %.3f) %s | %.1f us = %.1f TFLOP/s algorithmic (single-buffered prototype)\n", M, N, K, maxerr,
         maxref, ok ? "OK" : "FAIL", ms * 1e3, 2.0 * M * N * K / (ms * 1e-3) / 1e12);
  cudaFree(dA); cudaFree(dW); cudaFree(dC);
  return ok ? 0 : 1;
}

template <int N>
static int run_wgrad(int B, int F) {
  std::vector<float> X((size_t)B * F), D((size_t)B * N), dW((size_t)F * N);
  srand(2);
  for (auto& x : X) x = (float)rand() / RAND_MAX * 2 - 1;
  for (auto& x : D) x = ((float)rand() / RAND_MAX * 2 - 1) * 1e-3f;         // small upstream gradients: exercises the scaling
  const float sx = pow2_scale_for(X, 11), sd = pow2_scale_for(D, 13);
  float *dX, *dD, *ddW;
  cudaMalloc(&dX, X.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&ddW, dW.size() * 4);
  cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(ddW, 0, dW.size() * 4);
  const size_t sm = 2 * (size_t)TILE_M * KC * 2 + 2 * (size_t)N * KC * 2;
  cudaFuncSetAttribute(k_wgrad16<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
  const int ftiles = F / TILE_M;
  int nsplit = (148 * 2) / ftiles;                                          // ~2 CTAs per SM
  int slice = ((B + nsplit - 1) / nsplit + KC - 1) / KC * KC;
  nsplit = (B + slice - 1) / slice;
  k_wgrad16<N><<<dim3(ftiles, nsplit), 128, sm>>>(dX, dD, ddW, B, F, slice, sx, sd);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("wgrad N=%d: CUDA error: %s\n", N, cudaGetErrorString(e)); return 1; }
  cudaMemcpy(dW.data(), ddW, dW.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  for (int f = 0; f < F; f += 7)
    for (int n = 0; n < N; n += 5) {
      double r = 0;
      for (int b = 0; b < B; ++b) r += (double)X[(size_t)b * F + f] * D[(size_t)b * N + n];
      maxerr = fmax(maxerr, fabs(r - dW[(size_t)f * N + n]));
      maxref = fmax(maxref, fabs(r));
    }
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0); cudaEventCreate(&t1);
  cudaEventRecord(t0);
  for (int i = 0; i < 20; ++i) k_wgrad16<N><<<dim3(ftiles, nsplit), 128, sm>>>(dX, dD, ddW, B, F, slice, sx, sd);
  cudaEventRecord(t1); cudaEventSynchronize(t1);
  float ms = 0;
  cudaEventElapsedTime(&ms, t0, t1);
  ms /= 20;
  const bool ok = maxerr <= 2e-6 * maxref * sqrt((double)B) / 8 + 1e-12;     // fp32-class: random-walk bound over B terms
  printf("wgrad B=%d F=%d N=%d (%d x %d CTAs): max abs err %.3e (max |ref| %.3e) %s | %.1f us = %.1f TFLOP/s algorithmic\n", B, F, N, ftiles,
         nsplit, maxerr, maxref, ok ? "OK" : "FAIL", ms * 1e3, 2.0 * B * F * N / (ms * 1e-3) / 1e12);
  cudaFree(dX); cudaFree(dD); cudaFree(ddW);
  return ok ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run<256>(16384, 256);      // actor hidden layer at the config-3 batch
  bad += run<128>(16384, 128);      // critic 128 x 128
  bad += run<64>(16384, 64);        // critic 64 x 64
  bad += run<256>(1000, 256);       // ragged last tile
  bad += run_wgrad<256>(16384, 256);   // dW2 of the actor
  bad += run_wgrad<128>(16384, 128);   // critic 128 x 128
  bad += run_wgrad<256>(5000, 256);    // ragged batch
  return bad;
}
