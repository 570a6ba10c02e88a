// Microbenchmark: issue rate of FFMA, packed FFMA2 (fma.rn.f32x2), F2FP pack (cvt.rn.f16x2.f32), LOP3+FADD split, per SMSP.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N_IT 4096
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, long long* clk, float seed) {
  float a[8]; unsigned long long p[8]; uint32_t h[8];
  for (int i = 0; i < 8; ++i) { a[i] = seed + i + threadIdx.x; p[i] = ((unsigned long long)__float_as_uint(a[i]) << 32) | __float_as_uint(a[i] * 0.5f); h[i] = 0; }
  const float m = 0.999f, c = 0.001f;
  unsigned long long m2 = ((unsigned long long)__float_as_uint(m) << 32) | __float_as_uint(m);
  unsigned long long c2 = ((unsigned long long)__float_as_uint(c) << 32) | __float_as_uint(c);
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < N_IT; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], m, c);
    } else if (MODE == 1) {
#pragma unroll
      for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(m2), "l"(c2));
    } else if (MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(a[i]), "f"(a[(i + 1) & 7])); a[i] = __uint_as_float(h[i] ^ 0x3f000000u); }
    } else if (MODE == 3) {   // mixed: 4 FFMA2 + 4 LOP3 + 4 FADD
#pragma unroll
      for (int i = 0; i < 4; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(m2), "l"(c2));
#pragma unroll
      for (int i = 0; i < 4; ++i) { float hi = __uint_as_float(__float_as_uint(a[i]) & 0xffffe000u); a[i + 4] = a[i] - hi + a[i + 4]; }
    } else if (MODE == 4) {   // mixed: 8 FFMA + 4 LOP3 + 4 FADD
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], m, c);
#pragma unroll
      for (int i = 0; i < 4; ++i) { float hi = __uint_as_float(__float_as_uint(a[i]) & 0xffffe000u); a[i + 4] = a[i] - hi + a[i + 4]; }
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i] + __uint_as_float((uint32_t)p[i]) + __uint_as_float((uint32_t)(p[i] >> 32)) + h[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* clk; cudaMalloc(&out, 148 * 4 * 256 * 4); cudaMalloc(&clk, 148 * 4 * 8);
  long long h[148 * 4];
  const char* names[] = {"FFMA x8", "FFMA2 x8", "F2FP+LOP3 x8", "4 FFMA2 + 4 LOP3 + 8 FADD", "8 FFMA + 4 LOP3 + 8 FADD"};
  for (int mode = 0; mode < 5; ++mode) for (int bps = 1; bps <= 2; ++bps) {
    int nb = 148 * bps;
    for (int rep = 0; rep < 2; ++rep) {
      if (mode == 0) k<0><<<nb, 256>>>(out, clk, 1.f); if (mode == 1) k<1><<<nb, 256>>>(out, clk, 1.f); if (mode == 2) k<2><<<nb, 256>>>(out, clk, 1.f);
      if (mode == 3) k<3><<<nb, 256>>>(out, clk, 1.f); if (mode == 4) k<4><<<nb, 256>>>(out, clk, 1.f);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(h, clk, nb * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < nb; ++i) avg += h[i]; avg /= nb;
    // warps per SMSP = 2*bps ; per iteration each warp issues the listed instrs
    printf("%-28s warps/SMSP=%d  clk/iter=%.2f  (per warp-iter per SMSP: %.2f clk)\n", names[mode], 2 * bps, avg / N_IT, avg / N_IT / (2 * bps));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
