// Latency of mbarrier.test_wait / try_wait (with and without suspend hint) on an already-completed phase, one warp.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(long long* out) {
  __shared__ uint64_t bar;
  const uint32_t mb = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb));
  __syncthreads();
  uint32_t done, acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < 64; ++i) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(done) : "r"(mb), "r"(1u) : "memory");
    acc += done;
  }
  long long t1 = clock64();
  for (int i = 0; i < 64; ++i) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(done) : "r"(mb), "r"(1u) : "memory");
    acc += done;
  }
  long long t2 = clock64();
  for (int i = 0; i < 64; ++i) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0,1,0,p;\n\t}" : "=r"(done) : "r"(mb), "r"(1u), "r"(20000u) : "memory");
    acc += done;
  }
  long long t3 = clock64();
  if (threadIdx.x == 0) { out[0] = (t1 - t0) / 64; out[1] = (t2 - t1) / 64; out[2] = (t3 - t2) / 64; out[3] = acc; }
}
int main() {
  long long* d; cudaMalloc(&d, 64); k<<<1, 32>>>(d); long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("test_wait %lld clk, try_wait %lld clk, try_wait+hint %lld clk (successes %lld/192) %s\n", h[0], h[1], h[2], h[3], cudaGetErrorString(cudaGetLastError()));
}
