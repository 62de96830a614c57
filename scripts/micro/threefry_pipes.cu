// Micro-benchmark behind the rotate choice in csrc/threefry.cuh: six threefry2x32-20 blocks per thread (what one
// Dice step costs) with the rotate done by SHF (funnel shift, ALU pipe), by a wide multiply (IMAD.WIDE, FMA pipe)
// in every other round, or in every round.  Build and run on the GPU box:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tfp scripts/micro/threefry_pipes.cu && /tmp/tfp
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
#define RS(r) x0 += x1; x1 = rotl32(x1, r); x1 ^= x0;
#define RW(r) x0 += x1; { unsigned long long w_; asm("mul.wide.u32 %0, %1, %2;" : "=l"(w_) : "r"(x1), "r"(1u << (r))); x1 = ((uint32_t)(w_ >> 32) | (uint32_t)w_) ^ x0; }

template <int MODE>
__device__ __forceinline__ void tf(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += k0; x1 += k1;
#define A(r) if (MODE == 2) { RW(r) } else { RS(r) }
#define B(r) if (MODE >= 1) { RW(r) } else { RS(r) }
  A(13) B(15) A(26) B(6)
  x0 += k1; x1 += k2 + 1u;
  A(17) B(29) A(16) B(24)
  x0 += k2; x1 += k0 + 2u;
  A(13) B(15) A(26) B(6)
  x0 += k0; x1 += k1 + 3u;
  A(17) B(29) A(16) B(24)
  x0 += k1; x1 += k2 + 4u;
  A(13) B(15) A(26) B(6)
  x0 += k2; x1 += k0 + 5u;
}

template <int MODE>
__global__ void __launch_bounds__(256) k(uint2* key, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint2 kk = key[i];
    // split -> (k_, sub); split(sub) -> (k1, k2); bits(k1), bits(k2): six blocks, as in Dice.step_agent
    uint32_t a0 = 0, a1 = 2, b0 = 1, b1 = 3;
    tf<MODE>(kk.x, kk.y, a0, a1); tf<MODE>(kk.x, kk.y, b0, b1);
    uint32_t c0 = 0, c1 = 2, d0 = 1, d1 = 3;
    tf<MODE>(a1, b1, c0, c1); tf<MODE>(a1, b1, d0, d1);
    uint32_t e0 = 0, e1 = 0, f0 = 0, f1 = 0;
    tf<MODE>(c0, d0, e0, e1); tf<MODE>(c1, d1, f0, f1);
    key[i] = make_uint2(e0 ^ a1, f0 ^ b1);
  }
}

template <int MODE>
float run(uint2* d, int n) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 3; ++w) k<MODE><<<148 * 8, 256>>>(d, n);
  cudaEventRecord(e0);
  for (int it = 0; it < 20; ++it) k<MODE><<<148 * 8, 256>>>(d, n);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 20;
}

int main() {
  const int n = 4750000;
  uint2* d; cudaMalloc(&d, sizeof(uint2) * n); cudaMemset(d, 1, sizeof(uint2) * n);
  uint2 h0, h1, h2;
  float t0 = run<0>(d, n); cudaMemcpy(&h0, d + 12345, 8, cudaMemcpyDeviceToHost);
  cudaMemset(d, 1, sizeof(uint2) * n);
  float t1 = run<1>(d, n); cudaMemcpy(&h1, d + 12345, 8, cudaMemcpyDeviceToHost);
  cudaMemset(d, 1, sizeof(uint2) * n);
  float t2 = run<2>(d, n); cudaMemcpy(&h2, d + 12345, 8, cudaMemcpyDeviceToHost);
  printf("{\"threads\": %d, \"blocks_per_thread\": 6, \"shf_all_ms\": %.4f, \"wide_every_other_ms\": %.4f, \"wide_all_ms\": %.4f, \"same_bits\": %s}\n",
         n, t0, t1, t2, (h0.x == h1.x && h0.y == h1.y && h0.x == h2.x && h0.y == h2.y) ? "true" : "false");
  return 0;
}
