// K4: threefry2x32 and the jax.random samplers built on it (device functions).
//
// Reproduces jax.random bit-for-bit in JAX's LEGACY counter layout
// (jax_threefry_partitionable=False), the only layout that reproduces the
// reference's printed notebook output (hello_world.ipynb:92-99).  jax is a
// third-party dependency of the reference (pyproject.toml:20-23), so this is a
// restatement of the published algorithm (Random123 threefry2x32-20).
//
// Layout: bits(key, n) for counters iota(n): h = ceil(n/2); block i (0<=i<h)
// = TF(key, (i, h+i)) [the padded counter of an odd n is 0]; element j is
// block j's first word for j < h, else block (j-h)'s second word.
#pragma once
#include <stdint.h>

namespace fgx {

struct Key {
  uint32_t a, b;
};

__host__ __device__ __forceinline__ uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

// Rotate = SHF (funnel shift).  A variant that rotates with a wide multiply on the FMA pipe in every other round
// (x * 2^r: low word = x << r, high word = x >> (32 - r), one LOP3 for (hi | lo) ^ x0) was measured and REJECTED:
// scripts/micro/threefry_pipes.cu, 4.75 M threads x 6 blocks on a B200: SHF 0.079 ms, every other round wide
// 0.083 ms, all rounds wide 0.096 ms (profiles/r02_threefry_pipes.json) -- IMAD.WIDE.U32 costs more issue slots
// than the SHF it replaces.
__host__ __device__ __forceinline__ void threefry2x32(uint32_t k0, uint32_t k1, uint32_t& x0, uint32_t& x1) {
  const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
  x0 += k0;
  x1 += k1;
#define FGX_TF_R(r) \
  x0 += x1;         \
  x1 = rotl32(x1, r); \
  x1 ^= x0;
  FGX_TF_R(13) FGX_TF_R(15) FGX_TF_R(26) FGX_TF_R(6)
  x0 += k1; x1 += k2 + 1u;
  FGX_TF_R(17) FGX_TF_R(29) FGX_TF_R(16) FGX_TF_R(24)
  x0 += k2; x1 += k0 + 2u;
  FGX_TF_R(13) FGX_TF_R(15) FGX_TF_R(26) FGX_TF_R(6)
  x0 += k0; x1 += k1 + 3u;
  FGX_TF_R(17) FGX_TF_R(29) FGX_TF_R(16) FGX_TF_R(24)
  x0 += k1; x1 += k2 + 4u;
  FGX_TF_R(13) FGX_TF_R(15) FGX_TF_R(26) FGX_TF_R(6)
  x0 += k2; x1 += k0 + 5u;
#undef FGX_TF_R
}

// element j of bits(key, n)
__host__ __device__ __forceinline__ uint32_t bits_elem(Key k, uint32_t n, uint32_t j) {
  const uint32_t h = (n + 1u) >> 1;
  const uint32_t i = j < h ? j : j - h;
  uint32_t x0 = i, x1 = h + i;
  if (x1 >= n) x1 = 0u;  // the pad counter of an odd n
  threefry2x32(k.a, k.b, x0, x1);
  return j < h ? x0 : x1;
}

// random_bits(key, (1,))
__host__ __device__ __forceinline__ uint32_t bits1(Key k) {
  uint32_t x0 = 0u, x1 = 0u;
  threefry2x32(k.a, k.b, x0, x1);
  return x0;
}

// split(key) -> (child0, child1): two blocks TF(key,(0,2)), TF(key,(1,3))
__host__ __device__ __forceinline__ void split2(Key k, Key& c0, Key& c1) {
  uint32_t a0 = 0u, a1 = 2u, b0 = 1u, b1 = 3u;
  threefry2x32(k.a, k.b, a0, a1);
  threefry2x32(k.a, k.b, b0, b1);
  c0.a = a0; c0.b = b0;
  c1.a = a1; c1.b = b1;
}

// child c of split(key, num)
__host__ __device__ __forceinline__ Key split_child(Key k, uint32_t num, uint32_t c) {
  Key r;
  r.a = bits_elem(k, 2u * num, 2u * c);
  r.b = bits_elem(k, 2u * num, 2u * c + 1u);
  return r;
}

// bits -> uniform float32 in [lo, hi): max(lo, f*(hi-lo)+lo), mul and add rounded separately
__host__ __device__ __forceinline__ float bits_to_uniform(uint32_t bits, float lo, float hi) {
  union { uint32_t u; float f; } v;
  v.u = (bits >> 9) | 0x3F800000u;
#ifdef __CUDA_ARCH__
  float f = __fsub_rn(v.f, 1.0f);
  float r = __fadd_rn(__fmul_rn(f, __fsub_rn(hi, lo)), lo);
  return fmaxf(lo, r);
#else
  float f = v.f - 1.0f;
  volatile float m = f * (hi - lo);
  float r = m + lo;
  return r > lo ? r : lo;
#endif
}

__host__ __device__ __forceinline__ float uniform1(Key k, float lo, float hi) {
  return bits_to_uniform(bits1(k), lo, hi);
}

// the integer map of jax.random.randint given the two 32-bit draws
__host__ __device__ __forceinline__ int32_t randint_from_bits(uint32_t hb, uint32_t lb, int32_t lo, int32_t hi) {
  uint32_t span = hi > lo ? (uint32_t)hi - (uint32_t)lo : 1u;
  uint32_t m = 65536u % span;
  m = (m * m) % span;
  uint32_t off = (hb % span) * m + (lb % span);
  off %= span;
  return (int32_t)((uint32_t)lo + off);
}

// The same map for a span that is fixed for a whole launch (a die's number of sides): the five `% span` of
// jax.random.randint cost ~20 instructions each as run-time divisions; with the host-side constants of
// RandintPlan each becomes a multiply-high, a multiply-subtract and one conditional correction.  Exact:
// magic = floor(2^32 / span) (saturated), so q = umulhi(x, magic) is the true quotient or one less, and
// x - q * span lies in [0, 2 span).
struct RandintPlan {
  uint32_t span, magic, mult;  // mult = ((2^16 % span)^2) % span
  int32_t lo;
};
__host__ __forceinline__ RandintPlan randint_plan(int32_t lo, int32_t hi) {
  RandintPlan p;
  p.span = hi > lo ? (uint32_t)hi - (uint32_t)lo : 1u;
  const uint64_t q = (1ull << 32) / p.span;
  p.magic = q > 0xFFFFFFFFull ? 0xFFFFFFFFu : (uint32_t)q;
  const uint32_t m = 65536u % p.span;
  p.mult = (uint32_t)(((uint64_t)m * m) % p.span);
  p.lo = lo;
  return p;
}
#ifdef __CUDACC__
__device__ __forceinline__ uint32_t mod_plan(uint32_t x, const RandintPlan& p) {
  const uint32_t r = x - __umulhi(x, p.magic) * p.span;
  return r >= p.span ? r - p.span : r;
}
__device__ __forceinline__ int32_t randint_from_bits_plan(uint32_t hb, uint32_t lb, const RandintPlan& p) {
  const uint32_t off = mod_plan(mod_plan(hb, p) * p.mult + mod_plan(lb, p), p);  // uint32 wrap-around as in jax
  return (int32_t)((uint32_t)p.lo + off);
}
__device__ __forceinline__ int32_t randint1_plan(Key k, const RandintPlan& p) {
  Key k1, k2;
  split2(k, k1, k2);
  return randint_from_bits_plan(bits1(k1), bits1(k2), p);
}
#endif

// randint(key, (1,), lo, hi)
__host__ __device__ __forceinline__ int32_t randint1(Key k, int32_t lo, int32_t hi) {
  Key k1, k2;
  split2(k, k1, k2);
  return randint_from_bits(bits1(k1), bits1(k2), lo, hi);
}

#ifdef __CUDACC__
// XLA's float32 erf_inv (Giles' polynomial), w = -log1p(-x*x)
__device__ __forceinline__ float erf_inv_f32(float x) {
  float w = -log1pf(-x * x);
  float p;
  if (w < 5.0f) {
    w = w - 2.5f;
    p = 2.81022636e-08f;
    p = 3.43273939e-07f + p * w;
    p = -3.5233877e-06f + p * w;
    p = -4.39150654e-06f + p * w;
    p = 0.00021858087f + p * w;
    p = -0.00125372503f + p * w;
    p = -0.00417768164f + p * w;
    p = 0.246640727f + p * w;
    p = 1.50140941f + p * w;
  } else {
    w = sqrtf(w) - 3.0f;
    p = -0.000200214257f;
    p = 0.000100950558f + p * w;
    p = 0.00134934322f + p * w;
    p = -0.00367342844f + p * w;
    p = 0.00573950773f + p * w;
    p = -0.0076224613f + p * w;
    p = 0.00943887047f + p * w;
    p = 1.00167406f + p * w;
    p = 2.83297682f + p * w;
  }
  return fabsf(x) == 1.0f ? x * __int_as_float(0x7f800000) : p * x;
}

__device__ __forceinline__ float bits_to_normal(uint32_t bits) {
  const float lo = __int_as_float(0xBF7FFFFF);  // nextafter(-1, 0)
  return 1.41421356237f * erf_inv_f32(bits_to_uniform(bits, lo, 1.0f));
}
#endif

}  // namespace fgx
