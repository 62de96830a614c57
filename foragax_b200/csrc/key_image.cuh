// Order-preserving uint32 image of a sort key (jnp.argsort's total order: -0.0 == +0.0, every NaN last;
// "descending" = ascending order of the negated key, agent_methods.py:76).  Shared by sort.cu and peer_pack.cu.
#pragma once
#include "common.cuh"

namespace fgx {

// order-preserving uint32 image of a raw 32-bit key word (float32 or int32)
__device__ __forceinline__ uint32_t image_of(uint32_t raw, int dtype, int descending) {
  if (dtype == FGX_F32) {
    float v = __uint_as_float(raw);
    if (descending) v = -v;
    if (v != v) return 0xFFFFFFFFu;  // every NaN sorts last
    if (v == 0.0f) v = 0.0f;         // -0.0 -> +0.0
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  }
  int32_t v = (int32_t)raw;
  if (descending) v = (int32_t)(0u - (uint32_t)v);
  return (uint32_t)v ^ 0x80000000u;
}
__device__ __forceinline__ uint32_t key_image(const void* keys, int dtype, int descending, int64_t i) {
  return image_of(__ldg(reinterpret_cast<const uint32_t*>(keys) + i), dtype, descending);
}

}  // namespace fgx
