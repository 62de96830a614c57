// K4 entry points: batched jax.random ops over a key batch uint32[n_keys,2].
// One thread per output element; consecutive threads write consecutive words.
#include "common.cuh"
#include "threefry.cuh"

namespace fgx {

enum { RNG_BITS = 0, RNG_SPLIT = 1, RNG_UNIFORM = 2, RNG_RANDINT = 3, RNG_NORMAL = 4 };

template <int MODE>
__global__ void __launch_bounds__(256) rng_kernel(const uint2* __restrict__ keys, int64_t n_keys, uint32_t n,
                                                  uint32_t* __restrict__ out, float flo, float fhi, int32_t ilo,
                                                  int32_t ihi) {
  const int64_t total = n_keys * (int64_t)n;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t kidx = t / n;
    const uint32_t j = (uint32_t)(t - kidx * n);
    const uint2 kk = __ldg(keys + kidx);
    Key k{kk.x, kk.y};
    if (MODE == RNG_BITS || MODE == RNG_SPLIT) {
      out[t] = bits_elem(k, n, j);
    } else if (MODE == RNG_UNIFORM) {
      reinterpret_cast<float*>(out)[t] = bits_to_uniform(bits_elem(k, n, j), flo, fhi);
    } else if (MODE == RNG_NORMAL) {
      reinterpret_cast<float*>(out)[t] = bits_to_normal(bits_elem(k, n, j));
    } else {  // RNG_RANDINT
      Key k1, k2;
      split2(k, k1, k2);
      reinterpret_cast<int32_t*>(out)[t] = randint_from_bits(bits_elem(k1, n, j), bits_elem(k2, n, j), ilo, ihi);
    }
  }
}

template <int MODE>
static int launch(const uint32_t* keys, int64_t n_keys, int64_t n, void* out, float flo, float fhi, int32_t ilo,
                  int32_t ihi, fgx_stream_t stream, const char* what) {
  FGX_REQUIRE(n_keys >= 0 && n >= 0, "%s: negative size", what);
  FGX_REQUIRE(n < (1ll << 31), "%s: n too large", what);
  if (n_keys == 0 || n == 0) return FGX_OK;
  FGX_REQUIRE(keys && out, "%s: null pointer", what);
  FGX_REQUIRE((reinterpret_cast<uintptr_t>(keys) & 7) == 0, "%s: keys must be 8-byte aligned", what);
  const int block = 256;
  const int grid = grid_for(n_keys * n, block, 8);
  rng_kernel<MODE><<<grid, block, 0, as_stream(stream)>>>(reinterpret_cast<const uint2*>(keys), n_keys, (uint32_t)n,
                                                        reinterpret_cast<uint32_t*>(out), flo, fhi, ilo, ihi);
  return check_launch(what);
}

}  // namespace fgx

extern "C" {

int fgx_threefry_split(const uint32_t* keys, int64_t n_keys, int32_t num, uint32_t* out, fgx_stream_t stream) {
  return fgx::launch<fgx::RNG_SPLIT>(keys, n_keys, 2ll * num, out, 0.f, 0.f, 0, 0, stream, "fgx_threefry_split");
}
int fgx_threefry_random_bits(const uint32_t* keys, int64_t n_keys, int32_t n, uint32_t* out, fgx_stream_t stream) {
  return fgx::launch<fgx::RNG_BITS>(keys, n_keys, n, out, 0.f, 0.f, 0, 0, stream, "fgx_threefry_random_bits");
}
int fgx_threefry_uniform(const uint32_t* keys, int64_t n_keys, int32_t n, float minval, float maxval, float* out,
                         fgx_stream_t stream) {
  return fgx::launch<fgx::RNG_UNIFORM>(keys, n_keys, n, out, minval, maxval, 0, 0, stream, "fgx_threefry_uniform");
}
int fgx_threefry_randint(const uint32_t* keys, int64_t n_keys, int32_t n, int32_t minval, int32_t maxval,
                         int32_t* out, fgx_stream_t stream) {
  return fgx::launch<fgx::RNG_RANDINT>(keys, n_keys, n, out, 0.f, 0.f, minval, maxval, stream,
                                       "fgx_threefry_randint");
}
int fgx_threefry_normal(const uint32_t* keys, int64_t n_keys, int32_t n, float* out, fgx_stream_t stream) {
  return fgx::launch<fgx::RNG_NORMAL>(keys, n_keys, n, out, 0.f, 0.f, 0, 0, stream, "fgx_threefry_normal");
}
}
