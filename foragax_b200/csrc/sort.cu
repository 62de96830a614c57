// K3 (part 2): the argsort of sort_agents as a device radix sort.
//
// Reference: agent_methods.py:74-76 -- argsort(q) or argsort(-q) (jnp.argsort =
// stable lax.sort of (key, iota); float total order with -0.0 == +0.0 and NaN
// last).  Here: keys are mapped to order-preserving uint32 images, then sorted
// with a stable LSD radix sort (8-bit digits) of (image, slot index):
//   prep     image + four global digit histograms + min/max image
//   minmax   counts image==min / image==max: a two-valued key (the common
//            sort on active_state or a 0/1 flag) needs ONE 1-bit pass
//   plan     per pass: trivial? (one bin holds everything) + digit bases
//   pass x4  tile histogram -> per-digit scan over tiles -> stable scatter with
//            warp match/ballot ranking; skipped on device when trivial
//   finish   copy to the caller's perm if the result landed in scratch
// Small inputs take a single-CTA bitonic sort of the 64-bit (image, index)
// composite (unique, hence stable) in shared memory: one launch.
#include "common.cuh"

namespace fgx {

constexpr int RS_BLOCK = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_BLOCK * RS_ITEMS;  // 4096
constexpr int RS_WARPS = RS_BLOCK / 32;
constexpr int SORT_SMALL_N = 4096;  // bitonic path: 4096 * 8 B = 32 KB static smem

struct SortState {
  uint32_t hist[4][256];
  uint32_t base[4][256];
  uint32_t kmin, kmax;
  uint32_t cnt_min, cnt_max;
  int32_t two_valued;
  int32_t trivial[4];
  int32_t pad[3];
};

// The (image, idx) ping-pong side that holds the current order before `pass`
// = parity of the non-trivial passes already executed; the idx buffer is
// still the implicit iota until the first executed pass.
__device__ __forceinline__ int passes_done(const SortState* st, int pass) {
  int c = 0;
  for (int q = 0; q < pass; ++q) c += st->trivial[q] ? 0 : 1;
  return c;
}

// warp-aggregated shared-memory histogram increment (keys are often two-valued:
// un-aggregated atomics would serialise 32-way on one bank)
__device__ __forceinline__ void hist_add(uint32_t* h, uint32_t d, bool valid) {
  const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : 0xFFFFFFFFu);
  if (valid && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&h[d], (uint32_t)__popc(peers));
}

__device__ __forceinline__ uint32_t key_image(const void* keys, int dtype, int descending, int64_t i) {
  if (dtype == FGX_F32) {
    float v = reinterpret_cast<const float*>(keys)[i];
    if (descending) v = -v;
    if (v != v) return 0xFFFFFFFFu;  // every NaN sorts last
    if (v == 0.0f) v = 0.0f;         // -0.0 -> +0.0
    const uint32_t b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  }
  int32_t v = reinterpret_cast<const int32_t*>(keys)[i];
  if (descending) v = (int32_t)(0u - (uint32_t)v);
  return (uint32_t)v ^ 0x80000000u;
}

__global__ void __launch_bounds__(256) sort_prep_kernel(const void* __restrict__ keys, int dtype, int descending,
                                                        int64_t n, uint32_t* __restrict__ img,
                                                        SortState* __restrict__ st) {
  __shared__ uint32_t h[4][256];
  __shared__ uint32_t smin, smax;
  for (int k = threadIdx.x; k < 1024; k += blockDim.x) (&h[0][0])[k] = 0;
  if (threadIdx.x == 0) { smin = 0xFFFFFFFFu; smax = 0u; }
  __syncthreads();
  uint32_t lmin = 0xFFFFFFFFu, lmax = 0u;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x; i0 < n; i0 += stride) {  // warp-uniform trip count
    const int64_t i = i0 + threadIdx.x;
    const bool valid = i < n;
    const uint32_t v = valid ? key_image(keys, dtype, descending, i) : 0u;
    if (valid) {
      img[i] = v;
      lmin = min(lmin, v);
      lmax = max(lmax, v);
    }
    hist_add(h[0], v & 255u, valid);
    hist_add(h[1], (v >> 8) & 255u, valid);
    hist_add(h[2], (v >> 16) & 255u, valid);
    hist_add(h[3], v >> 24, valid);
  }
  lmin = __reduce_min_sync(0xffffffffu, lmin);
  lmax = __reduce_max_sync(0xffffffffu, lmax);
  if ((threadIdx.x & 31) == 0) { atomicMin(&smin, lmin); atomicMax(&smax, lmax); }
  __syncthreads();
  for (int k = threadIdx.x; k < 1024; k += blockDim.x) {
    const uint32_t c = (&h[0][0])[k];
    if (c) atomicAdd(&st->hist[0][0] + k, c);
  }
  if (threadIdx.x == 0) { atomicMin(&st->kmin, smin); atomicMax(&st->kmax, smax); }
}

__global__ void __launch_bounds__(256) sort_minmax_kernel(const uint32_t* __restrict__ img, int64_t n,
                                                          SortState* __restrict__ st) {
  const uint32_t kmin = st->kmin, kmax = st->kmax;
  uint32_t a = 0, b = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t v = __ldg(img + i);
    a += v == kmin;
    b += v == kmax;
  }
  a = __reduce_add_sync(0xffffffffu, a);
  b = __reduce_add_sync(0xffffffffu, b);
  if ((threadIdx.x & 31) == 0) {
    if (a) atomicAdd(&st->cnt_min, a);
    if (b) atomicAdd(&st->cnt_max, b);
  }
}

__global__ void __launch_bounds__(256) sort_plan_kernel(int64_t n, SortState* __restrict__ st) {
  __shared__ uint32_t sh[256];
  __shared__ int triv;
  const int t = threadIdx.x;
  const bool two = st->kmin != st->kmax && (uint64_t)st->cnt_min + st->cnt_max == (uint64_t)n;
  const bool one = st->kmin == st->kmax;
  for (int p = 0; p < 4; ++p) {
    uint32_t c = st->hist[p][t];
    if (two && p == 0) c = t == 0 ? st->cnt_min : (t == 1 ? st->cnt_max : 0u);
    if (t == 0) triv = 0;
    __syncthreads();
    if ((uint64_t)c == (uint64_t)n) triv = 1;
    sh[t] = c;
    __syncthreads();
    if (t == 0) {  // 256-entry exclusive scan; tiny
      uint32_t run = 0;
      for (int d = 0; d < 256; ++d) { const uint32_t v = sh[d]; sh[d] = run; run += v; }
    }
    __syncthreads();
    st->base[p][t] = sh[t];
    if (t == 0) st->trivial[p] = (one || (two && p > 0)) ? 1 : triv;
    __syncthreads();
  }
  if (t == 0) st->two_valued = two ? 1 : 0;
}

__device__ __forceinline__ uint32_t digit_of(uint32_t v, int pass, bool two, uint32_t kmax) {
  return two ? (v == kmax ? 1u : 0u) : ((v >> (8 * pass)) & 255u);
}

// tile digit counts -> block_hist[digit][tile]
__global__ void __launch_bounds__(RS_BLOCK) sort_hist_kernel(const uint32_t* __restrict__ img0,
                                                             const uint32_t* __restrict__ img1, int64_t n,
                                                             int64_t num_tiles, int pass,
                                                             const SortState* __restrict__ st,
                                                             uint32_t* __restrict__ block_hist) {
  if (st->trivial[pass]) return;
  const uint32_t* img = (passes_done(st, pass) & 1) ? img1 : img0;
  const bool two = st->two_valued;
  const uint32_t kmax = st->kmax;
  __shared__ uint32_t h[256];
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = tile * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
      const int64_t i = base + r * RS_BLOCK + threadIdx.x;
      const bool valid = i < n;
      hist_add(h, valid ? digit_of(__ldg(img + i), pass, two, kmax) : 0u, valid);
    }
    __syncthreads();
    block_hist[(int64_t)threadIdx.x * num_tiles + tile] = h[threadIdx.x];
    __syncthreads();
  }
}

// one CTA per digit: exclusive scan of block_hist[digit][0..num_tiles) in place
__global__ void __launch_bounds__(256) sort_scan_kernel(int64_t num_tiles, int pass, const SortState* __restrict__ st,
                                                        uint32_t* __restrict__ block_hist) {
  if (st->trivial[pass]) return;
  if (st->two_valued && blockIdx.x > 1) return;
  uint32_t* row = block_hist + (int64_t)blockIdx.x * num_tiles;
  __shared__ uint32_t wsum[8];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int64_t b = 0; b < num_tiles; b += 256) {
    const int64_t t = b + threadIdx.x;
    const uint32_t v = t < num_tiles ? row[t] : 0u;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (l >= d) inc += o;
    }
    if (l == 31) wsum[w] = inc;
    __syncthreads();
    uint32_t wb = carry;
    for (int k = 0; k < w; ++k) wb += wsum[k];
    if (t < num_tiles) row[t] = wb + inc - v;
    __syncthreads();
    if (threadIdx.x == 255) carry = wb + inc;
    __syncthreads();
  }
}

// stable scatter of one pass
__global__ void __launch_bounds__(RS_BLOCK) sort_scatter_kernel(uint32_t* __restrict__ img0, uint32_t* __restrict__ img1,
                                                                int32_t* __restrict__ idx0, int32_t* __restrict__ idx1,
                                                                int64_t n, int64_t num_tiles, int pass,
                                                                const SortState* __restrict__ st,
                                                                const uint32_t* __restrict__ block_hist) {
  if (st->trivial[pass]) return;
  const int done = passes_done(st, pass);
  const int cur = done & 1;
  const uint32_t* kin = cur ? img1 : img0;
  uint32_t* kout = cur ? img0 : img1;
  const int32_t* iin = cur ? idx1 : idx0;
  int32_t* iout = cur ? idx0 : idx1;
  const bool iota = done == 0;
  const bool two = st->two_valued;
  const uint32_t kmax = st->kmax;
  __shared__ uint32_t wcnt[RS_WARPS][256];
  __shared__ uint32_t gbase[256];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const unsigned lt = lanemask_lt();
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    for (int k = threadIdx.x; k < RS_WARPS * 256; k += RS_BLOCK) (&wcnt[0][0])[k] = 0;
    gbase[threadIdx.x] = st->base[pass][threadIdx.x] + block_hist[(int64_t)threadIdx.x * num_tiles + tile];
    __syncthreads();
    const int64_t wbase = tile * RS_TILE + (int64_t)w * (32 * RS_ITEMS);
    uint32_t key[RS_ITEMS];
    uint32_t lpos[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
      const int64_t i = wbase + r * 32 + l;
      const bool valid = i < n;
      key[r] = valid ? kin[i] : 0xFFFFFFFFu;
      // invalid lanes use a digit value outside 0..255 so they never match a real one
      const uint32_t d = valid ? digit_of(key[r], pass, two, kmax) : 256u;
      const unsigned peers = __match_any_sync(0xffffffffu, d);
      const int leader = __ffs(peers) - 1;
      uint32_t old = 0;
      if (valid && l == leader) {
        old = wcnt[w][d];
        wcnt[w][d] = old + __popc(peers);
      }
      old = __shfl_sync(0xffffffffu, old, leader);
      lpos[r] = old + __popc(peers & lt);
      __syncwarp();
    }
    __syncthreads();
    {  // exclusive scan over warps for digit = threadIdx.x
      uint32_t run = 0;
#pragma unroll
      for (int k = 0; k < RS_WARPS; ++k) {
        const uint32_t v = wcnt[k][threadIdx.x];
        wcnt[k][threadIdx.x] = run;
        run += v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
      const int64_t i = wbase + r * 32 + l;
      if (i < n) {
        const uint32_t d = digit_of(key[r], pass, two, kmax);
        const uint32_t pos = gbase[d] + wcnt[w][d] + lpos[r];
        kout[pos] = key[r];
        iout[pos] = iota ? (int32_t)i : iin[i];
      }
    }
    __syncthreads();
  }
}

// result -> perm (idx0 aliases perm): copy only if the final order sits in idx1, or is still iota
__global__ void __launch_bounds__(256) sort_finish_kernel(const int32_t* __restrict__ idx1, int32_t* __restrict__ perm,
                                                          int64_t n, const SortState* __restrict__ st) {
  const int done = passes_done(st, 4);
  const bool iota = done == 0;
  if (!iota && (done & 1) == 0) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    perm[i] = iota ? (int32_t)i : idx1[i];
}

// ---- small path: single CTA, bitonic sort of (image << 32 | index) ----------
__global__ void __launch_bounds__(1024) sort_small_kernel(const void* __restrict__ keys, int dtype, int descending,
                                                          int n, int32_t* __restrict__ perm) {
  __shared__ unsigned long long s[SORT_SMALL_N];
  int m = 1;
  while (m < n) m <<= 1;
  for (int i = threadIdx.x; i < m; i += blockDim.x)
    s[i] = i < n ? (((unsigned long long)key_image(keys, dtype, descending, i) << 32) | (unsigned)i)
                 : 0xFFFFFFFFFFFFFFFFull;
  __syncthreads();
  for (int k = 2; k <= m; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long a = s[i], b = s[ixj];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { s[i] = b; s[ixj] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) perm[i] = (int32_t)(s[i] & 0xFFFFFFFFull);
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace fgx

extern "C" {

size_t fgx_argsort_scratch_bytes(int64_t n) {
  using namespace fgx;
  if (n < 0) n = 0;
  const int64_t tiles = ceil_div(n, RS_TILE);
  size_t b = align_up(sizeof(SortState), 256);
  b += 2 * align_up((size_t)n * 4, 256);            // image ping-pong
  b += align_up((size_t)n * 4, 256);                // idx1 (idx0 is the caller's perm)
  b += align_up((size_t)tiles * 256 * 4 + 4, 256);  // block_hist[256][tiles]
  return b;
}

int fgx_argsort(const void* keys, int32_t key_dtype, int64_t n, int32_t descending, int32_t* perm, void* scratch,
                size_t scratch_bytes, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n >= 0 && n < (1ll << 31), "fgx_argsort: n out of range");
  FGX_REQUIRE(key_dtype == FGX_F32 || key_dtype == FGX_I32, "fgx_argsort: key dtype must be FGX_F32 or FGX_I32");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(keys && perm, "fgx_argsort: null pointer");
  cudaStream_t s = as_stream(stream);
  if (n <= SORT_SMALL_N) {
    sort_small_kernel<<<1, 1024, 0, s>>>(keys, key_dtype, descending ? 1 : 0, (int)n, perm);
    return check_launch("fgx_argsort(small)");
  }
  FGX_REQUIRE(scratch && scratch_bytes >= fgx_argsort_scratch_bytes(n), "fgx_argsort: scratch too small");
  FGX_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "fgx_argsort: scratch must be 256-byte aligned");
  const int64_t tiles = ceil_div(n, RS_TILE);
  char* p = reinterpret_cast<char*>(scratch);
  SortState* st = reinterpret_cast<SortState*>(p);
  p += align_up(sizeof(SortState), 256);
  uint32_t* img0 = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)n * 4, 256);
  uint32_t* img1 = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)n * 4, 256);
  int32_t* idx1 = reinterpret_cast<int32_t*>(p);
  p += align_up((size_t)n * 4, 256);
  uint32_t* block_hist = reinterpret_cast<uint32_t*>(p);
  int32_t* idx0 = perm;

  cudaError_t e = cudaMemsetAsync(st, 0, sizeof(SortState), s);
  if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "fgx_argsort: %s", cudaGetErrorString(e));
  // kmin starts at 0xFFFFFFFF
  e = cudaMemsetAsync(&st->kmin, 0xFF, sizeof(uint32_t), s);
  if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "fgx_argsort: %s", cudaGetErrorString(e));
  const int g = grid_for(n, 256, 8);
  sort_prep_kernel<<<g, 256, 0, s>>>(keys, key_dtype, descending ? 1 : 0, n, img0, st);
  sort_minmax_kernel<<<g, 256, 0, s>>>(img0, n, st);
  sort_plan_kernel<<<1, 256, 0, s>>>(n, st);
  const int gt = (int)(tiles < (int64_t)sm_count() * 4 ? tiles : (int64_t)sm_count() * 4);
  for (int pass = 0; pass < 4; ++pass) {
    sort_hist_kernel<<<gt, RS_BLOCK, 0, s>>>(img0, img1, n, tiles, pass, st, block_hist);
    sort_scan_kernel<<<256, 256, 0, s>>>(tiles, pass, st, block_hist);
    sort_scatter_kernel<<<gt, RS_BLOCK, 0, s>>>(img0, img1, idx0, idx1, n, tiles, pass, st, block_hist);
  }
  sort_finish_kernel<<<g, 256, 0, s>>>(idx1, perm, n, st);
  return check_launch("fgx_argsort");
}
}
