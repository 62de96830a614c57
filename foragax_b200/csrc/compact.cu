// K3 (part 1): select_agents as a stable two-way partition of slot indices.
//
// Reference: agent_methods.py:66-71 -- mask -> where(mask,1.,0.) -> full stable
// argsort(-mask) -> (count, permutation).  The permutation is "selected slots
// ascending, then unselected slots ascending", so the O(N log N) comparison
// sort of the reference is a 1-bit stable partition here:
//   pass A  per-tile selected counts (warp ballots + popc), last CTA scans them
//   pass B  re-evaluates the predicate (the leaves are L2-resident after pass
//           A), ranks each slot with ballot/popc prefix sums and writes
//           perm[rank] or perm[total + i - rank].
// The predicate is evaluated in-kernel from the agent leaves (up to four
// comparisons combined by a truth table), so no mask array is materialised.
// Algorithmic HBM bytes: predicate leaves read once + 4 B/slot written.
#include "common.cuh"

namespace fgx {

constexpr int SEL_BLOCK = 256;
constexpr int SEL_ITEMS = 8;
constexpr int SEL_TILE = SEL_BLOCK * SEL_ITEMS;  // 2048 slots per tile
constexpr int SEL_WARPS = SEL_BLOCK / 32;

struct SelScratch {
  unsigned int done;   // CTAs finished with pass A
  int32_t total;       // selected count
  int32_t pad[2];
  // followed by int32 tile_offset[num_tiles]
};

__device__ __forceinline__ bool eval_term(const fgx_pred_term_t& t, int64_t i) {
  const int64_t e = i * (int64_t)t.stride;
  if (t.dtype == FGX_F32) {
    const float v = __ldg(reinterpret_cast<const float*>(t.data) + e);
    const float c = __uint_as_float(t.imm_bits);
    switch (t.op) {
      case FGX_OP_EQ: return v == c;
      case FGX_OP_NE: return v != c;
      case FGX_OP_LT: return v < c;
      case FGX_OP_LE: return v <= c;
      case FGX_OP_GT: return v > c;
      case FGX_OP_GE: return v >= c;
      default: return v != 0.0f;
    }
  } else if (t.dtype == FGX_I32) {
    const int32_t v = __ldg(reinterpret_cast<const int32_t*>(t.data) + e);
    const int32_t c = (int32_t)t.imm_bits;
    switch (t.op) {
      case FGX_OP_EQ: return v == c;
      case FGX_OP_NE: return v != c;
      case FGX_OP_LT: return v < c;
      case FGX_OP_LE: return v <= c;
      case FGX_OP_GT: return v > c;
      case FGX_OP_GE: return v >= c;
      default: return v != 0;
    }
  } else {  // FGX_U8 (bool leaves)
    const int32_t v = __ldg(reinterpret_cast<const uint8_t*>(t.data) + e);
    const int32_t c = (int32_t)t.imm_bits;
    switch (t.op) {
      case FGX_OP_EQ: return v == c;
      case FGX_OP_NE: return v != c;
      case FGX_OP_LT: return v < c;
      case FGX_OP_LE: return v <= c;
      case FGX_OP_GT: return v > c;
      case FGX_OP_GE: return v >= c;
      default: return v != 0;
    }
  }
}

__device__ __forceinline__ bool eval_pred(const fgx_predicate_t& p, int64_t i) {
  unsigned idx = 0;
#pragma unroll
  for (int t = 0; t < 4; ++t)
    if (t < p.n_terms) idx |= (eval_term(p.term[t], i) ? 1u : 0u) << t;
  return (p.lut >> idx) & 1u;
}

// element handled by (warp w, round r, lane l) of a tile: contiguous per warp
__device__ __forceinline__ int64_t tile_elem(int64_t tile, int w, int r, int l) {
  return tile * SEL_TILE + w * (32 * SEL_ITEMS) + r * 32 + l;
}

__device__ __forceinline__ int tile_count(const fgx_predicate_t& p, int64_t tile, int64_t n, int* warp_sums) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  int c = 0;
#pragma unroll
  for (int r = 0; r < SEL_ITEMS; ++r) {
    const int64_t i = tile_elem(tile, w, r, l);
    const bool s = i < n && eval_pred(p, i);
    c += __popc(__ballot_sync(0xffffffffu, s));
  }
  if (l == 0) warp_sums[w] = c;
  __syncthreads();
  int tot = 0;
#pragma unroll
  for (int k = 0; k < SEL_WARPS; ++k) tot += warp_sums[k];
  __syncthreads();
  return tot;
}

__device__ __forceinline__ void tile_scatter(const fgx_predicate_t& p, int64_t tile, int64_t n, int32_t tile_off,
                                             int32_t total, int32_t* __restrict__ perm, int* warp_sums) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  unsigned ball[SEL_ITEMS];
  int c = 0;
#pragma unroll
  for (int r = 0; r < SEL_ITEMS; ++r) {
    const int64_t i = tile_elem(tile, w, r, l);
    const bool s = i < n && eval_pred(p, i);
    ball[r] = __ballot_sync(0xffffffffu, s);
    c += __popc(ball[r]);
  }
  if (l == 0) warp_sums[w] = c;
  __syncthreads();
  int before = tile_off;
#pragma unroll
  for (int k = 0; k < SEL_WARPS; ++k)
    if (k < w) before += warp_sums[k];
  __syncthreads();
  const unsigned lt = lanemask_lt();
#pragma unroll
  for (int r = 0; r < SEL_ITEMS; ++r) {
    const int64_t i = tile_elem(tile, w, r, l);
    if (i < n) {
      const int rank = before + __popc(ball[r] & lt);  // selected slots before i
      const bool s = (ball[r] >> l) & 1u;
      const int64_t pos = s ? (int64_t)rank : (int64_t)total + (i - rank);
      perm[pos] = (int32_t)i;
    }
    before += __popc(ball[r]);
  }
}

// pass A: per-tile counts; the last CTA to finish turns them into exclusive offsets
__global__ void __launch_bounds__(SEL_BLOCK) select_count_kernel(const __grid_constant__ fgx_predicate_t p, int64_t n,
                                                                int64_t num_tiles, SelScratch* __restrict__ sc,
                                                                int32_t* __restrict__ tile_off,
                                                                int32_t* __restrict__ count_out) {
  __shared__ int warp_sums[SEL_WARPS];
  __shared__ bool is_last;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int c = tile_count(p, tile, n, warp_sums);
    if (threadIdx.x == 0) tile_off[tile] = c;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = atomicAdd(&sc->done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // single-CTA exclusive scan of tile counts (num_tiles <= n/2048)
  __shared__ int carry;
  __shared__ int wsum[SEL_WARPS];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int64_t base = 0; base < num_tiles; base += SEL_BLOCK) {
    const int64_t t = base + threadIdx.x;
    const int v = t < num_tiles ? __ldcg(tile_off + t) : 0;
    int inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int o = __shfl_up_sync(0xffffffffu, inc, d);
      if (l >= d) inc += o;
    }
    if (l == 31) wsum[w] = inc;
    __syncthreads();
    int wbase = carry;
    for (int k = 0; k < w; ++k) wbase += wsum[k];
    if (t < num_tiles) tile_off[t] = wbase + inc - v;
    __syncthreads();
    if (threadIdx.x == SEL_BLOCK - 1) carry = wbase + inc;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    sc->total = carry;
    *count_out = carry;
    sc->done = 0;  // re-arm for the next call on this scratch
  }
}

// pass B: scatter
__global__ void __launch_bounds__(SEL_BLOCK) select_scatter_kernel(const __grid_constant__ fgx_predicate_t p, int64_t n,
                                                                  int64_t num_tiles,
                                                                  const SelScratch* __restrict__ sc,
                                                                  const int32_t* __restrict__ tile_off,
                                                                  int32_t* __restrict__ perm) {
  __shared__ int warp_sums[SEL_WARPS];
  const int32_t total = sc->total;
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x)
    tile_scatter(p, tile, n, tile_off[tile], total, perm, warp_sums);
}

// small sets (a few tiles): one CTA, one launch
__global__ void __launch_bounds__(SEL_BLOCK) select_small_kernel(const __grid_constant__ fgx_predicate_t p, int64_t n,
                                                                int64_t num_tiles, int32_t* __restrict__ perm,
                                                                int32_t* __restrict__ count_out) {
  __shared__ int warp_sums[SEL_WARPS];
  __shared__ int offs[64];
  int total = 0;
  for (int64_t tile = 0; tile < num_tiles; ++tile) {
    const int c = tile_count(p, tile, n, warp_sums);
    if (threadIdx.x == 0) offs[tile] = total;
    total += c;
  }
  __syncthreads();
  for (int64_t tile = 0; tile < num_tiles; ++tile) tile_scatter(p, tile, n, offs[tile], total, perm, warp_sums);
  if (threadIdx.x == 0) *count_out = total;
}

constexpr int64_t SEL_SMALL_TILES = 8;  // <= 16384 slots: single-CTA path

}  // namespace fgx

extern "C" {

size_t fgx_select_scratch_bytes(int64_t n) {
  if (n < 0) n = 0;
  const int64_t tiles = fgx::ceil_div(n, fgx::SEL_TILE);
  return sizeof(fgx::SelScratch) + (size_t)(tiles + 1) * sizeof(int32_t);
}

int fgx_select(const fgx_predicate_t* pred_host, int64_t n, int32_t* perm, int32_t* count, void* scratch,
               size_t scratch_bytes, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(pred_host, "fgx_select: null predicate");
  FGX_REQUIRE(n >= 0 && n < (1ll << 31), "fgx_select: n out of range");
  FGX_REQUIRE(count, "fgx_select: null count");
  FGX_REQUIRE(pred_host->n_terms >= 0 && pred_host->n_terms <= 4, "fgx_select: n_terms must be 0..4");
  for (int t = 0; t < pred_host->n_terms; ++t) {
    FGX_REQUIRE(pred_host->term[t].data, "fgx_select: term %d has a null leaf", t);
    FGX_REQUIRE(pred_host->term[t].dtype == FGX_U8 || pred_host->term[t].dtype == FGX_I32 ||
                    pred_host->term[t].dtype == FGX_F32,
                "fgx_select: term %d has an unsupported dtype", t);
    FGX_REQUIRE(pred_host->term[t].op >= FGX_OP_EQ && pred_host->term[t].op <= FGX_OP_NONZERO,
                "fgx_select: term %d has an unknown op", t);
  }
  cudaStream_t s = as_stream(stream);
  if (n == 0) {
    cudaError_t e = cudaMemsetAsync(count, 0, sizeof(int32_t), s);
    return e == cudaSuccess ? FGX_OK : fail(FGX_ERR_CUDA, "fgx_select: %s", cudaGetErrorString(e));
  }
  FGX_REQUIRE(perm, "fgx_select: null perm");
  const int64_t tiles = ceil_div(n, SEL_TILE);
  if (tiles <= SEL_SMALL_TILES) {
    select_small_kernel<<<1, SEL_BLOCK, 0, s>>>(*pred_host, n, tiles, perm, count);
    return check_launch("fgx_select(small)");
  }
  FGX_REQUIRE(scratch && scratch_bytes >= fgx_select_scratch_bytes(n), "fgx_select: scratch too small");
  FGX_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 15) == 0, "fgx_select: scratch must be 16-byte aligned");
  SelScratch* sc = reinterpret_cast<SelScratch*>(scratch);
  int32_t* tile_off = reinterpret_cast<int32_t*>(sc + 1);
  // `done` must be zero on entry; the count kernel re-arms it, but the first use of a fresh scratch needs this
  cudaError_t e = cudaMemsetAsync(sc, 0, sizeof(SelScratch), s);
  if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "fgx_select: %s", cudaGetErrorString(e));
  const int grid = (int)(tiles < (int64_t)sm_count() * 8 ? tiles : (int64_t)sm_count() * 8);
  select_count_kernel<<<grid, SEL_BLOCK, 0, s>>>(*pred_host, n, tiles, sc, tile_off, count);
  select_scatter_kernel<<<grid, SEL_BLOCK, 0, s>>>(*pred_host, n, tiles, sc, tile_off, perm);
  return check_launch("fgx_select");
}
}
