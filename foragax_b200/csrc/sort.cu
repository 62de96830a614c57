// K3 (part 2): the argsort of sort_agents as a device radix sort.
//
// Reference: agent_methods.py:74-76 -- argsort(q) or argsort(-q) (jnp.argsort =
// stable lax.sort of (key, iota); float total order with -0.0 == +0.0 and NaN
// last).  Here: keys are mapped to order-preserving uint32 images, then
//   1. stats    per-CTA (min, #min, max, #max) of the images in ONE read of the keys;
//   2. two-valued fast path (the common sort on active_state or a 0/1 flag): if every
//      key is the global min or max the sort IS a stable 2-way partition.  Up to ~77 M slots
//      steps 1 + 2 are ONE cooperative launch that reads the keys once: per 256-slot group a
//      warp keeps the ballots of "key == group minimum" in shared memory and publishes the
//      (min, #min, max, #max) record of its CTA; after a grid-wide sync every CTA derives the
//      global extrema and its offset, and scatters -- the permutation AND, for
//      fgx_sort_rows (= sort_agents, agent_methods.py:74-79), the narrow payload rows of every
//      leaf: rows are read coalesced at slot i and written once at their sorted position, so
//      the permutation never makes a round trip through memory and no gather pass runs.  The
//      general path below is skipped on device;
//   3. general path: stable LSD radix sort of (image, slot index), 8-bit digits:
//      prep (images + 4 digit histograms), plan (digit bases, trivial passes),
//      per pass: tile histogram -> per-digit scan over tiles -> stable scatter with
//      warp match/ballot ranking; finish copies to the caller's perm if needed -- all phases in
//      ONE cooperative launch separated by grid-wide syncs (FGX_SORT_MULTI=1: a launch per phase).
// Small inputs order the unique 64-bit (image, index) composite (hence stable): n <= 256 by a
// single-CTA bitonic network in shared memory, n <= 16384 by an all-pairs rank sort over every SM.
#include <cooperative_groups.h>

#include "common.cuh"
#include "key_image.cuh"
#include "rows.cuh"

namespace cg = cooperative_groups;

namespace fgx {

constexpr int RS_BLOCK = 256;
constexpr int RS_ITEMS = 16;
constexpr int RS_TILE = RS_BLOCK * RS_ITEMS;  // 4096
constexpr int RS_WARPS = RS_BLOCK / 32;
constexpr int SORT_SMALL_N = 256;  // bitonic path (needs no scratch): 256 * 8 B = 2 KB static smem
constexpr int TV_ITEMS = 8;         // two-valued partition: 2048 slots per tile
constexpr int TV_TILE = RS_BLOCK * TV_ITEMS;
constexpr int TV_MAX_GRID = 4096;

struct SortState {
  uint32_t hist[4][256];
  uint32_t base[4][256];
  uint32_t kmin, kmax;
  uint32_t cnt_min, cnt_max;
  int32_t done;  // the two-valued fast path produced the result
  int32_t trivial[4];
  int32_t pad[3];
};

// SortState fields written by an earlier phase are read with ld.cg: in the fused kernel the writer is
// another CTA of the SAME launch (L1 is not coherent; every phase boundary is a grid-wide sync)
__device__ __forceinline__ int passes_done(const SortState* st, int pass) {
  int c = 0;
  for (int q = 0; q < pass; ++q) c += __ldcg(&st->trivial[q]) ? 0 : 1;
  return c;
}

// warp-aggregated shared-memory histogram increment
__device__ __forceinline__ void hist_add(uint32_t* h, uint32_t d, bool valid) {
  const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : 0xFFFFFFFFu);
  if (valid && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&h[d], (uint32_t)__popc(peers));
}

// ---- 1. one pass of statistics: per-CTA (min, #min, max, #max) of its contiguous tile range -----------
// (min, count) is a monoid -- combine keeps the smaller value and adds the counts on a tie -- so the
// global extrema AND how many keys hold them come out of a single read of the keys: the scatter pass
// derives them from the per-CTA records, no separate min/max pass.
struct TvStat {
  uint32_t mn, cmn, mx, cmx;
};
__device__ __forceinline__ void tv_merge(TvStat& a, uint32_t mn, uint32_t cmn, uint32_t mx, uint32_t cmx) {
  a.cmn = mn < a.mn ? cmn : (mn == a.mn ? a.cmn + cmn : a.cmn);
  a.mn = min(a.mn, mn);
  a.cmx = mx > a.mx ? cmx : (mx == a.mx ? a.cmx + cmx : a.cmx);
  a.mx = max(a.mx, mx);
}

__device__ __forceinline__ int64_t tv_elem(int64_t tile, int w, int r, int l) {
  return tile * TV_TILE + w * (32 * TV_ITEMS) + r * 32 + l;
}

__global__ void __launch_bounds__(RS_BLOCK)
    sort_two_stats_kernel(const void* __restrict__ keys, int dtype, int descending, int64_t n, int64_t num_tiles,
                          int64_t tiles_per_cta, TvStat* __restrict__ cta_stat) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int64_t t0 = blockIdx.x * tiles_per_cta, t1 = min(num_tiles, t0 + tiles_per_cta);
  TvStat a{0xFFFFFFFFu, 0u, 0u, 0u};
  const uint32_t* __restrict__ kw = reinterpret_cast<const uint32_t*>(keys);
  for (int64_t tile = t0; tile < t1; ++tile) {
    uint32_t raw[TV_ITEMS];  // all loads first (no control flow between them), then the compares
#pragma unroll
    for (int r = 0; r < TV_ITEMS; ++r) raw[r] = __ldg(kw + min(tv_elem(tile, w, r, l), n - 1));
#pragma unroll
    for (int r = 0; r < TV_ITEMS; ++r)
      if (tv_elem(tile, w, r, l) < n) {
        const uint32_t v = image_of(raw[r], dtype, descending);
        tv_merge(a, v, 1u, v, 1u);
      }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    tv_merge(a, __shfl_xor_sync(0xffffffffu, a.mn, d), __shfl_xor_sync(0xffffffffu, a.cmn, d),
             __shfl_xor_sync(0xffffffffu, a.mx, d), __shfl_xor_sync(0xffffffffu, a.cmx, d));
  __shared__ TvStat ws[RS_WARPS];
  if (l == 0) ws[w] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < RS_WARPS; ++k) tv_merge(a, ws[k].mn, ws[k].cmn, ws[k].mx, ws[k].cmx);
    cta_stat[blockIdx.x] = a;
  }
}

// ---- 2. two-valued fast path: stable partition by (image == min) -------------------------
__global__ void __launch_bounds__(RS_BLOCK)
    sort_two_scatter_kernel(const void* __restrict__ keys, int dtype, int descending, int64_t n, int64_t num_tiles,
                            int64_t tiles_per_cta, SortState* __restrict__ st, const TvStat* __restrict__ cta_stat,
                            int32_t* __restrict__ perm) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __shared__ TvStat ws[RS_WARPS];
  __shared__ int wbefore[RS_WARPS];
  // global extrema and their multiplicities from the per-CTA records (every CTA recomputes them: <= 4096 records)
  TvStat a{0xFFFFFFFFu, 0u, 0u, 0u};
  for (int k = threadIdx.x; k < (int)gridDim.x; k += RS_BLOCK) {
    const TvStat c = cta_stat[k];
    tv_merge(a, c.mn, c.cmn, c.mx, c.cmx);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    tv_merge(a, __shfl_xor_sync(0xffffffffu, a.mn, d), __shfl_xor_sync(0xffffffffu, a.cmn, d),
             __shfl_xor_sync(0xffffffffu, a.mx, d), __shfl_xor_sync(0xffffffffu, a.cmx, d));
  if (l == 0) ws[w] = a;
  __syncthreads();
  a = ws[0];
  for (int k = 1; k < RS_WARPS; ++k) tv_merge(a, ws[k].mn, ws[k].cmn, ws[k].mx, ws[k].cmx);
  const uint32_t kmin = a.mn;
  const uint64_t covered = a.mn == a.mx ? (uint64_t)a.cmn : (uint64_t)a.cmn + a.cmx;
  if (covered != (uint64_t)n) return;  // more than two values: general path (uniform over the grid)
  if (blockIdx.x == 0 && threadIdx.x == 0) st->done = 1;
  const int total = (int)a.cmn;
  // keys equal to the minimum in the CTAs before this one
  int before = 0;
  for (int k = threadIdx.x; k < (int)blockIdx.x; k += RS_BLOCK) {
    const TvStat c = cta_stat[k];
    before += c.mn == kmin ? (int)c.cmn : 0;
  }
  before = __reduce_add_sync(0xffffffffu, before);
  if (l == 0) wbefore[w] = before;
  __syncthreads();
  int off = 0;
  for (int k = 0; k < RS_WARPS; ++k) off += wbefore[k];
  __shared__ int wsum[RS_WARPS];
  const unsigned lt = lanemask_lt();
  const uint32_t* __restrict__ kw = reinterpret_cast<const uint32_t*>(keys);
  const int64_t t0 = blockIdx.x * tiles_per_cta, t1 = min(num_tiles, t0 + tiles_per_cta);
  for (int64_t tile = t0; tile < t1; ++tile) {
    uint32_t raw[TV_ITEMS];  // all loads first, then the compares and ballots
#pragma unroll
    for (int r = 0; r < TV_ITEMS; ++r) raw[r] = __ldg(kw + min(tv_elem(tile, w, r, l), n - 1));
    unsigned flags = 0;
#pragma unroll
    for (int r = 0; r < TV_ITEMS; ++r)
      flags |= (tv_elem(tile, w, r, l) < n && image_of(raw[r], dtype, descending) == kmin ? 1u : 0u) << r;
    unsigned ball[TV_ITEMS];
    int c = 0;
#pragma unroll
    for (int r = 0; r < TV_ITEMS; ++r) {
      ball[r] = __ballot_sync(0xffffffffu, (flags >> r) & 1u);
      c += __popc(ball[r]);
    }
    if (l == 0) wsum[w] = c;
    __syncthreads();
    int rank0 = off, tile_total = 0;
#pragma unroll
    for (int k = 0; k < RS_WARPS; ++k) {
      if (k < w) rank0 += wsum[k];
      tile_total += wsum[k];
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < TV_ITEMS; ++r) {
      const int64_t i = tv_elem(tile, w, r, l);
      if (i < n) {
        const int rank = rank0 + __popc(ball[r] & lt);
        const bool s = (ball[r] >> l) & 1u;
        perm[s ? (int64_t)rank : (int64_t)total + (i - rank)] = (int32_t)i;
      }
      rank0 += __popc(ball[r]);
    }
    off += tile_total;
  }
}

// ---- 1 + 2 fused: keys read once, narrow payload rows moved in the same pass -----------------
constexpr int TO_MAX_GPW = 32;  // 256-slot groups per warp the ballots of which fit (8 KB + 1 KB of shared memory)

// rows of one leaf for the 8 slots of a lane (one base pointer, immediate offsets): BATCH loads in flight,
// then BATCH stores (16-byte rows go four at a time to stay inside the register budget).  Only leaves whose
// row is ONE 4 / 8 / 16-byte word ride in the fused kernel; anything else is gathered afterwards.
template <typename T, int BATCH>
__device__ __forceinline__ void scatter_full8(const fgx_leaf_t& lf, int64_t first, const int (&pos)[TV_ITEMS]) {
  const T* __restrict__ sp = reinterpret_cast<const T*>(lf.src) + first;
  T* __restrict__ dst = reinterpret_cast<T*>(lf.dst);
#pragma unroll
  for (int r0 = 0; r0 < TV_ITEMS; r0 += BATCH) {
    T v[BATCH];
#pragma unroll
    for (int r = 0; r < BATCH; ++r) v[r] = __ldg(sp + (r0 + r) * 32);
#pragma unroll
    for (int r = 0; r < BATCH; ++r) dst[pos[r0 + r]] = v[r];
  }
}
// two 4-byte leaves at once: all 16 loads are issued before the first store
__device__ __forceinline__ void scatter_full8_pair(const fgx_leaf_t& la, const fgx_leaf_t& lb, int64_t first,
                                                   const int (&pos)[TV_ITEMS]) {
  const uint32_t* __restrict__ pa = reinterpret_cast<const uint32_t*>(la.src) + first;
  const uint32_t* __restrict__ pb = reinterpret_cast<const uint32_t*>(lb.src) + first;
  uint32_t* __restrict__ da = reinterpret_cast<uint32_t*>(la.dst);
  uint32_t* __restrict__ db = reinterpret_cast<uint32_t*>(lb.dst);
  uint32_t va[TV_ITEMS], vb[TV_ITEMS];
#pragma unroll
  for (int r = 0; r < TV_ITEMS; ++r) va[r] = __ldg(pa + r * 32);
#pragma unroll
  for (int r = 0; r < TV_ITEMS; ++r) vb[r] = __ldg(pb + r * 32);
#pragma unroll
  for (int r = 0; r < TV_ITEMS; ++r) da[pos[r]] = va[r];
#pragma unroll
  for (int r = 0; r < TV_ITEMS; ++r) db[pos[r]] = vb[r];
}
// the one partial group at the end of the array
template <typename T>
__device__ __forceinline__ void scatter_tail8(const fgx_leaf_t& lf, int64_t first, int64_t n,
                                              const int (&pos)[TV_ITEMS]) {
  const T* __restrict__ sp = reinterpret_cast<const T*>(lf.src) + first;
  T* __restrict__ dst = reinterpret_cast<T*>(lf.dst);
#pragma unroll
  for (int r = 0; r < TV_ITEMS; ++r)
    if (first + r * 32 < n) dst[pos[r]] = __ldg(sp + r * 32);
}

__global__ void __launch_bounds__(RS_BLOCK, 4)
    sort_two_onepass_kernel(const void* __restrict__ keys, int dtype, int descending, int64_t n, int64_t groups, int gpw,
                            TvStat* __restrict__ cta_stat, SortState* __restrict__ st, int32_t* __restrict__ perm,
                            const __grid_constant__ LeafTable tab, int n_leaves, int n_leaves4) {
  __shared__ unsigned balls[RS_WARPS][TO_MAX_GPW][TV_ITEMS];
  __shared__ uint32_t gmin[RS_WARPS][TO_MAX_GPW];
  __shared__ TvStat ws[RS_WARPS];
  __shared__ TvStat gs[RS_WARPS];
  __shared__ int wbefore[RS_WARPS];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int64_t g0 = ((int64_t)blockIdx.x * RS_WARPS + w) * gpw;
  const int mine = (int)max((int64_t)0, min((int64_t)gpw, groups - g0));  // warp-uniform
  const uint32_t* __restrict__ kw = reinterpret_cast<const uint32_t*>(keys);
  TvStat a{0xFFFFFFFFu, 0u, 0u, 0u};
#pragma unroll 1
  for (int k = 0; k < mine; ++k) {
    const int64_t base = (g0 + k) * (32 * TV_ITEMS);
    uint32_t v[TV_ITEMS];  // all loads first, then the images
#pragma unroll
    for (int r = 0; r < TV_ITEMS; ++r) v[r] = __ldg(kw + min(base + r * 32 + l, n - 1));
    uint32_t lo = 0xFFFFFFFFu, hi = 0u;
    unsigned valid = 0;
#pragma unroll
    for (int r = 0; r < TV_ITEMS; ++r) {
      v[r] = image_of(v[r], dtype, descending);
      if (base + r * 32 + l < n) {
        valid |= 1u << r;
        lo = min(lo, v[r]);
        hi = max(hi, v[r]);
      }
    }
    const uint32_t mn = __reduce_min_sync(0xffffffffu, lo), mx = __reduce_max_sync(0xffffffffu, hi);
    uint32_t cmn = 0, cmx = 0;
#pragma unroll
    for (int r = 0; r < TV_ITEMS; ++r) {
      const unsigned b = __ballot_sync(0xffffffffu, ((valid >> r) & 1u) && v[r] == mn);
      if (l == r) balls[w][k][r] = b;
      cmn += __popc(b);
    }
    if (mx != mn) {
#pragma unroll
      for (int r = 0; r < TV_ITEMS; ++r) cmx += __popc(__ballot_sync(0xffffffffu, ((valid >> r) & 1u) && v[r] == mx));
    } else {
      cmx = cmn;
    }
    if (l == 0) gmin[w][k] = mn;
    tv_merge(a, mn, cmn, mx, cmx);
  }
  if (l == 0) ws[w] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int q = 1; q < RS_WARPS; ++q) tv_merge(a, ws[q].mn, ws[q].cmn, ws[q].mx, ws[q].cmx);
    uint4 rec = make_uint4(a.mn, a.cmn, a.mx, a.cmx);
    __stcg(reinterpret_cast<uint4*>(cta_stat) + blockIdx.x, rec);
  }
  cg::this_grid().sync();
  // global extrema and their multiplicities from the per-CTA records (every CTA recomputes them)
  TvStat g{0xFFFFFFFFu, 0u, 0u, 0u};
  for (int q = threadIdx.x; q < (int)gridDim.x; q += RS_BLOCK) {
    const uint4 c = __ldcg(reinterpret_cast<const uint4*>(cta_stat) + q);
    tv_merge(g, c.x, c.y, c.z, c.w);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1)
    tv_merge(g, __shfl_xor_sync(0xffffffffu, g.mn, d), __shfl_xor_sync(0xffffffffu, g.cmn, d),
             __shfl_xor_sync(0xffffffffu, g.mx, d), __shfl_xor_sync(0xffffffffu, g.cmx, d));
  if (l == 0) gs[w] = g;
  __syncthreads();
  g = gs[0];
  for (int q = 1; q < RS_WARPS; ++q) tv_merge(g, gs[q].mn, gs[q].cmn, gs[q].mx, gs[q].cmx);
  const uint32_t kmin = g.mn;
  const uint64_t covered = g.mn == g.mx ? (uint64_t)g.cmn : (uint64_t)g.cmn + g.cmx;
  if (covered != (uint64_t)n) return;  // more than two values: general path (uniform over the grid)
  if (blockIdx.x == 0 && threadIdx.x == 0) st->done = 1;
  const int total = (int)g.cmn;
  // keys equal to the minimum in the CTAs before this one, then in this CTA's warps before this one
  int before = 0;
  for (int q = threadIdx.x; q < (int)blockIdx.x; q += RS_BLOCK) {
    const uint4 c = __ldcg(reinterpret_cast<const uint4*>(cta_stat) + q);
    before += c.x == kmin ? (int)c.y : 0;
  }
  before = __reduce_add_sync(0xffffffffu, before);
  if (l == 0) wbefore[w] = before;
  __syncthreads();
  int off = 0;
#pragma unroll
  for (int q = 0; q < RS_WARPS; ++q) off += wbefore[q] + (q < w && ws[q].mn == kmin ? (int)ws[q].cmn : 0);
  const unsigned lt = lanemask_lt();
#pragma unroll 1
  for (int k = 0; k < mine; ++k) {
    const int64_t base = (g0 + k) * (32 * TV_ITEMS);
    const bool sel = gmin[w][k] == kmin;
    int pos[TV_ITEMS];
#pragma unroll
    for (int r = 0; r < TV_ITEMS; ++r) {
      const unsigned b = sel ? balls[w][k][r] : 0u;
      const int64_t i = base + r * 32 + l;
      const int rank = off + __popc(b & lt);
      pos[r] = (b >> l) & 1u ? rank : total + (int)(i - rank);
      if (i < n) perm[pos[r]] = (int32_t)i;
      off += __popc(b);
    }
    const bool full = base + 32 * TV_ITEMS <= n;  // warp-uniform
    int t = 0;
    if (full) {
      // 4-byte leaves (sorted to the front of the table) go two at a time: 16 loads in flight per thread instead of
      // 8 -- at 1,024 threads per SM that is the difference between ~32 KB and ~64 KB of reads in flight per SM
      for (; t + 1 < n_leaves4; t += 2) scatter_full8_pair(tab.leaf[t], tab.leaf[t + 1], base + l, pos);
    }
#pragma unroll 1
    for (; t < n_leaves; ++t) {
      const fgx_leaf_t& lf = tab.leaf[t];
      const int vec = tab.vec[t];
      if (full) {
        if (vec == 4) scatter_full8<uint32_t, 8>(lf, base + l, pos);
        else if (vec == 8) scatter_full8<uint2, 8>(lf, base + l, pos);
        else scatter_full8<uint4, 4>(lf, base + l, pos);
      } else {
        if (vec == 4) scatter_tail8<uint32_t>(lf, base + l, n, pos);
        else if (vec == 8) scatter_tail8<uint2>(lf, base + l, n, pos);
        else scatter_tail8<uint4>(lf, base + l, n, pos);
      }
    }
  }
}

// ---- 3. general path --------------------------------------------------------------------
__device__ __forceinline__ void sort_prep_body(const void* __restrict__ keys, int dtype, int descending, int64_t n,
                                               uint32_t* __restrict__ img, SortState* __restrict__ st) {
  __shared__ uint32_t h[4][256];
  for (int k = threadIdx.x; k < 1024; k += blockDim.x) (&h[0][0])[k] = 0;
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x; i0 < n; i0 += stride) {  // warp-uniform trip count
    const int64_t i = i0 + threadIdx.x;
    const bool valid = i < n;
    const uint32_t v = valid ? key_image(keys, dtype, descending, i) : 0u;
    if (valid) img[i] = v;
    hist_add(h[0], v & 255u, valid);
    hist_add(h[1], (v >> 8) & 255u, valid);
    hist_add(h[2], (v >> 16) & 255u, valid);
    hist_add(h[3], v >> 24, valid);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < 1024; k += blockDim.x) {
    const uint32_t c = (&h[0][0])[k];
    if (c) atomicAdd(&st->hist[0][0] + k, c);
  }
}

__device__ __forceinline__ void sort_plan_body(int64_t n, SortState* __restrict__ st) {
  __shared__ uint32_t sh[256];
  __shared__ int triv;
  const int t = threadIdx.x;
  for (int p = 0; p < 4; ++p) {
    const uint32_t c = __ldcg(&st->hist[p][t]);
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
    if (t == 0) st->trivial[p] = triv;
    __syncthreads();
  }
}

__device__ __forceinline__ uint32_t digit_of(uint32_t v, int pass) { return (v >> (8 * pass)) & 255u; }

// tile digit counts -> block_hist[digit][tile]
__device__ __forceinline__ void sort_hist_body(const uint32_t* img0, const uint32_t* img1, int64_t n,
                                               int64_t num_tiles, int pass, const SortState* st,
                                               uint32_t* block_hist) {
  const uint32_t* img = (passes_done(st, pass) & 1) ? img1 : img0;
  __shared__ uint32_t h[256];
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = tile * RS_TILE;
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
      const int64_t i = base + r * RS_BLOCK + threadIdx.x;
      const bool valid = i < n;
      hist_add(h, valid ? digit_of(__ldcg(img + i), pass) : 0u, valid);
    }
    __syncthreads();
    block_hist[(int64_t)threadIdx.x * num_tiles + tile] = h[threadIdx.x];
    __syncthreads();
  }
}

// one CTA per digit: exclusive scan of block_hist[digit][0..num_tiles) in place
__device__ __forceinline__ void sort_scan_body(int64_t num_tiles, uint32_t* block_hist) {
  __shared__ uint32_t wsum[8];
  __shared__ uint32_t carry;
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int digit = blockIdx.x; digit < 256; digit += gridDim.x) {
    uint32_t* row = block_hist + (int64_t)digit * num_tiles;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int64_t b = 0; b < num_tiles; b += 256) {
      const int64_t t = b + threadIdx.x;
      const uint32_t v = t < num_tiles ? __ldcg(row + t) : 0u;
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
}

// stable scatter of one pass
__device__ __forceinline__ void sort_scatter_body(uint32_t* img0, uint32_t* img1, int32_t* idx0, int32_t* idx1,
                                                  int64_t n, int64_t num_tiles, int pass, const SortState* st,
                                                  const uint32_t* block_hist) {
  const int done = passes_done(st, pass);
  const int cur = done & 1;
  const uint32_t* kin = cur ? img1 : img0;
  uint32_t* kout = cur ? img0 : img1;
  const int32_t* iin = cur ? idx1 : idx0;
  int32_t* iout = cur ? idx0 : idx1;
  const bool iota = done == 0;
  __shared__ uint32_t wcnt[RS_WARPS][256];
  __shared__ uint32_t gbase[256];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const unsigned lt = lanemask_lt();
  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    for (int k = threadIdx.x; k < RS_WARPS * 256; k += RS_BLOCK) (&wcnt[0][0])[k] = 0;
    gbase[threadIdx.x] = __ldcg(&st->base[pass][threadIdx.x]) + __ldcg(&block_hist[(int64_t)threadIdx.x * num_tiles + tile]);
    __syncthreads();
    const int64_t wbase = tile * RS_TILE + (int64_t)w * (32 * RS_ITEMS);
    uint32_t key[RS_ITEMS];
    uint32_t lpos[RS_ITEMS];
#pragma unroll
    for (int r = 0; r < RS_ITEMS; ++r) {
      const int64_t i = wbase + r * 32 + l;
      const bool valid = i < n;
      key[r] = valid ? __ldcg(kin + i) : 0xFFFFFFFFu;
      // invalid lanes use a digit value outside 0..255 so they never match a real one
      const uint32_t d = valid ? digit_of(key[r], pass) : 256u;
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
        const uint32_t d = digit_of(key[r], pass);
        const uint32_t pos = gbase[d] + wcnt[w][d] + lpos[r];
        kout[pos] = key[r];
        iout[pos] = iota ? (int32_t)i : __ldcg(iin + i);
      }
    }
    __syncthreads();
  }
}

// result -> perm (idx0 aliases perm): copy only if the final order sits in idx1, or is still iota
__device__ __forceinline__ void sort_finish_body(const int32_t* idx1, int32_t* perm, int64_t n, const SortState* st) {
  const int done = passes_done(st, 4);
  const bool iota = done == 0;
  if (!iota && (done & 1) == 0) return;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    perm[i] = iota ? (int32_t)i : __ldcg(idx1 + i);
}

// The general path as kernels.  Default: ONE cooperative launch with grid-wide syncs between the phases,
// so that a sort the two-valued path already finished costs one empty launch instead of fifteen (a tick
// replayed as a CUDA graph pays ~2 us per empty node).  FGX_SORT_MULTI=1 selects one launch per phase.
__global__ void __launch_bounds__(RS_BLOCK, 2) sort_general_kernel(const void* __restrict__ keys, int dtype, int descending,
                                                                int64_t n, int64_t num_tiles, uint32_t* img0,
                                                                uint32_t* img1, int32_t* idx0, int32_t* idx1,
                                                                uint32_t* block_hist, SortState* st) {
  if (__ldcg(&st->done)) return;  // uniform over the grid: written by an earlier launch
  cg::grid_group grid = cg::this_grid();
  sort_prep_body(keys, dtype, descending, n, img0, st);
  grid.sync();
  if (blockIdx.x == 0) sort_plan_body(n, st);
  grid.sync();
  for (int pass = 0; pass < 4; ++pass) {
    if (__ldcg(&st->trivial[pass])) continue;
    sort_hist_body(img0, img1, n, num_tiles, pass, st, block_hist);
    grid.sync();
    sort_scan_body(num_tiles, block_hist);
    grid.sync();
    sort_scatter_body(img0, img1, idx0, idx1, n, num_tiles, pass, st, block_hist);
    grid.sync();
  }
  sort_finish_body(idx1, idx0, n, st);
}

__global__ void __launch_bounds__(256) sort_prep_kernel(const void* __restrict__ keys, int dtype, int descending,
                                                        int64_t n, uint32_t* __restrict__ img,
                                                        SortState* __restrict__ st) {
  if (st->done) return;
  sort_prep_body(keys, dtype, descending, n, img, st);
}
__global__ void __launch_bounds__(256) sort_plan_kernel(int64_t n, SortState* __restrict__ st) {
  if (st->done) return;
  sort_plan_body(n, st);
}
__global__ void __launch_bounds__(RS_BLOCK) sort_hist_kernel(const uint32_t* img0, const uint32_t* img1, int64_t n,
                                                             int64_t num_tiles, int pass, const SortState* st,
                                                             uint32_t* block_hist) {
  if (st->done || st->trivial[pass]) return;
  sort_hist_body(img0, img1, n, num_tiles, pass, st, block_hist);
}
__global__ void __launch_bounds__(256) sort_scan_kernel(int64_t num_tiles, int pass, const SortState* st,
                                                        uint32_t* block_hist) {
  if (st->done || st->trivial[pass]) return;
  sort_scan_body(num_tiles, block_hist);
}
__global__ void __launch_bounds__(RS_BLOCK) sort_scatter_kernel(uint32_t* img0, uint32_t* img1, int32_t* idx0,
                                                                int32_t* idx1, int64_t n, int64_t num_tiles, int pass,
                                                                const SortState* st, const uint32_t* block_hist) {
  if (st->done || st->trivial[pass]) return;
  sort_scatter_body(img0, img1, idx0, idx1, n, num_tiles, pass, st, block_hist);
}
__global__ void __launch_bounds__(256) sort_finish_kernel(const int32_t* idx1, int32_t* perm, int64_t n,
                                                          const SortState* st) {
  if (st->done) return;
  sort_finish_body(idx1, perm, n, st);
}

// ---- small paths (latency-bound sizes: one or two graph nodes instead of the 15 of the radix path) ----
// n <= 256: single CTA, bitonic sort of the 64-bit composite (image << 32 | index) in shared memory.
// The composites are unique, so the (unstable) network yields the stable order; every thread owns a
// compare-exchange PAIR (i, i + j).
__global__ void __launch_bounds__(128) sort_small_kernel(const void* __restrict__ keys, int dtype, int descending,
                                                         int n, int32_t* __restrict__ perm) {
  __shared__ unsigned long long s[SORT_SMALL_N];
  int m = 1;
  while (m < n) m <<= 1;
  for (int i = threadIdx.x; i < m; i += blockDim.x)
    s[i] = i < n ? (((unsigned long long)key_image(keys, dtype, descending, i) << 32) | (unsigned)i)
                 : 0xFFFFFFFFFFFFFFFFull;
  __syncthreads();
  const int half = m >> 1;
  for (int k = 2; k <= m; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      const int t = threadIdx.x;
      if (t < half) {
        const int i = 2 * t - (t & (j - 1));  // bit j of i is clear
        const unsigned long long a = s[i], b = s[i + j];
        const bool up = (i & k) == 0;
        if ((a > b) == up) { s[i] = b; s[i + j] = a; }
      }
      __syncthreads();
    }
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) perm[i] = (int32_t)(s[i] & 0xFFFFFFFFull);
}

// 256 < n <= 16384: all-pairs RANK sort over the whole GPU.  rank(i) = #{j : composite_j < composite_i}
// is the final position of slot i (the composites are unique), so n^2 64-bit compares -- a few
// microseconds of integer work spread over every SM -- replace log^2 n dependent shared-memory steps
// on one SM.  grid = (i-tiles of 1024 slots) x (j-splits); partial ranks are added atomically and the
// LAST CTA of an i-tile to arrive (ticket counter, no waiting) writes perm[rank[i]] = i.
constexpr int RK_BLOCK = 256, RK_IPT = 4, RK_ITILE = RK_BLOCK * RK_IPT, RK_JMAX = 2048;
constexpr int SORT_RANK_N = 16384;

__global__ void __launch_bounds__(RK_BLOCK) sort_rank_kernel(const void* __restrict__ keys, int dtype, int descending,
                                                             int n, int jchunk, int32_t* __restrict__ rank,
                                                             int32_t* __restrict__ tickets,
                                                             int32_t* __restrict__ perm) {
  __shared__ __align__(16) unsigned long long sj[RK_JMAX];
  __shared__ int last;
  const int j0 = blockIdx.y * jchunk, cnt = min(n, j0 + jchunk) - j0, padded = (cnt + 7) & ~7;
  for (int k = threadIdx.x; k < padded; k += RK_BLOCK)  // padding compares greater than every slot
    sj[k] = k < cnt ? (((unsigned long long)key_image(keys, dtype, descending, j0 + k) << 32) | (unsigned)(j0 + k))
                    : 0xFFFFFFFFFFFFFFFFull;
  unsigned long long ci[RK_IPT];
  int c[RK_IPT];
#pragma unroll
  for (int q = 0; q < RK_IPT; ++q) {
    const int i = blockIdx.x * RK_ITILE + q * RK_BLOCK + threadIdx.x;
    ci[q] = i < n ? (((unsigned long long)key_image(keys, dtype, descending, i) << 32) | (unsigned)i) : 0ull;
    c[q] = 0;
  }
  __syncthreads();
  for (int k = 0; k < padded; k += 8) {
    ulonglong2 v[4];
#pragma unroll
    for (int h = 0; h < 4; ++h) v[h] = *reinterpret_cast<const ulonglong2*>(&sj[k + 2 * h]);  // broadcast reads
#pragma unroll
    for (int h = 0; h < 4; ++h)
#pragma unroll
      for (int q = 0; q < RK_IPT; ++q) c[q] += (v[h].x < ci[q] ? 1 : 0) + (v[h].y < ci[q] ? 1 : 0);
  }
#pragma unroll
  for (int q = 0; q < RK_IPT; ++q) {
    const int i = blockIdx.x * RK_ITILE + q * RK_BLOCK + threadIdx.x;
    if (i < n && c[q]) atomicAdd(&rank[i], c[q]);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) last = atomicAdd(&tickets[blockIdx.x], 1) == (int)gridDim.y - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
#pragma unroll
  for (int q = 0; q < RK_IPT; ++q) {
    const int i = blockIdx.x * RK_ITILE + q * RK_BLOCK + threadIdx.x;
    if (i < n) perm[__ldcg(&rank[i])] = i;
  }
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace fgx

extern "C" {

size_t fgx_argsort_scratch_bytes(int64_t n) {
  using namespace fgx;
  if (n < 0) n = 0;
  const int64_t tiles = ceil_div(n, RS_TILE);
  size_t b = align_up(sizeof(SortState), 256);
  b += align_up((size_t)TV_MAX_GRID * sizeof(TvStat), 256);  // per-CTA records of the two-valued path
  b += 2 * align_up((size_t)n * 4, 256);            // image ping-pong
  b += align_up((size_t)n * 4, 256);                // idx1 (idx0 is the caller's perm)
  b += align_up((size_t)tiles * 256 * 4 + 4, 256);  // block_hist[256][tiles]
  return b;
}

// argsort (+ optional payload move).  `narrow`/`wide`: the leaves of fgx_sort_rows (nn = nw = 0 for a bare argsort).
static int sort_impl(const char* what, const void* keys, int32_t key_dtype, int64_t n, int32_t descending,
                     int32_t* perm, void* scratch, size_t scratch_bytes, const fgx::LeafTable& narrow, int nn,
                     const fgx::LeafTable& wide, int nw, cudaStream_t s) {
  using namespace fgx;
  const int desc = descending ? 1 : 0;
  // leaves the fused two-valued kernel can move (a row = one 4 / 8 / 16-byte word) and the other narrow ones
  LeafTable fused{}, rest{};
  int nf = 0, nr = 0;
  for (int k = 0; k < nn; ++k) {
    const bool one_word = narrow.leaf[k].row_bytes == narrow.vec[k] && narrow.vec[k] >= 4;
    LeafTable& t = one_word ? fused : rest;
    int& c = one_word ? nf : nr;
    t.leaf[c] = narrow.leaf[k];
    t.vec[c++] = narrow.vec[k];
  }
  // the fused kernel wants the 4-byte leaves first (it moves them in pairs)
  int nf4 = 0;
  for (int k = 0; k < nf; ++k)
    if (fused.vec[k] == 4) {
      const fgx_leaf_t lf = fused.leaf[k];
      const int32_t vv = fused.vec[k];
      fused.leaf[k] = fused.leaf[nf4];
      fused.vec[k] = fused.vec[nf4];
      fused.leaf[nf4] = lf;
      fused.vec[nf4++] = vv;
    }
  if (n <= SORT_SMALL_N) {
    sort_small_kernel<<<1, 128, 0, s>>>(keys, key_dtype, desc, (int)n, perm);
    if (nn || nw) return launch_gather(narrow, nn, wide, nw, perm, n, n, nullptr, s);
    return check_launch(what);
  }
  FGX_REQUIRE(scratch && scratch_bytes >= fgx_argsort_scratch_bytes(n), "%s: scratch too small", what);
  FGX_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "%s: scratch must be 256-byte aligned", what);
  if (n <= SORT_RANK_N) {
    const int itiles = (int)ceil_div(n, RK_ITILE);
    int64_t js = ceil_div((int64_t)2 * sm_count(), itiles);
    if (js > ceil_div(n, 64)) js = ceil_div(n, 64);
    int jchunk = (int)ceil_div(n, js);
    jchunk = (jchunk + 7) & ~7;
    if (jchunk > RK_JMAX) jchunk = RK_JMAX;
    js = ceil_div(n, jchunk);
    int32_t* rank = reinterpret_cast<int32_t*>(scratch);  // rank[n] then one ticket per i-tile
    cudaError_t e = cudaMemsetAsync(rank, 0, sizeof(int32_t) * (size_t)(n + itiles), s);
    if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    sort_rank_kernel<<<dim3(itiles, (unsigned)js), RK_BLOCK, 0, s>>>(keys, key_dtype, desc, (int)n, jchunk, rank,
                                                                    rank + n, perm);
    if (nn || nw) return launch_gather(narrow, nn, wide, nw, perm, n, n, nullptr, s);
    return check_launch(what);
  }
  const int64_t tiles = ceil_div(n, RS_TILE);
  char* p = reinterpret_cast<char*>(scratch);
  SortState* st = reinterpret_cast<SortState*>(p);
  p += align_up(sizeof(SortState), 256);
  TvStat* cta_stat = reinterpret_cast<TvStat*>(p);
  p += align_up((size_t)TV_MAX_GRID * sizeof(TvStat), 256);
  uint32_t* img0 = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)n * 4, 256);
  uint32_t* img1 = reinterpret_cast<uint32_t*>(p);
  p += align_up((size_t)n * 4, 256);
  int32_t* idx1 = reinterpret_cast<int32_t*>(p);
  p += align_up((size_t)n * 4, 256);
  uint32_t* block_hist = reinterpret_cast<uint32_t*>(p);
  int32_t* idx0 = perm;

  cudaError_t e = cudaMemsetAsync(st, 0, sizeof(SortState), s);
  if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  const int g = grid_for(n, 256, 8);
  // two-valued fast path
  static const bool two_pass = [] { const char* v = getenv("FGX_SORT_TWO_PASS"); return v && atoi(v) == 1; }();
  static int tv_per_sm = -1;
  if (tv_per_sm < 0) {
    int q = 0;
    cudaError_t oe = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, sort_two_onepass_kernel, RS_BLOCK, 0);
    tv_per_sm = (oe == cudaSuccess && q >= 1) ? q : 0;
  }
  int64_t tv_coop = (int64_t)sm_count() * tv_per_sm;
  if (tv_coop > TV_MAX_GRID) tv_coop = TV_MAX_GRID;
  const int64_t groups = ceil_div(n, 32 * TV_ITEMS);
  bool fused_rows = false;
  if (!two_pass && tv_coop >= 1 && ceil_div(groups, tv_coop * RS_WARPS) <= TO_MAX_GPW) {
    // ONE cooperative launch: stats + partition (+ the narrow payload rows) with the keys read once
    int gpw = (int)ceil_div(groups, tv_coop * RS_WARPS);
    const int tv_grid = (int)ceil_div(groups, (int64_t)gpw * RS_WARPS);
    const void* keys_arg = keys;
    int dtype_arg = key_dtype, desc_arg = desc, nn_arg = nf;
    int64_t n_arg = n, groups_arg = groups;
    LeafTable tab = fused;
    void* args[] = {&keys_arg, &dtype_arg, &desc_arg, &n_arg, &groups_arg, &gpw, &cta_stat, &st, &perm, &tab, &nn_arg, &nf4};
    e = cudaLaunchCooperativeKernel((const void*)sort_two_onepass_kernel, dim3((unsigned)tv_grid), dim3(RS_BLOCK), args,
                                    0, s);
    if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "%s: cooperative launch: %s", what, cudaGetErrorString(e));
    fused_rows = true;
  } else {
    // contiguous tile range per CTA: a statistics pass, then the scatter pass
    const int64_t tv_tiles = ceil_div(n, TV_TILE);
    int64_t cap = (int64_t)sm_count() * 8;
    if (cap > TV_MAX_GRID) cap = TV_MAX_GRID;
    const int64_t per = ceil_div(tv_tiles, tv_tiles < cap ? tv_tiles : cap);
    const int tv_grid = (int)ceil_div(tv_tiles, per);
    sort_two_stats_kernel<<<tv_grid, RS_BLOCK, 0, s>>>(keys, key_dtype, desc, n, tv_tiles, per, cta_stat);
    sort_two_scatter_kernel<<<tv_grid, RS_BLOCK, 0, s>>>(keys, key_dtype, desc, n, tv_tiles, per, st, cta_stat, perm);
  }
  // general path (returns at once when st->done)
  static const bool multi = [] { const char* v = getenv("FGX_SORT_MULTI"); return v && atoi(v) == 1; }();
  if (!multi) {
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sort_general_kernel, RS_BLOCK, 0);
    if (e != cudaSuccess || per_sm < 1) return fail(FGX_ERR_CUDA, "%s: occupancy query failed", what);
    int64_t coop = (int64_t)sm_count() * per_sm;  // all CTAs must be co-resident for the grid-wide syncs
    const int64_t useful = ceil_div(n, RS_BLOCK * 4);
    if (coop > useful) coop = useful < 1 ? 1 : useful;
    const void* keys_arg = keys;
    int dtype_arg = key_dtype, desc_arg = desc;
    int64_t n_arg = n, tiles_arg = tiles;
    void* args[] = {&keys_arg, &dtype_arg, &desc_arg, &n_arg, &tiles_arg, &img0, &img1, &idx0, &idx1, &block_hist, &st};
    e = cudaLaunchCooperativeKernel((const void*)sort_general_kernel, dim3((unsigned)coop), dim3(RS_BLOCK), args, 0, s);
    if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "%s: cooperative launch: %s", what, cudaGetErrorString(e));
  } else {
    sort_prep_kernel<<<g, 256, 0, s>>>(keys, key_dtype, desc, n, img0, st);
    sort_plan_kernel<<<1, 256, 0, s>>>(n, st);
    const int gt = (int)(tiles < (int64_t)sm_count() * 4 ? tiles : (int64_t)sm_count() * 4);
    for (int pass = 0; pass < 4; ++pass) {
      sort_hist_kernel<<<gt, RS_BLOCK, 0, s>>>(img0, img1, n, tiles, pass, st, block_hist);
      sort_scan_kernel<<<256, 256, 0, s>>>(tiles, pass, st, block_hist);
      sort_scatter_kernel<<<gt, RS_BLOCK, 0, s>>>(img0, img1, idx0, idx1, n, tiles, pass, st, block_hist);
    }
    sort_finish_kernel<<<g, 256, 0, s>>>(idx1, perm, n, st);
  }
  // payload: the fused launch moved the one-word rows iff it finished the sort (st->done); all other rows gather
  if (nf) {
    const int rc = launch_gather(fused, nf, wide, 0, perm, n, n, fused_rows ? &st->done : nullptr, s);
    if (rc != FGX_OK) return rc;
  }
  if (nr || nw) return launch_gather(rest, nr, wide, nw, perm, n, n, nullptr, s);
  return check_launch(what);
}

int fgx_argsort(const void* keys, int32_t key_dtype, int64_t n, int32_t descending, int32_t* perm, void* scratch,
                size_t scratch_bytes, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n >= 0 && n < (1ll << 31), "fgx_argsort: n out of range");
  FGX_REQUIRE(key_dtype == FGX_F32 || key_dtype == FGX_I32, "fgx_argsort: key dtype must be FGX_F32 or FGX_I32");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(keys && perm, "fgx_argsort: null pointer");
  static const LeafTable none{};
  return sort_impl("fgx_argsort", keys, key_dtype, n, descending, perm, scratch, scratch_bytes, none, 0, none, 0,
                   as_stream(stream));
}

int fgx_sort_rows(const void* keys, int32_t key_dtype, int64_t n, int32_t descending, int32_t* perm,
                  const fgx_leaf_t* leaves_host, int32_t n_leaves, void* scratch, size_t scratch_bytes,
                  fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n >= 0 && n < (1ll << 31), "fgx_sort_rows: n out of range");
  FGX_REQUIRE(key_dtype == FGX_F32 || key_dtype == FGX_I32, "fgx_sort_rows: key dtype must be FGX_F32 or FGX_I32");
  FGX_REQUIRE(n_leaves >= 0 && n_leaves <= FGX_MAX_LEAVES, "fgx_sort_rows: at most %d leaves per call", FGX_MAX_LEAVES);
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(keys && perm && (n_leaves == 0 || leaves_host), "fgx_sort_rows: null pointer");
  LeafTable narrow{}, wide{};
  int nn = 0, nw = 0;
  const int rc = classify_leaves("fgx_sort_rows", leaves_host, n_leaves, &narrow, &nn, &wide, &nw);
  if (rc != FGX_OK) return rc;
  return sort_impl("fgx_sort_rows", keys, key_dtype, n, descending, perm, scratch, scratch_bytes, narrow, nn, wide, nw,
                   as_stream(stream));
}
}
