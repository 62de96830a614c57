// K3 (part 1): select_agents as a stable two-way partition of slot indices.
//
// Reference: agent_methods.py:66-71 -- mask -> where(mask,1.,0.) -> full stable
// argsort(-mask) -> (count, permutation).  The permutation is "selected slots
// ascending, then unselected slots ascending", so the O(N log N) comparison
// sort of the reference is a 1-bit stable partition here:
//   one pass (default, up to ~77 M slots): ONE cooperative launch.  Every warp owns a contiguous run
//           of 256-slot groups: it evaluates the predicate ONCE (leaves read once), keeps the
//           ballots in shared memory and publishes its CTA's count; after a grid-wide sync every
//           CTA sums the counts before it (its offset) and all of them (the total -- the position
//           of an unselected slot depends on it, which is why a chained look-back scan alone
//           cannot finish the job) and scatters perm[rank] / perm[total + i - rank] from the
//           saved ballots.  HBM traffic = leaves once + 4 B/slot.
//   two pass (larger sets): pass A counts per CTA, pass B re-evaluates the predicate and scatters.
//   small   up to 16,384 slots: ONE 1024-thread CTA holds every flag in registers
//           (one round of loads, one block scan, one scatter).
// The predicate is evaluated in-kernel from the agent leaves (up to four
// comparisons combined by a truth table), so no mask array is materialised.
// Algorithmic HBM bytes: predicate leaves read once + 4 B/slot written.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace fgx {

constexpr int SEL_BLOCK = 256;
constexpr int SEL_ITEMS = 8;
constexpr int SEL_TILE = SEL_BLOCK * SEL_ITEMS;  // 2048 slots per tile
constexpr int SEL_WARPS = SEL_BLOCK / 32;

// A comparison on a value already in a register (bits = raw 32-bit word, or the byte zero-extended)
__device__ __forceinline__ bool compare_bits(uint32_t bits, int dtype, int op, uint32_t imm) {
  if (dtype == FGX_F32) {
    const float v = __uint_as_float(bits), c = __uint_as_float(imm);
    switch (op) {
      case FGX_OP_EQ: return v == c;
      case FGX_OP_NE: return v != c;
      case FGX_OP_LT: return v < c;
      case FGX_OP_LE: return v <= c;
      case FGX_OP_GT: return v > c;
      case FGX_OP_GE: return v >= c;
      default: return v != 0.0f;
    }
  }
  const int32_t v = (int32_t)bits, c = (int32_t)imm;
  switch (op) {
    case FGX_OP_EQ: return v == c;
    case FGX_OP_NE: return v != c;
    case FGX_OP_LT: return v < c;
    case FGX_OP_LE: return v <= c;
    case FGX_OP_GT: return v > c;
    case FGX_OP_GE: return v >= c;
    default: return v != 0;
  }
}

// element handled by (warp w, round r, lane l) of a tile: contiguous per warp
__device__ __forceinline__ int64_t tile_elem(int64_t tile, int w, int r, int l) {
  return tile * SEL_TILE + w * (32 * SEL_ITEMS) + r * 32 + l;
}

// Truth values of ONE term for a lane's ITEMS slots start + r*32 + l, as a bit mask.  The loads are
// issued back to back with no control flow between them (all in flight together); the comparison
// runs on registers afterwards.
template <typename T, int ITEMS>
__device__ __forceinline__ unsigned term_mask(const fgx_pred_term_t& t, int64_t start, int64_t n, int l) {
  const T* __restrict__ base = reinterpret_cast<const T*>(t.data);
  uint32_t raw[ITEMS];
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const int64_t i = min(start + r * 32 + l, n - 1);
    raw[r] = (uint32_t)__ldg(base + i * t.stride);
  }
  unsigned m = 0;
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) m |= (compare_bits(raw[r], t.dtype, t.op, t.imm_bits) ? 1u : 0u) << r;
  return m;
}

// selected flags of a lane's ITEMS slots (bit r = slot start + r*32 + l)
template <int ITEMS>
__device__ __forceinline__ unsigned flags_at(const fgx_predicate_t& p, int64_t start, int64_t n, int l) {
  unsigned tm[4] = {0u, 0u, 0u, 0u};
#pragma unroll
  for (int t = 0; t < 4; ++t)
    if (t < p.n_terms)
      tm[t] = p.term[t].dtype == FGX_U8 ? term_mask<uint8_t, ITEMS>(p.term[t], start, n, l)
                                        : term_mask<uint32_t, ITEMS>(p.term[t], start, n, l);
  unsigned f = 0;
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const unsigned idx = ((tm[0] >> r) & 1u) | (((tm[1] >> r) & 1u) << 1) | (((tm[2] >> r) & 1u) << 2) |
                         (((tm[3] >> r) & 1u) << 3);
    const bool in = start + r * 32 + l < n;
    f |= (in ? (p.lut >> idx) & 1u : 0u) << r;
  }
  return f;
}

__device__ __forceinline__ unsigned tile_flags(const fgx_predicate_t& p, int64_t tile, int64_t n, int w, int l) {
  return flags_at<SEL_ITEMS>(p, tile_elem(tile, w, 0, 0), n, l);
}

__device__ __forceinline__ int tile_scatter(const fgx_predicate_t& p, int64_t tile, int64_t n, int32_t tile_off,
                                            int32_t total, int32_t* __restrict__ perm, int* warp_sums) {
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const unsigned flags = tile_flags(p, tile, n, w, l);
  unsigned ball[SEL_ITEMS];
  int c = 0;
#pragma unroll
  for (int r = 0; r < SEL_ITEMS; ++r) {
    ball[r] = __ballot_sync(0xffffffffu, (flags >> r) & 1u);
    c += __popc(ball[r]);
  }
  if (l == 0) warp_sums[w] = c;
  __syncthreads();
  int before = tile_off, tile_total = 0;
#pragma unroll
  for (int k = 0; k < SEL_WARPS; ++k) {
    if (k < w) before += warp_sums[k];
    tile_total += warp_sums[k];
  }
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
  return tile_total;
}

// Both passes give CTA b the CONTIGUOUS tile range [b*T, (b+1)*T), so only per-CTA counts are
// exchanged: pass B derives its starting rank (sum of the counts of the CTAs before it) and the
// global total with two block reductions over <= SEL_MAX_GRID ints -- no serial scan, no atomics.
constexpr int SEL_MAX_GRID = 4096;

__device__ __forceinline__ int block_sum(int v, int* warp_sums) {
  v = __reduce_add_sync(0xffffffffu, v);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int k = 0; k < SEL_WARPS; ++k) t += warp_sums[k];
  __syncthreads();
  return t;
}

// pass A: selected count of each CTA's tile range
__global__ void __launch_bounds__(SEL_BLOCK) select_count_kernel(const __grid_constant__ fgx_predicate_t p, int64_t n,
                                                                int64_t num_tiles, int64_t tiles_per_cta,
                                                                int32_t* __restrict__ cta_count) {
  __shared__ int warp_sums[SEL_WARPS];
  const int64_t t0 = blockIdx.x * tiles_per_cta, t1 = min(num_tiles, t0 + tiles_per_cta);
  // per-thread partial counts over the CTA's tiles (no barrier between tiles: their loads overlap), one
  // block reduction at the end
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  int c = 0;
  for (int64_t tile = t0; tile < t1; ++tile) c += __popc(tile_flags(p, tile, n, w, l));
  c = block_sum(c, warp_sums);
  if (threadIdx.x == 0) cta_count[blockIdx.x] = c;
}

// pass B: scatter
__global__ void __launch_bounds__(SEL_BLOCK) select_scatter_kernel(const __grid_constant__ fgx_predicate_t p, int64_t n,
                                                                  int64_t num_tiles, int64_t tiles_per_cta,
                                                                  const int32_t* __restrict__ cta_count,
                                                                  int32_t* __restrict__ perm,
                                                                  int32_t* __restrict__ count_out) {
  __shared__ int warp_sums[SEL_WARPS];
  int before = 0, all = 0;
  for (int k = threadIdx.x; k < (int)gridDim.x; k += SEL_BLOCK) {
    const int v = cta_count[k];
    all += v;
    if (k < (int)blockIdx.x) before += v;
  }
  const int total = block_sum(all, warp_sums);
  int off = block_sum(before, warp_sums);
  if (blockIdx.x == 0 && threadIdx.x == 0) *count_out = total;
  const int64_t t0 = blockIdx.x * tiles_per_cta, t1 = min(num_tiles, t0 + tiles_per_cta);
  for (int64_t tile = t0; tile < t1; ++tile) off += tile_scatter(p, tile, n, off, total, perm, warp_sums);
}

// Fast path of the common predicate shape -- ONE comparison on a contiguous 4-byte leaf (draw == 1, flag != 0,
// energy > t, ...): one base pointer with immediate offsets and the operator resolved to a closed-range test once
// per thread instead of a switch per slot.  It lives in its own kernel
// instantiation: with the general evaluator inlined next to a switch over dtype x operator the one-pass kernel was
// 17 k instructions long and spent 13 % of its 30 us waiting for instruction fetches (ncu: no_instruction).
// every operator is a closed-range test, possibly negated: EQ [c, c], NE not [c, c], LT [-inf, pred(c)], LE [-inf, c],
// GT [succ(c), +inf], GE [c, +inf], NONZERO not [0, 0].  Floats: b >= lo && b <= hi is false for a NaN b or a NaN
// bound, which is exactly what the ordered comparisons give (and NE / NONZERO, negated, then hold); -0 == +0 as in
// IEEE.  Integers: (uint32)(b - lo) <= (uint32)(hi - lo).
struct FastRange {
  uint32_t lo, hi;  // bit patterns (float) or values (int32)
  bool neg, is_float;
};
__device__ __forceinline__ FastRange fast_range(const fgx_pred_term_t& t) {
  FastRange r;
  r.is_float = t.dtype == FGX_F32;
  r.neg = t.op == FGX_OP_NE || t.op == FGX_OP_NONZERO;
  if (r.is_float) {
    const float c = t.op == FGX_OP_NONZERO ? 0.0f : __uint_as_float(t.imm_bits);
    const float inf = __int_as_float(0x7f800000);
    float lo = c, hi = c;
    if (t.op == FGX_OP_LT) { lo = -inf; hi = nextafterf(c, -inf); }
    else if (t.op == FGX_OP_LE) lo = -inf;
    else if (t.op == FGX_OP_GT) { lo = nextafterf(c, inf); hi = inf; }
    else if (t.op == FGX_OP_GE) hi = inf;
    // x < -inf and x > +inf hold for no x: an empty range (lo > hi) -- nextafter(-inf, -inf) would keep -inf in it
    if ((t.op == FGX_OP_LT && c == -inf) || (t.op == FGX_OP_GT && c == inf)) { lo = 1.0f; hi = 0.0f; }
    r.lo = __float_as_uint(lo);
    r.hi = __float_as_uint(hi);
  } else {
    const int32_t c = t.op == FGX_OP_NONZERO ? 0 : (int32_t)t.imm_bits;
    const int32_t mn = (int32_t)0x80000000, mx = 0x7fffffff;
    int32_t lo = c, hi = c;
    bool empty = false;
    if (t.op == FGX_OP_LT) { lo = mn; hi = c - 1; empty = c == mn; }
    else if (t.op == FGX_OP_LE) lo = mn;
    else if (t.op == FGX_OP_GT) { lo = c + 1; hi = mx; empty = c == mx; }
    else if (t.op == FGX_OP_GE) hi = mx;
    if (empty) { r.neg = !r.neg; lo = mn; hi = mx; }  // nothing matches = not (everything)
    r.lo = (uint32_t)lo;
    r.hi = (uint32_t)hi - (uint32_t)lo;  // width
  }
  return r;
}
__device__ __forceinline__ unsigned fast_flags(const fgx_predicate_t& p, const FastRange& fr, int64_t base, int64_t n, int l) {
  const uint32_t* __restrict__ q = reinterpret_cast<const uint32_t*>(p.term[0].data) + base + l;
  const bool full = base + 32 * SEL_ITEMS <= n;  // warp-uniform
  uint32_t raw[SEL_ITEMS];
  unsigned valid = 0;
#pragma unroll
  for (int r = 0; r < SEL_ITEMS; ++r) {
    const bool ok = full || base + r * 32 + l < n;
    raw[r] = ok ? __ldg(q + r * 32) : 0u;
    valid |= (ok ? 1u : 0u) << r;
  }
  unsigned m = 0;
  if (fr.is_float) {
    const float lo = __uint_as_float(fr.lo), hi = __uint_as_float(fr.hi);
#pragma unroll
    for (int r = 0; r < SEL_ITEMS; ++r) {
      const float b = __uint_as_float(raw[r]);
      m |= ((b >= lo && b <= hi) ? 1u : 0u) << r;
    }
  } else {
#pragma unroll
    for (int r = 0; r < SEL_ITEMS; ++r) m |= ((raw[r] - fr.lo <= fr.hi) ? 1u : 0u) << r;
  }
  if (fr.neg) m = ~m & 0xFFu;
  // 1-term truth table: bit 1 = value when the term holds, bit 0 = value when it does not; slots beyond n: 0
  return (((p.lut & 2u) ? m : 0u) | ((p.lut & 1u) ? (~m & 0xFFu) : 0u)) & valid;
}
template <bool FAST>
__device__ __forceinline__ unsigned group_flags(const fgx_predicate_t& p, const FastRange& fr, int64_t base, int64_t n, int l) {
  if constexpr (FAST) return fast_flags(p, fr, base, n, l);
  else return flags_at<SEL_ITEMS>(p, base, n, l);
}

// ---- one pass: cooperative launch, predicate evaluated once, ballots parked in shared memory ----------
constexpr int OP_GROUP = 32 * SEL_ITEMS;  // 256 slots: lane l holds rows r = 0..7 (slot g*256 + r*32 + l)
constexpr int OP_MAX_GPW = 32;            // groups per warp the ballots of which fit (8 KB of shared memory)

template <bool FAST>
__global__ void __launch_bounds__(SEL_BLOCK) select_onepass_kernel(const __grid_constant__ fgx_predicate_t p, int64_t n,
                                                                  int64_t groups, int gpw,
                                                                  int32_t* __restrict__ cta_count,
                                                                  int32_t* __restrict__ perm,
                                                                  int32_t* __restrict__ count_out) {
  const FastRange relmask = FAST ? fast_range(p.term[0]) : FastRange{};
  __shared__ unsigned balls[SEL_WARPS][OP_MAX_GPW][SEL_ITEMS];
  __shared__ int wcount[SEL_WARPS];
  __shared__ int warp_sums[SEL_WARPS];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int64_t g0 = ((int64_t)blockIdx.x * SEL_WARPS + w) * gpw;
  const int mine = (int)max((int64_t)0, min((int64_t)gpw, groups - g0));  // groups of this warp (warp-uniform)
  int c = 0;
  int k = 0;
  for (; k + 2 <= mine; k += 2) {  // two groups per trip: 16 independent loads per term in flight
    const unsigned f0 = group_flags<FAST>(p, relmask, (g0 + k) * OP_GROUP, n, l);
    const unsigned f1 = group_flags<FAST>(p, relmask, (g0 + k + 1) * OP_GROUP, n, l);
#pragma unroll
    for (int r = 0; r < SEL_ITEMS; ++r) {
      const unsigned b0 = __ballot_sync(0xffffffffu, (f0 >> r) & 1u), b1 = __ballot_sync(0xffffffffu, (f1 >> r) & 1u);
      if (l == r) { balls[w][k][r] = b0; balls[w][k + 1][r] = b1; }
      c += __popc(b0) + __popc(b1);
    }
  }
  if (k < mine) {
    const unsigned f0 = group_flags<FAST>(p, relmask, (g0 + k) * OP_GROUP, n, l);
#pragma unroll
    for (int r = 0; r < SEL_ITEMS; ++r) {
      const unsigned b0 = __ballot_sync(0xffffffffu, (f0 >> r) & 1u);
      if (l == r) balls[w][k][r] = b0;
      c += __popc(b0);
    }
  }
  if (l == 0) wcount[w] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int q = 0; q < SEL_WARPS; ++q) t += wcount[q];
    __stcg(cta_count + blockIdx.x, t);
  }
  cg::this_grid().sync();
  int before = 0, all = 0;
  for (int q = threadIdx.x; q < (int)gridDim.x; q += SEL_BLOCK) {
    const int v = __ldcg(cta_count + q);
    all += v;
    if (q < (int)blockIdx.x) before += v;
  }
  const int total = block_sum(all, warp_sums);
  int off = block_sum(before, warp_sums);
  if (blockIdx.x == 0 && threadIdx.x == 0) *count_out = total;
#pragma unroll
  for (int q = 0; q < SEL_WARPS; ++q)
    if (q < w) off += wcount[q];
  const unsigned lt = lanemask_lt();
  for (k = 0; k < mine; ++k) {
    const int64_t base = (g0 + k) * OP_GROUP;
    const bool full = base + OP_GROUP <= n;
    const int i0 = (int)base + l;  // n < 2^31
#pragma unroll
    for (int r = 0; r < SEL_ITEMS; ++r) {
      const unsigned b = balls[w][k][r];
      const int i = i0 + r * 32;
      if (full || i < n) {
        const int rank = off + __popc(b & lt);
        perm[(b >> l) & 1u ? rank : total + (i - rank)] = i;
      }
      off += __popc(b);
    }
  }
}

// small sets (<= 16384 slots): ONE CTA of 1024 threads holds every slot's flag in registers -- one round
// of loads, one block scan, one scatter (config C2's select calls are latency-bound, not bandwidth-bound)
constexpr int SEL_SMALL_BLOCK = 1024;
constexpr int64_t SEL_SMALL_MAX = 16384;

template <int ITEMS>
__global__ void __launch_bounds__(SEL_SMALL_BLOCK) select_small_kernel(const __grid_constant__ fgx_predicate_t p,
                                                                      int64_t n, int32_t* __restrict__ perm,
                                                                      int32_t* __restrict__ count_out) {
  __shared__ int warp_sums[32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int64_t start = (int64_t)w * (32 * ITEMS);
  const unsigned flags = flags_at<ITEMS>(p, start, n, l);
  unsigned ball[ITEMS];
  int c = 0;
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    ball[r] = __ballot_sync(0xffffffffu, (flags >> r) & 1u);
    c += __popc(ball[r]);
  }
  if (l == 0) warp_sums[w] = c;
  __syncthreads();
  const int mine = warp_sums[l];  // 32 warps: lane k holds warp k's count
  const int total = __reduce_add_sync(0xffffffffu, mine);
  int before = __reduce_add_sync(0xffffffffu, l < w ? mine : 0);
  const unsigned lt = lanemask_lt();
#pragma unroll
  for (int r = 0; r < ITEMS; ++r) {
    const int64_t i = start + r * 32 + l;
    if (i < n) {
      const int rank = before + __popc(ball[r] & lt);
      const bool s = (ball[r] >> l) & 1u;
      perm[s ? (int64_t)rank : (int64_t)total + (i - rank)] = (int32_t)i;
    }
    before += __popc(ball[r]);
  }
  if (threadIdx.x == 0) *count_out = total;
}

}  // namespace fgx

extern "C" {

size_t fgx_select_scratch_bytes(int64_t n) {
  (void)n;
  return (size_t)fgx::SEL_MAX_GRID * sizeof(int32_t);
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
  if (n <= SEL_SMALL_MAX) {
    const int64_t items = ceil_div(n, SEL_SMALL_BLOCK);
    if (items <= 1) select_small_kernel<1><<<1, SEL_SMALL_BLOCK, 0, s>>>(*pred_host, n, perm, count);
    else if (items <= 2) select_small_kernel<2><<<1, SEL_SMALL_BLOCK, 0, s>>>(*pred_host, n, perm, count);
    else if (items <= 4) select_small_kernel<4><<<1, SEL_SMALL_BLOCK, 0, s>>>(*pred_host, n, perm, count);
    else if (items <= 8) select_small_kernel<8><<<1, SEL_SMALL_BLOCK, 0, s>>>(*pred_host, n, perm, count);
    else select_small_kernel<16><<<1, SEL_SMALL_BLOCK, 0, s>>>(*pred_host, n, perm, count);
    return check_launch("fgx_select(small)");
  }
  FGX_REQUIRE(scratch && scratch_bytes >= fgx_select_scratch_bytes(n), "fgx_select: scratch too small");
  FGX_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 3) == 0, "fgx_select: scratch must be 4-byte aligned");
  int32_t* cta_count = reinterpret_cast<int32_t*>(scratch);
  {  // one cooperative pass whenever the ballots of a warp's groups fit in shared memory
    static const bool two_pass = [] { const char* v = getenv("FGX_SELECT_TWO_PASS"); return v && atoi(v) == 1; }();
    fgx_predicate_t pred = *pred_host;
    // one comparison on a contiguous, 4-byte-aligned 4-byte leaf: the compact fast-path kernel
    const bool fast = pred.n_terms == 1 && pred.term[0].stride == 1 && pred.term[0].dtype != FGX_U8 &&
                      (reinterpret_cast<uintptr_t>(pred.term[0].data) & 3) == 0;
    const void* kern = fast ? (const void*)select_onepass_kernel<true> : (const void*)select_onepass_kernel<false>;
    static int per_sm_of[2] = {-1, -1};
    int& per_sm = per_sm_of[fast ? 1 : 0];
    if (per_sm < 0) {
      int q = 0;
      cudaError_t oe = fast ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, select_onepass_kernel<true>, SEL_BLOCK, 0)
                            : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&q, select_onepass_kernel<false>, SEL_BLOCK, 0);
      per_sm = (oe == cudaSuccess && q >= 1) ? q : 0;
    }
    int64_t coop = (int64_t)sm_count() * per_sm;
    if (coop > SEL_MAX_GRID) coop = SEL_MAX_GRID;
    const int64_t groups = ceil_div(n, OP_GROUP);
    if (!two_pass && coop >= 1 && ceil_div(groups, coop * SEL_WARPS) <= OP_MAX_GPW) {
      int gpw = (int)ceil_div(groups, coop * SEL_WARPS);
      const int grid = (int)ceil_div(groups, (int64_t)gpw * SEL_WARPS);
      int64_t n_arg = n, groups_arg = groups;
      void* args[] = {&pred, &n_arg, &groups_arg, &gpw, &cta_count, &perm, &count};
      cudaError_t e = cudaLaunchCooperativeKernel(kern, dim3((unsigned)grid), dim3(SEL_BLOCK), args, 0, s);
      if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "fgx_select: cooperative launch: %s", cudaGetErrorString(e));
      return check_launch("fgx_select(one pass)");
    }
  }
  int64_t cap = (int64_t)sm_count() * 8;
  if (cap > SEL_MAX_GRID) cap = SEL_MAX_GRID;
  const int64_t per = ceil_div(tiles, tiles < cap ? tiles : cap);
  const int grid = (int)ceil_div(tiles, per);
  select_count_kernel<<<grid, SEL_BLOCK, 0, s>>>(*pred_host, n, tiles, per, cta_count);
  select_scatter_kernel<<<grid, SEL_BLOCK, 0, s>>>(*pred_host, n, tiles, per, cta_count, perm, count);
  return check_launch("fgx_select");
}
}
