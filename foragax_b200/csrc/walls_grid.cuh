// Shared pieces of the spatially culled wall ray cast (see raycast_walls_grid.cu for the derivation of the
// cull): the grid descriptor, the per-(ray, wall) test and the per-ray walk over a cell's wall list.  Used by
// raycast_walls_grid_kernel and by the fused sense + step kernel of the Forager (foraging.cu).
#pragma once
#include "common.cuh"
#include "geometry.cuh"

namespace fgx {

struct WGrid {
  float x0, y0, cs, inv_cs;
  int gx, gy;
  float thick, eps, scale;  // band half-widths (perpendicular) and the coordinate scale they assume
  float margin;             // thick - ray_length: how far beyond a ray's bounding box a wall can still matter
  float tol_line;           // per-(agent, wall) line band of the run-time pre-test (>= eps; see walls_pretest)
};

struct WState {
  uint32_t total;     // entries written by the fill pass
  int32_t fallback;   // != 0: lists unusable (overflow / out-of-scale wall): every agent scans all walls
};

// what the sensing side needs of a built grid
struct WallsGridView {
  WGrid g;
  const float *wx1, *wy1, *wx2, *wy2;
  int n_walls;
  const uint32_t* start;
  const int32_t* entries;
  const float4* wall4;  // (p2x, p2y, q2x, q2y) per wall, packed by the grid build
  const uint4* cellrec; // two uint4 per cell: {start, count, w0, w1}, {w2, w3, w4, w5}: the list range AND its first
                        // entries in one 32-byte record, so the common short list costs one dependent load less
  const float4* wallr;  // pre-test records, two float4 per wall: (min x, min y, max x, max y) of the wall and its
                        // unit-normal line (a, b, c, line band): a x + b y + c = signed distance to the wall's line
  const WState* st;
};

// line_intercept + distance of the proper-crossing candidate d1 (space_methods.py:98-103,124-128)
__device__ __forceinline__ float crossing_dist(float p1x, float p1y, float q1x, float q1y, float p2x, float p2y,
                                               float q2x, float q2y) {
  const float ax = f_sub(q1x, p1x), ay = f_sub(q1y, p1y);
  const float c1 = f_sub(f_mul(ax, f_sub(p2y, p1y)), f_mul(ay, f_sub(p2x, p1x)));
  const float c2 = f_sub(f_mul(ax, f_sub(q2y, p1y)), f_mul(ay, f_sub(q2x, p1x)));
  const float r = __fdiv_rn(c1, f_add(f_sub(c1, c2), 1e-6f));
  const float ix = f_add(p2x, f_mul(r, f_sub(q2x, p2x)));
  const float iy = f_add(p2y, f_mul(r, f_sub(q2y, p2y)));
  return norm2(f_sub(ix, p1x), f_sub(iy, p1y));
}

// ray_wall_sensor (space_methods.py:96-143) given the four orientation determinants the caller already holds, for
// the rare pairs with a ZERO determinant (the collinear cases c2..c5 each need an o == 0).
static __device__ __noinline__ float ray_wall_exact_o(float p1x, float p1y, float q1x, float q1y, float L, float p2x,
                                                      float p2y, float q2x, float q2y, float v1, float v2, float v3,
                                                      float v4) {
  const int o1 = orient_cls(v1), o2 = orient_cls(v2), o3 = orient_cls(v3), o4 = orient_cls(v4);
  float d1 = L;
  if (o1 != o2 && o3 != o4) d1 = crossing_dist(p1x, p1y, q1x, q1y, p2x, p2y, q2x, q2y);
  float d2 = L, d3 = L, d4 = L, d5 = L;
  if (o1 == 0 && on_segment(p1x, p1y, p2x, p2y, q1x, q1y)) d2 = norm2(f_sub(p1x, p2x), f_sub(p1y, p2y));
  if (o2 == 0 && on_segment(p1x, p1y, q2x, q2y, q1x, q1y)) d3 = norm2(f_sub(p1x, q2x), f_sub(p1y, q2y));
  if (o3 == 0 && on_segment(p2x, p2y, p1x, p1y, q2x, q2y)) d4 = 0.0f;
  if (o4 == 0 && on_segment(p2x, p2y, q1x, q1y, q2x, q2y)) d5 = norm2(f_sub(p1x, q1x), f_sub(p1y, q1y));
  return min_nan(d1, min_nan(d2, min_nan(d3, min_nan(d4, d5))));
}

// one (ray, wall) pair, exactly as the all-pairs kernel decides it: sure miss <=> no candidate distance of
// ray_wall_sensor can apply (max(o1*o2, o3*o4) > 0 and all four determinants nonzero); otherwise the op-for-op
// restatement.  (distance, wall) minima are lexicographic = argmin's first-index rule for any visiting order.
// When no determinant is zero (NaN counts as nonzero: class 2, space_methods.py:113) only the proper-crossing
// candidate d1 can apply, and it does iff the classes differ pairwise, i.e. iff (o1 > 0) != (o2 > 0) and
// (o3 > 0) != (o4 > 0); that common case runs inline.  Without a crossing the pair yields exactly L, which can
// never replace the running minimum: bidx >= 0 implies best < L (an update needs d < best <= L, or a tie with
// an earlier update), so the early return is exact.
__device__ __forceinline__ void test_wall(float p1x, float p1y, float q1x, float q1y, float L, float p2x, float p2y,
                                          float q2x, float q2y, int w, float& best, int& bidx) {
  const float o1 = orient_val(p1x, p1y, q1x, q1y, p2x, p2y), o2 = orient_val(p1x, p1y, q1x, q1y, q2x, q2y);
  const float o3 = orient_val(p2x, p2y, q2x, q2y, p1x, p1y), o4 = orient_val(p2x, p2y, q2x, q2y, q1x, q1y);
  const float A = f_mul(o1, o2), B = f_mul(o3, o4);
  if (fmaxf(A, B) > 0.0f && fabsf(f_mul(A, B)) > 0.0f) return;
  float d;
  if (o1 != 0.0f && o2 != 0.0f && o3 != 0.0f && o4 != 0.0f) {
    if (((o1 > 0.0f) == (o2 > 0.0f)) || ((o3 > 0.0f) == (o4 > 0.0f))) return;
    d = min_nan(crossing_dist(p1x, p1y, q1x, q1y, p2x, p2y, q2x, q2y), L);
  } else {
    d = ray_wall_exact_o(p1x, p1y, q1x, q1y, L, p2x, p2y, q2x, q2y, o1, o2, o3, o4);
  }
  if (d < best || (d == best && bidx >= 0 && w < bidx)) {
    best = d;
    bidx = w;
  }
}

// The ray's bounding box widened by g.margin (what the per-ray pre-test compares wall boxes with).
struct RayBox {
  float x0, y0, x1, y1;
};
__device__ __forceinline__ RayBox ray_box(float p1x, float p1y, float q1x, float q1y, float margin) {
  return RayBox{fminf(p1x, q1x) - margin, fminf(p1y, q1y) - margin, fmaxf(p1x, q1x) + margin, fmaxf(p1y, q1y) + margin};
}

// Run-time pre-test of a LISTED wall for one ray -- the list-building rule of raycast_walls_grid.cu evaluated per
// (ray, wall) instead of per (cell, wall).  A wall can give d < L for the ray p1 -> q1 only if
//   (a) a wall point lies in the ray's bounding box: c2 / c3 / c5 (space_methods.py:130-140) test exactly that with
//       on_segment, and a proper crossing (c1) of two segments lies in both boxes -- the boxes are compared with a
//       margin of 5e-4 * S (S = coordinate scale), hundreds of times the rounding of the determinants; or
//   (b) rounding can fake a crossing: c1 with a wrong sign class needs a determinant within the noise (~4e-7 * S in
//       distance units) of zero, i.e. the wall's line passes within that noise of a ray endpoint or the ray's line
//       of a wall endpoint, with the true crossing of the two LINES outside the widened box -- the lines are then
//       parallel to within noise / margin and p1 sits within (2 L / margin + 1) * noise of the wall's line;
//       tol_line is 8x that (and never below the list builder's eps); c4 (d = 0) needs o3 == 0 exactly, i.e. p1 ON
//       the line.
// So: skip <=> boxes disjoint AND p1 farther than the line band from the wall's line.  The distance comes from the
// wall's unit-normal line equation (two FMAs; its own rounding, a few ulp of S, is added to the band by the grid
// build), and only walls that survive load their end points.  NaN operands fail every comparison -> not skipped;
// a zero-length wall has the line (0, 0, 0) -> distance 0 -> never skipped by (b).
__device__ __forceinline__ bool walls_pretest_skip(const RayBox& rb, const float4& bb, const float4& ln, float px,
                                                   float py) {
  const bool disjoint = bb.z < rb.x0 || bb.x > rb.x1 || bb.w < rb.y0 || bb.y > rb.y1;
  return disjoint && fabsf(fmaf(ln.x, px, fmaf(ln.y, py, ln.z))) > ln.w;
}
__device__ __forceinline__ float4 wall_line_record(float ax, float ay, float bx, float by, float tol_line, float scale) {
  const float dx = bx - ax, dy = by - ay, len = sqrtf(dx * dx + dy * dy);
  const float inv = len > 0.0f ? 1.0f / len : 0.0f;
  const float na = -dy * inv, nb = dx * inv;  // unit normal; (0, 0) for a zero-length wall
  return make_float4(na, nb, -(na * ax + nb * ay), tol_line * 1.05f + 8e-7f * scale);
}

// min / first-argmin over the walls for the ray p1 -> q1 (best = L, bidx = -1 on entry); `fallback`: the value of
// v.st->fallback (read once per thread by the caller)
__device__ __forceinline__ void walls_grid_cast(const WallsGridView& v, bool fallback, float p1x, float p1y, float q1x,
                                                float q1y, float L, float& best, int& bidx) {
  // strictly inside the grid (NaN positions fail and take the all-walls loop)
  const float fx = (p1x - v.g.x0) * v.g.inv_cs, fy = (p1y - v.g.y0) * v.g.inv_cs;
  const bool inside = fx >= 0.0f && fy >= 0.0f && fx < (float)v.g.gx && fy < (float)v.g.gy;
  const RayBox rb = ray_box(p1x, p1y, q1x, q1y, v.g.margin);
  if (inside && !fallback) {
    const int cell = min((int)fy, v.g.gy - 1) * v.g.gx + min((int)fx, v.g.gx - 1);
    const uint4 c0 = __ldg(v.cellrec + 2 * (int64_t)cell), c1 = __ldg(v.cellrec + 2 * (int64_t)cell + 1);
    const uint32_t e0 = c0.x, n_list = c0.y;
    // two walls per trip: their pre-test records (32 KB for 1,000 walls) live in L1, so deeper batches only cost
    // registers; the first six indices came with the cell record, longer lists continue in entries[]
    for (uint32_t u0 = 0; u0 < n_list; u0 += 2) {
      int w[2];
      if (u0 == 0) { w[0] = (int)c0.z; w[1] = (int)c0.w; }
      else if (u0 == 2) { w[0] = (int)c1.x; w[1] = (int)c1.y; }
      else if (u0 == 4) { w[0] = (int)c1.z; w[1] = (int)c1.w; }
      else {
        w[0] = v.entries[e0 + u0];
        w[1] = v.entries[e0 + min(u0 + 1, n_list - 1)];
      }
      float4 bb[2], ln[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const float4* __restrict__ r = v.wallr + 2 * (int64_t)w[u];
        bb[u] = __ldg(r);
        ln[u] = __ldg(r + 1);
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        if (u0 + u >= n_list || walls_pretest_skip(rb, bb[u], ln[u], p1x, p1y)) continue;
        const float4 wl = __ldg(v.wall4 + w[u]);
        test_wall(p1x, p1y, q1x, q1y, L, wl.x, wl.y, wl.z, wl.w, w[u], best, bidx);
      }
    }
  } else {
    for (int w = 0; w < v.n_walls; ++w) {
      const float4 wl = make_float4(v.wx1[w], v.wy1[w], v.wx2[w], v.wy2[w]);
      const float4 bb = make_float4(fminf(wl.x, wl.z), fminf(wl.y, wl.w), fmaxf(wl.x, wl.z), fmaxf(wl.y, wl.w));
      if (walls_pretest_skip(rb, bb, wall_line_record(wl.x, wl.y, wl.z, wl.w, v.g.tol_line, v.g.scale), p1x, p1y)) continue;
      test_wall(p1x, p1y, q1x, q1y, L, wl.x, wl.y, wl.z, wl.w, w, best, bidx);
    }
  }
}
__device__ __forceinline__ void walls_grid_cast(const WallsGridView& v, float p1x, float p1y, float q1x, float q1y,
                                                float L, float& best, int& bidx) {
  walls_grid_cast(v, v.st->fallback != 0, p1x, p1y, q1x, q1y, L, best, bidx);
}

// ---- n_rays == 32, warp-cooperative: a warp owns an agent (lane = ray) and walks its cell's wall list together ----
// Used by raycast_walls_grid_warp32_kernel (raycast_walls_grid.cu, where the scheme is described) and by the
// ray-casting warps of the fused sense + step kernel (foraging.cu).
template <int TILE = 32>
struct WgwTileSmem {  // per warp: a parked tile of up to TILE agents
  float4 pose[TILE];      // x, y, ang, active
  uint32_t rec[TILE][8];  // the agent's cell record {start, count, w0..w5}
};
struct WgwSurvSmem {  // per warp: the survivors of one pre-test trip
  float4 bb[32];  // bounding box
  float4 wl[32];  // end points
  int2 w[32];     // {wall index, agent outside the wall's line band}
};
struct WgwTile {  // lane k's agent of a tile, in registers while its loads are in flight
  float x, y, ang, act;
  uint4 r0, r1;
};
constexpr uint32_t WGW_ALL_WALLS = 0xffffffffu;  // cell record count: take the all-walls loop (outside the grid / fallback)

__device__ __forceinline__ float warp_min_f32(float v) {
  float r;
  asm volatile("redux.sync.min.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float warp_max_f32(float v) {
  float r;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(r) : "f"(v));
  return r;
}

// linspace(lo, hi, 32)[lane] = lo * (1 - s) + hi * s, s = lane / 31; the last element is hi (linspace_elem)
__device__ __forceinline__ void wgw_lane_weights(uint32_t lane, float& sj, float& omsj) {
  sj = __fdiv_rn((float)lane, 31.0f);
  omsj = f_sub(1.0f, sj);
}
// MUTABLE: the pose arrays are written by other warps of the running kernel (the fused sense + step kernel updates
// positions in place, each agent after its own rays) -> coherent loads instead of the read-only path
template <bool MUTABLE = false>
__device__ __forceinline__ void wgw_load_pose(const float* ax, const float* ay, const float* aang, const float* aactive,
                                              uint32_t end, uint32_t agent, WgwTile& t) {
  t.act = 0.0f; t.x = 0.0f; t.y = 0.0f; t.ang = 0.0f;
  if (agent < end) {
    t.act = aactive != nullptr ? __ldg(aactive + agent) : 1.0f;
    if (MUTABLE) {
      t.x = __ldcg(ax + agent); t.y = __ldcg(ay + agent); t.ang = __ldcg(aang + agent);
    } else {
      t.x = __ldg(ax + agent); t.y = __ldg(ay + agent); t.ang = __ldg(aang + agent);
    }
  }
}
__device__ __forceinline__ void wgw_load_rec(const WallsGridView& v, bool fallback, WgwTile& t) {
  t.r0 = make_uint4(0u, 0u, 0u, 0u);
  t.r1 = make_uint4(0u, 0u, 0u, 0u);
  if (t.act == 0.0f) return;
  const float fx = (t.x - v.g.x0) * v.g.inv_cs, fy = (t.y - v.g.y0) * v.g.inv_cs;  // as walls_grid_cast
  const bool inside = fx >= 0.0f && fy >= 0.0f && fx < (float)v.g.gx && fy < (float)v.g.gy;
  if (inside && !fallback) {
    const int cell = min((int)fy, v.g.gy - 1) * v.g.gx + min((int)fx, v.g.gx - 1);
    t.r0 = __ldg(v.cellrec + 2 * (int64_t)cell);
    t.r1 = __ldg(v.cellrec + 2 * (int64_t)cell + 1);
  } else {
    t.r0.y = WGW_ALL_WALLS;
  }
}
template <class SM>
__device__ __forceinline__ void wgw_park(SM& sm, uint32_t lane, const WgwTile& t, uint32_t tile = 32u) {
  if (lane < tile) {
    sm.pose[lane] = make_float4(t.x, t.y, t.ang, t.act);
    *reinterpret_cast<uint4*>(&sm.rec[lane][0]) = t.r0;
    *reinterpret_cast<uint4*>(&sm.rec[lane][4]) = t.r1;
  }
}

// What an agent's cast needs from memory, requested ahead of its arithmetic: the pose and list header (shared
// memory), and lane u's listed wall of the first 32 -- its index, pre-test record and end points (they depend on the
// cell only).  Lanes beyond the list repeat its last entry, an empty list reads wall 0's records.
struct WgwPre {
  float4 pose;
  uint32_t e0, n_list;
  int w;
  float4 bb, ln, wl;
};
template <class SM>
__device__ __forceinline__ void wgw_prefetch(const WallsGridView& v, const SM& sm, uint32_t k, uint32_t lane, WgwPre& p) {
  p.pose = sm.pose[k];
  const uint2 hdr = *reinterpret_cast<const uint2*>(&sm.rec[k][0]);
  p.e0 = hdr.x;
  p.n_list = hdr.y;
  if (p.pose.w == 0.0f) return;
  int w = (int)sm.rec[k][2u + min(lane, 5u)];
  if (p.n_list > 6u && p.n_list != WGW_ALL_WALLS) w = __ldg(v.entries + p.e0 + min(lane, p.n_list - 1u));
  const float4* __restrict__ r = v.wallr + 2 * (int64_t)w;
  p.w = w;
  p.bb = __ldg(r);
  p.ln = __ldg(r + 1);
  p.wl = __ldg(v.wall4 + w);
}

// ray `lane` of the prefetched agent: min / first-argmin over the walls (best = 0, bidx = -1 for an inactive
// agent: sim.py:459-460).  Must be called by all 32 lanes.
__device__ __forceinline__ void wgw_cast_agent(const WallsGridView& v, WgwSurvSmem& sm, uint32_t lane, float L, float span,
                                               float sj, float omsj, const WgwPre& p, float& best, int& bidx) {
  bidx = -1;
  if (p.pose.w == 0.0f) {
    best = 0.0f;
    return;
  }
  best = L;
  const float p1x = p.pose.x, p1y = p.pose.y;
  const uint32_t n_list = p.n_list;
  // ray_generator (space_methods.py:56-69)
  const float lo = f_sub(p.pose.z, span), hi = f_add(p.pose.z, span);
  const float th = lane == 31u ? hi : f_add(f_mul(lo, omsj), f_mul(hi, sj));
  float sn, cs;
  sincosf(th, &sn, &cs);
  const float q1x = f_add(p1x, f_mul(cs, L)), q1y = f_add(p1y, f_mul(sn, L));
  if (n_list == WGW_ALL_WALLS) {
    walls_grid_cast(v, true, p1x, p1y, q1x, q1y, L, best, bidx);
  } else if (n_list != 0u) {
    const RayBox rb = ray_box(p1x, p1y, q1x, q1y, v.g.margin);
    // union of the 32 ray boxes (x - m is monotone in x, so the reduction commutes with the widening)
    const RayBox fb = RayBox{fminf(p1x, warp_min_f32(q1x)) - v.g.margin, fminf(p1y, warp_min_f32(q1y)) - v.g.margin,
                             fmaxf(p1x, warp_max_f32(q1x)) + v.g.margin, fmaxf(p1y, warp_max_f32(q1y)) + v.g.margin};
    int w = p.w;
    float4 bb = p.bb, ln = p.ln, wl = p.wl;
    for (uint32_t u0 = 0;;) {
      const bool have = u0 + lane < n_list;
      const bool disjoint = bb.z < fb.x0 || bb.x > fb.x1 || bb.w < fb.y0 || bb.y > fb.y1;
      const bool far = fabsf(fmaf(ln.x, p1x, fmaf(ln.y, p1y, ln.z))) > ln.w;
      uint32_t m = __ballot_sync(0xffffffffu, have && !(disjoint && far));
      if (m) {
        sm.bb[lane] = bb;
        sm.wl[lane] = wl;
        sm.w[lane] = make_int2(w, far ? 1 : 0);
        __syncwarp();
        do {
          const uint32_t src = 31u - (uint32_t)__clz(m);
          m ^= 1u << src;
          const float4 sb = sm.bb[src];
          const int2 sw = sm.w[src];
          const bool dj = sb.z < rb.x0 || sb.x > rb.x1 || sb.w < rb.y0 || sb.y > rb.y1;
          if (dj && sw.y) continue;
          const float4 swl = sm.wl[src];
          test_wall(p1x, p1y, q1x, q1y, L, swl.x, swl.y, swl.z, swl.w, sw.x, best, bidx);
        } while (m);
        __syncwarp();
      }
      u0 += 32u;
      if (u0 >= n_list) break;
      w = __ldg(v.entries + p.e0 + min(u0 + lane, n_list - 1u));
      const float4* __restrict__ r = v.wallr + 2 * (int64_t)w;
      bb = __ldg(r);
      ln = __ldg(r + 1);
      wl = __ldg(v.wall4 + w);
    }
  }
}

// host side (raycast_walls_grid.cu): validates, carves the scratch and -- unless reuse_grid -- builds the wall
// lists on `stream`; fills `view`.  Returns an fgx status code.
size_t wall_grid_scratch_bytes(int64_t n_walls, float x_min, float x_max, float y_min, float y_max, float ray_length);
int wall_grid_prepare(const char* what, const float* wx1, const float* wy1, const float* wx2, const float* wy2,
                      int64_t n_walls, float x_min, float x_max, float y_min, float y_max, float ray_length,
                      void* scratch, size_t scratch_bytes, int reuse_grid, cudaStream_t stream, WallsGridView* view);

}  // namespace fgx
