// K1 (part 2): ray casting against nearby agents -- agents x rays x circle targets.
//
// Reference: ray_agent_sensor (space_methods.py:71-94) evaluated for every (ray, target) and
// reduced per ray with min / first-argmin (EcoEvoJax/sim.py:444-447).  The reference is all-pairs
// (O(N^2 R)); here a uniform grid with cell = (ray_length + max_rad) / 3 bounds the search to the cells of
// the 7x7 neighbourhood that can hold a target in reach and inside the (widened) ray fan.  Every cull is
// conservative, not approximate: a target farther than ray_length + rad from the ray origin, or farther
// than rad from the cone spanned by a group of rays, cannot produce t < ray_length (the rounding error
// of t = -b - sqrt(b*b - c) is ~1e-6 relative, the culls carry 1e-4 relative margins), so the result
// equals the all-pairs evaluation bit for bit; ties are resolved by the lexicographic (distance, target
// index) minimum, i.e. argmin's first-index rule, independent of the grid order.
//
//   1. counting sort of the targets by cell (histogram with atomics, scan, scatter) and of the
//      sensing agents by cell;
//   2. one warp per sensing agent, walking the cell-sorted agent list (the warps of a CTA sense
//      neighbouring agents and share their targets through L1/L2; no shared staging, no barrier).
//      Per neighbourhood row, lanes = cells: keep the cells within reach and inside the fan; their
//      targets are contiguous.  Broad phase, lanes = targets: s = o - c, c = |s|^2 - r^2, reject
//      inactive / out-of-reach targets, then test the circle against the cone of each of the 4 ray
//      groups (lanes 8g..8g+7 own the rays of group g) and append it to the shared-memory list of every
//      group it can touch (ballot + popc compaction).  Narrow phase, lanes = rays: each group walks its
//      own list, two entries per trip; a cheap necessary condition (b <= 0, h >= (-b - best - slack)^2)
//      guards the exactly rounded sqrt, and rows are visited nearest first so the minima settle early;
//   3. optional merge with the wall result of fgx_raycast_walls in entity order [agents, walls].
#include "common.cuh"
#include "geometry.cuh"

#include "scan.cuh"

namespace fgx {

constexpr int RA_BLOCK = 256;
constexpr int RA_WARPS = RA_BLOCK / 32;  // agents per CTA (one warp each)
constexpr int RA_G = 4;                  // ray groups per warp
constexpr int RA_GW = 32 / RA_G;         // lanes per group
constexpr int RA_MAXRB = 4;              // rays per lane (n_rays <= 128)
constexpr int RA_CAP = 64;               // survivor list entries per (warp, group)
constexpr int RA_MINB = 4;                // CTAs per SM the register budget is held to
constexpr int RA_SUB = 3;                // at most this many cells per (ray_length + max_rad); search radius K cells

struct GridDesc {
  float x0, y0, inv_cs;
  int gx, gy;
  int K;  // neighbourhood radius in cells: K * cell >= ray_length + max_rad
};

__device__ __forceinline__ int cell_coord(float v, float v0, float inv_cs, int g) {
  const float f = floorf((v - v0) * inv_cs);
  // clamp (monotone): |a - b| <= cell size still implies |cell(a) - cell(b)| <= 1; NaN -> 0
  return f >= (float)(g - 1) ? g - 1 : (f > 0.0f ? (int)f : 0);
}
__device__ __forceinline__ int cell_of(const GridDesc& g, float x, float y) {
  return cell_coord(y, g.y0, g.inv_cs, g.gy) * g.gx + cell_coord(x, g.x0, g.inv_cs, g.gx);
}

// histogram of points per cell; remembers each point's cell.  Points whose `active` flag is 0 (targets only:
// an inactive target can never be hit, space_methods.py:92) get cell -1 and stay out of the grid.
__global__ void __launch_bounds__(256) grid_count_kernel(GridDesc g, const float* __restrict__ x,
                                                         const float* __restrict__ y,
                                                         const float* __restrict__ active, int64_t stride, int64_t n,
                                                         int32_t* __restrict__ cell, uint32_t* __restrict__ count) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (active != nullptr && active[i * stride] == 0.0f) {
      cell[i] = -1;
      continue;
    }
    const int c = cell_of(g, x[i * stride], y[i * stride]);
    cell[i] = c;
    atomicAdd(count + c, 1u);
  }
}

__global__ void __launch_bounds__(256) grid_scatter_targets_kernel(const float* __restrict__ tx, const float* __restrict__ ty,
                                                                   const float* __restrict__ tr, const float* __restrict__ ta,
                                                                   int64_t stride, int64_t m, const int32_t* __restrict__ cell,
                                                                   const uint32_t* __restrict__ tstart,
                                                                   uint32_t* __restrict__ cursor, float4* __restrict__ sorted,
                                                                   int32_t* __restrict__ sorted_idx) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < m; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = cell[i];
    if (c < 0) continue;  // inactive target
    const uint32_t pos = tstart[c] + atomicAdd(cursor + c, 1u);
    sorted[pos] = make_float4(tx[i * stride], ty[i * stride], tr[i * stride], ta[i * stride]);
    sorted_idx[pos] = (int32_t)i;
  }
}

// sensing agents in cell order: their index and a (x, y, heading, active | cell) record, so that the sensing
// kernel reads everything it needs about agent number i of the order with independent, coalesced loads
struct SensorRec {
  float x, y, ang;
  int32_t cell;  // -1: inactive sensor
};

__global__ void __launch_bounds__(256)
    grid_scatter_agents_kernel(int64_t n, const float* __restrict__ x, const float* __restrict__ y,
                               const float* __restrict__ ang, const float* __restrict__ active,
                               const int32_t* __restrict__ cell, const uint32_t* __restrict__ astart,
                               uint32_t* __restrict__ cursor, int32_t* __restrict__ order, SensorRec* __restrict__ rec) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = cell[i];
    const uint32_t pos = astart[c] + atomicAdd(cursor + c, 1u);
    order[pos] = (int32_t)i;
    rec[pos] = SensorRec{x[i], y[i], ang[i], (active == nullptr || active[i] != 0.0f) ? c : -1};
  }
}

struct RaycastAgentsArgs {
  GridDesc g;
  int64_t n, m;
  int n_rays;
  float span, L, max_rad;
  const uint32_t* tstart;
  const int32_t* order;
  const SensorRec* rec;  // sensing agents in cell order
  const float4* targets;
  const int32_t* tidx;
  int merge_walls;
  float* dist;
  int32_t* hit;
};

// rays <-> lanes: the warp's lanes form RA_G groups of RA_GW lanes; group g owns the contiguous rays
// [g*RG, (g+1)*RG), RG = ceil(n_rays / RA_G), lane w of the group the rays g*RG + w + RA_GW*q
__device__ __forceinline__ int ray_of(int g, int w, int q, int RG) { return g * RG + q * RA_GW + w; }

// max over the rectangle [x0,x1] x [y0,y1] of the linear form cx*x + cy*y
__device__ __forceinline__ float rect_max(float cx, float cy, float x0, float x1, float y0, float y1) {
  return cx * (cx > 0.0f ? x1 : x0) + cy * (cy > 0.0f ? y1 : y0);
}

template <int RPL>
__global__ void __launch_bounds__(RA_BLOCK, RA_MINB) raycast_agents_kernel(const RaycastAgentsArgs a) {
  // per warp and ray group: survivors of the broad phase as (sx, sy, c, target index); the odd row
  // stride keeps the four groups' concurrent 16-byte reads on different banks
  __shared__ float4 lists[RA_WARPS][RA_G][RA_CAP + 3];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane / RA_GW, w = lane % RA_GW;
  const unsigned lt = lanemask_lt();
  float4(*mine)[RA_CAP + 3] = lists[warp];
  const int RG = (a.n_rays + RA_G - 1) / RA_G;
  const float L = a.L;
  // far-target bound: a hit needs |s| <= L + r, i.e. c = |s|^2 - r^2 <= L^2 + 2 L r
  const float reach = (L + a.max_rad) * 1.0001f;
  const bool fan_convex = a.span < 1.5f;  // the whole fan is a convex cone: cells can be culled against it
  const float cs = 1.0f / a.g.inv_cs, cell_margin = 1e-3f * cs + 1e-3f;
  const int K = a.g.K;
  const int64_t n_warps = (int64_t)gridDim.x * RA_WARPS;
  int64_t i = blockIdx.x * (int64_t)RA_WARPS + warp;
  int next_agent = 0;
  int4 next_rec = make_int4(0, 0, 0, -1);
  if (i < a.n) {
    next_agent = a.order[i];
    next_rec = *reinterpret_cast<const int4*>(a.rec + i);
  }
  for (; i < a.n; i += n_warps) {
    const int64_t agent = next_agent;
    const int4 rec = next_rec;
    if (i + n_warps < a.n) {  // the next agent's record is in flight while this one is processed
      next_agent = a.order[i + n_warps];
      next_rec = *reinterpret_cast<const int4*>(a.rec + i + n_warps);
    }
    const int cell = rec.w;
    if (cell < 0) {  // inactive sensor: zeros, as the reference's cond
      for (int r = lane; r < a.n_rays; r += 32) {
        a.dist[agent * a.n_rays + r] = 0.0f;
        a.hit[agent * a.n_rays + r] = -1;
      }
      continue;
    }
    const float ox = __int_as_float(rec.x), oy = __int_as_float(rec.y), ang = __int_as_float(rec.z);
    float dx[RPL], dy[RPL], best[RPL], kb[RPL];
    int bidx[RPL];
    {
      const float lo_a = f_sub(ang, a.span), hi_a = f_add(ang, a.span);
#pragma unroll
      for (int q = 0; q < RPL; ++q) {
        const int r = ray_of(g, w, q, RG);
        const bool ok = q * RA_GW + w < RG && r < a.n_rays;
        sincosf(linspace_elem(lo_a, hi_a, a.n_rays, ok ? r : 0), &dy[q], &dx[q]);
        best[q] = L;
        kb[q] = L * 1.00001f;
        bidx[q] = -1;
      }
    }
    // group cones: group k's rays lie between bd[k] (its first ray) and bd[k+1] (the next group's first ray,
    // or the fan's last ray); the directions come from the lanes that own those rays
    float bdx[RA_G + 1], bdy[RA_G + 1];
    bool gok[RA_G];
    {
      const int g_last = (a.n_rays - 1) / RG, li = a.n_rays - 1 - g_last * RG, qb = li / RA_GW;
      float ex = dx[0], ey = dy[0];
#pragma unroll
      for (int q = 1; q < RPL; ++q)
        if (q == qb) { ex = dx[q]; ey = dy[q]; }
      bdx[RA_G] = __shfl_sync(0xffffffffu, ex, g_last * RA_GW + li % RA_GW);
      bdy[RA_G] = __shfl_sync(0xffffffffu, ey, g_last * RA_GW + li % RA_GW);
#pragma unroll
      for (int k = 0; k < RA_G; ++k) {
        gok[k] = k * RG < a.n_rays;
        const float fx = __shfl_sync(0xffffffffu, dx[0], k * RA_GW), fy = __shfl_sync(0xffffffffu, dy[0], k * RA_GW);
        bdx[k] = gok[k] ? fx : bdx[RA_G];
        bdy[k] = gok[k] ? fy : bdy[RA_G];
      }
    }
    const float fax = bdx[0], fay = bdy[0], fbx = bdx[RA_G], fby = bdy[RA_G], fmx = fax + fbx, fmy = fay + fby;

    int cnt[RA_G];
    float gm[RA_G];  // largest current minimum among a group's rays: a circle whose nearest point is farther cannot improve any
#pragma unroll
    for (int k = 0; k < RA_G; ++k) {
      cnt[k] = 0;
      gm[k] = L;
    }

    // narrow phase, lanes = rays: every group walks its own survivor list, two entries per trip
    auto drain = [&]() {
      __syncwarp();
      int n_mine = cnt[0], n_max = cnt[0];
#pragma unroll
      for (int k = 1; k < RA_G; ++k) {
        if (g == k) n_mine = cnt[k];
        n_max = max(n_max, cnt[k]);
      }
      for (int e = 0; e < n_max; e += 2) {
        const float4 t0 = mine[g][e], t1 = mine[g][e + 1];  // stale slots are masked by in0 / in1
        const bool in0 = e < n_mine, in1 = e + 1 < n_mine;
#pragma unroll
        for (int q = 0; q < RPL; ++q) {
          const float b0 = f_add(f_mul(t0.x, dx[q]), f_mul(t0.y, dy[q]));
          const float b1 = f_add(f_mul(t1.x, dx[q]), f_mul(t1.y, dy[q]));
          const float h0 = f_sub(f_mul(b0, b0), t0.z), h1 = f_sub(f_mul(b1, b1), t1.z);
          // cheap necessary condition for "tt >= 0 and tt <= best": b <= 0 and sqrt(h) >= -b - best, with a
          // 1e-5 relative slack that dwarfs the rounding of the exact evaluation below (kb = best * 1.00001)
          const float v0 = fmaf(b0, -0.99999f, -kb[q]), v1 = fmaf(b1, -0.99999f, -kb[q]);
          const bool c0 = in0 & (h0 >= 0.0f) & (b0 <= 0.0f) & ((v0 <= 0.0f) | (h0 >= v0 * v0 * 0.99999f));
          const bool c1 = in1 & (h1 >= 0.0f) & (b1 <= 0.0f) & ((v1 <= 0.0f) | (h1 >= v1 * v1 * 0.99999f));
          if (c0 | c1) {
            if (c0) {
              const float tt = f_sub(-b0, __fsqrt_rn(h0));
              const int bid = __float_as_int(t0.w);
              if (tt >= 0.0f && (tt < best[q] || (tt == best[q] && bid < bidx[q]))) {
                best[q] = tt;
                bidx[q] = bid;
              }
            }
            if (c1) {
              const float tt = f_sub(-b1, __fsqrt_rn(h1));
              const int bid = __float_as_int(t1.w);
              if (tt >= 0.0f && (tt < best[q] || (tt == best[q] && bid < bidx[q]))) {
                best[q] = tt;
                bidx[q] = bid;
              }
            }
            kb[q] = best[q] * 1.00001f;
          }
        }
      }
      float mb = 0.0f;  // lanes that own no ray contribute nothing
#pragma unroll
      for (int q = 0; q < RPL; ++q)
        if (q * RA_GW + w < RG && ray_of(g, w, q, RG) < a.n_rays) mb = fmaxf(mb, best[q]);
#pragma unroll
      for (int d = 1; d < RA_GW; d <<= 1) mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, d));
#pragma unroll
      for (int k = 0; k < RA_G; ++k) {
        gm[k] = __shfl_sync(0xffffffffu, mb, k * RA_GW);
        cnt[k] = 0;
      }
      __syncwarp();
    };

    const int cx = cell % a.g.gx, cy = cell / a.g.gx;
    // Cells are visited in rings around the sensor's cell (ring r = Chebyshev distance r): the minima settle
    // on the nearest targets first, after every ring the groups' largest minima are refreshed (drain), and
    // the reach that the farther cells, targets and rings are culled against shrinks accordingly.
    for (int ring = 0; ring <= K; ++ring) {
      float gmax = gm[0];
#pragma unroll
      for (int k = 1; k < RA_G; ++k) gmax = fmaxf(gmax, gm[k]);
      const float rdyn = fminf(reach, (gmax + a.max_rad) * 1.0001f + 1e-3f);
      // every cell of ring r >= 2 is at least (r - 1) cells away from any point of the sensor's cell
      if (ring >= 2 && (ring - 1) * cs - cell_margin > rdyn) break;
      for (int dy = -ring; dy <= ring; ++dy) {
        const int ry = cy + dy;
        if (ry < 0 || ry >= a.g.gy) continue;
        const bool full_row = dy == -ring || dy == ring;  // top / bottom row of the ring, else its two side cells
        // a 3 x 3 neighbourhood (K == 1) takes the sensor's whole row at once: three segments instead of five
        if (K == 1 && ring == 1 && !full_row) continue;
        const int wide0 = (K == 1 && ring == 0) ? 1 : 0;
        for (int part = 0; part < (full_row ? 1 : 2); ++part) {
          const int c_first = (full_row || part == 0) ? cx - ring - wide0 : cx + ring;
          const int c_last = full_row ? cx + ring + wide0 : c_first;
          // cells that can hold a target in reach and (convex fan) inside the widened fan; lane <-> cell
          const int ccx = c_first + lane;
          bool keep = ccx <= c_last && ccx >= 0 && ccx < a.g.gx;
          if (keep) {
            // rectangle of target centres in the cell, relative to the sensor; border cells are unbounded outwards
            const float x0 = ccx == 0 ? -1e30f : (a.g.x0 + ccx * cs - cell_margin) - ox;
            const float x1 = ccx == a.g.gx - 1 ? 1e30f : (a.g.x0 + (ccx + 1) * cs + cell_margin) - ox;
            const float y0 = ry == 0 ? -1e30f : (a.g.y0 + ry * cs - cell_margin) - oy;
            const float y1 = ry == a.g.gy - 1 ? 1e30f : (a.g.y0 + (ry + 1) * cs + cell_margin) - oy;
            const float ddx = fmaxf(fmaxf(x0, -x1), 0.0f), ddy = fmaxf(fmaxf(y0, -y1), 0.0f);
            keep = ddx * ddx + ddy * ddy <= rdyn * rdyn;
            if (fan_convex) {
              const float R = a.max_rad * 1.0001f + 1e-3f;
              keep = keep && rect_max(-fay, fax, x0, x1, y0, y1) >= -R    // cross(d_first, p) >= -R
                          && rect_max(fby, -fbx, x0, x1, y0, y1) >= -R    // cross(p, d_last) >= -R
                          && rect_max(fmx, fmy, x0, x1, y0, y1) >= -2.0f * R;  // not behind the sensor
            }
          }
          const unsigned km = __ballot_sync(0xffffffffu, keep);
          if (km == 0u) continue;
          const int c_lo = ry * a.g.gx + c_first + (__ffs(km) - 1), c_hi = ry * a.g.gx + c_first + (31 - __clz(km));
          const uint32_t r0 = a.tstart[c_lo], r1 = a.tstart[c_hi + 1];  // the row's cells are contiguous
          for (uint32_t base = r0; base < r1; base += 32) {
            // broad phase: lane <-> target
            const uint32_t k = base + lane;
            const bool in = k < r1;
            const float4 t = a.targets[in ? k : r1 - 1];
            const int id = a.tidx[in ? k : r1 - 1];
            const float sx = f_sub(ox, t.x), sy = f_sub(oy, t.y);
            const float c = f_sub(f_add(f_mul(sx, sx), f_mul(sy, sy)), f_mul(t.z, t.z));
            // reach: a hit at distance <= m needs |s| <= m + r, i.e. c = |s|^2 - r^2 <= m (m + 2 r); m = the group's
            // largest current minimum (ray_length at first), tested per group below
            const bool pass = in && t.w != 0.0f;
            const float two_r = t.z + t.z;
            // a circle farther than r from the cone spanned by a group's rays cannot be hit by any of them.
            // v = target - sensor = -s:  cross(bd, v) = bdy*sx - bdx*sy,  cross(v, bd) = -cross(bd, v)
            const float slack = 1e-4f * (fabsf(sx) + fabsf(sy)) + 1e-6f;  // >> the rounding of t = -b - sqrt(b*b - c)
            const float tol = t.z * 1.0001f + slack;
            float cr[RA_G + 1];
#pragma unroll
            for (int k = 0; k <= RA_G; ++k) cr[k] = bdy[k] * sx - bdx[k] * sy;
            const bool front_fan = -(fmx * sx + fmy * sy) >= -2.0f * tol;  // not behind the sensor (convex fan)
            const float4 entry = make_float4(sx, sy, c, __int_as_float(id));
            bool full = false;
#pragma unroll
            for (int gg = 0; gg < RA_G; ++gg) {
              bool pg = pass && gok[gg] && cr[gg] >= -tol && cr[gg + 1] <= tol &&
                        !(c > (gm[gg] + slack) * (gm[gg] + slack + two_r) * 1.0001f);
              if (fan_convex) pg = pg && front_fan;
              else pg = pg && -((bdx[gg] + bdx[gg + 1]) * sx + (bdy[gg] + bdy[gg + 1]) * sy) >= -2.0f * tol;
              const unsigned m = __ballot_sync(0xffffffffu, pg);
              if (pg) mine[gg][cnt[gg] + __popc(m & lt)] = entry;
              cnt[gg] += __popc(m);
              full = full || cnt[gg] > RA_CAP - 32;
            }
            if (full) drain();
          }
        }
      }
      drain();  // per ring: refreshes the groups' minima before the farther cells are culled against them
    }

#pragma unroll
    for (int q = 0; q < RPL; ++q) {
      const int r = ray_of(g, w, q, RG);
      if (!(q * RA_GW + w < RG && r < a.n_rays)) continue;
      const int64_t o = agent * a.n_rays + r;
      float d = best[q];
      int h = bidx[q];
      if (a.merge_walls) {
        // entity order [agents, walls]: an agent wins a tie; wall index w -> m + w
        const float wd = a.dist[o];
        const int wh = a.hit[o];
        if (!(bidx[q] >= 0 && best[q] <= wd)) {
          d = wd;
          h = wh >= 0 ? (int)(a.m + wh) : -1;
        }
      }
      a.dist[o] = d;
      a.hit[o] = h;
    }
  }
}

static inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

struct RaScratch {
  uint32_t *tcount, *acount, *tstart, *astart, *partial;
  int32_t *tcell, *acell, *order, *tidx;
  SensorRec* rec;
  float4* sorted;
};
static size_t carve(char* base, int64_t n, int64_t m, int n_cells, RaScratch* s) {
  size_t o = 0;
  auto take = [&](size_t bytes) { char* p = base ? base + o : nullptr; o += al256(bytes); return p; };
  char* p;
  p = take((size_t)n_cells * 4); if (s) s->tcount = (uint32_t*)p;
  p = take((size_t)n_cells * 4); if (s) s->acount = (uint32_t*)p;
  p = take((size_t)(n_cells + 1) * 4); if (s) s->tstart = (uint32_t*)p;
  p = take((size_t)(n_cells + 1) * 4); if (s) s->astart = (uint32_t*)p;
  p = take((size_t)m * 4); if (s) s->tcell = (int32_t*)p;
  p = take((size_t)n * 4); if (s) s->acell = (int32_t*)p;
  p = take((size_t)n * 4); if (s) s->order = (int32_t*)p;
  p = take((size_t)m * 4); if (s) s->tidx = (int32_t*)p;
  p = take((size_t)m * 16); if (s) s->sorted = (float4*)p;
  p = take((size_t)n * 16); if (s) s->rec = (SensorRec*)p;
  p = take(coop_scan_scratch_bytes(2)); if (s) s->partial = (uint32_t*)p;
  return o;
}

static int grid_dims(float x_min, float x_max, float y_min, float y_max, float ray_length, float max_rad,
                     int64_t n_targets, GridDesc* g) {
  const float reach = (ray_length + max_rad) * 1.0001f;  // a target farther than this (per axis) cannot be hit
  if (!(reach > 0.0f) || !(x_max > x_min) || !(y_max > y_min)) return -1;
  // up to RA_SUB cells per reach: the fan and range culls then skip most of the (2K+1)^2 neighbourhood -- as
  // long as a cell still holds ~16 targets (a row of cells is one 32-lane trip of the broad phase)
  const float per_reach_cell = (float)n_targets / ((x_max - x_min) * (y_max - y_min)) * reach * reach;
  int sub = (int)sqrtf(per_reach_cell / 16.0f);
  sub = sub < 1 ? 1 : (sub > RA_SUB ? RA_SUB : sub);
  float c = reach / sub;
  const float ext = fmaxf(x_max - x_min, y_max - y_min);
  if (ext / c > 2048.0f) c = ext / 2048.0f;  // at most 2048 cells per axis
  g->x0 = x_min; g->y0 = y_min; g->inv_cs = 1.0f / c;
  g->gx = (int)ceilf((x_max - x_min) / c); g->gy = (int)ceilf((y_max - y_min) / c);
  if (g->gx < 1) g->gx = 1;
  if (g->gy < 1) g->gy = 1;
  // |cell(a) - cell(b)| <= K whenever |a - b| <= reach (monotone clamp included; reach carries the margin)
  g->K = (int)ceilf(reach / c * 0.99999f);
  if (g->K < 1) g->K = 1;
  return g->K <= 15 ? 0 : -1;  // one lane per cell of a neighbourhood row
}

}  // namespace fgx

extern "C" {

size_t fgx_raycast_agents_scratch_bytes(int64_t n_agents, int64_t n_targets, float x_min, float x_max, float y_min,
                                        float y_max, float ray_length, float max_rad) {
  fgx::GridDesc g;
  if (fgx::grid_dims(x_min, x_max, y_min, y_max, ray_length, max_rad, n_targets, &g)) return 0;
  return fgx::carve(nullptr, n_agents < 0 ? 0 : n_agents, n_targets < 0 ? 0 : n_targets, g.gx * g.gy, nullptr);
}

int fgx_raycast_agents(const float* x, const float* y, const float* ang, const float* active, int64_t n_agents,
                       int32_t n_rays, float ray_span, float ray_length, const float* tx, const float* ty,
                       const float* trad, const float* tactive, int64_t target_stride, int64_t n_targets,
                       float x_min, float x_max, float y_min, float y_max, float max_rad, int32_t merge_walls,
                       float* dist, int32_t* hit, void* scratch, size_t scratch_bytes, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n_agents >= 0 && n_targets >= 0 && n_agents < (1ll << 31) && n_targets < (1ll << 31),
              "fgx_raycast_agents: size out of range");
  FGX_REQUIRE(n_rays > 0 && n_rays <= 32 * RA_MAXRB && target_stride >= 1,
              "fgx_raycast_agents: n_rays must be 1..%d, stride >= 1", 32 * RA_MAXRB);
  if (n_agents == 0) return FGX_OK;
  FGX_REQUIRE(x && y && ang && dist && hit, "fgx_raycast_agents: null pointer");
  FGX_REQUIRE(n_targets == 0 || (tx && ty && trad && tactive), "fgx_raycast_agents: null target array");
  GridDesc g;
  FGX_REQUIRE(grid_dims(x_min, x_max, y_min, y_max, ray_length, max_rad, n_targets, &g) == 0,
              "fgx_raycast_agents: bad arena bounds / ray_length / max_rad");
  const int n_cells = g.gx * g.gy;
  const size_t need = carve(nullptr, n_agents, n_targets, n_cells, nullptr);
  FGX_REQUIRE(scratch && scratch_bytes >= need, "fgx_raycast_agents: scratch too small");
  FGX_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "fgx_raycast_agents: scratch must be 256-byte aligned");
  RaScratch s;
  carve(reinterpret_cast<char*>(scratch), n_agents, n_targets, n_cells, &s);
  cudaStream_t st = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(s.tcount, 0, al256((size_t)n_cells * 4) * 2, st);  // tcount + acount
  if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "fgx_raycast_agents: %s", cudaGetErrorString(e));
  if (n_targets) grid_count_kernel<<<grid_for(n_targets, 256, 8), 256, 0, st>>>(g, tx, ty, tactive, target_stride, n_targets, s.tcell, s.tcount);
  grid_count_kernel<<<grid_for(n_agents, 256, 8), 256, 0, st>>>(g, x, y, nullptr, 1, n_agents, s.acell, s.acount);
  {  // exclusive scans of both histograms in one cooperative launch (counts are reset: they become the cursors)
    const int rc = launch_coop_scan<2>("fgx_raycast_agents", n_cells, ScanJob{s.tcount, s.tstart}, ScanJob{s.acount, s.astart},
                                       s.partial, nullptr, 0u, nullptr, st);
    if (rc != FGX_OK) return rc;
  }
  if (n_targets)
    grid_scatter_targets_kernel<<<grid_for(n_targets, 256, 8), 256, 0, st>>>(tx, ty, trad, tactive, target_stride,
                                                                             n_targets, s.tcell, s.tstart, s.tcount,
                                                                             s.sorted, s.tidx);
  grid_scatter_agents_kernel<<<grid_for(n_agents, 256, 8), 256, 0, st>>>(n_agents, x, y, ang, active, s.acell, s.astart,
                                                                        s.acount, s.order, s.rec);
  RaycastAgentsArgs a{g, n_agents, n_targets, n_rays, ray_span, ray_length, max_rad,
                      s.tstart, s.order, s.rec, s.sorted, s.tidx, merge_walls ? 1 : 0, dist, hit};
  // one warp per agent, walking the cell-sorted agent list (neighbouring warps share their targets in L1/L2)
  const int64_t ctas = ceil_div(n_agents, RA_WARPS);
  const int64_t cap = (int64_t)sm_count() * 16;
  const int grid = (int)(ctas < cap ? ctas : cap);
  const int rg = (n_rays + RA_G - 1) / RA_G;
  switch ((rg + RA_GW - 1) / RA_GW) {
    case 1: raycast_agents_kernel<1><<<grid, RA_BLOCK, 0, st>>>(a); break;
    case 2: raycast_agents_kernel<2><<<grid, RA_BLOCK, 0, st>>>(a); break;
    case 3: raycast_agents_kernel<3><<<grid, RA_BLOCK, 0, st>>>(a); break;
    default: raycast_agents_kernel<4><<<grid, RA_BLOCK, 0, st>>>(a); break;
  }
  return check_launch("fgx_raycast_agents");
}
}
