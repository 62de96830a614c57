// K1 with a spatial cull: the same agents x rays x walls reduction as fgx_raycast_walls, evaluated only
// against the walls that can possibly produce a distance below ray_length -- bit-identical results.
//
// Reference: ray_wall_sensor (space_methods.py:96-143) + min / first-argmin over walls
// (EcoEvoJax/sim.py:444-447).  The reference is all-pairs; a 1 M-agent arena with 1,000 walls has ~6 walls
// anywhere near an agent.  A wall can yield d < L for a ray from p1 only if
//   (a) the wall SEGMENT is within L (+ margin) of p1: proper crossings (d1) and the collinear
//       touch cases d2, d3, d5 put a wall point inside the ray's bounding box; or
//   (b) p1 lies within the float32 noise of the wall's infinite LINE: then the computed orientation
//       signs o3 / o4 are arbitrary, `o1 != o2 && o3 != o4` can hold spuriously and the intercept
//       formula (with its 1e-6 denominator bias) can return a point near p1 for a wall that is far
//       away; likewise d4 = 0 needs o3 == 0 exactly.  The all-pairs kernel reproduces these cases,
//       so the cull must not drop them.
// Orientation determinants are accurate to ~4e-7 * S in distance units (S = coordinate scale), so the
// cull uses eps_line = 1e-4 * S for (b) and margin = 5e-4 * S for (a) -- hundreds of times the noise.
// Every listed (ray, wall) pair is evaluated by the same op-for-op code as the all-pairs kernel
// (sure-miss filter on the four determinants, then ray_wall_exact), and ties resolve to the lowest
// wall index, so the output equals fgx_raycast_walls bit for bit.
//
//   1. wall lists per cell of a uniform grid: one thread per (wall, grid column along the wall's major
//      axis); in a column the wall's line occupies a short run of cells -- a band of half-width
//      L + margin where the column is within reach of the segment, eps_line elsewhere (count pass,
//      scan, fill pass);
//   2. one thread per (agent, ray): the lanes of a warp share an agent (32 rays), read its cell's wall
//      list (broadcast loads, wall data L1-resident) and test their own ray.
// Agents outside the grid bounds, walls with coordinates beyond the declared scale and list overflow
// fall back to the all-walls loop for the affected agents / launch (exact, just slower).
#include <algorithm>

#include "scan.cuh"
#include "walls_grid.cuh"

namespace fgx {

__device__ __forceinline__ int wg_coord(float v, float v0, float inv_cs, int n) {
  const float f = floorf((v - v0) * inv_cs);
  return f >= (float)(n - 1) ? n - 1 : (f > 0.0f ? (int)f : 0);
}

// cells of major-axis column `cu` whose agents may need wall (ax,ay)-(bx,by).  One thread per (wall, column):
// in a column the wall's line occupies a short run of cells.  emit(cell) may be called for a cell by two
// neighbouring columns' pads (harmless: evaluating a wall twice changes nothing).
template <class F>
__device__ __forceinline__ void wall_column_cells(const WGrid& g, float ax, float ay, float bx, float by, int cu, F emit) {
  const bool xmajor = fabsf(bx - ax) >= fabsf(by - ay);
  const int nu = xmajor ? g.gx : g.gy, nv = xmajor ? g.gy : g.gx;
  if (cu >= nu) return;
  float au = xmajor ? ax : ay, av = xmajor ? ay : ax, bu = xmajor ? bx : by, bv = xmajor ? by : bx;
  if (au > bu) {
    float t = au; au = bu; bu = t;
    t = av; av = bv; bv = t;
  }
  const float du = bu - au;
  const float m = du > 0.0f ? (bv - av) / du : 0.0f;  // |m| <= 1 along the major axis
  const float wv = sqrtf(1.0f + m * m);                // perpendicular half-width -> extent along the minor axis
  const float u0 = xmajor ? g.x0 : g.y0, v0 = xmajor ? g.y0 : g.x0;
  const float pad = 0.02f * g.cs + 1e-5f * g.scale;    // cell assignment of agents on a cell edge, rounding here
  // agents OUTSIDE the grid take the all-walls loop in the sensing kernel, so columns are the plain intervals
  const float ulo = u0 + cu * g.cs - pad, uhi = u0 + (cu + 1) * g.cs + pad;
  const bool thick = uhi >= au - g.thick && ulo <= bu + g.thick;
  const float hw = (thick ? g.thick : g.eps) * wv + pad;
  const float v1 = av + m * (ulo - au), v2 = av + m * (uhi - au);
  const float vlo = fminf(v1, v2) - hw, vhi = fmaxf(v1, v2) + hw;
  if (vhi < v0 || vlo > v0 + nv * g.cs) return;
  const int c0 = wg_coord(vlo, v0, g.inv_cs, nv), c1 = wg_coord(vhi, v0, g.inv_cs, nv);
  for (int cv = c0; cv <= c1; ++cv) emit(xmajor ? cv * g.gx + cu : cu * g.gx + cv);
}

__device__ __forceinline__ bool wall_in_scale(const WGrid& g, float ax, float ay, float bx, float by) {
  const float s = g.scale;  // NaN coordinates fail every comparison -> out of scale
  return fabsf(ax) <= s && fabsf(ay) <= s && fabsf(bx) <= s && fabsf(by) <= s;
}

// grid = (columns / 128, walls): thread <-> (wall, major-axis column)
__global__ void __launch_bounds__(128) wall_count_kernel(WGrid g, const float* __restrict__ wx1,
                                                         const float* __restrict__ wy1, const float* __restrict__ wx2,
                                                         const float* __restrict__ wy2, int n_walls,
                                                         uint32_t* __restrict__ count, float4* __restrict__ wall4,
                                                         float4* __restrict__ wallr, WState* __restrict__ st) {
  const int cu = blockIdx.x * blockDim.x + threadIdx.x;
  for (int w = blockIdx.y; w < n_walls; w += gridDim.y) {
    const float ax = wx1[w], ay = wy1[w], bx = wx2[w], by = wy2[w];
    if (cu == 0) {  // the records the sensing side reads: end points; bounding box + unit-normal line with its band
      wall4[w] = make_float4(ax, ay, bx, by);
      wallr[2 * w] = make_float4(fminf(ax, bx), fminf(ay, by), fmaxf(ax, bx), fmaxf(ay, by));
      wallr[2 * w + 1] = wall_line_record(ax, ay, bx, by, g.tol_line, g.scale);
    }
    if (!wall_in_scale(g, ax, ay, bx, by)) {
      if (cu == 0) st->fallback = 1;
      continue;
    }
    wall_column_cells(g, ax, ay, bx, by, cu, [&](int cell) { atomicAdd(count + cell, 1u); });
  }
}

__global__ void __launch_bounds__(128) wall_fill_kernel(WGrid g, const float* __restrict__ wx1,
                                                        const float* __restrict__ wy1, const float* __restrict__ wx2,
                                                        const float* __restrict__ wy2, int n_walls,
                                                        const uint32_t* __restrict__ start, uint32_t* __restrict__ cursor,
                                                        int32_t* __restrict__ entries, const WState* __restrict__ st) {
  if (st->fallback) return;
  const int cu = blockIdx.x * blockDim.x + threadIdx.x;
  for (int w = blockIdx.y; w < n_walls; w += gridDim.y)
    wall_column_cells(g, wx1[w], wy1[w], wx2[w], wy2[w], cu,
                      [&](int cell) { entries[start[cell] + atomicAdd(cursor + cell, 1u)] = w; });
}

// cell records for the sensing side: list range + the first six wall indices of every cell (unused index slots
// repeat a valid index -- 0 when the list is empty -- so that speculative record loads stay in bounds)
__global__ void __launch_bounds__(256) wall_cellrec_kernel(int n_cells, const uint32_t* __restrict__ start,
                                                           const int32_t* __restrict__ entries,
                                                           uint4* __restrict__ cellrec, const WState* __restrict__ st) {
  if (st->fallback) return;
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n_cells; c += gridDim.x * blockDim.x) {
    const uint32_t e0 = start[c], n = start[c + 1] - e0;
    uint32_t w[6];
#pragma unroll
    for (int u = 0; u < 6; ++u) w[u] = n ? (uint32_t)entries[e0 + min((uint32_t)u, n - 1)] : 0u;
    cellrec[2 * c] = make_uint4(e0, n, w[0], w[1]);
    cellrec[2 * c + 1] = make_uint4(w[2], w[3], w[4], w[5]);
  }
}

struct WallsGridArgs {
  WallsGridView v;
  const float *ax, *ay, *aang, *aactive;
  int64_t n;
  int n_rays;
  float span, L;
  float* dist;
  int32_t* hit;
};

// NR > 0: rays per agent known at compile time (index split by shifts / constant division); NR == 0: run time.
// 32-bit flat (agent, ray) index; linspace steps j / (n_rays - 1) come from a per-CTA shared-memory table (the
// IEEE division -- whose j = 0 lane takes the slow path of the software divide -- runs once per CTA, not per ray).
constexpr int WG_STEP_TABLE = 256;

// MINB: resident CTAs per SM the register allocation aims for (4 = 64 registers; 3 = 80; FGX_WG_OCC picks, default 4)
template <int NR, int MINB>
__global__ void __launch_bounds__(256, MINB) raycast_walls_grid_kernel(const WallsGridArgs a, uint32_t total) {
  __shared__ float s_step[WG_STEP_TABLE];
  const uint32_t nr = NR > 0 ? (uint32_t)NR : (uint32_t)a.n_rays;
  const bool table = nr <= (uint32_t)WG_STEP_TABLE;
  if (table) {
    for (uint32_t j = threadIdx.x; j < nr; j += blockDim.x)
      s_step[j] = nr > 1 ? __fdiv_rn((float)j, (float)(nr - 1)) : 0.0f;
    __syncthreads();
  }
  const float L = a.L, span = a.span;
  const bool fallback = a.v.st->fallback != 0;  // written by the grid build (an earlier launch)
  // software pipeline over the grid-stride loop: the pose of the NEXT (agent, ray) of this thread is loaded while
  // the current one is processed, which takes the first of the dependent loads (pose -> cell record -> wall
  // records -> end points) off every agent's critical path
  const uint32_t stride = gridDim.x * blockDim.x;
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  float n_act = 1.0f, n_x = 0.0f, n_y = 0.0f, n_ang = 0.0f;
  if (t < total) {
    const uint32_t ag = t / nr;
    n_act = a.aactive != nullptr ? __ldg(a.aactive + ag) : 1.0f;
    n_x = __ldg(a.ax + ag); n_y = __ldg(a.ay + ag); n_ang = __ldg(a.aang + ag);
  }
  for (; t < total; t += stride) {
    const uint32_t agent = t / nr;
    const uint32_t j = t - agent * nr;
    const float act = n_act, p1x = n_x, p1y = n_y, ang = n_ang;
    if (t + stride < total) {
      const uint32_t ag = (t + stride) / nr;
      n_act = a.aactive != nullptr ? __ldg(a.aactive + ag) : 1.0f;
      n_x = __ldg(a.ax + ag); n_y = __ldg(a.ay + ag); n_ang = __ldg(a.aang + ag);
    }
    if (act == 0.0f) {  // inactive agents sense nothing (sim.py:459-460)
      a.dist[t] = 0.0f;
      if (a.hit) a.hit[t] = -1;
      continue;
    }
    // ray_generator (space_methods.py:56-69): linspace(ang - span, ang + span, n)[j] = lo*(1-s) + hi*s, last = hi
    const float lo = f_sub(ang, span), hi = f_add(ang, span);
    float th;
    if (!table) th = linspace_elem(lo, hi, (int)nr, (int)j);
    else if (nr == 1) th = lo;
    else if (j == nr - 1) th = hi;
    else {
      const float sj = s_step[j];
      th = f_add(f_mul(lo, f_sub(1.0f, sj)), f_mul(hi, sj));
    }
    float sn, cs;
    sincosf(th, &sn, &cs);
    const float q1x = f_add(p1x, f_mul(cs, L)), q1y = f_add(p1y, f_mul(sn, L));
    float best = L;
    int bidx = -1;
    walls_grid_cast(a.v, fallback, p1x, p1y, q1x, q1y, L, best, bidx);
    a.dist[t] = best;
    if (a.hit) a.hit[t] = bidx;
  }
}

// ---- n_rays == 32: a warp owns an agent (lane = ray) and walks the cell's wall list cooperatively ----------------
// Same outputs, bit for bit, as raycast_walls_grid_kernel: only which (ray, wall) pairs are skipped BEFORE the exact
// test changes, and every skip is implied by walls_pretest_skip.
//   A. tiles of 8 / 16 / 32 consecutive agents go round-robin over the resident warps (neighbouring warps write
//      neighbouring output rows; one contiguous agent range per warp measured 9 % slower): lane k loads the pose of
//      agent k of the tile (coalesced), finds its cell and loads its 32-byte cell record; the tile is parked in
//      shared memory.  The loads of the NEXT tile
//      are issued while the current one is processed, so that the dependent pose -> cell record chain (two DRAM
//      latencies per agent in the per-thread kernel) is off the critical path.
//   B. per agent: lane j computes ray j (pose broadcast from shared memory; the linspace weights of lane j live in
//      registers); the FAN's bounding box is a warp reduction of the ray end points (redux.sync.{min,max}.f32).
//      B1: lane u pre-tests listed wall u against the fan box and the line band (up to 32 walls per trip; the
//          per-thread kernel needs a trip per two walls) -> ballot of survivors, parked in shared memory.
//      B2: per survivor: lane j pre-tests it against its own ray box, then runs the exact test.
// fan box = union of the 32 ray boxes, so "disjoint from the fan box" implies "disjoint from the ray box" of
// walls_pretest_skip; the line band condition is per agent and is evaluated once, by the lane that owns the wall.
constexpr int WGW_WARPS = 8;

template <int MINB>
__global__ void __launch_bounds__(32 * WGW_WARPS, MINB)
    raycast_walls_grid_warp32_kernel(const WallsGridArgs a, uint32_t n_agents, uint32_t tile, uint32_t n_tiles) {
  __shared__ WgwTileSmem<32> s_tile[WGW_WARPS];
  __shared__ WgwSurvSmem s_surv[WGW_WARPS];
  const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
  WgwTileSmem<32>& sm = s_tile[wid];
  const float L = a.L, span = a.span;
  const WallsGridView& v = a.v;
  const bool fallback = v.st->fallback != 0;
  float sj, omsj;
  wgw_lane_weights(lane, sj, omsj);
  const uint32_t stride = gridDim.x * WGW_WARPS;
  uint32_t t = blockIdx.x * WGW_WARPS + wid;
  WgwTile nx;
  if (t < n_tiles) {
    wgw_load_pose(a.ax, a.ay, a.aang, a.aactive, min(n_agents, (t + 1u) * tile), t * tile + lane, nx);
    wgw_load_rec(v, fallback, nx);
  }
  for (; t < n_tiles; t += stride) {
    const uint32_t base = t * tile;
    const uint32_t cnt = min(tile, n_agents - base);
    wgw_park(sm, lane, nx);
    __syncwarp();
    const uint32_t tn = t + stride;
    if (tn < n_tiles) wgw_load_pose(a.ax, a.ay, a.aang, a.aactive, min(n_agents, (tn + 1u) * tile), tn * tile + lane, nx);
    for (uint32_t k = 0; k < cnt; ++k) {
      if (k == (cnt >> 1) && tn < n_tiles) wgw_load_rec(v, fallback, nx);
      const uint32_t o = (base + k) * 32u + lane;
      float best;
      int bidx;
      WgwPre pre;
      wgw_prefetch(v, sm, k, lane, pre);
      wgw_cast_agent(v, s_surv[wid], lane, L, span, sj, omsj, pre, best, bidx);
      a.dist[o] = best;
      if (a.hit) a.hit[o] = bidx;
    }
    __syncwarp();
  }
}

// the same for flat indices beyond 2^31 (64-bit index arithmetic)
__global__ void __launch_bounds__(256) raycast_walls_grid_kernel_big(const WallsGridArgs a) {
  const int64_t total = a.n * a.n_rays;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t agent = t / a.n_rays;
    const int j = (int)(t - agent * a.n_rays);
    if (a.aactive != nullptr && a.aactive[agent] == 0.0f) {
      a.dist[t] = 0.0f;
      if (a.hit) a.hit[t] = -1;
      continue;
    }
    const float p1x = a.ax[agent], p1y = a.ay[agent], ang = a.aang[agent], L = a.L;
    float s, c;
    sincosf(linspace_elem(f_sub(ang, a.span), f_add(ang, a.span), a.n_rays, j), &s, &c);  // ray_generator
    const float q1x = f_add(p1x, f_mul(c, L)), q1y = f_add(p1y, f_mul(s, L));
    float best = L;
    int bidx = -1;
    walls_grid_cast(a.v, p1x, p1y, q1x, q1y, L, best, bidx);
    a.dist[t] = best;
    if (a.hit) a.hit[t] = bidx;
  }
}

// statistics of the cull for a set of agents: out[0] = sum over active agents of the walls their rays are tested
// against (the cell's list; n_walls for agents that take the all-walls loop), out[1] = active agents
__global__ void __launch_bounds__(256) walls_grid_listed_kernel(WallsGridView v, const float* __restrict__ ax,
                                                                const float* __restrict__ ay,
                                                                const float* __restrict__ aactive, int64_t n,
                                                                unsigned long long* __restrict__ out) {
  unsigned long long listed = 0, act = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (aactive != nullptr && aactive[i] == 0.0f) continue;
    ++act;
    const float fx = (ax[i] - v.g.x0) * v.g.inv_cs, fy = (ay[i] - v.g.y0) * v.g.inv_cs;
    const bool inside = fx >= 0.0f && fy >= 0.0f && fx < (float)v.g.gx && fy < (float)v.g.gy;
    if (inside && v.st->fallback == 0) {
      const int cell = min((int)fy, v.g.gy - 1) * v.g.gx + min((int)fx, v.g.gx - 1);
      listed += v.start[cell + 1] - v.start[cell];
    } else {
      listed += (unsigned long long)v.n_walls;
    }
  }
  for (int d = 16; d > 0; d >>= 1) {
    listed += __shfl_xor_sync(0xffffffffu, listed, d);
    act += __shfl_xor_sync(0xffffffffu, act, d);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out, listed);
    atomicAdd(out + 1, act);
  }
}

static inline size_t wg_al(size_t x) { return (x + 255) / 256 * 256; }

static int wgrid_dims(float x_min, float x_max, float y_min, float y_max, float L, WGrid* g) {
  if (!(L > 0.0f) || !(x_max > x_min) || !(y_max > y_min)) return -1;
  const float ext = fmaxf(x_max - x_min, y_max - y_min);
  const float scale = 2.0f * fmaxf(fmaxf(fabsf(x_min), fabsf(x_max)), fmaxf(fabsf(y_min), fabsf(y_max)));
  // finer cells shorten the per-cell lists (the thin line bands scale with the cell size); FGX_WG_DIV tunes
  static const float div = [] { const char* v = getenv("FGX_WG_DIV"); const float d = v ? (float)atof(v) : 0.0f; return d >= 1.0f ? d : 8.0f; }();
  float cs = L / div;
  if (ext / cs > 1024.0f) cs = ext / 1024.0f;
  g->x0 = x_min; g->y0 = y_min; g->cs = cs; g->inv_cs = 1.0f / cs;
  g->gx = (int)ceilf((x_max - x_min) / cs); g->gy = (int)ceilf((y_max - y_min) / cs);
  if (g->gx < 1) g->gx = 1;
  if (g->gy < 1) g->gy = 1;
  g->scale = scale;
  g->eps = 1e-4f * scale + 1e-6f;
  g->thick = L * 1.0001f + 5e-4f * scale + 1e-6f;
  g->margin = g->thick - L;
  // run-time line band (walls_pretest_skip): 8 x (2 L / margin + 1) x the determinant noise of 4e-7 * scale, at least eps
  const float noise = 4e-7f * scale;
  g->tol_line = fmaxf(g->eps, 8.0f * (2.0f * L / g->margin + 1.0f) * noise);
  return 0;
}

// worst case of wall_cells: per major-axis column the band covers (2 * half-width * sqrt2 + 2 pads) / cs + 2 cells
static size_t wgrid_capacity(const WGrid& g, int64_t n_walls) {
  const int nmaj = g.gx > g.gy ? g.gx : g.gy;
  const double thick_cells = 2.0 * 1.4143 * g.thick / g.cs + 4.0, thin_cells = 2.0 * 1.4143 * g.eps / g.cs + 4.0;
  // thick columns: the segment's extent is unknown on the host; allow every wall a reach-wide band over a
  // quarter of the grid plus the thin line over the rest, and let the device flag an overflow (fallback = exact)
  const double per_wall = thin_cells * nmaj + thick_cells * (nmaj / 4.0 + 2.0 * g.thick / g.cs + 2.0);
  double cap = per_wall * (double)n_walls;
  if (cap > 64e6) cap = 64e6;
  if (cap < 1024) cap = 1024;
  return (size_t)cap;
}

size_t wall_grid_scratch_bytes(int64_t n_walls, float x_min, float x_max, float y_min, float y_max, float ray_length) {
  WGrid g;
  if (n_walls < 0 || wgrid_dims(x_min, x_max, y_min, y_max, ray_length, &g)) return 0;
  const size_t cells = (size_t)g.gx * g.gy;
  return wg_al(sizeof(WState)) + 2 * wg_al((cells + 1) * 4) + wg_al(coop_scan_scratch_bytes(1)) + wg_al((cells + 1) * 32) +
         wg_al((size_t)n_walls * 16 + 16) + wg_al((size_t)n_walls * 32 + 16) + wg_al(wgrid_capacity(g, n_walls) * 4);
}

int wall_grid_prepare(const char* what, const float* wx1, const float* wy1, const float* wx2, const float* wy2,
                      int64_t n_walls, float x_min, float x_max, float y_min, float y_max, float ray_length,
                      void* scratch, size_t scratch_bytes, int reuse_grid, cudaStream_t s, WallsGridView* view) {
  FGX_REQUIRE(n_walls >= 0 && n_walls < (1ll << 31), "%s: bad wall count", what);
  FGX_REQUIRE(n_walls == 0 || (wx1 && wy1 && wx2 && wy2), "%s: null wall array", what);
  WGrid g;
  FGX_REQUIRE(wgrid_dims(x_min, x_max, y_min, y_max, ray_length, &g) == 0, "%s: bad arena bounds / ray_length", what);
  const size_t need = wall_grid_scratch_bytes(n_walls, x_min, x_max, y_min, y_max, ray_length);
  FGX_REQUIRE(scratch && scratch_bytes >= need, "%s: wall-grid scratch too small", what);
  FGX_REQUIRE((reinterpret_cast<uintptr_t>(scratch) & 255) == 0, "%s: wall-grid scratch must be 256-byte aligned", what);
  const int n_cells = g.gx * g.gy;
  const size_t capacity = wgrid_capacity(g, n_walls);
  char* p = reinterpret_cast<char*>(scratch);
  WState* st = reinterpret_cast<WState*>(p);
  p += wg_al(sizeof(WState));
  uint32_t* count = reinterpret_cast<uint32_t*>(p);
  p += wg_al(((size_t)n_cells + 1) * 4);
  uint32_t* start = reinterpret_cast<uint32_t*>(p);
  p += wg_al(((size_t)n_cells + 1) * 4);
  uint32_t* partial = reinterpret_cast<uint32_t*>(p);
  p += wg_al(coop_scan_scratch_bytes(1));
  uint4* cellrec = reinterpret_cast<uint4*>(p);
  p += wg_al(((size_t)n_cells + 1) * 32);
  float4* wall4 = reinterpret_cast<float4*>(p);
  p += wg_al((size_t)n_walls * 16 + 16);
  float4* wallr = reinterpret_cast<float4*>(p);
  p += wg_al((size_t)n_walls * 32 + 16);
  int32_t* entries = reinterpret_cast<int32_t*>(p);
  cudaError_t e = cudaSuccess;
  if (reuse_grid) {
    // the scratch holds the grid a previous call built from the same walls, bounds and ray_length
  } else if (n_walls) {
    e = cudaMemsetAsync(scratch, 0, wg_al(sizeof(WState)) + wg_al(((size_t)n_cells + 1) * 4), s);
    if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    const int nmaj = g.gx > g.gy ? g.gx : g.gy;
    const dim3 wg((unsigned)ceil_div(nmaj, 128), (unsigned)(n_walls < 32768 ? n_walls : 32768));
    wall_count_kernel<<<wg, 128, 0, s>>>(g, wx1, wy1, wx2, wy2, (int)n_walls, count, wall4, wallr, st);
    {  // cooperative exclusive scan; the counts become the fill cursors (reset to 0); overflow -> exact fallback
      const int rc = launch_coop_scan<1>(what, n_cells, ScanJob{count, start}, ScanJob{nullptr, nullptr}, partial,
                                         &st->total, (uint32_t)capacity, &st->fallback, s);
      if (rc != FGX_OK) return rc;
    }
    wall_fill_kernel<<<wg, 128, 0, s>>>(g, wx1, wy1, wx2, wy2, (int)n_walls, start, count, entries, st);
    wall_cellrec_kernel<<<grid_for(n_cells, 256, 4), 256, 0, s>>>(n_cells, start, entries, cellrec, st);
  } else {
    // no walls: empty lists everywhere (start, cursors and cell records all zero)
    e = cudaMemsetAsync(scratch, 0,
                        wg_al(sizeof(WState)) + 2 * wg_al(((size_t)n_cells + 1) * 4) + wg_al(coop_scan_scratch_bytes(1)) +
                            wg_al(((size_t)n_cells + 1) * 32),
                        s);
    if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  }
  *view = WallsGridView{g, wx1, wy1, wx2, wy2, (int)n_walls, start, entries, wall4, cellrec, wallr, st};
  return FGX_OK;
}

}  // namespace fgx

extern "C" {

size_t fgx_raycast_walls_grid_scratch_bytes(int64_t n_walls, float x_min, float x_max, float y_min, float y_max,
                                            float ray_length) {
  return fgx::wall_grid_scratch_bytes(n_walls, x_min, x_max, y_min, y_max, ray_length);
}

int fgx_raycast_walls_grid(const float* x, const float* y, const float* ang, const float* active, int64_t n_agents,
                           int32_t n_rays, float ray_span, float ray_length, const float* wx1, const float* wy1,
                           const float* wx2, const float* wy2, int64_t n_walls, float x_min, float x_max, float y_min,
                           float y_max, float* dist, int32_t* hit, void* scratch, size_t scratch_bytes,
                           int32_t reuse_grid, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n_agents >= 0, "fgx_raycast_walls_grid: bad sizes");
  FGX_REQUIRE(n_rays > 0, "fgx_raycast_walls_grid: n_rays must be > 0");
  cudaStream_t s = as_stream(stream);
  WallsGridArgs a{};
  // the grid is built (reuse_grid == 0) even for an empty agent set: the caller caches the scratch as "built"
  const int rc = wall_grid_prepare("fgx_raycast_walls_grid", wx1, wy1, wx2, wy2, n_walls, x_min, x_max, y_min, y_max,
                                   ray_length, scratch, scratch_bytes, reuse_grid, s, &a.v);
  if (rc != FGX_OK) return rc;
  if (n_agents == 0) return FGX_OK;
  FGX_REQUIRE(x && y && ang && dist, "fgx_raycast_walls_grid: null pointer");
  a.ax = x; a.ay = y; a.aang = ang; a.aactive = active;
  a.n = n_agents; a.n_rays = n_rays; a.span = ray_span; a.L = ray_length;
  a.dist = dist; a.hit = hit;
  const int64_t total = n_agents * (int64_t)n_rays;
  const int grid = grid_for(total, 256, 8);
  if (total >= (1ll << 31)) {
    raycast_walls_grid_kernel_big<<<grid, 256, 0, s>>>(a);
  } else {
    const uint32_t t32 = (uint32_t)total;
    static const int occ = [] { const char* v = getenv("FGX_WG_OCC"); return v && atoi(v) == 3 ? 3 : 4; }();
    // FGX_WG_WARP=0 keeps the per-thread kernel for 32 rays; the warp kernel runs 3 CTAs per SM (78 registers; the
    // 64-register build spills and measured slower) unless FGX_WG_OCC=4
    static const int warp_coop = [] { const char* v = getenv("FGX_WG_WARP"); return v ? atoi(v) : 1; }();
    static const int wocc = [] { const char* v = getenv("FGX_WG_OCC"); return v && atoi(v) == 4 ? 4 : 3; }();
    if (n_rays == 32 && warp_coop) {
      // tiles of 32 agents amortise the lane-parallel loads best; smaller ones keep >= 4 tiles per resident warp
      // when the agent set is small (the tail of a static round-robin is one tile).  FGX_WG_TILE overrides.
      static const int tile_env = [] { const char* v = getenv("FGX_WG_TILE"); return v ? atoi(v) : 0; }();
      const int64_t resident = (int64_t)sm_count() * wocc * WGW_WARPS;
      const uint32_t tile = tile_env == 8 || tile_env == 16 || tile_env == 32 ? (uint32_t)tile_env
                            : n_agents >= 128 * resident ? 32u : n_agents >= 64 * resident ? 16u : 8u;
      const uint32_t n_tiles = (uint32_t)ceil_div(n_agents, tile);
      const int wgrid = (int)std::min<int64_t>(ceil_div(n_tiles, WGW_WARPS), (int64_t)sm_count() * wocc);
      if (wocc == 3) raycast_walls_grid_warp32_kernel<3><<<wgrid, 32 * WGW_WARPS, 0, s>>>(a, (uint32_t)n_agents, tile, n_tiles);
      else raycast_walls_grid_warp32_kernel<4><<<wgrid, 32 * WGW_WARPS, 0, s>>>(a, (uint32_t)n_agents, tile, n_tiles);
      return check_launch("fgx_raycast_walls_grid");
    }
    switch (n_rays) {
      case 32:
        if (occ == 3) raycast_walls_grid_kernel<32, 3><<<grid, 256, 0, s>>>(a, t32);
        else raycast_walls_grid_kernel<32, 4><<<grid, 256, 0, s>>>(a, t32);
        break;
      case 16: raycast_walls_grid_kernel<16, 4><<<grid, 256, 0, s>>>(a, t32); break;
      case 9: raycast_walls_grid_kernel<9, 4><<<grid, 256, 0, s>>>(a, t32); break;
      case 8: raycast_walls_grid_kernel<8, 4><<<grid, 256, 0, s>>>(a, t32); break;
      default: raycast_walls_grid_kernel<0, 4><<<grid, 256, 0, s>>>(a, t32); break;
    }
  }
  return check_launch("fgx_raycast_walls_grid");
}

int fgx_raycast_walls_grid_listed(const float* x, const float* y, const float* active, int64_t n_agents, const float* wx1,
                                  const float* wy1, const float* wx2, const float* wy2, int64_t n_walls, float x_min,
                                  float x_max, float y_min, float y_max, float ray_length, void* scratch,
                                  size_t scratch_bytes, int32_t reuse_grid, int64_t* out2, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n_agents >= 0 && out2, "fgx_raycast_walls_grid_listed: bad arguments");
  cudaStream_t s = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(out2, 0, 2 * sizeof(int64_t), s);
  if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "fgx_raycast_walls_grid_listed: %s", cudaGetErrorString(e));
  WallsGridView v{};
  const int rc = wall_grid_prepare("fgx_raycast_walls_grid_listed", wx1, wy1, wx2, wy2, n_walls, x_min, x_max, y_min,
                                   y_max, ray_length, scratch, scratch_bytes, reuse_grid, s, &v);
  if (rc != FGX_OK) return rc;
  if (n_agents == 0) return FGX_OK;
  FGX_REQUIRE(x && y, "fgx_raycast_walls_grid_listed: null pointer");
  walls_grid_listed_kernel<<<grid_for(n_agents, 256, 8), 256, 0, s>>>(v, x, y, active, n_agents,
                                                                      reinterpret_cast<unsigned long long*>(out2));
  return check_launch("fgx_raycast_walls_grid_listed");
}
}
