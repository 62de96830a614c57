// K1 (part 2): ray casting against nearby agents -- agents x rays x circle targets.
//
// Reference: ray_agent_sensor (space_methods.py:71-94) evaluated for every (ray, target) and
// reduced per ray with min / first-argmin (EcoEvoJax/sim.py:444-447).  The reference is all-pairs
// (O(N^2 R)); here a uniform grid with cell >= ray_length + max_rad bounds the search to the 3x3
// neighbourhood of the sensing agent's cell.  The cull is exact, not approximate: a target farther
// than ray_length + rad from the ray origin cannot produce t < ray_length (the rounding error of
// t = -b - sqrt(b*b - c) is ~1e-6 relative, the cell carries a 1e-4 relative margin), so the result
// equals the all-pairs evaluation bit for bit; ties are resolved by the lexicographic (distance,
// target index) minimum, i.e. argmin's first-index rule, independent of the grid order.
//
//   1. counting sort of the targets by cell (histogram with atomics, scan, scatter) and of the
//      sensing agents by cell;
//   2. one CTA task = 8 agents of ONE cell, one warp per agent; the targets of the 3x3
//      neighbourhood are staged through shared memory in chunks and shared by the 8 warps.
//      Broad phase, lanes = targets: each lane takes one staged target, forms s = o - c and
//      c = |s|^2 - r^2 and rejects it if it is inactive, farther than ray_length + r, or outside the
//      ray fan by more than its radius.  Narrow phase, lanes = rays: the few survivors (ballot) are
//      broadcast one by one with shuffles and every lane tests its own ray against them;
//   3. optional merge with the wall result of fgx_raycast_walls in entity order [agents, walls].
#include "common.cuh"
#include "geometry.cuh"

namespace fgx {

constexpr int RA_BLOCK = 256;
constexpr int RA_WARPS = RA_BLOCK / 32;  // agents per CTA task
constexpr int RA_MAXRB = 4;              // rays per lane (n_rays <= 128)
constexpr int RA_CHUNK = 512;  // targets staged per shared-memory chunk

struct GridDesc {
  float x0, y0, inv_cs;
  int gx, gy;
};

__device__ __forceinline__ int cell_coord(float v, float v0, float inv_cs, int g) {
  const float f = floorf((v - v0) * inv_cs);
  // clamp (monotone): |a - b| <= cell size still implies |cell(a) - cell(b)| <= 1; NaN -> 0
  return f >= (float)(g - 1) ? g - 1 : (f > 0.0f ? (int)f : 0);
}
__device__ __forceinline__ int cell_of(const GridDesc& g, float x, float y) {
  return cell_coord(y, g.y0, g.inv_cs, g.gy) * g.gx + cell_coord(x, g.x0, g.inv_cs, g.gx);
}

// histogram of points per cell; remembers each point's cell
__global__ void __launch_bounds__(256) grid_count_kernel(GridDesc g, const float* __restrict__ x,
                                                         const float* __restrict__ y, int64_t stride, int64_t n,
                                                         int32_t* __restrict__ cell, uint32_t* __restrict__ count) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = cell_of(g, x[i * stride], y[i * stride]);
    cell[i] = c;
    atomicAdd(count + c, 1u);
  }
}

// single CTA: exclusive scans of the target counts, the agent counts and the per-cell CTA-task counts
__global__ void __launch_bounds__(1024) grid_scan_kernel(int n_cells, int groups, uint32_t* __restrict__ tcount,
                                                         uint32_t* __restrict__ acount, uint32_t* __restrict__ tstart,
                                                         uint32_t* __restrict__ astart, uint32_t* __restrict__ bstart) {
  __shared__ uint32_t carry[3];
  __shared__ uint32_t wsum[3][32];
  if (threadIdx.x < 3) carry[threadIdx.x] = 0;
  __syncthreads();
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  for (int base = 0; base < n_cells; base += 1024) {
    const int c = base + threadIdx.x;
    uint32_t v[3] = {0u, 0u, 0u};
    if (c < n_cells) {
      v[0] = tcount[c];
      v[1] = acount[c];
      v[2] = (v[1] + RA_WARPS - 1) / RA_WARPS;  // CTA tasks of this cell (8 agents each)
    }
    uint32_t inc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      inc[k] = v[k];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const uint32_t o = __shfl_up_sync(0xffffffffu, inc[k], d);
        if (l >= d) inc[k] += o;
      }
      if (l == 31) wsum[k][w] = inc[k];
    }
    __syncthreads();
    uint32_t* outs[3] = {tstart, astart, bstart};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      uint32_t wb = carry[k];
      for (int q = 0; q < w; ++q) wb += wsum[k][q];
      if (c < n_cells) outs[k][c] = wb + inc[k] - v[k];
      inc[k] += wb;
    }
    __syncthreads();
    if (threadIdx.x == 1023) {
#pragma unroll
      for (int k = 0; k < 3; ++k) carry[k] = inc[k];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    tstart[n_cells] = carry[0];
    astart[n_cells] = carry[1];
    bstart[n_cells] = carry[2];
  }
  // the counts double as scatter cursors: reset them
  for (int c = threadIdx.x; c < n_cells; c += 1024) {
    tcount[c] = 0;
    acount[c] = 0;
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
    const uint32_t pos = tstart[c] + atomicAdd(cursor + c, 1u);
    sorted[pos] = make_float4(tx[i * stride], ty[i * stride], tr[i * stride], ta[i * stride]);
    sorted_idx[pos] = (int32_t)i;
  }
}

__global__ void __launch_bounds__(256) grid_scatter_agents_kernel(int64_t n, const int32_t* __restrict__ cell,
                                                                  const uint32_t* __restrict__ astart,
                                                                  uint32_t* __restrict__ cursor,
                                                                  int32_t* __restrict__ order) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = cell[i];
    order[astart[c] + atomicAdd(cursor + c, 1u)] = (int32_t)i;
  }
}

struct RaycastAgentsArgs {
  GridDesc g;
  const float *ax, *ay, *aang, *aactive;  // sensing agents [n]
  int64_t n, m;
  int n_rays, groups;
  float span, L;
  const uint32_t *tstart, *astart, *bstart;
  const int32_t* order;
  const float4* targets;
  const int32_t* tidx;
  int merge_walls;
  float* dist;
  int32_t* hit;
};

template <int RB>
__global__ void __launch_bounds__(RA_BLOCK) raycast_agents_kernel(const RaycastAgentsArgs a) {
  __shared__ float4 st[RA_CHUNK];
  __shared__ int32_t si[RA_CHUNK];
  const int n_cells = a.g.gx * a.g.gy;
  const uint32_t total_tasks = a.bstart[n_cells];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float L = a.L;
  // far-target bound: a hit needs |s| <= L + r, i.e. c = |s|^2 - r^2 <= L^2 + 2 L r
  const float L2 = f_mul(L, L), twoL = f_mul(2.0f, L);
  // fan cull (only for a convex fan): a circle farther than r from the cone of half-angle `span`
  // around the heading cannot be hit by any ray of the fan
  const bool fan_cull = a.span < 1.5f;
  float cs_span = 1.0f, sn_span = 0.0f;
  if (fan_cull) sincosf(a.span, &sn_span, &cs_span);
  for (uint32_t task = blockIdx.x; task < total_tasks; task += gridDim.x) {
    int lo = 0, hi = n_cells;  // cell of this task: last c with bstart[c] <= task
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (a.bstart[mid] <= task) lo = mid; else hi = mid;
    }
    const int cell = lo;
    const uint32_t slot = (task - a.bstart[cell]) * RA_WARPS + warp;
    const uint32_t a0 = a.astart[cell], a1 = a.astart[cell + 1];
    const bool valid = a0 + slot < a1;                     // warp-uniform
    const int64_t agent = valid ? a.order[a0 + slot] : 0;
    const bool act = valid && (a.aactive == nullptr || a.aactive[agent] != 0.0f);
    const float ox = valid ? a.ax[agent] : 0.0f, oy = valid ? a.ay[agent] : 0.0f;
    const float ang = valid ? a.aang[agent] : 0.0f;
    float hu, hv;  // heading
    sincosf(ang, &hv, &hu);
    float dx[RB], dy[RB], best[RB];
    int bidx[RB];
    {
      const float lo_a = f_sub(ang, a.span), hi_a = f_add(ang, a.span);
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        sincosf(linspace_elem(lo_a, hi_a, a.n_rays, min(lane + 32 * q, a.n_rays - 1)), &dy[q], &dx[q]);
        best[q] = L;
        bidx[q] = -1;
      }
    }
    const int cx = cell % a.g.gx, cy = cell / a.g.gx;
    for (int ry = max(cy - 1, 0); ry <= min(cy + 1, a.g.gy - 1); ++ry) {
      const int c_lo = ry * a.g.gx + max(cx - 1, 0), c_hi = ry * a.g.gx + min(cx + 1, a.g.gx - 1);
      const uint32_t r0 = a.tstart[c_lo], r1 = a.tstart[c_hi + 1];  // the row's cells are contiguous
      for (uint32_t base = r0; base < r1; base += RA_CHUNK) {
        const int cnt = (int)min((uint32_t)RA_CHUNK, r1 - base);
        __syncthreads();
        for (int k = threadIdx.x; k < cnt; k += RA_BLOCK) {
          st[k] = a.targets[base + k];
          si[k] = a.tidx[base + k];
        }
        __syncthreads();
        if (!act) continue;
        for (int k0 = 0; k0 < cnt; k0 += 32) {
          // broad phase: lane <-> target
          const int k = k0 + lane;
          const float4 t = st[min(k, cnt - 1)];
          const float sx = f_sub(ox, t.x), sy = f_sub(oy, t.y);
          const float c = f_sub(f_add(f_mul(sx, sx), f_mul(sy, sy)), f_mul(t.z, t.z));
          bool pass = k < cnt && t.w != 0.0f && !(c > f_mul(f_add(L2, f_mul(twoL, t.z)), 1.0001f));
          if (fan_cull) {
            // v = c - o = -s in the heading frame: along = v.h, across = |v x h|
            const float along = -(sx * hu + sy * hv), across = fabsf(sx * hv - sy * hu);
            const float outside = across * cs_span - along * sn_span;  // distance to the fan's boundary line
            pass = pass && !(outside > t.z * 1.0001f + 1e-4f * (fabsf(sx) + fabsf(sy)) + 1e-6f);
          }
          unsigned m = __ballot_sync(0xffffffffu, pass);
          const int id = si[min(k, cnt - 1)];
          // narrow phase: lane <-> ray, survivors broadcast one at a time
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1u;
            const float bsx = __shfl_sync(0xffffffffu, sx, src), bsy = __shfl_sync(0xffffffffu, sy, src);
            const float bc = __shfl_sync(0xffffffffu, c, src);
            const int bid = __shfl_sync(0xffffffffu, id, src);
#pragma unroll
            for (int q = 0; q < RB; ++q) {
              const float b = f_add(f_mul(bsx, dx[q]), f_mul(bsy, dy[q]));
              const float h = f_sub(f_mul(b, b), bc);
              if (h >= 0.0f) {
                const float tt = f_sub(-b, __fsqrt_rn(h));
                if (tt >= 0.0f && (tt < best[q] || (tt == best[q] && bid < bidx[q]))) {
                  best[q] = tt;
                  bidx[q] = bid;
                }
              }
            }
          }
        }
      }
    }
    if (valid) {
#pragma unroll
      for (int q = 0; q < RB; ++q) {
        const int rj = lane + 32 * q;
        if (rj >= a.n_rays) continue;
        const int64_t o = agent * a.n_rays + rj;
        float d = act ? best[q] : 0.0f;
        int h = act ? bidx[q] : -1;
        if (a.merge_walls && act) {
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
}

static inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

struct RaScratch {
  uint32_t *tcount, *acount, *tstart, *astart, *bstart;
  int32_t *tcell, *acell, *order, *tidx;
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
  p = take((size_t)(n_cells + 1) * 4); if (s) s->bstart = (uint32_t*)p;
  p = take((size_t)m * 4); if (s) s->tcell = (int32_t*)p;
  p = take((size_t)n * 4); if (s) s->acell = (int32_t*)p;
  p = take((size_t)n * 4); if (s) s->order = (int32_t*)p;
  p = take((size_t)m * 4); if (s) s->tidx = (int32_t*)p;
  p = take((size_t)m * 16); if (s) s->sorted = (float4*)p;
  return o;
}

static int grid_dims(float x_min, float x_max, float y_min, float y_max, float ray_length, float max_rad,
                     GridDesc* g) {
  const float cs = (ray_length + max_rad) * 1.0001f;
  if (!(cs > 0.0f) || !(x_max > x_min) || !(y_max > y_min)) return -1;
  int gx = (int)ceilf((x_max - x_min) / cs), gy = (int)ceilf((y_max - y_min) / cs);
  gx = gx < 1 ? 1 : (gx > 1024 ? 1024 : gx);
  gy = gy < 1 ? 1 : (gy > 1024 ? 1024 : gy);
  // if the clamp made cells LARGER than cs that is still conservative (cell >= cs is all that matters)
  const float csx = (x_max - x_min) / gx, csy = (y_max - y_min) / gy;
  const float c = fmaxf(cs, fmaxf(csx, csy));
  g->x0 = x_min; g->y0 = y_min; g->inv_cs = 1.0f / c;
  g->gx = (int)ceilf((x_max - x_min) / c); g->gy = (int)ceilf((y_max - y_min) / c);
  if (g->gx < 1) g->gx = 1;
  if (g->gy < 1) g->gy = 1;
  return 0;
}

}  // namespace fgx

extern "C" {

size_t fgx_raycast_agents_scratch_bytes(int64_t n_agents, int64_t n_targets, float x_min, float x_max, float y_min,
                                        float y_max, float ray_length, float max_rad) {
  fgx::GridDesc g;
  if (fgx::grid_dims(x_min, x_max, y_min, y_max, ray_length, max_rad, &g)) return 0;
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
  FGX_REQUIRE(grid_dims(x_min, x_max, y_min, y_max, ray_length, max_rad, &g) == 0,
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
  const int groups = 1;  // one warp per agent
  if (n_targets) grid_count_kernel<<<grid_for(n_targets, 256, 8), 256, 0, st>>>(g, tx, ty, target_stride, n_targets, s.tcell, s.tcount);
  grid_count_kernel<<<grid_for(n_agents, 256, 8), 256, 0, st>>>(g, x, y, 1, n_agents, s.acell, s.acount);
  grid_scan_kernel<<<1, 1024, 0, st>>>(n_cells, groups, s.tcount, s.acount, s.tstart, s.astart, s.bstart);
  if (n_targets)
    grid_scatter_targets_kernel<<<grid_for(n_targets, 256, 8), 256, 0, st>>>(tx, ty, trad, tactive, target_stride,
                                                                             n_targets, s.tcell, s.tstart, s.tcount,
                                                                             s.sorted, s.tidx);
  grid_scatter_agents_kernel<<<grid_for(n_agents, 256, 8), 256, 0, st>>>(n_agents, s.acell, s.astart, s.acount,
                                                                        s.order);
  RaycastAgentsArgs a{g, x, y, ang, active, n_agents, n_targets, n_rays, groups, ray_span, ray_length,
                      s.tstart, s.astart, s.bstart, s.order, s.sorted, s.tidx, merge_walls ? 1 : 0, dist, hit};
  // upper bound of CTA tasks: every cell rounds up by at most one
  const int64_t max_tasks = ceil_div(n_agents, RA_WARPS) + n_cells;
  const int64_t cap = (int64_t)sm_count() * 8;
  const int grid = (int)(max_tasks < cap ? max_tasks : cap);
  switch ((n_rays + 31) / 32) {
    case 1: raycast_agents_kernel<1><<<grid, RA_BLOCK, 0, st>>>(a); break;
    case 2: raycast_agents_kernel<2><<<grid, RA_BLOCK, 0, st>>>(a); break;
    case 3: raycast_agents_kernel<3><<<grid, RA_BLOCK, 0, st>>>(a); break;
    default: raycast_agents_kernel<4><<<grid, RA_BLOCK, 0, st>>>(a); break;
  }
  return check_launch("fgx_raycast_agents");
}
}
