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
  const WState* st;
};

// one (ray, wall) pair, exactly as the all-pairs kernel decides it: sure miss <=> no candidate distance of
// ray_wall_sensor can apply (max(o1*o2, o3*o4) > 0 and all four determinants nonzero); otherwise the op-for-op
// restatement.  (distance, wall) minima are lexicographic = argmin's first-index rule for any visiting order.
__device__ __forceinline__ void test_wall(float p1x, float p1y, float q1x, float q1y, float L, float p2x, float p2y,
                                          float q2x, float q2y, int w, float& best, int& bidx) {
  const float o1 = orient_val(p1x, p1y, q1x, q1y, p2x, p2y), o2 = orient_val(p1x, p1y, q1x, q1y, q2x, q2y);
  const float o3 = orient_val(p2x, p2y, q2x, q2y, p1x, p1y), o4 = orient_val(p2x, p2y, q2x, q2y, q1x, q1y);
  const float A = f_mul(o1, o2), B = f_mul(o3, o4);
  if (fmaxf(A, B) > 0.0f && fabsf(f_mul(A, B)) > 0.0f) return;
  const float d = ray_wall_exact(p1x, p1y, q1x, q1y, L, p2x, p2y, q2x, q2y);
  if (d < best || (d == best && bidx >= 0 && w < bidx)) {
    best = d;
    bidx = w;
  }
}

// min / first-argmin over the walls for the ray p1 -> q1 (best = L, bidx = -1 on entry)
__device__ __forceinline__ void walls_grid_cast(const WallsGridView& v, float p1x, float p1y, float q1x, float q1y,
                                                float L, float& best, int& bidx) {
  // strictly inside the grid (NaN positions fail and take the all-walls loop)
  const float fx = (p1x - v.g.x0) * v.g.inv_cs, fy = (p1y - v.g.y0) * v.g.inv_cs;
  const bool inside = fx >= 0.0f && fy >= 0.0f && fx < (float)v.g.gx && fy < (float)v.g.gy;
  if (inside && v.st->fallback == 0) {
    const int cell = min((int)fy, v.g.gy - 1) * v.g.gx + min((int)fx, v.g.gx - 1);
    const uint32_t e0 = v.start[cell], e1 = v.start[cell + 1];
    for (uint32_t e = e0; e < e1; e += 4) {  // four independent (index -> wall record) load chains in flight
      int w[4];
      float4 wl[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) w[u] = v.entries[min(e + u, e1 - 1)];
#pragma unroll
      for (int u = 0; u < 4; ++u) wl[u] = v.wall4[w[u]];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (e + u < e1) test_wall(p1x, p1y, q1x, q1y, L, wl[u].x, wl[u].y, wl[u].z, wl[u].w, w[u], best, bidx);
    }
  } else {
    for (int w = 0; w < v.n_walls; ++w)
      test_wall(p1x, p1y, q1x, q1y, L, v.wx1[w], v.wy1[w], v.wx2[w], v.wy2[w], w, best, bidx);
  }
}

// host side (raycast_walls_grid.cu): validates, carves the scratch and -- unless reuse_grid -- builds the wall
// lists on `stream`; fills `view`.  Returns an fgx status code.
size_t wall_grid_scratch_bytes(int64_t n_walls, float x_min, float x_max, float y_min, float y_max, float ray_length);
int wall_grid_prepare(const char* what, const float* wx1, const float* wy1, const float* wx2, const float* wy2,
                      int64_t n_walls, float x_min, float x_max, float y_min, float y_max, float ray_length,
                      void* scratch, size_t scratch_bytes, int reuse_grid, cudaStream_t stream, WallsGridView* view);

}  // namespace fgx
