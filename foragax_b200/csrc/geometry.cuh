// Exact float32 restatement of the reference's 2-D geometry (space_methods.py:10-143)
// as device functions.  Every operation is an explicitly rounded __f*_rn intrinsic so
// that nvcc never contracts mul+add into FMA: the orientation predicates are discrete
// (val == 0 / > 0 / else, space_methods.py:112-113) and must see the same roundings as
// an un-contracted evaluation.
#pragma once
#include <cuda_runtime.h>

namespace fgx {

__device__ __forceinline__ float f_sub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float f_add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float f_mul(float a, float b) { return __fmul_rn(a, b); }

// space_methods.py:112: val = (q.y-p.y)*(r.x-q.x) - (q.x-p.x)*(r.y-q.y)
__device__ __forceinline__ float orient_val(float px, float py, float qx, float qy, float rx, float ry) {
  return f_sub(f_mul(f_sub(qy, py), f_sub(rx, qx)), f_mul(f_sub(qx, px), f_sub(ry, qy)));
}
// space_methods.py:113: select(val == 0, 0, select(val > 0, 1, 2)); NaN -> 2
__device__ __forceinline__ int orient_cls(float v) { return v == 0.0f ? 0 : (v > 0.0f ? 1 : 2); }

// space_methods.py:105-110 (and :13-17): q inside the bounding box of p, r
__device__ __forceinline__ bool on_segment(float px, float py, float qx, float qy, float rx, float ry) {
  return qx <= fmaxf(px, rx) && qx >= fminf(px, rx) && qy <= fmaxf(py, ry) && qy >= fminf(py, ry);
}

__device__ __forceinline__ float norm2(float dx, float dy) {
  return __fsqrt_rn(f_add(f_mul(dx, dx), f_mul(dy, dy)));
}

// jnp.minimum propagates NaN
__device__ __forceinline__ float min_nan(float a, float b) {
  return (a != a || b != b) ? __int_as_float(0x7fc00000) : fminf(a, b);
}

// ray_wall_sensor (space_methods.py:96-143) for ray p1->q1 (q1 = origin + dir*L already
// rounded by the caller) and wall p2-q2.  Returns L when nothing is hit.
static __device__ __noinline__ float ray_wall_exact(float p1x, float p1y, float q1x, float q1y, float L, float p2x,
                                             float p2y, float q2x, float q2y) {
  const int o1 = orient_cls(orient_val(p1x, p1y, q1x, q1y, p2x, p2y));
  const int o2 = orient_cls(orient_val(p1x, p1y, q1x, q1y, q2x, q2y));
  const int o3 = orient_cls(orient_val(p2x, p2y, q2x, q2y, p1x, p1y));
  const int o4 = orient_cls(orient_val(p2x, p2y, q2x, q2y, q1x, q1y));
  float d1 = L, d2 = L, d3 = L, d4 = L, d5 = L;
  if (o1 != o2 && o3 != o4) {
    // line_intercept (space_methods.py:98-103)
    const float ax = f_sub(q1x, p1x), ay = f_sub(q1y, p1y);
    const float c1 = f_sub(f_mul(ax, f_sub(p2y, p1y)), f_mul(ay, f_sub(p2x, p1x)));
    const float c2 = f_sub(f_mul(ax, f_sub(q2y, p1y)), f_mul(ay, f_sub(q2x, p1x)));
    const float r = __fdiv_rn(c1, f_add(f_sub(c1, c2), 1e-6f));
    const float ix = f_add(p2x, f_mul(r, f_sub(q2x, p2x)));
    const float iy = f_add(p2y, f_mul(r, f_sub(q2y, p2y)));
    d1 = norm2(f_sub(ix, p1x), f_sub(iy, p1y));
  }
  if (o1 == 0 && on_segment(p1x, p1y, p2x, p2y, q1x, q1y)) d2 = norm2(f_sub(p1x, p2x), f_sub(p1y, p2y));
  if (o2 == 0 && on_segment(p1x, p1y, q2x, q2y, q1x, q1y)) d3 = norm2(f_sub(p1x, q2x), f_sub(p1y, q2y));
  if (o3 == 0 && on_segment(p2x, p2y, p1x, p1y, q2x, q2y)) d4 = 0.0f;
  if (o4 == 0 && on_segment(p2x, p2y, q1x, q1y, q2x, q2y)) d5 = norm2(f_sub(p1x, q1x), f_sub(p1y, q1y));
  return min_nan(d1, min_nan(d2, min_nan(d3, min_nan(d4, d5))));
}

// ray_agent_sensor (space_methods.py:71-94): ray-circle intersection
__device__ __forceinline__ float ray_agent_exact(float ox, float oy, float dx, float dy, float L, float ax, float ay,
                                                 float ar, float aactive) {
  if (aactive == 0.0f) return L;
  const float sx = f_sub(ox, ax), sy = f_sub(oy, ay);
  const float b = f_add(f_mul(sx, dx), f_mul(sy, dy));
  const float c = f_sub(f_add(f_mul(sx, sx), f_mul(sy, sy)), f_mul(ar, ar));
  float h = f_sub(f_mul(b, b), c);
  h = h < 0.0f ? -1.0f : __fsqrt_rn(h);
  float t = h >= 0.0f ? f_sub(-b, h) : L;
  t = t < 0.0f ? L : t;
  return min_nan(t, L);
}

// do_intersect of check_collision (space_methods.py:10-43): wall p1-q1 vs move p2->q2
__device__ __forceinline__ bool segments_intersect(float p1x, float p1y, float q1x, float q1y, float p2x, float p2y,
                                                   float q2x, float q2y) {
  const int o1 = orient_cls(orient_val(p1x, p1y, q1x, q1y, p2x, p2y));
  const int o2 = orient_cls(orient_val(p1x, p1y, q1x, q1y, q2x, q2y));
  const int o3 = orient_cls(orient_val(p2x, p2y, q2x, q2y, p1x, p1y));
  const int o4 = orient_cls(orient_val(p2x, p2y, q2x, q2y, q1x, q1y));
  return (o1 != o2 && o3 != o4) || (o1 == 0 && on_segment(p1x, p1y, p2x, p2y, q1x, q1y)) ||
         (o2 == 0 && on_segment(p1x, p1y, q2x, q2y, q1x, q1y)) ||
         (o3 == 0 && on_segment(p2x, p2y, p1x, p1y, q2x, q2y)) ||
         (o4 == 0 && on_segment(p2x, p2y, q1x, q1y, q2x, q2y));
}

// jnp.linspace(a, b, n)[j] in float32: a*(1-s) + b*s with s = j/(n-1); last element is b
__device__ __forceinline__ float linspace_elem(float a, float b, int n, int j) {
  if (n == 1) return a;
  if (j == n - 1) return b;
  const float s = __fdiv_rn((float)j, (float)(n - 1));
  return f_add(f_mul(a, f_sub(1.0f, s)), f_mul(b, s));
}

}  // namespace fgx
