// K1: ray casting -- agents x rays against wall segments.
//
// Reference: space_methods.py:56-69 (ray_generator), :96-143 (ray_wall_sensor), and the
// min / first-argmin / "-1 if min >= L" reduction of the example-level composition
// (examples/.../EcoEvoJax/sim.py:444-447).  Work = agents x rays x walls intersection
// tests; FP32/issue bound (walls live in shared memory, 16 B/agent in, 8 B/(agent,ray) out).
//
// Kernel design (sm_100a)
//  * wall segments are bulk-copied global -> shared with cp.async.bulk (TMA, completion on
//    an mbarrier), then re-laid out once per CTA as float4 (p2x,p2y,q2x,q2y) + float2
//    (q2-p2), so the inner loop issues one LDS.128 + one LDS.64 per wall (warp broadcast);
//  * a thread owns RPT rays of one agent (16 when the ray count allows: 218 registers, one CTA per
//    SM, 16-way instruction-level parallelism; q1 and q1-p1 in registers); per wall the
//    agent-only determinant o3 is computed once and amortised over the RPT rays; o4
//    reuses the two differences of o2 (fp negation is exact), so a test costs 13 rounded
//    FP32 ops + 3 products + the sign/zero test;
//  * every (ray, wall) pair evaluates all four orientation determinants.  A pair is a
//    *candidate* unless the signs prove a miss: max(o1*o2, o3*o4) > 0 and no zero / NaN
//    among them -- exactly the negation of c1||c2||c3||c4||c5 (space_methods.py:127-140).
//    Candidates (a fraction of a percent) are queued per thread in shared memory and
//    resolved after the scan by the op-for-op restatement ray_wall_exact(), in ascending
//    wall order with a strict '<', which is jnp.argmin's first-minimum rule;
//  * no FMA contraction anywhere (file compiled with -fmad=false, predicates use __f*_rn).
#include <stdlib.h>

#include "common.cuh"
#include "geometry.cuh"
#include "tma.cuh"

namespace fgx {

constexpr int RC_BLOCK = 256;
constexpr int RC_CHUNK = 2048;  // max walls resident in shared memory at a time
constexpr int RC_QCAP = 16;     // queued candidate walls per thread before a flush

struct WallsSoA {
  const float* x1;
  const float* y1;
  const float* x2;
  const float* y2;
};

// Stage walls [base, base+cnt) into shared memory: TMA bulk copy of the four SoA rows into
// `stage` (row stride `chunk`), then one pass re-laying them out as
//   wall4[i] = (p2x, p2y, q2x, q2y)       wdir[i] = (-(q2y-p2y), q2x-p2x)
// the operand pairs of the packed f32x2 inner loop.  Called by all threads.
__device__ __forceinline__ void stage_walls(const WallsSoA& w, int64_t base, int cnt, int chunk, bool tma_ok,
                                            float* stage, float4* wall4, float2* wdir, uint64_t* bar,
                                            uint32_t& parity) {
  const int vec = tma_ok ? (cnt & ~3) : 0;  // part moved by TMA (multiple of 16 bytes)
  if (vec > 0 && threadIdx.x == 0) {
    mbar_expect_tx(bar, 4u * vec * 4u);
    tma_bulk_g2s(stage + 0 * chunk, w.x1 + base, vec * 4u, bar);
    tma_bulk_g2s(stage + 1 * chunk, w.y1 + base, vec * 4u, bar);
    tma_bulk_g2s(stage + 2 * chunk, w.x2 + base, vec * 4u, bar);
    tma_bulk_g2s(stage + 3 * chunk, w.y2 + base, vec * 4u, bar);
  }
  for (int i = vec + threadIdx.x; i < cnt; i += blockDim.x) {  // tail / unaligned fallback
    stage[0 * chunk + i] = __ldg(w.x1 + base + i);
    stage[1 * chunk + i] = __ldg(w.y1 + base + i);
    stage[2 * chunk + i] = __ldg(w.x2 + base + i);
    stage[3 * chunk + i] = __ldg(w.y2 + base + i);
  }
  if (vec > 0) {
    mbar_wait(bar, parity);
    parity ^= 1u;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cnt; i += blockDim.x) {
    const float ax = stage[i], ay = stage[chunk + i], bx = stage[2 * chunk + i], by = stage[3 * chunk + i];
    wall4[i] = make_float4(ax, ay, bx, by);
    wdir[i] = make_float2(-f_sub(by, ay), f_sub(bx, ax));
  }
  __syncthreads();  // also: `stage` is free again (it doubles as the candidate queue)
}

// Per-thread ray state, laid out as the f32x2 operand pairs of the inner loop.
template <int RPT>
struct RaySet {
  float p1x, p1y, L;
  float2 nq1[RPT];  // (-q1x, -q1y): ray end point, negated (negation is exact)
  float2 rr[RPT];   // (q1y-p1y, -(q1x-p1x))
  float best[RPT];
  int hit[RPT];
};

// Resolve queued candidate walls exactly, in ascending wall order.  Inlined (one call site):
// the ray state must stay in registers, so it is never passed by reference to a real call.
template <int RPT>
__device__ __forceinline__ void flush_queue(RaySet<RPT>& r, const uint32_t* q, int qn, const float4* wall4,
                                            int64_t chunk_base) {
  for (int e = 0; e < qn; ++e) {
    const int wl = (int)q[e * blockDim.x];
    const float4 w = wall4[wl];
#pragma unroll
    for (int j = 0; j < RPT; ++j) {
      const float d = ray_wall_exact(r.p1x, r.p1y, -r.nq1[j].x, -r.nq1[j].y, r.L, w.x, w.y, w.z, w.w);
      if (d < r.best[j]) {
        r.best[j] = d;
        r.hit[j] = (int)(chunk_base + wl);
      }
    }
  }
}

// Blackwell packed FP32 (SASS FADD2 / FMUL2): two correctly rounded ops per issue slot
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }

// sure miss <=> one pair has equal non-zero signs and nothing is zero / NaN  (C = A*B)
__device__ __forceinline__ bool sure_miss(float A, float B, float C) {
  return (fmaxf(A, B) > 0.0f) && (fabsf(C) > 0.0f);
}

template <int RPT, int MINB, int BLOCK>
__global__ void __launch_bounds__(BLOCK, MINB)
    raycast_walls_kernel(const float* __restrict__ ax, const float* __restrict__ ay, const float* __restrict__ aang,
                         const float* __restrict__ aactive, int64_t n_agents, int n_rays, float span, float L,
                         WallsSoA walls, int64_t n_walls, int chunk, int tma_ok, float* __restrict__ dist,
                         int32_t* __restrict__ hit) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  float4* wall4 = reinterpret_cast<float4*>(smem_raw + 16);
  float2* wdir = reinterpret_cast<float2*>(wall4 + chunk);
  float* stage = reinterpret_cast<float*>(wdir + chunk);                   // [4][chunk], staging only
  uint32_t* queue = reinterpret_cast<uint32_t*>(stage) + threadIdx.x;      // [RC_QCAP][RC_BLOCK], scan only

  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  uint32_t parity = 0;

  const int groups = (n_rays + RPT - 1) / RPT;  // ray groups per agent
  const int64_t total = n_agents * groups;
  const int64_t num_tiles = (total + BLOCK - 1) / BLOCK;
  const int n_chunks = (int)((n_walls + chunk - 1) / chunk);
  if (n_chunks == 1) stage_walls(walls, 0, (int)n_walls, chunk, tma_ok, stage, wall4, wdir, bar, parity);

  for (int64_t tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
    const int64_t task = tile * BLOCK + threadIdx.x;
    const bool valid = task < total;
    const int64_t agent = valid ? task / groups : 0;
    const int ray0 = valid ? (int)(task - agent * groups) * RPT : 0;
    const bool act = valid && (aactive == nullptr || aactive[agent] != 0.0f);

    RaySet<RPT> r;
    r.L = L;
    r.p1x = valid ? ax[agent] : 0.0f;
    r.p1y = valid ? ay[agent] : 0.0f;
    const float2 np1 = make_float2(-r.p1x, -r.p1y);
    {
      // ray_generator (space_methods.py:56-69)
      const float ang = valid ? aang[agent] : 0.0f;
      const float lo = f_sub(ang, span), hi = f_add(ang, span);
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const int rj = min(ray0 + j, n_rays - 1);
        float s, c;
        sincosf(linspace_elem(lo, hi, n_rays, rj), &s, &c);
        const float q1x = f_add(r.p1x, f_mul(c, L));
        const float q1y = f_add(r.p1y, f_mul(s, L));
        r.nq1[j] = make_float2(-q1x, -q1y);
        r.rr[j] = make_float2(f_sub(q1y, r.p1y), -f_sub(q1x, r.p1x));
        r.best[j] = L;
        r.hit[j] = -1;
      }
    }

    for (int ch = 0; ch < n_chunks; ++ch) {
      const int64_t base = (int64_t)ch * chunk;
      const int cnt = (int)min((int64_t)chunk, n_walls - base);
      if (n_chunks > 1) {
        __syncthreads();  // everyone done with the previous chunk (and its queue)
        stage_walls(walls, base, cnt, chunk, tma_ok, stage, wall4, wdir, bar, parity);
      }
      if (act) {
        int qn = 0;
        float4 wl = wall4[0];
        float2 wd = wdir[0];
        for (int w = 0;; ++w) {
          if (qn == RC_QCAP || (w == cnt && qn)) {  // the only flush site
            flush_queue<RPT>(r, queue, qn, wall4, base);
            qn = 0;
          }
          if (w == cnt) break;
          const float2 P2 = make_float2(wl.x, wl.y), Q2 = make_float2(wl.z, wl.w), WD = wd;
          {  // prefetch the next wall while this one is tested
            const int wn = min(w + 1, cnt - 1);
            wl = wall4[wn];
            wd = wdir[wn];
          }
          // o3 = orientation(p2, q2, p1) = wy*(p1x-q2x) - wx*(p1y-q2y): agent-only, shared by the rays
          const float2 m3 = mul2(WD, add2(Q2, np1));
          const float o3 = f_add(m3.x, m3.y);
          bool cand = false;
          if constexpr (RPT % 2 == 0) {
            const float2 o33 = make_float2(o3, o3);
#pragma unroll
            for (int j = 0; j < RPT; j += 2) {
              float2 o1, o2, o4;
              {
                const float2 ab = add2(P2, r.nq1[j]), uv = add2(Q2, r.nq1[j]);
                const float2 m1 = mul2(r.rr[j], ab), m2 = mul2(r.rr[j], uv), m4 = mul2(WD, uv);
                o1.x = f_add(m1.x, m1.y);  // (q1y-p1y)*(p2x-q1x) - (q1x-p1x)*(p2y-q1y)
                o2.x = f_add(m2.x, m2.y);
                o4.x = f_add(m4.x, m4.y);  // wy*(q1x-q2x) - wx*(q1y-q2y)
              }
              {
                const float2 ab = add2(P2, r.nq1[j + 1]), uv = add2(Q2, r.nq1[j + 1]);
                const float2 m1 = mul2(r.rr[j + 1], ab), m2 = mul2(r.rr[j + 1], uv), m4 = mul2(WD, uv);
                o1.y = f_add(m1.x, m1.y);
                o2.y = f_add(m2.x, m2.y);
                o4.y = f_add(m4.x, m4.y);
              }
              const float2 A = mul2(o1, o2), B = mul2(o33, o4);
              const float2 C = mul2(A, B);
              cand |= !sure_miss(A.x, B.x, C.x);
              cand |= !sure_miss(A.y, B.y, C.y);
            }
          } else {
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
              const float2 ab = add2(P2, r.nq1[j]), uv = add2(Q2, r.nq1[j]);
              const float2 m1 = mul2(r.rr[j], ab), m2 = mul2(r.rr[j], uv), m4 = mul2(WD, uv);
              const float o1 = f_add(m1.x, m1.y), o2 = f_add(m2.x, m2.y), o4 = f_add(m4.x, m4.y);
              const float A = f_mul(o1, o2), B = f_mul(o3, o4);
              cand |= !sure_miss(A, B, f_mul(A, B));
            }
          }
          if (cand) {
            queue[qn * BLOCK] = (uint32_t)w;
            ++qn;
          }
        }
      }
    }

    if (valid) {
#pragma unroll
      for (int j = 0; j < RPT; ++j) {
        const int rj = ray0 + j;
        if (rj < n_rays) {
          const int64_t o = agent * n_rays + rj;
          // inactive agents sense nothing: zeros (sim.py:459-460), hit -1
          dist[o] = act ? r.best[j] : 0.0f;
          if (hit) hit[o] = act ? r.hit[j] : -1;
        }
      }
    }
  }
}

static size_t raycast_smem_bytes(int chunk, int block) {
  const size_t stage = (size_t)chunk * 16, queue = (size_t)RC_QCAP * block * 4;
  return 16 + (size_t)chunk * (16 + 8) + (stage > queue ? stage : queue);
}

template <int RPT, int MINB = 2, int BLOCK = RC_BLOCK>
static int launch_raycast(const float* x, const float* y, const float* ang, const float* active, int64_t n,
                          int n_rays, float span, float L, WallsSoA w, int64_t n_walls, float* dist, int32_t* hit,
                          cudaStream_t s) {
  int chunk = (int)(n_walls < RC_CHUNK ? n_walls : RC_CHUNK);
  chunk = chunk < 4 ? 4 : (chunk + 3) & ~3;  // stage rows stay 16-byte aligned
  const size_t smem = raycast_smem_bytes(chunk, BLOCK);
  auto kern = raycast_walls_kernel<RPT, MINB, BLOCK>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)raycast_smem_bytes(RC_CHUNK, BLOCK));
  if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "fgx_raycast_walls: %s", cudaGetErrorString(e));
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, BLOCK, smem);
  if (e != cudaSuccess || per_sm < 1) per_sm = 1;
  const int groups = (n_rays + RPT - 1) / RPT;
  const int64_t tiles = ceil_div(n * groups, BLOCK);
  const int64_t cap = (int64_t)sm_count() * per_sm;
  const int grid = (int)(tiles < cap ? tiles : cap);
  const uintptr_t al = reinterpret_cast<uintptr_t>(w.x1) | reinterpret_cast<uintptr_t>(w.y1) |
                       reinterpret_cast<uintptr_t>(w.x2) | reinterpret_cast<uintptr_t>(w.y2);
  const int tma_ok = (al & 15) == 0;
  kern<<<grid, BLOCK, smem, s>>>(x, y, ang, active, n, n_rays, span, L, w, n_walls, chunk, tma_ok, dist, hit);
  return check_launch("fgx_raycast_walls");
}

// ---- element-wise mirrors of the jit_* helpers (unit-level parity; not the hot loop) ----
__global__ void __launch_bounds__(256) ray_generator_kernel(const float* __restrict__ ang, int64_t n, int n_rays,
                                                            float span, float* __restrict__ dx,
                                                            float* __restrict__ dy) {
  const int64_t total = n * n_rays;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = t / n_rays;
    const int j = (int)(t - a * n_rays);
    const float g = ang[a];
    float s, c;
    sincosf(linspace_elem(f_sub(g, span), f_add(g, span), n_rays, j), &s, &c);
    dx[t] = c;
    dy[t] = s;
  }
}

__global__ void __launch_bounds__(256) ray_wall_sensor_kernel(const float* __restrict__ ox, const float* __restrict__ oy,
                                                              const float* __restrict__ dx, const float* __restrict__ dy,
                                                              int64_t m, float L, WallsSoA w, int64_t n_walls,
                                                              float* __restrict__ out) {
  const int64_t total = m * n_walls;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / n_walls, k = t - i * n_walls;
    const float px = ox[i], py = oy[i];
    const float qx = f_add(px, f_mul(dx[i], L)), qy = f_add(py, f_mul(dy[i], L));
    out[t] = ray_wall_exact(px, py, qx, qy, L, w.x1[k], w.y1[k], w.x2[k], w.y2[k]);
  }
}

__global__ void __launch_bounds__(256) ray_agent_sensor_kernel(const float* __restrict__ ox, const float* __restrict__ oy,
                                                               const float* __restrict__ dx, const float* __restrict__ dy,
                                                               int64_t m, float L, const float* __restrict__ ex,
                                                               const float* __restrict__ ey, const float* __restrict__ er,
                                                               const float* __restrict__ eact, int64_t n_ent,
                                                               float* __restrict__ out) {
  const int64_t total = m * n_ent;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / n_ent, k = t - i * n_ent;
    out[t] = ray_agent_exact(ox[i], oy[i], dx[i], dy[i], L, ex[k], ey[k], er[k], eact[k]);
  }
}

// check_collision (space_methods.py:10-43): one warp per move, lanes stride the walls, ballot = any()
__global__ void __launch_bounds__(256) check_collision_kernel(WallsSoA w, int64_t n_walls, const float* __restrict__ oldx,
                                                              const float* __restrict__ oldy,
                                                              const float* __restrict__ newx,
                                                              const float* __restrict__ newy, int64_t m,
                                                              uint8_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp; i < m; i += nwarps) {
    const float px = oldx[i], py = oldy[i], qx = newx[i], qy = newy[i];
    bool any = false;
    for (int64_t k = lane; k < n_walls; k += 32)
      any |= segments_intersect(w.x1[k], w.y1[k], w.x2[k], w.y2[k], px, py, qx, qy);
    const unsigned b = __ballot_sync(0xffffffffu, any);
    if (lane == 0) out[i] = b ? 1 : 0;
  }
}

}  // namespace fgx

extern "C" {

int fgx_raycast_walls(const float* x, const float* y, const float* ang, const float* active, int64_t n_agents,
                      int32_t n_rays, float ray_span, float ray_length, const float* wx1, const float* wy1,
                      const float* wx2, const float* wy2, int64_t n_walls, float* dist, int32_t* hit,
                      fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n_agents >= 0 && n_walls >= 0, "fgx_raycast_walls: negative size");
  FGX_REQUIRE(n_rays > 0, "fgx_raycast_walls: n_rays must be > 0");
  FGX_REQUIRE(n_walls < (1ll << 31), "fgx_raycast_walls: too many walls");
  if (n_agents == 0) return FGX_OK;
  FGX_REQUIRE(x && y && ang && dist, "fgx_raycast_walls: null pointer");
  FGX_REQUIRE(n_walls == 0 || (wx1 && wy1 && wx2 && wy2), "fgx_raycast_walls: null wall array");
  WallsSoA w{wx1, wy1, wx2, wy2};
  cudaStream_t s = as_stream(stream);
  // rays per thread: the largest group size that divides the ray count well
  if (const char* v = getenv("FGX_RC_VARIANT")) {  // tuning hook (profiles/): RPT x min CTAs per SM
    switch (atoi(v)) {
      case 1: return launch_raycast<8, 3>(x, y, ang, active, n_agents, n_rays, ray_span, ray_length, w, n_walls, dist, hit, s);
      case 2: return launch_raycast<4, 3>(x, y, ang, active, n_agents, n_rays, ray_span, ray_length, w, n_walls, dist, hit, s);
      case 3: return launch_raycast<4, 4>(x, y, ang, active, n_agents, n_rays, ray_span, ray_length, w, n_walls, dist, hit, s);
      case 4: return launch_raycast<4, 5>(x, y, ang, active, n_agents, n_rays, ray_span, ray_length, w, n_walls, dist, hit, s);
      case 5: return launch_raycast<2, 6>(x, y, ang, active, n_agents, n_rays, ray_span, ray_length, w, n_walls, dist, hit, s);
      case 6: return launch_raycast<4, 2>(x, y, ang, active, n_agents, n_rays, ray_span, ray_length, w, n_walls, dist, hit, s);
      case 7: return launch_raycast<8, 2>(x, y, ang, active, n_agents, n_rays, ray_span, ray_length, w, n_walls, dist, hit, s);
      default: break;
    }
  }
  if (n_rays % 16 == 0) return launch_raycast<16, 1>(x, y, ang, active, n_agents, n_rays, ray_span, ray_length, w, n_walls, dist, hit, s);
  if (n_rays % 8 == 0) return launch_raycast<8>(x, y, ang, active, n_agents, n_rays, ray_span, ray_length, w, n_walls, dist, hit, s);
  if (n_rays % 4 == 0) return launch_raycast<4>(x, y, ang, active, n_agents, n_rays, ray_span, ray_length, w, n_walls, dist, hit, s);
  if (n_rays % 3 == 0) return launch_raycast<3>(x, y, ang, active, n_agents, n_rays, ray_span, ray_length, w, n_walls, dist, hit, s);
  if (n_rays % 2 == 0) return launch_raycast<2>(x, y, ang, active, n_agents, n_rays, ray_span, ray_length, w, n_walls, dist, hit, s);
  if (n_rays >= 5) return launch_raycast<4>(x, y, ang, active, n_agents, n_rays, ray_span, ray_length, w, n_walls, dist, hit, s);
  return launch_raycast<1>(x, y, ang, active, n_agents, n_rays, ray_span, ray_length, w, n_walls, dist, hit, s);
}

int fgx_ray_generator(const float* ang, int64_t n, int32_t n_rays, float ray_span, float* dir_x, float* dir_y,
                      fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n >= 0 && n_rays > 0, "fgx_ray_generator: bad sizes");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(ang && dir_x && dir_y, "fgx_ray_generator: null pointer");
  ray_generator_kernel<<<grid_for(n * n_rays, 256, 8), 256, 0, as_stream(stream)>>>(ang, n, n_rays, ray_span, dir_x,
                                                                                  dir_y);
  return check_launch("fgx_ray_generator");
}

int fgx_ray_wall_sensor(const float* ox, const float* oy, const float* dx, const float* dy, int64_t n_rays_total,
                        float ray_length, const float* wx1, const float* wy1, const float* wx2, const float* wy2,
                        int64_t n_walls, float* out, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n_rays_total >= 0 && n_walls >= 0, "fgx_ray_wall_sensor: negative size");
  if (n_rays_total == 0 || n_walls == 0) return FGX_OK;
  FGX_REQUIRE(ox && oy && dx && dy && wx1 && wy1 && wx2 && wy2 && out, "fgx_ray_wall_sensor: null pointer");
  WallsSoA w{wx1, wy1, wx2, wy2};
  ray_wall_sensor_kernel<<<grid_for(n_rays_total * n_walls, 256, 8), 256, 0, as_stream(stream)>>>(
      ox, oy, dx, dy, n_rays_total, ray_length, w, n_walls, out);
  return check_launch("fgx_ray_wall_sensor");
}

int fgx_ray_agent_sensor(const float* ox, const float* oy, const float* dx, const float* dy, int64_t n_rays_total,
                         float ray_length, const float* ex, const float* ey, const float* erad, const float* eactive,
                         int64_t n_entities, float* out, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n_rays_total >= 0 && n_entities >= 0, "fgx_ray_agent_sensor: negative size");
  if (n_rays_total == 0 || n_entities == 0) return FGX_OK;
  FGX_REQUIRE(ox && oy && dx && dy && ex && ey && erad && eactive && out, "fgx_ray_agent_sensor: null pointer");
  ray_agent_sensor_kernel<<<grid_for(n_rays_total * n_entities, 256, 8), 256, 0, as_stream(stream)>>>(
      ox, oy, dx, dy, n_rays_total, ray_length, ex, ey, erad, eactive, n_entities, out);
  return check_launch("fgx_ray_agent_sensor");
}

int fgx_check_collision(const float* wx1, const float* wy1, const float* wx2, const float* wy2, int64_t n_walls,
                        const float* old_x, const float* old_y, const float* new_x, const float* new_y, int64_t n,
                        uint8_t* out, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n >= 0 && n_walls >= 0, "fgx_check_collision: negative size");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(old_x && old_y && new_x && new_y && out, "fgx_check_collision: null pointer");
  FGX_REQUIRE(n_walls == 0 || (wx1 && wy1 && wx2 && wy2), "fgx_check_collision: null wall array");
  WallsSoA w{wx1, wy1, wx2, wy2};
  check_collision_kernel<<<grid_for(n * 32, 256, 8), 256, 0, as_stream(stream)>>>(w, n_walls, old_x, old_y, new_x,
                                                                                 new_y, n, out);
  return check_launch("fgx_check_collision");
}
}
