// K2 for the foraging example (examples/many_agent_systems/foraging_agents/EcoEvoJax/sim.py):
// the fused Forager step (WilsonCowan policy + point-mass physics + energy bookkeeping), the
// Resource step, the all-pairs eating interaction and the ray-sensor composition.
//
// forager_step is HBM-bound: every agent owns its RNN weights (J n*n, E n*obs, D act*n, tau,
// B, Z), so a step streams ~12 KB per active agent (n=40, obs=28) and does ~6 kFLOP on it --
// a batched GEMV, no tensor-core shape.  One warp per agent: lane 0 pulls the agent's
// weight rows global -> shared with cp.async.bulk (TMA, mbarrier completion; every byte is
// read exactly once, fully coalesced), then lane r owns output row r and walks it with a
// rotated column order so that the 32 lanes hit 32 different banks.  16 warps per SM keep
// ~190 KB of weight rows in flight per SM.
#include <stdlib.h>

#include "common.cuh"
#include "geometry.cuh"
#include "threefry.cuh"
#include "tma.cuh"
#include "walls_grid.cuh"

namespace fgx {

constexpr int FS_WARPS = 8;  // warps per CTA of the forager step

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }
__device__ __forceinline__ int align4(int x) { return (x + 3) & ~3; }

struct WcLayout {  // per-warp shared-memory slots (in floats), each 16-byte aligned
  int J, E, D, tau, B, Z, tz, tnz, obs, total;
};
__host__ __device__ inline WcLayout wc_layout(int n, int nobs, int nact) {
  WcLayout l;
  int o = 0;
  auto take = [&](int cnt) { int r = o; o += (cnt + 3) & ~3; return r; };
  l.J = take(n * n);
  l.E = take(n * nobs);
  l.D = take(nact * n);
  l.tau = take(n);
  l.B = take(n);
  l.Z = take(n);
  l.tz = take(n);
  l.tnz = take(n);
  l.obs = take(nobs);
  l.total = o;
  return l;
}

// Dot products of TWO weight rows (lane-rows r and r+32, the second optional) with the shared
// vector `v`, all in shared memory.  The 32 lanes walk 32 different rows at once, so the column
// order is chosen to keep them on different banks:
//   len % 4 == 0 : 16-byte chunks (LDS.128), chunk order rotated by r  (pitch 16*k bytes)
//   len odd      : plain order -- the odd pitch already spreads the banks, v[j] is a broadcast
//   otherwise    : scalar walk rotated by r
// Two rows per lane and four accumulators per row give the FMA chain enough independent work to
// hide the LDS latency without relying on other warps.
template <int LEN>
__device__ __forceinline__ void row_dot2(const float* rowA, const float* rowB, const float* v, int len_rt, int r,
                                         float& outA, float& outB) {
  const int len = LEN ? LEN : len_rt;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f, b0 = 0.f, b1 = 0.f, b2 = 0.f, b3 = 0.f;
  if ((len & 3) == 0) {
    const int nch = len >> 2;
    const float4* A4 = reinterpret_cast<const float4*>(rowA);
    const float4* B4 = reinterpret_cast<const float4*>(rowB);
    const float4* v4 = reinterpret_cast<const float4*>(v);
    int c = r % nch;
#pragma unroll(LEN ? (LEN / 4 < 16 ? LEN / 4 : 4) : 2)
    for (int t = 0; t < nch; ++t) {
      const float4 x = v4[c], wa = A4[c], wb = B4[c];
      a0 += wa.x * x.x; a1 += wa.y * x.y; a2 += wa.z * x.z; a3 += wa.w * x.w;
      b0 += wb.x * x.x; b1 += wb.y * x.y; b2 += wb.z * x.z; b3 += wb.w * x.w;
      c = c + 1 == nch ? 0 : c + 1;
    }
  } else if (len & 1) {
    int j = 0;
    for (; j + 4 <= len; j += 4) {
      // v is 16-byte aligned (every slot of WcLayout is) and the same for all lanes: one broadcast LDS.128 per four
      // columns; the odd row pitch keeps the rows on scalar loads.  Same products in the same order as before.
      const float4 x = *reinterpret_cast<const float4*>(v + j);
      const float x0 = x.x, x1 = x.y, x2 = x.z, x3 = x.w;
      a0 += rowA[j] * x0; a1 += rowA[j + 1] * x1; a2 += rowA[j + 2] * x2; a3 += rowA[j + 3] * x3;
      b0 += rowB[j] * x0; b1 += rowB[j + 1] * x1; b2 += rowB[j + 2] * x2; b3 += rowB[j + 3] * x3;
    }
    for (; j < len; ++j) {
      a0 += rowA[j] * v[j];
      b0 += rowB[j] * v[j];
    }
  } else {
    int j = r % len;
    for (int t = 0; t < len; ++t) {
      a0 += rowA[j] * v[j];
      b0 += rowB[j] * v[j];
      j = j + 1 == len ? 0 : j + 1;
    }
  }
  outA = (a0 + a1) + (a2 + a3);
  outB = (b0 + b1) + (b2 + b3);
}

// one leaf row of one agent -> shared memory: TMA bulk copy when 16-byte clean, else lanes copy
__device__ __forceinline__ uint32_t leaf_tma_bytes(const float* base, int row_floats) {
  return ((reinterpret_cast<uintptr_t>(base) & 15) == 0 && (row_floats & 3) == 0) ? (uint32_t)row_floats * 4u : 0u;
}

// Iterator over the ACTIVE agents of this warp's sequence a0, a0+GW, a0+2GW, ...: one coalesced
// probe of 32 candidates (ballot) serves up to 32 agents, so inactive runs cost nothing per slot.
// ROUNDS: also track `round`, the index of the returned agent in the warp's sequence (the fused kernel's step warps
// wait for the ray warps by round)
template <bool ROUNDS>
struct ActiveScan {
  int64_t base, probe_base;
  unsigned mask;
  uint32_t next_round = 0u, probe_round = 0u, round = 0u;
  __device__ __forceinline__ int64_t next(const float* __restrict__ active, int64_t GW, int64_t n_agents, int lane) {
    while (mask == 0u) {
      if (base >= n_agents) return n_agents;
      const int64_t mine = base + lane * GW;
      mask = __ballot_sync(0xffffffffu, mine < n_agents && active[mine] != 0.0f);
      probe_base = base;
      base += 32 * GW;
      if (ROUNDS) {
        probe_round = next_round;
        next_round += 32u;
      }
    }
    const int b = __ffs(mask) - 1;
    mask &= mask - 1u;
    if (ROUNDS) round = probe_round + (uint32_t)b;
    return probe_base + (int64_t)b * GW;
  }
};

// Fused sensing (the sensor feeds the policy directly): when enabled, the warp that owns an
// agent casts the agent's rays against the wall grid right after issuing the copies of its weights -- in the
// shadow of their HBM latency, where the step kernel has idle issue slots -- writes the distances into the
// agent's obs slot in shared memory (no [N, R] round trip for the policy) and to the dist / hit outputs.
struct SenseArgs {
  int enabled;
  WallsGridView v;
  int n_rays;
  float span, L;
  float* dist;
  int32_t* hit;
};

// ---- warp-specialised fused sensing (RW > 0 extra warps per CTA; 32 rays) ------------------------------------------
// The step warps of a CTA consume, in round j, the 8 consecutive agents first + j * GW + (0..7) (first = 8 * blockIdx,
// GW = 8 * gridDim).  The CTA's RW ray warps produce exactly those tiles ahead of them -- warp r the rounds r, r + RW,
// ... -- with the warp-cooperative walk of walls_grid.cuh, write dist / hit to global memory and publish the number of
// the last finished round in shared memory (fence.cta + volatile flag).  A step warp issues the weight copies of its
// agent, then waits for the agent's round and reads its 32 distances back as the observation.  The ray warps never
// wait for the step warps, so there is no cycle; the step kernel is HBM-bound and leaves > 80 % of the issue slots idle,
// which is what the (issue-bound) ray cast runs in.  Positions are updated in place by the step warps: an agent's pose
// is read by its ray warp before the flag of its round is published, i.e. before its step can start.
struct RaySmem {  // per ray warp: two parked tiles (a tile = the 8 agents of one round) + the survivor area
  WgwTileSmem<FS_WARPS> tile[2];
  WgwSurvSmem surv;
};

template <int RW>
__device__ __forceinline__ void ray_warps_role(const SenseArgs& sense, const fgx_forager_state_t& st,
                                               const float* __restrict__ active, uint32_t n, RaySmem* wsm,
                                               volatile uint32_t* done, uint32_t r, uint32_t lane) {
  // two parked tiles per warp: while the agents of one are cast, the next one lands in the other, so that the
  // first agent of every tile can be prefetched like the others
  WgwTileSmem<FS_WARPS>* sm2 = wsm[r].tile;
  WgwSurvSmem& surv = wsm[r].surv;
  const WallsGridView& v = sense.v;
  const bool fallback = v.st->fallback != 0;
  float sj, omsj;
  wgw_lane_weights(lane, sj, omsj);
  const uint32_t GW = gridDim.x * FS_WARPS, first = blockIdx.x * FS_WARPS;
  const uint32_t rounds = first < n ? (n - first + GW - 1u) / GW : 0u;
  uint32_t j = r;
  if (j >= rounds) return;
  WgwTile nx;
  {
    const uint32_t base = first + j * GW;
    wgw_load_pose<true>(st.X, st.Y, st.ang, active, min(n, base + FS_WARPS), base + lane, nx);
    wgw_load_rec(v, fallback, nx);
    wgw_park(sm2[0], lane, nx, FS_WARPS);
    __syncwarp();
  }
  uint32_t buf = 0;
  WgwPre cur, nxt;
  wgw_prefetch(v, sm2[0], 0, lane, cur);
  for (; j < rounds; j += RW) {
    const uint32_t base = first + j * GW;
    const uint32_t cnt = min((uint32_t)FS_WARPS, n - base);
    const uint32_t jn = j + RW;
    const uint32_t base_n = first + jn * GW;
    const bool more = jn < rounds;
    if (more) wgw_load_pose<true>(st.X, st.Y, st.ang, active, min(n, base_n + FS_WARPS), base_n + lane, nx);
    // (an L2 bulk prefetch of the next agent's J / E rows by the step warps, cp.async.bulk.prefetch.L2 a turn ahead, was
    // measured here: fused 2.21 -> 2.25 ms, plain step kernel 2.00 -> 2.11 ms -- rejected)
#pragma unroll 1  // one copy of the per-agent code: eight would push the step warps' loop out of the instruction cache
    for (uint32_t k = 0; k < cnt; ++k) {
      // next tile: cell records requested a quarter into this tile, parked three quarters in, so that ...
      if (more && k == (cnt >> 2)) wgw_load_rec(v, fallback, nx);
      if (more && k == cnt - 1u - (cnt >> 2)) {
        wgw_park(sm2[buf ^ 1u], lane, nx, FS_WARPS);
        __syncwarp();
      }
      // ... the loads of the NEXT agent (this tile's, or the first of the next tile) fly during this agent's cast
      if (k + 1u < cnt) wgw_prefetch(v, sm2[buf], k + 1u, lane, nxt);
      else if (more) wgw_prefetch(v, sm2[buf ^ 1u], 0, lane, nxt);
      const uint32_t o = (base + k) * 32u + lane;
      float best;
      int bidx;
      wgw_cast_agent(v, surv, lane, sense.L, sense.span, sj, omsj, cur, best, bidx);
      sense.dist[o] = best;
      if (sense.hit) sense.hit[o] = bidx;
      cur = nxt;
    }
    __threadfence_block();
    __syncwarp();
    if (lane == 0) done[r] = j + 1u;
    buf ^= 1u;
  }
}

// NN / NOBS / NACT > 0 fix the policy shape at compile time (constant trip counts, immediate
// shared-memory offsets, no shape branches); 0 = read it from the params (any shape).
// RW > 0: RW more warps per CTA cast the rays (ray_warps_role above); two CTAs per SM then cap the registers.
// RW == -1: the warp that steps an agent casts its rays itself (the older fused mode).  RW == 0: no sensing code at all.
template <int NN, int NOBS, int NACT, int RW = 0>
__global__ void __launch_bounds__((FS_WARPS + (RW > 0 ? RW : 2)) * 32, 2)
    forager_step_kernel(const fgx_forager_params_t p, int64_t n_agents, const float* __restrict__ active,
                        const float* __restrict__ obs_in, const float* __restrict__ energy_in,
                        const fgx_forager_state_t st, const fgx_wc_policy_t pol, float* __restrict__ age,
                        float* __restrict__ actions_out, int slots_flags, const SenseArgs sense) {
  // actions_out != nullptr: policy-only mode (WilsonCowan.step_policy on its own): obs has all n_obs
  // columns, actions are written out and the Forager physics is skipped.
  // slots == 2: each warp owns two weight slots -- while it computes agent i from one, the TMA copies
  // of its next active agent land in the other (mbarrier per slot).  slots == 1: one slot per warp and
  // twice the warps per SM; the copies of the other warps overlap this warp's compute.
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int slots = slots_flags & 0xff;
  const bool stream_weights = (slots_flags >> 8) & 1;  // weights are read once per launch: L2 evict-first
  const uint64_t l2pol = l2_policy_evict_first();
  const int n = NN ? NN : p.n_neurons, nobs = NOBS ? NOBS : p.n_obs, nact = NACT ? NACT : p.n_act;
  const WcLayout L = wc_layout(n, nobs, nact);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwarps = RW ? FS_WARPS : (int)(blockDim.x >> 5);  // step warps
  volatile uint32_t* ray_done = nullptr;
  if constexpr (RW > 0) {
    // tail of the dynamic shared memory: RW parked-tile areas + the per-ray-warp round counters
    unsigned char* tail = smem_raw + 16 * FS_WARPS + (size_t)FS_WARPS * slots * L.total * sizeof(float);
    RaySmem* wsm = reinterpret_cast<RaySmem*>(tail);
    ray_done = reinterpret_cast<volatile uint32_t*>(tail + RW * sizeof(RaySmem));
    if (threadIdx.x < RW) ray_done[threadIdx.x] = 0u;
    __syncthreads();
    if (warp >= FS_WARPS) {
      ray_warps_role<RW>(sense, st, active, (uint32_t)n_agents, wsm, ray_done, (uint32_t)(warp - FS_WARPS), (uint32_t)lane);
      return;
    }
  }
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw) + 2 * warp;
  float* slot0 = reinterpret_cast<float*>(smem_raw + 16 * FS_WARPS) + (size_t)warp * slots * L.total;
  if (lane == 0) { mbar_init(bars, 1); mbar_init(bars + 1, 1); }
  __syncwarp();
  uint32_t parity = 0u;  // bit s = phase of slot s's mbarrier (a register: an array indexed by s lived in local memory)

  const int obs_cols = actions_out ? nobs : nobs - 1;
  const uint32_t bJ = leaf_tma_bytes(pol.J, n * n), bE = leaf_tma_bytes(pol.E, n * nobs),
                 bD = leaf_tma_bytes(pol.D, nact * n), bT = leaf_tma_bytes(pol.tau, n),
                 bB = leaf_tma_bytes(pol.B, n), bZ = leaf_tma_bytes(pol.Z, n),
                 bO = RW != 0 ? 0u : leaf_tma_bytes(obs_in, obs_cols);
  const uint32_t tx = bJ + bE + bD + bT + bB + bZ + bO;

  // lane l < 11 owns state leaf l of an agent, lane 11 its age, lane 12 its energy_in: the scalars are
  // fetched when the agent's copies are issued and shared by shuffles when it is computed
  float* const* leaves = reinterpret_cast<float* const*>(&st);
  const float* my_leaf = actions_out ? nullptr : (lane < 11 ? leaves[lane] : (lane == 11 ? age : (lane == 12 ? energy_in : nullptr)));

  auto issue = [&](int64_t a, int s) -> float {
    float* sm = slot0 + (size_t)s * L.total;
    uint64_t* bar = bars + s;
    const float* gJ = pol.J + a * n * n;
    const float* gE = pol.E + a * n * nobs;
    const float* gD = pol.D + a * nact * n;
    const float* gT = pol.tau + a * n;
    const float* gB = pol.B + a * n;
    const float* gZ = pol.Z + a * n;
    const float* gO = obs_in + a * obs_cols;
    if (lane == 0 && tx) {
      mbar_expect_tx(bar, tx);
      if (stream_weights) {
        if (bJ) tma_bulk_g2s_hint(sm + L.J, gJ, bJ, bar, l2pol);
        if (bE) tma_bulk_g2s_hint(sm + L.E, gE, bE, bar, l2pol);
        if (bD) tma_bulk_g2s_hint(sm + L.D, gD, bD, bar, l2pol);
        if (bT) tma_bulk_g2s_hint(sm + L.tau, gT, bT, bar, l2pol);
        if (bB) tma_bulk_g2s_hint(sm + L.B, gB, bB, bar, l2pol);
      } else {
        if (bJ) tma_bulk_g2s(sm + L.J, gJ, bJ, bar);
        if (bE) tma_bulk_g2s(sm + L.E, gE, bE, bar);
        if (bD) tma_bulk_g2s(sm + L.D, gD, bD, bar);
        if (bT) tma_bulk_g2s(sm + L.tau, gT, bT, bar);
        if (bB) tma_bulk_g2s(sm + L.B, gB, bB, bar);
      }
      if (bZ) tma_bulk_g2s(sm + L.Z, gZ, bZ, bar);
      if (bO) tma_bulk_g2s(sm + L.obs, gO, bO, bar);
    }
    if (!bJ) for (int i = lane; i < n * n; i += 32) sm[L.J + i] = __ldg(gJ + i);
    if (!bE) for (int i = lane; i < n * nobs; i += 32) sm[L.E + i] = __ldg(gE + i);
    if (!bD) for (int i = lane; i < nact * n; i += 32) sm[L.D + i] = __ldg(gD + i);
    if (!bT) for (int i = lane; i < n; i += 32) sm[L.tau + i] = __ldg(gT + i);
    if (!bB) for (int i = lane; i < n; i += 32) sm[L.B + i] = __ldg(gB + i);
    if (!bZ) for (int i = lane; i < n; i += 32) sm[L.Z + i] = gZ[i];
    if (RW == 0 && !bO) for (int k = lane; k < obs_cols; k += 32) sm[L.obs + k] = __ldg(gO + k);
    return my_leaf ? my_leaf[a] : 0.0f;
  };
  // ray_generator + ray_wall_sensor + min / argmin for agent g (pose in the scalars `svv` just fetched), one lane
  // per ray, while the agent's weight copies are in flight; results go to obs slot s and to dist / hit
  auto cast_rays = [&](int64_t g, float svv, int s) {
    float* sm = slot0 + (size_t)s * L.total;
    const float px = __shfl_sync(0xffffffffu, svv, 0), py = __shfl_sync(0xffffffffu, svv, 3);
    const float ang = __shfl_sync(0xffffffffu, svv, 6);
    const float lo = f_sub(ang, sense.span), hi = f_add(ang, sense.span);
    for (int j = lane; j < sense.n_rays; j += 32) {
      float sn, cs;
      sincosf(linspace_elem(lo, hi, sense.n_rays, j), &sn, &cs);
      const float q1x = f_add(px, f_mul(cs, sense.L)), q1y = f_add(py, f_mul(sn, sense.L));
      float best = sense.L;
      int bidx = -1;
      walls_grid_cast(sense.v, px, py, q1x, q1y, sense.L, best, bidx);
      sm[L.obs + j] = best;
      sense.dist[g * sense.n_rays + j] = best;
      if (sense.hit) sense.hit[g * sense.n_rays + j] = bidx;
    }
  };

  const int64_t gw = (int64_t)blockIdx.x * nwarps + warp, GW = (int64_t)gridDim.x * nwarps;
  ActiveScan<(RW > 0)> scan{gw, gw, 0u};
  int64_t a = scan.next(active, GW, n_agents, lane);
  int s = 0;
  float sv = 0.0f;
  uint32_t round = scan.round;
  [[maybe_unused]] float obs_cur = 0.0f;
  [[maybe_unused]] bool have_cur = false;
  if (a < n_agents) {
    sv = issue(a, 0);
    if constexpr (RW == -1) cast_rays(a, sv, 0);
  }
  while (a < n_agents) {
    const int64_t a_next = scan.next(active, GW, n_agents, lane);
    const uint32_t round_next = scan.round;
    float sv_next = 0.0f;
    // the NEXT agent's observation, requested a whole turn ahead when its rays are already done (the usual case: the
    // ray warps run ahead), so that the L2 round trip of the 128-byte row is off the warp's critical path
    [[maybe_unused]] float obs_ahead = 0.0f;
    [[maybe_unused]] bool have_ahead = false;
    if constexpr (RW > 0) {
      if (a_next < n_agents && ray_done[round_next % RW] > round_next) {  // warp-uniform: one shared-memory word
        __threadfence_block();
        obs_ahead = *reinterpret_cast<const volatile float*>(sense.dist + a_next * 32 + lane);
        have_ahead = true;
      }
    }
    if (slots == 2 && a_next < n_agents) {  // slot s^1 was released by the last __syncwarp
      sv_next = issue(a_next, s ^ 1);
      if constexpr (RW == -1) cast_rays(a_next, sv_next, s ^ 1);
    }
    float* sm = slot0 + (size_t)s * L.total;
    float* gZ = pol.Z + a * n;
    const float energy = __shfl_sync(0xffffffffu, sv, 7);
    __syncwarp();
    if constexpr (RW > 0) {
      // the agent's rays: its 32 distances are the observation -- fetched a turn ahead (obs_cur), else wait for its
      // round now
      if (!have_cur) {
        const uint32_t need = round + 1u;
        volatile uint32_t* d = ray_done + (round % RW);
        while (*d < need) __nanosleep(256);
        __threadfence_block();
        obs_cur = *reinterpret_cast<const volatile float*>(sense.dist + a * 32 + lane);
      }
      sm[L.obs + lane] = obs_cur;
    }
    if (tx) {
      mbar_wait(bars + s, (parity >> s) & 1u);
      parity ^= 1u << s;
    }
    if (!actions_out && lane == 0) sm[L.obs + nobs - 1] = energy;  // obs = concat(sensor obs, own energy)
    // WilsonCowan.step_policy (wilson_cowan.py:34-51)
    // The per-agent arithmetic is a chain of short dependent sequences (tanh, sigmoid, shuffle reductions) and a warp
    // issues one instruction per ~9 cycles through it, so independent pieces are written side by side: the two tanh
    // of a lane's neurons, the two rows' epilogues, the two actions' reductions.  Same operations per value, same bits.
    for (int j = lane; j < n; j += 64) {
      const bool two = j + 32 < n;
      const float z0 = sm[L.Z + j], z1 = sm[L.Z + (two ? j + 32 : j)];
      const float t0 = tanhf(z0), t1 = tanhf(z1);
      sm[L.tz + j] = t0;
      if (two) sm[L.tz + j + 32] = t1;
    }
    __syncwarp();
    for (int rA = lane; rA < n; rA += 64) {  // lane-rows rA and rA+32 together
      const bool hasB = rA + 32 < n;
      const int rB = hasB ? rA + 32 : rA;
      float sJ[2], sE[2];
      row_dot2<NN>(sm + L.J + rA * n, sm + L.J + rB * n, sm + L.tz, n, rA, sJ[0], sJ[1]);
      row_dot2<NOBS>(sm + L.E + rA * nobs, sm + L.E + rB * nobs, sm + L.obs, nobs, rA, sE[0], sE[1]);
      // both rows' epilogues as straight-line code (a lane without a second row repeats its first and does not store)
      const float zA = sm[L.Z + rA], zB = sm[L.Z + rB];
      const float gA = 10.0f * sigmoidf_(sm[L.tau + rA]), gB = 10.0f * sigmoidf_(sm[L.tau + rB]);
      float dzA = ((sJ[0] + sE[0]) + sm[L.B + rA]) - zA, dzB = ((sJ[1] + sE[1]) + sm[L.B + rB]) - zB;
      dzA = dzA * gA;
      dzB = dzB * gB;
      const float nzA = zA + p.policy_dt * dzA, nzB = zB + p.policy_dt * dzB;
      const float tA = tanhf(nzA), tB = tanhf(nzB);
      gZ[rA] = nzA;
      sm[L.tnz + rA] = tA;
      if (hasB) {
        gZ[rB] = nzB;
        sm[L.tnz + rB] = tB;
      }
    }
    __syncwarp();
    float act0 = 0.0f, act1 = 0.0f;  // the Forager uses actions[0], actions[1] (sim.py:86-87)
    {
      // actions 0 and 1 together (nact >= 2 is checked by the launchers of the Forager step; the policy-only entry
      // point may have a single action: its second sum then repeats the first)
      const int k1 = nact > 1 ? 1 : 0;
      float d0 = 0.0f, d1 = 0.0f;
      for (int j = lane; j < n; j += 32) {
        const float t = sm[L.tnz + j];
        d0 += sm[L.D + j] * t;
        d1 += sm[L.D + k1 * n + j] * t;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        d0 += __shfl_xor_sync(0xffffffffu, d0, o);
        d1 += __shfl_xor_sync(0xffffffffu, d1, o);
      }
      act0 = p.action_scale * tanhf(d0);
      act1 = p.action_scale * tanhf(d1);
      if (actions_out && lane == 0) {
        actions_out[a * nact] = act0;
        if (nact > 1) actions_out[a * nact + 1] = act1;
      }
    }
    for (int k = 2; k < nact; ++k) {  // further actions (policy-only mode with nact > 2)
      float d = 0.0f;
      for (int j = lane; j < n; j += 32) d += sm[L.D + k * n + j] * sm[L.tnz + j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
      const float v = p.action_scale * tanhf(d);
      if (actions_out && lane == 0) actions_out[a * nact + k] = v;
    }
    const float X = __shfl_sync(0xffffffffu, sv, 0), Xv = __shfl_sync(0xffffffffu, sv, 1);
    const float Y = __shfl_sync(0xffffffffu, sv, 3), Yv = __shfl_sync(0xffffffffu, sv, 4);
    const float tar = __shfl_sync(0xffffffffu, sv, 9), tbd = __shfl_sync(0xffffffffu, sv, 10);
    const float age0 = __shfl_sync(0xffffffffu, sv, 11), e_in = __shfl_sync(0xffffffffu, sv, 12);
    if (!actions_out) {
      // Forager.step_agent physics (sim.py:89-128): every lane evaluates the same scalars (no divergence), then
      // lane l < 12 stores the leaf it owns (my_leaf: the 11 state leaves and age) -- one store instruction
      // instead of thirteen single-lane ones
      const float dt = p.dt;
      float Xn = f_add(X, f_mul(Xv, dt)), Yn = f_add(Y, f_mul(Yv, dt));
      float Xvn = f_add(Xv, f_mul(act0, dt)), Yvn = f_add(Yv, f_mul(act1, dt));
      if (Xn > p.x_max || Xn < p.x_min) { Xn = X; Xvn = -Xv; }
      if (Yn > p.y_max || Yn < p.y_min) { Yn = Y; Yvn = -Yv; }
      float en = f_sub(f_add(energy, e_in), f_mul(p.energy_cost_dt, f_add(fabsf(act0), fabsf(act1))));
      en = fminf(en, p.energy_max);
      if (p.clip_min) en = fmaxf(en, p.energy_min);
      // leaf order of fgx_forager_state_t: X X_vel X_acc Y Y_vel Y_acc ang energy rad t_above t_below | age
      float val = Xn;
      val = lane == 1 ? Xvn : val;
      val = lane == 2 ? act0 : val;
      val = lane == 3 ? Yn : val;
      val = lane == 4 ? Yvn : val;
      val = lane == 5 ? act1 : val;
      val = lane == 6 ? atan2f(Yvn, Xvn) : val;
      val = (lane == 7 || lane == 8) ? en : val;
      val = lane == 9 ? (en > p.repr_thresh ? f_add(tar, dt) : 0.0f) : val;
      val = lane == 10 ? (en < p.die_thresh ? f_add(tbd, dt) : 0.0f) : val;
      val = lane == 11 ? f_add(age0, dt) : val;
      if (lane < 12) const_cast<float*>(my_leaf)[a] = val;
    }
    __syncwarp();  // all lanes are done with slot s before the copies of a later agent land in it
    if (slots == 1) {
      if (a_next < n_agents) {
        sv_next = issue(a_next, 0);
        if constexpr (RW == -1) cast_rays(a_next, sv_next, 0);
      }
    } else {
      s ^= 1;
    }
    a = a_next;
    sv = sv_next;
    round = round_next;
    obs_cur = obs_ahead;
    have_cur = have_ahead;
  }
}

// Resource.step_agent under vmap (sim.py:264-303)
__global__ void __launch_bounds__(256)
    resource_step_kernel(const fgx_resource_params_t p, int64_t n, const float* __restrict__ active,
                         const float* __restrict__ energy_out, const fgx_resource_state_t st,
                         const float* __restrict__ growth, const float* __restrict__ decay, float* __restrict__ age) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (active[i] == 0.0f) continue;
    const float dt = p.dt, old = st.energy[i];
    const float g = f_sub(f_mul(growth[i], old), f_mul(decay[i], f_mul(old, old)));
    float en = f_sub(f_add(old, f_mul(dt, g)), energy_out[i]);
    en = fminf(fmaxf(en, 0.0f), 5.0f);
    const float prob = sigmoidf_(__fdiv_rn(f_mul(4.0f, f_sub(p.x_max, st.X[i])), f_sub(p.x_max, p.x_min)));
    const uint2 kk = reinterpret_cast<const uint2*>(st.key)[i];
    Key nk, bk;
    split2(Key{kk.x, kk.y}, nk, bk);
    const bool luck = uniform1(bk, 0.0f, 1.0f) < prob;  // bernoulli(key, p) = uniform(key) < p
    const bool old_b = st.blossom[i] != 0;
    const bool blossom = old_b ? true : (luck && en > p.blossom_thresh);
    st.energy[i] = en;
    st.rad[i] = en;
    st.blossom[i] = blossom ? 1 : 0;
    reinterpret_cast<uint2*>(st.key)[i] = make_uint2(nk.a, nk.b);
    st.time_above_blossom[i] = blossom ? f_add(st.time_above_blossom[i], dt) : 0.0f;
    age[i] = f_add(age[i], dt);
  }
}

// forager_resource_interaction (sim.py:358-375): energy_in[f] = sum_r e(f,r), energy_out[r] = sum_f e(f,r),
// e(f,r) = energy_r if |f - r| < rad_r + 0.1 else 0.  One CTA per forager: threads take consecutive
// resources (coalesced, IA_U independent loads in flight); the few hits go to a shared list that
// warp 0 puts back into ascending resource order before folding, so energy_in is the serial sum
// over r whatever the launch shape (a list overflow falls back to an ordered rescan).  energy_out is
// accumulated with atomics: all addends of one resource are equal, so that sum is order-free too.
constexpr int IA_BLOCK = 256, IA_U = 4, IA_CAP = 128;

__global__ void __launch_bounds__(IA_BLOCK)
    interaction_kernel(int64_t nf, const float* __restrict__ fx, const float* __restrict__ fy, int64_t nr,
                       const float* __restrict__ rx, const float* __restrict__ ry, const float* __restrict__ rrad,
                       const float* __restrict__ ren, float* __restrict__ energy_in, float* __restrict__ energy_out) {
  __shared__ int hit_idx[IA_CAP];
  __shared__ float hit_val[IA_CAP], ordered[IA_CAP];
  __shared__ int n_hits;
  const int64_t f = blockIdx.x;
  const float px = fx[f], py = fy[f];
  if (threadIdx.x == 0) n_hits = 0;
  __syncthreads();
  for (int64_t base = 0; base < nr; base += IA_BLOCK * IA_U) {
    float x[IA_U], y[IA_U], z[IA_U], w[IA_U];
#pragma unroll
    for (int u = 0; u < IA_U; ++u) {
      const int64_t r = base + u * IA_BLOCK + threadIdx.x;
      const bool ok = r < nr;
      x[u] = ok ? rx[r] : 0.0f;
      y[u] = ok ? ry[r] : 0.0f;
      z[u] = ok ? rrad[r] : -2.0f;
      w[u] = ok ? ren[r] : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < IA_U; ++u) {
      const float d = norm2(f_sub(x[u], px), f_sub(y[u], py));
      if (d < f_add(z[u], 0.1f)) {
        const int64_t r = base + u * IA_BLOCK + threadIdx.x;
        const int pos = atomicAdd(&n_hits, 1);
        if (pos < IA_CAP) { hit_idx[pos] = (int)r; hit_val[pos] = w[u]; }
        atomicAdd(energy_out + r, w[u]);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x >= 32) return;
  const int h = n_hits, lane = threadIdx.x;
  float acc = 0.0f;
  if (h <= IA_CAP) {
    for (int e = lane; e < h; e += 32) {
      int rank = 0;
      for (int k = 0; k < h; ++k) rank += hit_idx[k] < hit_idx[e] ? 1 : 0;
      ordered[rank] = hit_val[e];
    }
    __syncwarp();
    if (lane == 0)
      for (int k = 0; k < h; ++k) acc = f_add(acc, ordered[k]);
  } else {  // more overlaps than the list holds: ordered rescan by one warp
    for (int64_t base = 0; base < nr; base += 32) {
      const int64_t r = base + lane;
      const bool ok = r < nr;
      const float d = ok ? norm2(f_sub(rx[r], px), f_sub(ry[r], py)) : 0.0f;
      const float w = ok ? ren[r] : 0.0f;
      unsigned m = __ballot_sync(0xffffffffu, ok && d < f_add(rrad[ok ? r : 0], 0.1f));
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        acc = f_add(acc, __shfl_sync(0xffffffffu, w, src));
      }
    }
  }
  if (lane == 0) energy_in[f] = acc;
}

// sensor_model (sim.py:388-468); entities = [foragers, resources, walls].  The reference prefilters
// agent entity e with column e of a distance matrix ordered [resources, foragers] (sim.py:381,416,435)
// -- reproduced: `partner` is that other agent.  One CTA per forager: the prefilter does not depend on
// the ray, so the broad phase runs with lanes = entities (ballot of |partner - forager| < L + rad), the
// CTA's warps interleaving blocks of entities, and only the survivors reach the narrow phase, lanes =
// rays.  argmin keeps the first minimum: strict < inside a warp (ascending entities), then a merge of
// the warps' (distance, entity) pairs that prefers the lower entity on ties.
struct SensorArgs {
  int64_t nf, nr, nw;
  const float *fx, *fy, *fang, *frad, *fen, *fact;
  const int32_t* ftype;
  const float *rx, *ry, *rrad, *ren, *ract;
  const int32_t* rtype;
  const float *wx1, *wy1, *wx2, *wy2;
  int n_rays;
  float span, L;
  float* obs;  // [nf, n_rays*3]
};

constexpr int SM_WARPS = 8, SM_U = 4;

__global__ void __launch_bounds__(SM_WARPS * 32) sensor_model_kernel(const SensorArgs a) {
  __shared__ float sbest[SM_WARPS][32];
  __shared__ long long sidx[SM_WARPS][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t f = blockIdx.x;  // one CTA per forager; its warps interleave blocks of 32*SM_U entities
  const bool act = a.fact[f] != 0.0f;
  const float ox = a.fx[f], oy = a.fy[f], L = a.L, g = a.fang[f];
  const int64_t ne = a.nf + a.nr;
  for (int j0 = 0; j0 < a.n_rays; j0 += 32) {  // one pass for up to 32 rays
    const int j = j0 + lane;
    const bool ray = j < a.n_rays;
    float best = L, s = 0.0f, c = 1.0f;
    int64_t bidx = -1;
    if (act) {
      sincosf(linspace_elem(f_sub(g, a.span), f_add(g, a.span), a.n_rays, ray ? j : 0), &s, &c);
      for (int64_t base = (int64_t)warp * (32 * SM_U); base < ne; base += (int64_t)SM_WARPS * 32 * SM_U) {
        float ex[SM_U], ey[SM_U], er[SM_U], ea[SM_U], qx[SM_U], qy[SM_U];
#pragma unroll
        for (int u = 0; u < SM_U; ++u) {
          const int64_t e = base + u * 32 + lane;
          const bool ok = e < ne;
          const bool isf = !ok || e < a.nf;  // padding lanes read forager 0 (always present)
          const int64_t k = ok ? (isf ? e : e - a.nf) : 0;
          ex[u] = (isf ? a.fx : a.rx)[k];
          ey[u] = (isf ? a.fy : a.ry)[k];
          er[u] = (isf ? a.frad : a.rrad)[k];
          ea[u] = (isf ? a.fact : a.ract)[k];
          // column e of concat(forager-resource dists, forager-forager dists)
          const bool pr = e < a.nr;
          const int64_t kp = ok ? (pr ? e : e - a.nr) : 0;
          qx[u] = (pr ? a.rx : a.fx)[kp];
          qy[u] = (pr ? a.ry : a.fy)[kp];
        }
#pragma unroll
        for (int u = 0; u < SM_U; ++u) {
          const float pd = norm2(f_sub(qx[u], ox), f_sub(qy[u], oy));
          const bool near = (base + u * 32 + lane < ne) && pd < f_add(L, er[u]);
          unsigned m = __ballot_sync(0xffffffffu, near);
          while (m) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const float tx = __shfl_sync(0xffffffffu, ex[u], src), ty = __shfl_sync(0xffffffffu, ey[u], src);
            const float tr = __shfl_sync(0xffffffffu, er[u], src), ta = __shfl_sync(0xffffffffu, ea[u], src);
            const float d = ray_agent_exact(ox, oy, c, s, L, tx, ty, tr, ta);
            if (d < best) { best = d; bidx = base + u * 32 + src; }
          }
        }
      }
    }
    sbest[warp][lane] = best;
    sidx[warp][lane] = bidx;
    __syncthreads();
    if (warp == 0 && ray) {
      // the serial scan keeps the LOWEST entity among equal minima: merge on (distance, index)
      for (int w = 1; w < SM_WARPS; ++w) {
        const float d = sbest[w][lane];
        const int64_t i = sidx[w][lane];
        if (d < best || (d == best && i >= 0 && i < bidx)) { best = d; bidx = i; }
      }
      if (act) {
        const float ex1 = f_add(ox, f_mul(c, L)), ey1 = f_add(oy, f_mul(s, L));
        for (int64_t w = 0; w < a.nw; ++w) {
          const float d = ray_wall_exact(ox, oy, ex1, ey1, L, a.wx1[w], a.wy1[w], a.wx2[w], a.wy2[w]);
          if (d < best) { best = d; bidx = ne + w; }
        }
      }
      float en = 0.0f, ty = 0.0f;
      if (act && bidx >= 0 && bidx < ne) {
        const bool isf = bidx < a.nf;
        const int64_t k = isf ? bidx : bidx - a.nf;
        en = isf ? a.fen[k] : a.ren[k];
        ty = (float)(isf ? a.ftype[k] : a.rtype[k]);
      }
      float* o = a.obs + f * (a.n_rays * 3) + j * 3;
      o[0] = act ? en : 0.0f;
      o[1] = act ? ty : 0.0f;
      o[2] = act ? best : 0.0f;
    }
    __syncthreads();
  }
}

}  // namespace fgx

extern "C" {

}  // extern "C" (templates below need C++ linkage)

// shared launcher: as many warps per CTA (<= FS_WARPS) as the per-warp weight slots allow
template <int NN, int NOBS, int NACT>
static int launch_forager_step_t(const char* what, const fgx_forager_params_t& p, int64_t n, const float* active,
                                 const float* obs, const float* energy_in, const fgx_forager_state_t& st,
                                 const fgx_wc_policy_t& pol, float* age, float* actions, cudaStream_t s,
                                 const fgx::SenseArgs& sense) {
  using namespace fgx;
  const WcLayout L = wc_layout(p.n_neurons, p.n_obs, p.n_act);
  // one weight slot per warp and two CTAs per SM when that fits (most memory-level parallelism);
  // FGX_FS_SLOTS=2 selects the double-buffered variant (tuning hook, see profiles/)
  int slots = 1;
  if (const char* v = getenv("FGX_FS_SLOTS")) slots = atoi(v) == 2 ? 2 : 1;
  const size_t head = 16 * FS_WARPS, per_warp = (size_t)slots * L.total * sizeof(float);
  const size_t budget = slots == 2 ? 226 * 1024 : 113 * 1024;
  int warps = FS_WARPS;
  while (warps > 1 && head + warps * per_warp > budget) --warps;
  FGX_REQUIRE(head + warps * per_warp <= 227 * 1024,
              "%s: policy too large for shared memory (%zu bytes per agent)", what, per_warp);
  size_t smem = head + warps * per_warp;
  // FGX_K2_EVICT_FIRST=0 switches the L2 evict-first hint of the weight copies off (A/B hook, see profiles/)
  static const int evict_first = [] { const char* v = getenv("FGX_K2_EVICT_FIRST"); return v && atoi(v) == 0 ? 0 : 1; }();
  if constexpr (NOBS == 33) {
    // warp-specialised fused sensing (sense.enabled == 2): RW ray warps beside the 8 step warps of every CTA
    if (sense.enabled == 2 && warps == FS_WARPS && slots == 1 && n < (1ll << 27)) {
      const char* rw_env = getenv("FGX_SENSE_RW");
      // 2 or 4 ray warps per CTA (registers are allotted per 4 warps: 3 would cost what 4 do); default 4
      const int rw = rw_env && atoi(rw_env) == 2 ? 2 : 4;
      const size_t smem_f = smem + rw * sizeof(RaySmem) + 16;
      auto launch = [&](auto kern_f) -> int {
        cudaError_t e = cudaFuncSetAttribute(kern_f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f);
        if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
        int per_sm = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern_f, (FS_WARPS + rw) * 32, smem_f);
        if (e != cudaSuccess || per_sm < 1) per_sm = 1;
        const int64_t need = ceil_div(n, FS_WARPS), cap = (int64_t)sm_count() * per_sm;
        kern_f<<<(int)(need < cap ? need : cap), (FS_WARPS + rw) * 32, smem_f, s>>>(
            p, n, active, obs, energy_in, st, pol, age, actions, slots | (evict_first << 8), sense);
        return check_launch(what);
      };
      return rw == 4 ? launch(forager_step_kernel<NN, NOBS, NACT, 4>) : launch(forager_step_kernel<NN, NOBS, NACT, 2>);
    }
  }
  auto launch_plain = [&](auto kern) -> int {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem);
    if (e != cudaSuccess || per_sm < 1) per_sm = 1;
    const int64_t need = ceil_div(n, warps);
    const int64_t cap = (int64_t)sm_count() * per_sm;
    kern<<<(int)(need < cap ? need : cap), warps * 32, smem, s>>>(p, n, active, obs, energy_in, st, pol, age, actions,
                                                                slots | (evict_first << 8), sense);
    return check_launch(what);
  };
  // the sensing code lives in its own instantiation (RW = -1): the plain step kernel carries none of it
  return sense.enabled ? launch_plain(forager_step_kernel<NN, NOBS, NACT, -1>)
                       : launch_plain(forager_step_kernel<NN, NOBS, NACT, 0>);
}

static int launch_forager_step(const char* what, const fgx_forager_params_t& p, int64_t n, const float* active,
                               const float* obs, const float* energy_in, const fgx_forager_state_t& st,
                               const fgx_wc_policy_t& pol, float* age, float* actions, cudaStream_t s,
                               const fgx::SenseArgs& sense = fgx::SenseArgs{}) {
#define FGX_FS_SHAPE(NN, NOBS, NACT)                                                              \
  if (p.n_neurons == NN && p.n_obs == NOBS && p.n_act == NACT)                                    \
    return launch_forager_step_t<NN, NOBS, NACT>(what, p, n, active, obs, energy_in, st, pol, age, actions, s, sense);
  FGX_FS_SHAPE(40, 28, 2)  // EcoEvoJax/sim.py:712
  FGX_FS_SHAPE(30, 28, 2)  // EcoEvoJax/sim.ipynb
  FGX_FS_SHAPE(40, 33, 2)  // bench C4: 32 ray distances + own energy
#undef FGX_FS_SHAPE
  return launch_forager_step_t<0, 0, 0>(what, p, n, active, obs, energy_in, st, pol, age, actions, s, sense);
}

extern "C" {

int fgx_forager_step(const fgx_forager_params_t* p, int64_t n, const float* active_state, const float* obs,
                     const float* energy_in, const fgx_forager_state_t* state, const fgx_wc_policy_t* policy,
                     float* age, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(p && state && policy, "fgx_forager_step: null struct");
  FGX_REQUIRE(n >= 0, "fgx_forager_step: bad n");
  FGX_REQUIRE(p->n_neurons > 0 && p->n_obs > 0 && p->n_act >= 2, "fgx_forager_step: bad policy shape");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(active_state && obs && energy_in && age, "fgx_forager_step: null pointer");
  return launch_forager_step("fgx_forager_step", *p, n, active_state, obs, energy_in, *state, *policy, age, nullptr,
                             as_stream(stream));
}

// rows of inactive agents: dist 0, hit -1 (the fused kernel only visits active agents)
__global__ void __launch_bounds__(256) fill_inactive_rays_kernel(int64_t n, int n_rays, const float* __restrict__ active,
                                                                 float* __restrict__ dist, int32_t* __restrict__ hit) {
  const int64_t total = n * n_rays;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    if (active[t / n_rays] == 0.0f) {
      dist[t] = 0.0f;
      if (hit) hit[t] = -1;
    }
  }
}

int fgx_forager_sense_step(const fgx_forager_params_t* p, int64_t n, const float* active_state, const float* energy_in,
                           const fgx_forager_state_t* state, const fgx_wc_policy_t* policy, float* age,
                           int32_t n_rays, float ray_span, float ray_length, const float* wx1, const float* wy1,
                           const float* wx2, const float* wy2, int64_t n_walls, float x_min, float x_max, float y_min,
                           float y_max, void* wall_scratch, size_t wall_scratch_bytes, int32_t reuse_grid, float* dist,
                           int32_t* hit, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(p && state && policy, "fgx_forager_sense_step: null struct");
  FGX_REQUIRE(n >= 0, "fgx_forager_sense_step: bad n");
  FGX_REQUIRE(p->n_neurons > 0 && p->n_obs > 0 && p->n_act >= 2, "fgx_forager_sense_step: bad policy shape");
  FGX_REQUIRE(n_rays > 0 && n_rays == p->n_obs - 1,
              "fgx_forager_sense_step: the policy observes the ray distances and its own energy: n_rays must be n_obs - 1");
  cudaStream_t s = as_stream(stream);
  SenseArgs sense{};
  // built (reuse_grid == 0) even for an empty agent set: the caller caches the scratch as "built" after this call
  const int rc = wall_grid_prepare("fgx_forager_sense_step", wx1, wy1, wx2, wy2, n_walls, x_min, x_max, y_min, y_max,
                                   ray_length, wall_scratch, wall_scratch_bytes, reuse_grid, s, &sense.v);
  if (rc != FGX_OK) return rc;
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(active_state && energy_in && age && dist, "fgx_forager_sense_step: null pointer");
  // mode 2 (default when it applies: 32 rays, the 40 / 33 / 2 policy): ray warps beside the step warps of every CTA;
  // mode 1 (FGX_SENSE_STEP_MODE=1, or any other shape): the warp that steps an agent casts its rays first
  const char* mode_env = getenv("FGX_SENSE_STEP_MODE");  // read per call (tests switch it)
  const bool specialised = !(mode_env && atoi(mode_env) == 1) && n_rays == 32 && p->n_neurons == 40 && p->n_obs == 33 && p->n_act == 2 &&
                           n < (1ll << 27) && !getenv("FGX_FS_SLOTS");
  sense.enabled = specialised ? 2 : 1;
  sense.n_rays = n_rays;
  sense.span = ray_span;
  sense.L = ray_length;
  sense.dist = dist;
  sense.hit = hit;
  // mode 1 visits active agents only; the ray warps of mode 2 write the rows of inactive agents themselves
  if (!specialised) fill_inactive_rays_kernel<<<grid_for(n * n_rays, 256, 8), 256, 0, s>>>(n, n_rays, active_state, dist, hit);
  return launch_forager_step("fgx_forager_sense_step", *p, n, active_state, nullptr, energy_in, *state, *policy, age,
                             nullptr, s, sense);
}

int fgx_wilson_cowan_step(int64_t n, int32_t n_neurons, int32_t n_obs, int32_t n_act, float policy_dt,
                          float action_scale, const float* active_state, const float* obs,
                          const fgx_wc_policy_t* policy, float* actions, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(policy && n >= 0, "fgx_wilson_cowan_step: bad arguments");
  FGX_REQUIRE(n_neurons > 0 && n_obs > 0 && n_act > 0, "fgx_wilson_cowan_step: bad policy shape");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(active_state && obs && actions, "fgx_wilson_cowan_step: null pointer");
  fgx_forager_params_t p{};
  p.n_neurons = n_neurons; p.n_obs = n_obs; p.n_act = n_act; p.policy_dt = policy_dt; p.action_scale = action_scale;
  fgx_forager_state_t st{};
  return launch_forager_step("fgx_wilson_cowan_step", p, n, active_state, obs, nullptr, st, *policy, nullptr, actions,
                             as_stream(stream));
}

int fgx_resource_step(const fgx_resource_params_t* p, int64_t n, const float* active_state, const float* energy_out,
                      const fgx_resource_state_t* state, const float* growth_rate, const float* decay_rate,
                      float* age, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(p && state, "fgx_resource_step: null struct");
  FGX_REQUIRE(n >= 0, "fgx_resource_step: bad n");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(active_state && energy_out && growth_rate && decay_rate && age, "fgx_resource_step: null pointer");
  resource_step_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(*p, n, active_state, energy_out, *state,
                                                                          growth_rate, decay_rate, age);
  return check_launch("fgx_resource_step");
}

int fgx_forager_resource_interaction(int64_t n_foragers, const float* fx, const float* fy, int64_t n_resources,
                                     const float* rx, const float* ry, const float* rrad, const float* renergy,
                                     float* energy_in, float* energy_out, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n_foragers >= 0 && n_foragers < (1ll << 31) && n_resources >= 0 && n_resources < (1ll << 31),
              "fgx_forager_resource_interaction: bad sizes");
  cudaStream_t s = as_stream(stream);
  if (n_resources > 0) {
    FGX_REQUIRE(energy_out, "fgx_forager_resource_interaction: null energy_out");
    cudaError_t e = cudaMemsetAsync(energy_out, 0, sizeof(float) * n_resources, s);
    if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "fgx_forager_resource_interaction: %s", cudaGetErrorString(e));
  }
  if (n_foragers == 0) return FGX_OK;
  FGX_REQUIRE(fx && fy && energy_in, "fgx_forager_resource_interaction: null pointer");
  FGX_REQUIRE(n_resources == 0 || (rx && ry && rrad && renergy), "fgx_forager_resource_interaction: null pointer");
  interaction_kernel<<<(unsigned)n_foragers, IA_BLOCK, 0, s>>>(n_foragers, fx, fy, n_resources, rx, ry, rrad,
                                                                   renergy, energy_in, energy_out);
  return check_launch("fgx_forager_resource_interaction");
}

int fgx_sensor_model(int64_t n_foragers, const float* fx, const float* fy, const float* fang, const float* frad,
                     const float* fenergy, const float* factive, const int32_t* ftype, int64_t n_resources,
                     const float* rx, const float* ry, const float* rrad, const float* renergy, const float* ractive,
                     const int32_t* rtype, int64_t n_walls, const float* wx1, const float* wy1, const float* wx2,
                     const float* wy2, int32_t n_rays, float ray_span, float ray_length, float* obs,
                     fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n_foragers >= 0 && n_foragers < (1ll << 31) && n_resources >= 0 && n_walls >= 0 && n_rays > 0,
              "fgx_sensor_model: bad sizes");
  if (n_foragers == 0) return FGX_OK;
  FGX_REQUIRE(fx && fy && fang && frad && fenergy && factive && ftype && obs, "fgx_sensor_model: null forager leaf");
  FGX_REQUIRE(n_resources == 0 || (rx && ry && rrad && renergy && ractive && rtype),
              "fgx_sensor_model: null resource leaf");
  FGX_REQUIRE(n_walls == 0 || (wx1 && wy1 && wx2 && wy2), "fgx_sensor_model: null wall array");
  SensorArgs a{n_foragers, n_resources, n_walls, fx, fy, fang, frad, fenergy, factive, ftype, rx, ry, rrad, renergy,
               ractive, rtype, wx1, wy1, wx2, wy2, n_rays, ray_span, ray_length, obs};
  sensor_model_kernel<<<(unsigned)n_foragers, SM_WARPS * 32, 0, as_stream(stream)>>>(a);
  return check_launch("fgx_sensor_model");
}
}
