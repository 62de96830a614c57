// K3 across GPUs: sort_agents(active_state, descending) of a block-sharded set as ONE kernel that does
// the compaction and its exchange -- every rank scatters its rows straight into the peers' buffers
// (stores over NVLink / NVSwitch into peer-mapped memory), at the slots the rows have in the GLOBAL stable
// order "all actives in slot order, then all inactives" (agent_methods.py:74-79 applied to the whole set).
//
// Inputs per rank: its local stable partition (fgx_select: front = active rows ascending, then the rest), the
// all-gathered front counts (device int64[world]; read on device, no host round trip) and, per leaf, the
// destination buffer of EVERY rank.  Position p of the local partition is row perm[p]; its global
// position is A_before + p for a front row and A_total + I_before + (p - count) otherwise, which fixes the
// owning rank and the slot.  Consecutive positions go to consecutive slots, so the peer stores coalesce.
// The caller orders the launches across ranks (a collective after the kernel: nobody reads a
// destination buffer before every rank's scatter has completed).  HBM / NVLink bytes: row_bytes per
// slot read locally, row_bytes per slot written (locally or to a peer), 4 B of index.
#include "common.cuh"

namespace fgx {

struct PeerTable {
  fgx_peer_leaf_t leaf[FGX_MAX_PEER_LEAVES];
  int32_t vec[FGX_MAX_PEER_LEAVES];
  int64_t bounds[FGX_MAX_PEERS + 1];
  int32_t rank, world;
};

struct PackPlan {
  int64_t a_before, a_total, i_before, count;
};

__device__ __forceinline__ PackPlan pack_plan(const PeerTable& tab, const int64_t* __restrict__ front_counts,
                                              const int32_t* __restrict__ n_front) {
  PackPlan pl{0, 0, 0, (int64_t)__ldg(n_front)};
  for (int k = 0; k < tab.world; ++k) {
    const int64_t c = __ldg(front_counts + k);
    if (k < tab.rank) pl.a_before += c;
    pl.a_total += c;
  }
  pl.i_before = tab.bounds[tab.rank] - pl.a_before;
  return pl;
}

// global position of local partition position p -> (owner rank, slot)
__device__ __forceinline__ void pack_dest(const PeerTable& tab, const PackPlan& pl, int64_t p, int& d, int64_t& slot) {
  const int64_t g = p < pl.count ? pl.a_before + p : pl.a_total + pl.i_before + (p - pl.count);
  d = 0;
  while (d + 1 < tab.world && g >= tab.bounds[d + 1]) ++d;
  slot = g - tab.bounds[d];
}

template <typename T>
__device__ __forceinline__ void pack_narrow(const PeerTable& tab, const PackPlan& pl, int li,
                                            const int32_t* __restrict__ perm, int64_t n) {
  const fgx_peer_leaf_t& lf = tab.leaf[li];
  const int64_t chunks = lf.row_bytes / (int64_t)sizeof(T);
  const T* __restrict__ src = reinterpret_cast<const T*>(lf.src);
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t p0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p0 < n; p0 += 4 * nthreads) {
    int64_t s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t p = p0 + u * nthreads;
      s[u] = p < n ? (int64_t)__ldg(perm + p) : 0;
    }
    if (chunks == 1) {
      T v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = __ldg(src + s[u]);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t p = p0 + u * nthreads;
        if (p < n) {
          int d;
          int64_t slot;
          pack_dest(tab, pl, p, d, slot);
          reinterpret_cast<T*>(lf.dst[d])[slot] = v[u];
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t p = p0 + u * nthreads;
        if (p < n) {
          int d;
          int64_t slot;
          pack_dest(tab, pl, p, d, slot);
          T* dst = reinterpret_cast<T*>(lf.dst[d]) + slot * chunks;
          for (int64_t c = 0; c < chunks; ++c) dst[c] = __ldg(src + s[u] * chunks + c);
        }
      }
    }
  }
}

// grid.y = leaf; narrow leaves: one thread per row
__global__ void __launch_bounds__(256) pack_narrow_kernel(const __grid_constant__ PeerTable tab,
                                                          const int32_t* __restrict__ perm,
                                                          const int32_t* __restrict__ n_front,
                                                          const int64_t* __restrict__ front_counts, int64_t n) {
  const PackPlan pl = pack_plan(tab, front_counts, n_front);
  switch (tab.vec[blockIdx.y]) {
    case 16: pack_narrow<uint4>(tab, pl, blockIdx.y, perm, n); break;
    case 8: pack_narrow<uint2>(tab, pl, blockIdx.y, perm, n); break;
    case 4: pack_narrow<uint32_t>(tab, pl, blockIdx.y, perm, n); break;
    case 2: pack_narrow<uint16_t>(tab, pl, blockIdx.y, perm, n); break;
    default: pack_narrow<uint8_t>(tab, pl, blockIdx.y, perm, n); break;
  }
}

// wide leaves (per-agent weight matrices): one warp per row, lanes stride the chunks
__global__ void __launch_bounds__(256) pack_wide_kernel(const __grid_constant__ PeerTable tab,
                                                        const int32_t* __restrict__ perm,
                                                        const int32_t* __restrict__ n_front,
                                                        const int64_t* __restrict__ front_counts, int64_t n) {
  const PackPlan pl = pack_plan(tab, front_counts, n_front);
  const fgx_peer_leaf_t& lf = tab.leaf[blockIdx.y];
  const int v = tab.vec[blockIdx.y];
  const int64_t chunks = lf.row_bytes / v;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t p = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; p < n; p += nwarps) {
    const int64_t s = __ldg(perm + p);
    int d;
    int64_t slot;
    pack_dest(tab, pl, p, d, slot);
    const char* __restrict__ src = reinterpret_cast<const char*>(lf.src) + s * lf.row_bytes;
    char* dst = reinterpret_cast<char*>(lf.dst[d]) + slot * lf.row_bytes;
    for (int64_t c = lane; c < chunks; c += 32) {
      if (v == 16) reinterpret_cast<uint4*>(dst)[c] = __ldg(reinterpret_cast<const uint4*>(src) + c);
      else if (v == 8) reinterpret_cast<uint2*>(dst)[c] = __ldg(reinterpret_cast<const uint2*>(src) + c);
      else if (v == 4) reinterpret_cast<uint32_t*>(dst)[c] = __ldg(reinterpret_cast<const uint32_t*>(src) + c);
      else if (v == 2) reinterpret_cast<uint16_t*>(dst)[c] = __ldg(reinterpret_cast<const uint16_t*>(src) + c);
      else dst[c] = __ldg(src + c);
    }
  }
}

}  // namespace fgx

extern "C" {

int fgx_pack_rows_peer(const fgx_peer_leaf_t* leaves_host, int32_t n_leaves, const int32_t* perm,
                       const int32_t* n_front, int64_t n_local, const int64_t* front_counts,
                       const int64_t* bounds_host, int32_t rank, int32_t world, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(world >= 1 && world <= FGX_MAX_PEERS && rank >= 0 && rank < world, "fgx_pack_rows_peer: bad rank / world");
  FGX_REQUIRE(n_leaves >= 0 && n_leaves <= FGX_MAX_PEER_LEAVES, "fgx_pack_rows_peer: at most %d leaves per call",
              FGX_MAX_PEER_LEAVES);
  FGX_REQUIRE(bounds_host, "fgx_pack_rows_peer: null bounds");
  for (int k = 0; k < world; ++k)
    FGX_REQUIRE(bounds_host[k] <= bounds_host[k + 1], "fgx_pack_rows_peer: shard bounds must ascend");
  FGX_REQUIRE(bounds_host[0] == 0 && bounds_host[rank + 1] - bounds_host[rank] == n_local,
              "fgx_pack_rows_peer: n_local does not match this rank's shard bounds");
  FGX_REQUIRE(n_local >= 0 && n_local < (1ll << 31), "fgx_pack_rows_peer: n_local out of range");
  if (n_local == 0 || n_leaves == 0) return FGX_OK;
  FGX_REQUIRE(leaves_host && perm && n_front && front_counts, "fgx_pack_rows_peer: null pointer");
  PeerTable narrow, wide;
  int nn = 0, nw = 0;
  for (int k = 0; k < n_leaves; ++k) {
    const fgx_peer_leaf_t& lf = leaves_host[k];
    FGX_REQUIRE(lf.src && lf.row_bytes > 0, "fgx_pack_rows_peer: leaf %d is malformed", k);
    uintptr_t bits = reinterpret_cast<uintptr_t>(lf.src) | (uintptr_t)lf.row_bytes;
    for (int d = 0; d < world; ++d) {
      FGX_REQUIRE(lf.dst[d], "fgx_pack_rows_peer: leaf %d has no destination on rank %d", k, d);
      FGX_REQUIRE(lf.dst[d] != lf.src, "fgx_pack_rows_peer: leaf %d aliases its destination", k);
      bits |= reinterpret_cast<uintptr_t>(lf.dst[d]);
    }
    int v = 16;
    while (v > 1 && (bits & (uintptr_t)(v - 1))) v >>= 1;
    PeerTable& t = lf.row_bytes / v <= 8 ? narrow : wide;
    int& cnt = lf.row_bytes / v <= 8 ? nn : nw;
    t.leaf[cnt] = lf;
    t.vec[cnt++] = v;
  }
  for (PeerTable* t : {&narrow, &wide}) {
    for (int k = 0; k <= world; ++k) t->bounds[k] = bounds_host[k];
    t->rank = rank;
    t->world = world;
  }
  cudaStream_t s = as_stream(stream);
  if (nn) {
    dim3 grid(grid_for(ceil_div(n_local, 4), 256, 8), nn);
    pack_narrow_kernel<<<grid, 256, 0, s>>>(narrow, perm, n_front, front_counts, n_local);
  }
  if (nw) {
    dim3 grid(grid_for(n_local * 32, 256, 8), nw);
    pack_wide_kernel<<<grid, 256, 0, s>>>(wide, perm, n_front, front_counts, n_local);
  }
  return check_launch("fgx_pack_rows_peer");
}
}

// ---- counts between ranks without the host or NCCL: mailboxes in peer-mapped memory ------------------------
namespace fgx {

struct Mailbox {
  unsigned long long flag[FGX_MAX_PEERS];        // flag[s] = last epoch rank s has released here
  long long payload[2][FGX_MAX_PEERS][2];        // [epoch parity][source rank][value]
};

struct PeerSync {
  Mailbox* box[FGX_MAX_PEERS];
  unsigned long long* epoch;
  int32_t* error;
  int32_t rank, world;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys(long long* p, long long v) {
  asm volatile("st.relaxed.sys.global.s64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ long long ld_relaxed_sys(const long long* p) {
  long long v;
  asm volatile("ld.relaxed.sys.global.s64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// One warp: lane d publishes to rank d, lane s waits for rank s.  Returns (in every lane) the values of all ranks
// through vals[j][s] in shared memory.  Payload slots alternate with the epoch parity: a rank can run at most one
// epoch ahead of a peer (it cannot pass epoch e + 1 before the peer has released e + 1, which the peer does only
// after it has read epoch e), so the slot written for e + 1 is never the one a peer still reads for e.
__device__ __forceinline__ void peer_exchange_warp(const PeerSync& sy, long long v0, long long v1,
                                                   long long (*vals)[FGX_MAX_PEERS]) {
  const int lane = threadIdx.x & 31;
  const unsigned long long e = *sy.epoch + 1ull;
  const int par = (int)(e & 1ull);
  if (lane < sy.world) {
    Mailbox* peer = sy.box[lane];
    st_relaxed_sys(&peer->payload[par][sy.rank][0], v0);
    st_relaxed_sys(&peer->payload[par][sy.rank][1], v1);
    __threadfence_system();  // also orders the peer stores of every earlier kernel of this stream before the flag
    st_release_sys(&peer->flag[sy.rank], e);
    const Mailbox* mine = sy.box[sy.rank];
    const unsigned long long t0 = globaltimer_ns();
    bool ok = true;
    while (ld_acquire_sys(&mine->flag[lane]) < e) {
      if (globaltimer_ns() - t0 > 5000000000ull) {  // 5 s: a peer is gone -- report, never hang
        ok = false;
        break;
      }
      __nanosleep(64);
    }
    if (!ok) atomicExch(sy.error, 1 + lane);
    vals[0][lane] = ld_relaxed_sys(&mine->payload[par][lane][0]);
    vals[1][lane] = ld_relaxed_sys(&mine->payload[par][lane][1]);
  } else if (lane < FGX_MAX_PEERS) {
    vals[0][lane] = 0;
    vals[1][lane] = 0;
  }
  __syncwarp();
  if (lane == 0) *sy.epoch = e;
}

__global__ void __launch_bounds__(32) peer_exchange_kernel(const PeerSync sy, const int32_t* __restrict__ v0,
                                                           const int32_t* __restrict__ v1,
                                                           long long* __restrict__ out_all) {
  __shared__ long long vals[2][FGX_MAX_PEERS];
  peer_exchange_warp(sy, v0 ? (long long)*v0 : 0ll, v1 ? (long long)*v1 : 0ll, vals);
  const int lane = threadIdx.x & 31;
  if (out_all != nullptr && lane < FGX_MAX_PEERS) {
    out_all[lane] = vals[0][lane];
    out_all[FGX_MAX_PEERS + lane] = vals[1][lane];
  }
}

__global__ void __launch_bounds__(32) peer_add_range_kernel(const PeerSync sy, const int32_t* __restrict__ n_sel,
                                                            const int32_t* __restrict__ n_act, long long n_local,
                                                            long long lo, long long n_total,
                                                            int32_t* __restrict__ out3) {
  __shared__ long long vals[2][FGX_MAX_PEERS];
  peer_exchange_warp(sy, (long long)*n_sel, (long long)*n_act, vals);
  if ((threadIdx.x & 31) == 0) {
    long long sel = 0, act = 0;
    for (int s = 0; s < sy.world; ++s) {
      sel += vals[0][s];
      act += vals[1][s];
    }
    const long long k = sel < n_total - act ? sel : n_total - act;
    auto clampll = [](long long v, long long a, long long b) { return v < a ? a : (v > b ? b : v); };
    const long long start = clampll(act - lo, 0, n_local), end = clampll(act + k - lo, 0, n_local);
    out3[0] = (int32_t)k;
    out3[1] = (int32_t)(end - start);
    out3[2] = (int32_t)(start + lo - act);
  }
}

// ---- positions for agent-agent sensing: pack -> barrier -> pull over NVLink, in slot order -----------------
struct GatherPlan {
  const float4* stage[FGX_MAX_PEERS];  // every rank's staging buffer (both parities) as mapped here
  long long bounds[FGX_MAX_PEERS + 1];
  long long n_max;
  int32_t world, layout;
};

// rows of this rank -> its staging buffer, parity of the NEXT epoch (the barrier that follows bumps the epoch)
__global__ void __launch_bounds__(256) peer_stage_positions_kernel(const unsigned long long* __restrict__ epoch, float4* stage,
                                                                   long long n_max, const float* __restrict__ x,
                                                                   const float* __restrict__ y, const float* __restrict__ r,
                                                                   const float* __restrict__ a, long long n) {
  float4* dst = stage + (long long)((*epoch + 1ull) & 1ull) * n_max;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    dst[i] = make_float4(x[i], y[i], r[i], a[i]);
}

// every global slot <- the owning rank's staged row (peer loads), parity of the CURRENT epoch
__global__ void __launch_bounds__(256) peer_pull_positions_kernel(const unsigned long long* __restrict__ epoch,
                                                                  const __grid_constant__ GatherPlan pl, long long n_total,
                                                                  float4* __restrict__ table) {
  const long long par = (long long)(*epoch & 1ull) * pl.n_max;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n_total; i += (long long)gridDim.x * blockDim.x) {
    int d;
    long long row;
    if (pl.layout == 1) {
      d = (int)(i % pl.world);
      row = i / pl.world;
    } else {
      d = 0;
      while (d + 1 < pl.world && i >= pl.bounds[d + 1]) ++d;
      row = i - pl.bounds[d];
    }
    table[i] = pl.stage[d][par + row];  // plain (coherent) load: the rows were published by the barrier's fences
  }
}

static int make_sync(const char* what, const fgx_peer_sync_t* sync, PeerSync* out) {
  FGX_REQUIRE(sync, "%s: null sync", what);
  FGX_REQUIRE(sync->world >= 1 && sync->world <= FGX_MAX_PEERS && sync->rank >= 0 && sync->rank < sync->world,
              "%s: bad rank / world", what);
  FGX_REQUIRE(sync->epoch && sync->error, "%s: null epoch / error", what);
  for (int d = 0; d < sync->world; ++d) {
    FGX_REQUIRE(sync->mailbox[d], "%s: rank %d has no mailbox", what, d);
    FGX_REQUIRE((reinterpret_cast<uintptr_t>(sync->mailbox[d]) & 15) == 0, "%s: mailbox %d must be 16-byte aligned", what, d);
    out->box[d] = reinterpret_cast<Mailbox*>(sync->mailbox[d]);
  }
  out->epoch = reinterpret_cast<unsigned long long*>(sync->epoch);
  out->error = sync->error;
  out->rank = sync->rank;
  out->world = sync->world;
  return FGX_OK;
}

}  // namespace fgx

extern "C" {

size_t fgx_peer_mailbox_bytes(void) { return (sizeof(fgx::Mailbox) + 255) / 256 * 256; }

int fgx_peer_exchange(const fgx_peer_sync_t* sync, const int32_t* v0, const int32_t* v1, int64_t* out_all,
                      fgx_stream_t stream) {
  using namespace fgx;
  PeerSync sy{};
  const int rc = make_sync("fgx_peer_exchange", sync, &sy);
  if (rc != FGX_OK) return rc;
  peer_exchange_kernel<<<1, 32, 0, as_stream(stream)>>>(sy, v0, v1, reinterpret_cast<long long*>(out_all));
  return check_launch("fgx_peer_exchange");
}

int fgx_peer_gather_positions(const fgx_peer_sync_t* sync, void* const* stage, int64_t n_max, const float* x,
                              const float* y, const float* rad, const float* active, int64_t n_local, int64_t n_total,
                              int32_t layout, const int64_t* bounds_host, float* table, fgx_stream_t stream) {
  using namespace fgx;
  PeerSync sy{};
  const int rc = make_sync("fgx_peer_gather_positions", sync, &sy);
  if (rc != FGX_OK) return rc;
  FGX_REQUIRE(stage && table && n_local >= 0 && n_total >= n_local && n_max >= n_local,
              "fgx_peer_gather_positions: bad arguments");
  FGX_REQUIRE(n_local == 0 || (x && y && rad && active), "fgx_peer_gather_positions: null leaf");
  FGX_REQUIRE(layout == 0 || layout == 1, "fgx_peer_gather_positions: layout must be 0 (blocks) or 1 (round robin)");
  GatherPlan pl{};
  pl.world = sy.world;
  pl.layout = layout;
  pl.n_max = n_max;
  for (int d = 0; d < sy.world; ++d) {
    FGX_REQUIRE(stage[d] && (reinterpret_cast<uintptr_t>(stage[d]) & 15) == 0,
                "fgx_peer_gather_positions: staging buffer %d missing / unaligned", d);
    pl.stage[d] = reinterpret_cast<const float4*>(stage[d]);
  }
  if (layout == 0) {
    FGX_REQUIRE(bounds_host && bounds_host[0] == 0 && bounds_host[sy.world] == n_total,
                "fgx_peer_gather_positions: block layout needs shard bounds covering [0, n_total)");
    for (int d = 0; d < sy.world; ++d) {
      FGX_REQUIRE(bounds_host[d] <= bounds_host[d + 1] && bounds_host[d + 1] - bounds_host[d] <= n_max,
                  "fgx_peer_gather_positions: bad shard bounds");
      pl.bounds[d] = bounds_host[d];
    }
    pl.bounds[sy.world] = bounds_host[sy.world];
    FGX_REQUIRE(bounds_host[sy.rank + 1] - bounds_host[sy.rank] == n_local, "fgx_peer_gather_positions: n_local != own block");
  } else {
    FGX_REQUIRE((n_total + sy.world - 1 - sy.rank) / sy.world == n_local && (n_total + sy.world - 1) / sy.world <= n_max,
                "fgx_peer_gather_positions: n_local does not match slot i -> rank i mod world");
  }
  cudaStream_t s = as_stream(stream);
  float4* own = reinterpret_cast<float4*>(stage[sy.rank]);
  if (n_local)
    peer_stage_positions_kernel<<<grid_for(n_local, 256, 4), 256, 0, s>>>(sy.epoch, own, n_max, x, y, rad, active, n_local);
  peer_exchange_kernel<<<1, 32, 0, s>>>(sy, nullptr, nullptr, nullptr);  // barrier: every rank's rows are staged
  if (n_total)
    peer_pull_positions_kernel<<<grid_for(n_total, 256, 8), 256, 0, s>>>(sy.epoch, pl, n_total, reinterpret_cast<float4*>(table));
  return check_launch("fgx_peer_gather_positions");
}

int fgx_peer_add_range(const fgx_peer_sync_t* sync, const int32_t* n_selected_local, const int32_t* n_active_local,
                       int64_t n_local, int64_t lo, int64_t n_total, int32_t* out3, fgx_stream_t stream) {
  using namespace fgx;
  PeerSync sy{};
  const int rc = make_sync("fgx_peer_add_range", sync, &sy);
  if (rc != FGX_OK) return rc;
  FGX_REQUIRE(n_selected_local && n_active_local && out3, "fgx_peer_add_range: null pointer");
  FGX_REQUIRE(n_local >= 0 && lo >= 0 && n_total >= lo + n_local, "fgx_peer_add_range: bad block");
  peer_add_range_kernel<<<1, 32, 0, as_stream(stream)>>>(sy, n_selected_local, n_active_local, n_local, lo, n_total,
                                                         out3);
  return check_launch("fgx_peer_add_range");
}
}

// ---- general keys: the global stable sort of a block-sharded set ("general float keys") ----
// Every rank sorts its keys locally (fgx_argsort), publishes the SORTED key images of its block (an
// all-gather: 4 B per slot), and finds the global position of each of its rows without any exchange of
// rows: position of local sorted row p with image k =  p  +  sum over earlier ranks of #{images <= k}
// +  sum over later ranks of #{images < k}   (ties keep rank order, then local order = the stable
// order of the concatenated set), each count one binary search.  The row is then written once,
// straight to its slot on the owning GPU, exactly like the 1-bit pack above.
#include "key_image.cuh"

namespace fgx {

struct SortTable {
  PeerTable t;
  const uint32_t* images;  // [world][n_pad] sorted key images of every rank's block
  int64_t n_pad;
};

__global__ void __launch_bounds__(256) sort_images_kernel(const void* __restrict__ keys, int dtype, int descending,
                                                          const int32_t* __restrict__ perm, int64_t n,
                                                          uint32_t* __restrict__ out) {
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x)
    out[p] = key_image(keys, dtype, descending, perm[p]);
}

// #{a[0..n) < k} (strict) or #{a[0..n) <= k}
__device__ __forceinline__ int64_t count_below(const uint32_t* __restrict__ a, int64_t n, uint32_t k, bool or_equal) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const uint32_t v = __ldg(a + mid);
    if (v < k || (or_equal && v == k)) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// global position of every local sorted row -> gpos[p]
__global__ void __launch_bounds__(256) sort_positions_kernel(const __grid_constant__ SortTable st, int64_t n,
                                                             int64_t* __restrict__ gpos) {
  const PeerTable& tab = st.t;
  for (int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t k = __ldg(st.images + (int64_t)tab.rank * st.n_pad + p);
    int64_t g = p;
    for (int s = 0; s < tab.world; ++s) {
      if (s == tab.rank) continue;
      g += count_below(st.images + (int64_t)s * st.n_pad, tab.bounds[s + 1] - tab.bounds[s], k, s < tab.rank);
    }
    gpos[p] = g;
  }
}

// row perm[p] of every leaf -> slot of global position gpos[p] on its owner (grid.y = leaf)
__global__ void __launch_bounds__(256) sort_scatter_peer_kernel(const __grid_constant__ PeerTable tab,
                                                                const int32_t* __restrict__ perm,
                                                                const int64_t* __restrict__ gpos, int64_t n) {
  const fgx_peer_leaf_t& lf = tab.leaf[blockIdx.y];
  const int v = tab.vec[blockIdx.y];
  const int64_t chunks = lf.row_bytes / v;
  // narrow rows: a thread per row; wide rows: a warp per row (lanes stride the chunks)
  const bool wide = chunks > 8;
  const int lane = threadIdx.x & 31;
  const int64_t workers = wide ? ((int64_t)gridDim.x * blockDim.x) >> 5 : (int64_t)gridDim.x * blockDim.x;
  const int64_t first = wide ? (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5
                             : blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  for (int64_t p = first; p < n; p += workers) {
    const int64_t s = __ldg(perm + p), g = __ldg(gpos + p);
    int d = 0;
    while (d + 1 < tab.world && g >= tab.bounds[d + 1]) ++d;
    const char* __restrict__ src = reinterpret_cast<const char*>(lf.src) + s * lf.row_bytes;
    char* dst = reinterpret_cast<char*>(lf.dst[d]) + (g - tab.bounds[d]) * lf.row_bytes;
    for (int64_t c = wide ? lane : 0; c < chunks; c += wide ? 32 : 1) {
      if (v == 16) reinterpret_cast<uint4*>(dst)[c] = __ldg(reinterpret_cast<const uint4*>(src) + c);
      else if (v == 8) reinterpret_cast<uint2*>(dst)[c] = __ldg(reinterpret_cast<const uint2*>(src) + c);
      else if (v == 4) reinterpret_cast<uint32_t*>(dst)[c] = __ldg(reinterpret_cast<const uint32_t*>(src) + c);
      else if (v == 2) reinterpret_cast<uint16_t*>(dst)[c] = __ldg(reinterpret_cast<const uint16_t*>(src) + c);
      else dst[c] = __ldg(src + c);
    }
  }
}

}  // namespace fgx

extern "C" {

int fgx_sort_key_images(const void* keys, int32_t key_dtype, int64_t n, int32_t descending, const int32_t* perm,
                        uint32_t* images, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n >= 0 && (key_dtype == FGX_F32 || key_dtype == FGX_I32), "fgx_sort_key_images: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(keys && perm && images, "fgx_sort_key_images: null pointer");
  sort_images_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(keys, key_dtype, descending ? 1 : 0, perm, n,
                                                                        images);
  return check_launch("fgx_sort_key_images");
}

int fgx_sort_rows_peer(const fgx_peer_leaf_t* leaves_host, int32_t n_leaves, const int32_t* perm, int64_t n_local,
                       const uint32_t* images_all, int64_t n_pad, const int64_t* bounds_host, int32_t rank,
                       int32_t world, int64_t* gpos_scratch, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(world >= 1 && world <= FGX_MAX_PEERS && rank >= 0 && rank < world, "fgx_sort_rows_peer: bad rank / world");
  FGX_REQUIRE(n_leaves >= 0 && n_leaves <= FGX_MAX_PEER_LEAVES, "fgx_sort_rows_peer: at most %d leaves per call",
              FGX_MAX_PEER_LEAVES);
  FGX_REQUIRE(bounds_host, "fgx_sort_rows_peer: null bounds");
  for (int k = 0; k < world; ++k) {
    FGX_REQUIRE(bounds_host[k] <= bounds_host[k + 1], "fgx_sort_rows_peer: shard bounds must ascend");
    FGX_REQUIRE(bounds_host[k + 1] - bounds_host[k] <= n_pad, "fgx_sort_rows_peer: n_pad smaller than a shard");
  }
  FGX_REQUIRE(bounds_host[0] == 0 && bounds_host[rank + 1] - bounds_host[rank] == n_local,
              "fgx_sort_rows_peer: n_local does not match this rank's shard bounds");
  FGX_REQUIRE(n_local >= 0 && n_local < (1ll << 31), "fgx_sort_rows_peer: n_local out of range");
  if (n_local == 0 || n_leaves == 0) return FGX_OK;
  FGX_REQUIRE(leaves_host && perm && images_all && gpos_scratch, "fgx_sort_rows_peer: null pointer");
  SortTable st;
  int cnt = 0;
  for (int k = 0; k < n_leaves; ++k) {
    const fgx_peer_leaf_t& lf = leaves_host[k];
    FGX_REQUIRE(lf.src && lf.row_bytes > 0, "fgx_sort_rows_peer: leaf %d is malformed", k);
    uintptr_t bits = reinterpret_cast<uintptr_t>(lf.src) | (uintptr_t)lf.row_bytes;
    for (int d = 0; d < world; ++d) {
      FGX_REQUIRE(lf.dst[d] && lf.dst[d] != lf.src, "fgx_sort_rows_peer: leaf %d has a bad destination on rank %d", k, d);
      bits |= reinterpret_cast<uintptr_t>(lf.dst[d]);
    }
    int v = 16;
    while (v > 1 && (bits & (uintptr_t)(v - 1))) v >>= 1;
    st.t.leaf[cnt] = lf;
    st.t.vec[cnt++] = v;
  }
  for (int k = 0; k <= world; ++k) st.t.bounds[k] = bounds_host[k];
  st.t.rank = rank;
  st.t.world = world;
  st.images = images_all;
  st.n_pad = n_pad;
  cudaStream_t s = as_stream(stream);
  sort_positions_kernel<<<grid_for(n_local, 256, 8), 256, 0, s>>>(st, n_local, gpos_scratch);
  dim3 grid(grid_for(n_local, 256, 8), cnt);
  sort_scatter_peer_kernel<<<grid, 256, 0, s>>>(st.t, perm, gpos_scratch, n_local);
  return check_launch("fgx_sort_rows_peer");
}
}
