// K3 (part 3): row movement of a struct-of-arrays agent pytree, and the active count.
//
// fgx_gather_rows is the payload half of sort_agents (agent_methods.py:77:
// tree_map(lambda x: jnp.take(x, sorted_ids, axis=0), agents)): every leaf is
// gathered by the same permutation in ONE launch (gridDim.y = leaf).  A row of
// a leaf is copied in the widest vector its size and alignment allow; threads
// of a warp copy consecutive chunks of one row (wide leaves: the per-agent RNN
// weights, 6.4 KB rows) or consecutive rows (scalar leaves).  HBM-bound:
// 2 * row_bytes per slot + 4 B of index.
#include "rows.cuh"
#include "threefry.cuh"

namespace fgx {

// Narrow leaves (row <= 8 vector chunks of T): one thread per row, FOUR rows per thread.  The index of a row is
// loaded ONCE and reused for every leaf (an earlier version ran one grid.y slice per leaf and re-read the
// 4 B/slot permutation for each of them: 7x the index traffic on a Dice set); per leaf all four gathers are in
// flight before the first store.  Wide leaves (per-agent weight matrices): one warp per row, lanes stride the
// 16-byte chunks (coalesced 512 B per step), grid.y = leaf.
__device__ __forceinline__ int64_t clamp_index(int64_t s, int64_t n) { return s < 0 ? 0 : (s >= n ? n - 1 : s); }

template <typename T>
__device__ __forceinline__ void gather_narrow4(const fgx_leaf_t& lf, const int64_t (&s)[4], const int64_t (&r)[4],
                                               int64_t n) {
  const int64_t chunks = lf.row_bytes / (int64_t)sizeof(T);
  const T* __restrict__ src = reinterpret_cast<const T*>(lf.src);
  T* __restrict__ dst = reinterpret_cast<T*>(lf.dst);
  if (chunks == 1) {
    T v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(src + s[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (r[u] < n) dst[r[u]] = v[u];
  } else {
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (r[u] < n)
        for (int64_t c = 0; c < chunks; ++c) dst[r[u] * chunks + c] = __ldg(src + s[u] * chunks + c);
  }
}

__global__ void __launch_bounds__(256) gather_narrow_kernel(const __grid_constant__ LeafTable tab, int n_leaves,
                                                            const int32_t* __restrict__ perm, int64_t n,
                                                            int64_t n_src, const int32_t* __restrict__ skip) {
  if (skip != nullptr && __ldcg(skip) != 0) return;  // uniform over the grid: written by an earlier launch
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  for (int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; row < n; row += 4 * nthreads) {
    int64_t s[4], r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      r[u] = row + u * nthreads;
      s[u] = r[u] < n ? clamp_index(__ldg(perm + r[u]), n_src) : 0;  // jnp.take clamps
    }
    for (int t = 0; t < n_leaves; ++t) {
      const fgx_leaf_t& lf = tab.leaf[t];
      switch (tab.vec[t]) {
        case 16: gather_narrow4<uint4>(lf, s, r, n); break;
        case 8: gather_narrow4<uint2>(lf, s, r, n); break;
        case 4: gather_narrow4<uint32_t>(lf, s, r, n); break;
        case 2: gather_narrow4<uint16_t>(lf, s, r, n); break;
        default: gather_narrow4<uint8_t>(lf, s, r, n); break;
      }
    }
  }
}

__global__ void __launch_bounds__(256) gather_wide_kernel(const __grid_constant__ LeafTable tab,
                                                          const int32_t* __restrict__ perm, int64_t n, int64_t n_src) {
  const fgx_leaf_t& lf = tab.leaf[blockIdx.y];
  const int v = tab.vec[blockIdx.y];
  const int64_t chunks = lf.row_bytes / v;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5; row < n; row += nwarps) {
    const int64_t s = clamp_index(__ldg(perm + row), n_src);
    if (v == 16) {
      const uint4* __restrict__ src = reinterpret_cast<const uint4*>(lf.src) + s * chunks;
      uint4* __restrict__ dst = reinterpret_cast<uint4*>(lf.dst) + row * chunks;
      for (int64_t c = lane; c < chunks; c += 32) dst[c] = __ldg(src + c);
    } else {  // rows that are not 16-byte clean: byte-granular vector of v bytes
      const char* __restrict__ src = reinterpret_cast<const char*>(lf.src) + s * lf.row_bytes;
      char* __restrict__ dst = reinterpret_cast<char*>(lf.dst) + row * lf.row_bytes;
      for (int64_t c = lane; c < chunks; c += 32) {
        if (v == 8) reinterpret_cast<uint2*>(dst)[c] = __ldg(reinterpret_cast<const uint2*>(src) + c);
        else if (v == 4) reinterpret_cast<uint32_t*>(dst)[c] = __ldg(reinterpret_cast<const uint32_t*>(src) + c);
        else if (v == 2) reinterpret_cast<uint16_t*>(dst)[c] = __ldg(reinterpret_cast<const uint16_t*>(src) + c);
        else dst[c] = __ldg(src + c);
      }
    }
  }
}

// out[0] += this CTA's count; clip_in != nullptr: the LAST CTA to finish (ticket in out[2]) also writes
// out[1] = min(*clip_in, n - out[0]) -- the caller's "at most as many new agents as free slots" clip
// (README.md:124-125) without a second launch.  All loads of a thread are issued before the adds.
__global__ void __launch_bounds__(256) count_active_kernel(const float* __restrict__ active, int64_t n,
                                                           int32_t* __restrict__ out,
                                                           const int32_t* __restrict__ clip_in) {
  int c = 0;
  // 16-byte loads when the leaf is 16-byte aligned (it is, unless the caller passes an offset view); n < 2^31
  const uint32_t stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t done = 0;
  if ((reinterpret_cast<uintptr_t>(active) & 15) == 0) {
    const float4* __restrict__ a4 = reinterpret_cast<const float4*>(active);
    const uint32_t n4 = (uint32_t)(n >> 2);
    for (uint32_t i = tid; i < n4; i += stride) {
      const float4 v = __ldg(a4 + i);
      c += (int)v.x + (int)v.y + (int)v.z + (int)v.w;  // sum(x, dtype=int32) converts each element first
    }
    done = n4 << 2;
  }
  for (uint32_t i = done + tid; i < (uint32_t)n; i += stride) c += (int)__ldg(active + i);
  c = __reduce_add_sync(0xffffffffu, c);
  __shared__ int ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int k = 0; k < 8; ++k) t += ws[k];
    if (t) atomicAdd(out, t);
    if (clip_in != nullptr) {
      __threadfence();
      if (atomicAdd(out + 2, 1) == (int)gridDim.x - 1) {
        __threadfence();
        const int64_t free_slots = n - (int64_t)__ldcg(out);
        const int64_t want = *clip_in;
        out[1] = (int32_t)(want < free_slots ? want : free_slots);
      }
    }
  }
}

// jnp.sum(x) of a float32 leaf as ONE launch with a fixed summation order (run-to-run deterministic): every CTA
// reduces a contiguous chunk (thread-serial, then a fixed shuffle tree), writes its partial, and the last CTA to
// finish (ticket) adds the partials in index order.
constexpr int SUM_MAX_GRID = 256;
__global__ void __launch_bounds__(256) sum_f32_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out,
                                                      float* __restrict__ partial, unsigned* __restrict__ ticket) {
  __shared__ float ws[8];
  __shared__ int last;
  const int64_t chunk = (n + gridDim.x - 1) / gridDim.x;
  const int64_t lo = blockIdx.x * chunk, hi = min(n, lo + chunk);
  float s = 0.0f;
  for (int64_t i = lo + threadIdx.x; i < hi; i += 256) s += x[i];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int k = 0; k < 8; ++k) t += ws[k];
    __stcg(partial + blockIdx.x, t);
    __threadfence();
    last = atomicAdd(ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    float t = 0.0f;
    for (unsigned k = 0; k < gridDim.x; ++k) t += __ldcg(partial + k);
    *out = t;
    *ticket = 0u;  // the scratch is reusable without another memset
  }
}

// the life-expectancy update of EcoEvoJax/sim.py:486-493 on device scalars:
// le = (age_before - age_after) / (n_removed + 0.01); le = where(le == 0, old, le)
__global__ void life_expectancy_kernel(const float* __restrict__ before, const float* __restrict__ after,
                                       const int32_t* __restrict__ n_removed, const float* __restrict__ old,
                                       float* __restrict__ out) {
  const float le = __fdiv_rn(__fsub_rn(*before, *after), __fadd_rn((float)*n_removed, 0.01f));
  *out = le == 0.0f ? *old : le;
}

// Recording ring (EcoEvoJax/sim.py:581-594 record_data: rec.<leaf>.at[frame].set(leaf) for eight leaves per frame):
// every leaf's n rows are copied into slot (frame mod frames) of its [frames, n, ...] ring in ONE launch
// (grid.y = leaf).  `frame` is a device scalar so that a recorded tick can live inside a CUDA graph.
__global__ void __launch_bounds__(256) record_frame_kernel(const __grid_constant__ LeafTable tab, int64_t n, int32_t frames,
                                                           const int32_t* __restrict__ frame) {
  const fgx_leaf_t& lf = tab.leaf[blockIdx.y];
  const int v = tab.vec[blockIdx.y];
  int32_t f = __ldg(frame) % frames;
  if (f < 0) f += frames;  // negative frames wrap like negative indices (jnp .at[])
  const int64_t bytes = lf.row_bytes * n, chunks = bytes / v;
  const char* __restrict__ src = reinterpret_cast<const char*>(lf.src);
  char* __restrict__ dst = reinterpret_cast<char*>(lf.dst) + (int64_t)f * bytes;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < chunks; c += (int64_t)gridDim.x * blockDim.x) {
    if (v == 16) reinterpret_cast<uint4*>(dst)[c] = __ldg(reinterpret_cast<const uint4*>(src) + c);
    else if (v == 8) reinterpret_cast<uint2*>(dst)[c] = __ldg(reinterpret_cast<const uint2*>(src) + c);
    else if (v == 4) reinterpret_cast<uint32_t*>(dst)[c] = __ldg(reinterpret_cast<const uint32_t*>(src) + c);
    else if (v == 2) reinterpret_cast<uint16_t*>(dst)[c] = __ldg(reinterpret_cast<const uint16_t*>(src) + c);
    else dst[c] = __ldg(src + c);
  }
}

// create_agents' argument construction (agent_methods.py:10-15) for the slots
// [first, first + n_local) of a set of n_total slots (a shard computes exactly its rows)
__global__ void __launch_bounds__(256) create_base_kernel(Key root, int64_t n_total, int64_t n_active, int64_t first,
                                                          int64_t stride, int64_t n_local, int32_t type,
                                                          int32_t* __restrict__ uid,
                                                          uint2* __restrict__ ckeys, float* __restrict__ active,
                                                          int32_t* __restrict__ atype) {
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < n_local; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = first + j * stride;
    const Key ck = split_child(root, (uint32_t)(n_total + 1), (uint32_t)(i + 1));
    uid[j] = (int32_t)i;
    ckeys[j] = make_uint2(ck.a, ck.b);
    active[j] = i < n_active ? 1.0f : 0.0f;
    atype[j] = type;
  }
}

int classify_leaves(const char* what, const fgx_leaf_t* leaves_host, int32_t n_leaves, LeafTable* narrow, int* nn,
                    LeafTable* wide, int* nw) {
  *nn = *nw = 0;
  for (int k = 0; k < n_leaves; ++k) {
    const fgx_leaf_t& lf = leaves_host[k];
    FGX_REQUIRE(lf.src && lf.dst && lf.row_bytes > 0, "%s: leaf %d is malformed", what, k);
    FGX_REQUIRE(lf.src != lf.dst, "%s: leaf %d aliases its destination", what, k);
    const int v = leaf_vec(lf);
    if (lf.row_bytes / v <= 8) {
      narrow->leaf[*nn] = lf;
      narrow->vec[(*nn)++] = v;
    } else {
      wide->leaf[*nw] = lf;
      wide->vec[(*nw)++] = v;
    }
  }
  return FGX_OK;
}

int launch_gather(const LeafTable& narrow, int nn, const LeafTable& wide, int nw, const int32_t* perm, int64_t n,
                  int64_t n_src, const int32_t* skip_narrow, cudaStream_t s) {
  if (nn) gather_narrow_kernel<<<grid_for(ceil_div(n, 4), 256, 8), 256, 0, s>>>(narrow, nn, perm, n, n_src, skip_narrow);
  if (nw) {
    dim3 grid(grid_for(n * 32, 256, 8), nw);
    gather_wide_kernel<<<grid, 256, 0, s>>>(wide, perm, n, n_src);
  }
  return check_launch("fgx_gather_rows");
}

}  // namespace fgx

extern "C" {

int fgx_gather_rows(const fgx_leaf_t* leaves_host, int32_t n_leaves, const int32_t* perm, int64_t n,
                    fgx_stream_t stream) {
  return fgx_gather_rows_from(leaves_host, n_leaves, perm, n, n, stream);
}

int fgx_gather_rows_from(const fgx_leaf_t* leaves_host, int32_t n_leaves, const int32_t* perm, int64_t n,
                         int64_t n_src, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n >= 0 && n < (1ll << 31) && n_src >= 1 && n_src < (1ll << 31), "fgx_gather_rows: n out of range");
  FGX_REQUIRE(n_leaves >= 0 && n_leaves <= FGX_MAX_LEAVES, "fgx_gather_rows: at most %d leaves per call",
              FGX_MAX_LEAVES);
  if (n == 0 || n_leaves == 0) return FGX_OK;
  FGX_REQUIRE(leaves_host && perm, "fgx_gather_rows: null pointer");
  LeafTable narrow, wide;
  int nn = 0, nw = 0;
  const int rc = classify_leaves("fgx_gather_rows", leaves_host, n_leaves, &narrow, &nn, &wide, &nw);
  if (rc != FGX_OK) return rc;
  return launch_gather(narrow, nn, wide, nw, perm, n, n_src, nullptr, as_stream(stream));
}

int fgx_count_active(const float* active_state, int64_t n, int32_t* out, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n >= 0 && n < (1ll << 31) && out, "fgx_count_active: bad arguments");
  cudaStream_t s = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(int32_t), s);
  if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "fgx_count_active: %s", cudaGetErrorString(e));
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(active_state, "fgx_count_active: null pointer");
  count_active_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(active_state, n, out, nullptr);
  return check_launch("fgx_count_active");
}

int fgx_count_active_clip(const float* active_state, int64_t n, const int32_t* n_add, int32_t* out3,
                          fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n >= 0 && n < (1ll << 31) && out3 && n_add, "fgx_count_active_clip: bad arguments");
  FGX_REQUIRE(n == 0 || active_state, "fgx_count_active_clip: null pointer");
  cudaStream_t s = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(out3, 0, 3 * sizeof(int32_t), s);
  if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "fgx_count_active_clip: %s", cudaGetErrorString(e));
  // n == 0: one CTA with nothing to count still writes the clip (min(n_add, 0) = 0)
  count_active_kernel<<<n ? grid_for(n, 256, 8) : 1, 256, 0, s>>>(active_state, n, out3, n_add);
  return check_launch("fgx_count_active_clip");
}

int fgx_create_base_shard(const uint32_t key_host[2], int64_t n_total, int64_t n_active, int64_t first,
                          int64_t stride, int64_t n_local, int32_t agent_type, int32_t* unique_id,
                          uint32_t* create_keys, float* active_state, int32_t* agent_types, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(key_host, "fgx_create_base: null key");
  FGX_REQUIRE(n_total >= 0 && n_active >= 0 && n_active <= n_total, "fgx_create_base: bad counts");
  FGX_REQUIRE(first >= 0 && stride >= 1 && n_local >= 0 && (n_local == 0 || first + (n_local - 1) * stride < n_total),
              "fgx_create_base: bad shard range");
  FGX_REQUIRE(n_total < (1ll << 30), "fgx_create_base: n_total too large for split(key, n+1)");
  if (n_local == 0) return FGX_OK;
  FGX_REQUIRE(unique_id && create_keys && active_state && agent_types, "fgx_create_base: null pointer");
  create_base_kernel<<<grid_for(n_local, 256, 8), 256, 0, as_stream(stream)>>>(
      Key{key_host[0], key_host[1]}, n_total, n_active, first, stride, n_local, agent_type, unique_id,
      reinterpret_cast<uint2*>(create_keys), active_state, agent_types);
  return check_launch("fgx_create_base");
}

int fgx_create_base(const uint32_t key_host[2], int64_t n_total, int64_t n_active, int32_t agent_type,
                    int32_t* unique_id, uint32_t* create_keys, float* active_state, int32_t* agent_types,
                    fgx_stream_t stream) {
  return fgx_create_base_shard(key_host, n_total, n_active, 0, 1, n_total, agent_type, unique_id, create_keys,
                               active_state, agent_types, stream);
}

int fgx_record_frame(const fgx_leaf_t* leaves_host, int32_t n_leaves, int64_t n, int32_t frames, const int32_t* frame,
                     fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n >= 0 && frames >= 1, "fgx_record_frame: bad sizes");
  FGX_REQUIRE(n_leaves >= 0 && n_leaves <= FGX_MAX_LEAVES, "fgx_record_frame: at most %d leaves per call", FGX_MAX_LEAVES);
  if (n == 0 || n_leaves == 0) return FGX_OK;
  FGX_REQUIRE(leaves_host && frame, "fgx_record_frame: null pointer");
  LeafTable tab{};
  int64_t most = 0;
  for (int k = 0; k < n_leaves; ++k) {
    const fgx_leaf_t& lf = leaves_host[k];
    FGX_REQUIRE(lf.src && lf.dst && lf.row_bytes > 0, "fgx_record_frame: leaf %d is malformed", k);
    tab.leaf[k] = lf;
    // chunk width: every frame slot starts at a multiple of n * row_bytes from the ring base
    const uintptr_t bits = reinterpret_cast<uintptr_t>(lf.src) | reinterpret_cast<uintptr_t>(lf.dst) |
                           (uintptr_t)(lf.row_bytes * n);
    int v = 16;
    while (v > 1 && (bits & (uintptr_t)(v - 1))) v >>= 1;
    tab.vec[k] = v;
    if (lf.row_bytes * n / v > most) most = lf.row_bytes * n / v;
  }
  dim3 grid(grid_for(most, 256, 4), n_leaves);
  record_frame_kernel<<<grid, 256, 0, as_stream(stream)>>>(tab, n, frames, frame);
  return check_launch("fgx_record_frame");
}

size_t fgx_sum_f32_scratch_bytes(void) { return (size_t)(fgx::SUM_MAX_GRID + 1) * 4; }

int fgx_sum_f32(const float* x, int64_t n, float* out, void* scratch, size_t scratch_bytes, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n >= 0 && out, "fgx_sum_f32: bad arguments");
  FGX_REQUIRE(scratch && scratch_bytes >= fgx_sum_f32_scratch_bytes() && (reinterpret_cast<uintptr_t>(scratch) & 3) == 0,
              "fgx_sum_f32: scratch too small / unaligned");
  FGX_REQUIRE(n == 0 || x, "fgx_sum_f32: null pointer");
  cudaStream_t s = as_stream(stream);
  float* partial = reinterpret_cast<float*>(scratch);
  unsigned* ticket = reinterpret_cast<unsigned*>(partial + SUM_MAX_GRID);
  cudaError_t e = cudaMemsetAsync(ticket, 0, 4, s);
  if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "fgx_sum_f32: %s", cudaGetErrorString(e));
  int64_t grid = ceil_div(n, 256 * 16);
  grid = grid < 1 ? 1 : (grid > SUM_MAX_GRID ? SUM_MAX_GRID : grid);
  sum_f32_kernel<<<(int)grid, 256, 0, s>>>(x, n, out, partial, ticket);
  return check_launch("fgx_sum_f32");
}

int fgx_life_expectancy(const float* age_before, const float* age_after, const int32_t* n_removed, const float* old_value,
                        float* out, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(age_before && age_after && n_removed && old_value && out, "fgx_life_expectancy: null pointer");
  life_expectancy_kernel<<<1, 1, 0, as_stream(stream)>>>(age_before, age_after, n_removed, old_value, out);
  return check_launch("fgx_life_expectancy");
}
}
