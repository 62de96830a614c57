// Exclusive prefix sums of per-cell counts for the two uniform-grid builds (wall lists, raycast_walls_grid.cu;
// target / sensor lists, raycast_agents.cu) as ONE cooperative launch over up to one CTA per SM.  The round-1
// builds scanned with a single 1024-thread CTA: 350 us for the 298 k cells of the C4 wall grid and 74 us for the
// agent grid, latency-bound on one SM.  Here every CTA sums a contiguous chunk, publishes its total, and after a
// grid-wide sync adds the totals of the chunks before it.
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"

namespace fgx {

struct ScanJob {
  uint32_t* count;  // in: per-cell counts; out: zero (they double as fill cursors)
  uint32_t* start;  // out: exclusive prefix, start[n] = total
};

// K arrays (1 or 2) of n cells; partial: uint32[K][gridDim.x] scratch.  total_out / capacity / overflow (optional,
// job 0 only): *total_out = total, *overflow = 1 if total > capacity.
template <int K>
__global__ void __launch_bounds__(1024) coop_scan_kernel(int n, ScanJob j0, ScanJob j1, uint32_t* __restrict__ partial,
                                                         uint32_t* __restrict__ total_out, uint32_t capacity,
                                                         int32_t* __restrict__ overflow) {
  __shared__ uint32_t wsum[K][32];
  __shared__ uint32_t red[K][32];
  const ScanJob jobs[2] = {j0, j1};
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  const int chunk = (n + (int)gridDim.x - 1) / (int)gridDim.x;
  const int per = (chunk + 1023) / 1024;  // consecutive cells per thread
  const int c0 = blockIdx.x * chunk + threadIdx.x * per;
  const int c1 = min(min(c0 + per, (int)(blockIdx.x + 1) * chunk), n);
  uint32_t tsum[K], tincl[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    uint32_t s = 0;
    for (int c = c0; c < c1; ++c) s += jobs[k].count[c];
    tsum[k] = s;
    uint32_t inc = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t o = __shfl_up_sync(0xffffffffu, inc, d);
      if (l >= d) inc += o;
    }
    tincl[k] = inc;
    if (l == 31) wsum[k][w] = inc;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    uint32_t before = 0, all = 0;
    for (int q = 0; q < 32; ++q) {
      const uint32_t v = wsum[k][q];
      all += v;
      if (q < w) before += v;
    }
    tincl[k] += before;  // inclusive prefix of this thread inside the CTA
    if (threadIdx.x == 0) __stcg(partial + k * gridDim.x + blockIdx.x, all);
  }
  cooperative_groups::this_grid().sync();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    uint32_t before = 0, all = 0;
    for (int q = threadIdx.x; q < (int)gridDim.x; q += 1024) {
      const uint32_t v = __ldcg(partial + k * gridDim.x + q);
      all += v;
      if (q < (int)blockIdx.x) before += v;
    }
    before = __reduce_add_sync(0xffffffffu, before);
    all = __reduce_add_sync(0xffffffffu, all);
    if (l == 0) { wsum[k][w] = before; red[k][w] = all; }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < K; ++k) {
    uint32_t cta_before = 0, total = 0;
    for (int q = 0; q < 32; ++q) { cta_before += wsum[k][q]; total += red[k][q]; }
    uint32_t run = cta_before + tincl[k] - tsum[k];
    for (int c = c0; c < c1; ++c) {
      const uint32_t v = jobs[k].count[c];
      jobs[k].start[c] = run;
      jobs[k].count[c] = 0;
      run += v;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      jobs[k].start[n] = total;
      if (k == 0) {
        if (total_out) *total_out = total;
        if (overflow && total > capacity) *overflow = 1;
      }
    }
  }
}

// bytes of `partial` for K arrays
inline size_t coop_scan_scratch_bytes(int k) { return (size_t)k * 1024 * sizeof(uint32_t); }

template <int K>
inline int launch_coop_scan(const char* what, int n, ScanJob j0, ScanJob j1, uint32_t* partial, uint32_t* total_out,
                            uint32_t capacity, int32_t* overflow, cudaStream_t s) {
  int grid = (n + 1023) / 1024;
  if (grid > sm_count()) grid = sm_count();
  if (grid > 1024) grid = 1024;
  if (grid < 1) grid = 1;
  void* args[] = {&n, &j0, &j1, &partial, &total_out, &capacity, &overflow};
  cudaError_t e = cudaLaunchCooperativeKernel((const void*)coop_scan_kernel<K>, dim3((unsigned)grid), dim3(1024), args, 0, s);
  if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "%s: cooperative scan launch: %s", what, cudaGetErrorString(e));
  return FGX_OK;
}

}  // namespace fgx
