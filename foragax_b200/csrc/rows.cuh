// Row movement shared by rows.cu (gather by a permutation) and sort.cu (the payload half of sort_agents
// fused into the two-valued partition): the leaf table passed to the kernels and its host-side builder.
#pragma once
#include "common.cuh"

namespace fgx {

struct LeafTable {
  fgx_leaf_t leaf[FGX_MAX_LEAVES];
  int32_t vec[FGX_MAX_LEAVES];  // bytes per copy chunk: 16 / 8 / 4 / 2 / 1
};

// widest vector (bytes) that divides the row size and both base addresses
inline int leaf_vec(const fgx_leaf_t& lf) {
  const uintptr_t bits = reinterpret_cast<uintptr_t>(lf.src) | reinterpret_cast<uintptr_t>(lf.dst) |
                         (uintptr_t)lf.row_bytes;
  int v = 16;
  while (v > 1 && (bits & (uintptr_t)(v - 1))) v >>= 1;
  return v;
}

// Validates the leaves and splits them into narrow rows (<= 8 vector chunks: one thread per row) and wide rows
// (per-agent weight matrices: one warp per row).  Returns an fgx status code.
int classify_leaves(const char* what, const fgx_leaf_t* leaves_host, int32_t n_leaves, LeafTable* narrow, int* nn,
                    LeafTable* wide, int* nw);

// dst[i] = src[clamp(perm[i])] for i < n, rows taken from leaves of n_src rows.  `skip` (device int32, may be
// null): the narrow launch returns at once when *skip != 0 (its rows were already moved by the fused sort).
int launch_gather(const LeafTable& narrow, int nn, const LeafTable& wide, int nw, const int32_t* perm, int64_t n,
                  int64_t n_src, const int32_t* skip_narrow, cudaStream_t s);

}  // namespace fgx
