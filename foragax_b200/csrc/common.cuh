// Shared host/device helpers for libforagax_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/foragax_b200.h"

#define FGX_OK 0
#define FGX_ERR_INVALID (-1)
#define FGX_ERR_CUDA (-2)
#define FGX_ERR_SCRATCH (-3)

namespace fgx {

// thread-local last-error text, read through fgx_last_error()
char* err_buf();
int fail(int code, const char* fmt, ...);
int sm_count();

inline int check_launch(const char* what) {
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(FGX_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  }
  return FGX_OK;
}

inline cudaStream_t as_stream(fgx_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// grid for a grid-stride kernel: enough CTAs to fill the machine, a multiple of the SM count
inline int grid_for(int64_t work_items, int block, int ctas_per_sm) {
  int64_t need = ceil_div(work_items, block);
  int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

#define FGX_REQUIRE(cond, ...)                                   \
  do {                                                           \
    if (!(cond)) return fgx::fail(FGX_ERR_INVALID, __VA_ARGS__); \
  } while (0)

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

}  // namespace fgx
