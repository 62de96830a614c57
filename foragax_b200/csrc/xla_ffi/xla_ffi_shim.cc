// XLA-FFI handler symbols around the plain C launchers of include/foragax_b200.h, for use from
// JAX (`jax.ffi.register_ffi_target` / `jax.ffi.ffi_call`): the "thin jax.ffi C-ABI layer" of the
// north star.
//
// NOT BUILT by the default Makefile: the XLA FFI headers (xla/ffi/api/ffi.h) ship inside jaxlib
// (`jax.ffi.include_dir()`), and neither jax nor jaxlib is installed in this image, so this file has
// never been compiled or run here.  Build where jaxlib is available:
//
//   g++ -O2 -fPIC -shared -std=c++17 -I$(python -c "import jax.ffi; print(jax.ffi.include_dir())") \
//       -I include -I /usr/local/cuda/include foragax_b200/csrc/xla_ffi/xla_ffi_shim.cc \
//       -L foragax_b200 -lforagax_b200 -Wl,-rpath,'$ORIGIN/..' -o foragax_b200/libforagax_b200_xla.so
//
// Each handler only unpacks buffers / attributes and forwards to the launcher on XLA's stream; all
// validation and error text come from the launcher (fgx_last_error()).  INTEGRATION.md shows the
// Python side.
#include <cuda_runtime.h>

#include "foragax_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

static ffi::Error Status(int rc) {
  if (rc == 0) return ffi::Error::Success();
  return ffi::Error(rc == -1 ? ffi::ErrorCode::kInvalidArgument : ffi::ErrorCode::kInternal, fgx_last_error());
}

// ---- select_agents (agent_methods.py:66-72): mask (f32/i32/u8, non-zero = selected) -> (count, perm)
static ffi::Error SelectImpl(cudaStream_t stream, ffi::AnyBuffer mask, ffi::ScratchAllocator scratch,
                             ffi::Result<ffi::Buffer<ffi::S32>> count, ffi::Result<ffi::Buffer<ffi::S32>> perm) {
  const int64_t n = static_cast<int64_t>(mask.element_count());
  fgx_predicate_t p{};
  p.n_terms = 1;
  p.lut = 0x2;  // selected iff term 0 is true
  p.term[0].data = mask.untyped_data();
  p.term[0].op = FGX_OP_NONZERO;
  p.term[0].stride = 1;
  switch (mask.element_type()) {
    case ffi::F32: p.term[0].dtype = FGX_F32; break;
    case ffi::S32: p.term[0].dtype = FGX_I32; break;
    case ffi::PRED:
    case ffi::U8: p.term[0].dtype = FGX_U8; break;
    default: return ffi::Error(ffi::ErrorCode::kInvalidArgument, "fgx_select: mask must be f32 / s32 / pred");
  }
  const size_t bytes = fgx_select_scratch_bytes(n);
  auto buf = scratch.Allocate(bytes, 256);
  if (!buf.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "fgx_select: scratch");
  return Status(fgx_select(&p, n, perm->typed_data(), count->typed_data(), *buf, bytes, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxSelect, SelectImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

// ---- the argsort of sort_agents (agent_methods.py:74-76): key f32/s32 [N] -> perm s32 [N]
static ffi::Error ArgsortImpl(cudaStream_t stream, ffi::AnyBuffer keys, ffi::ScratchAllocator scratch,
                              bool descending, ffi::Result<ffi::Buffer<ffi::S32>> perm) {
  const int64_t n = static_cast<int64_t>(keys.element_count());
  const int32_t dt = keys.element_type() == ffi::F32 ? FGX_F32 : FGX_I32;
  if (keys.element_type() != ffi::F32 && keys.element_type() != ffi::S32)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "fgx_argsort: key must be f32 / s32");
  const size_t bytes = fgx_argsort_scratch_bytes(n);
  auto buf = scratch.Allocate(bytes, 256);
  if (!buf.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "fgx_argsort: scratch");
  return Status(fgx_argsort(keys.untyped_data(), dt, n, descending ? 1 : 0, perm->typed_data(), *buf, bytes, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxArgsort, ArgsortImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Attr<bool>("descending")
                                  .Ret<ffi::Buffer<ffi::S32>>());

// ---- one leaf of the sort_agents payload (agent_methods.py:77): out[i] = leaf[perm[i]]
static ffi::Error TakeRowsImpl(cudaStream_t stream, ffi::AnyBuffer leaf, ffi::Buffer<ffi::S32> perm,
                               ffi::Result<ffi::AnyBuffer> out) {
  const int64_t n = static_cast<int64_t>(perm.element_count());
  fgx_leaf_t l{leaf.untyped_data(), out->untyped_data(), n ? static_cast<int64_t>(leaf.size_bytes()) / n : 0};
  return Status(fgx_gather_rows(&l, 1, perm.typed_data(), n, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxTakeRows, TakeRowsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::AnyBuffer>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::AnyBuffer>());

// ---- ray_generator + ray_wall_sensor + min/argmin over walls (space_methods.py:56-69,96-143)
static ffi::Error RaycastWallsImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> y,
                                   ffi::Buffer<ffi::F32> ang, ffi::Buffer<ffi::F32> active,
                                   ffi::Buffer<ffi::F32> wx1, ffi::Buffer<ffi::F32> wy1, ffi::Buffer<ffi::F32> wx2,
                                   ffi::Buffer<ffi::F32> wy2, int32_t n_rays, float ray_span, float ray_length,
                                   ffi::Result<ffi::Buffer<ffi::F32>> dist, ffi::Result<ffi::Buffer<ffi::S32>> hit) {
  return Status(fgx_raycast_walls(x.typed_data(), y.typed_data(), ang.typed_data(), active.typed_data(),
                                  static_cast<int64_t>(x.element_count()), n_rays, ray_span, ray_length,
                                  wx1.typed_data(), wy1.typed_data(), wx2.typed_data(), wy2.typed_data(),
                                  static_cast<int64_t>(wx1.element_count()), dist->typed_data(), hit->typed_data(),
                                  stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxRaycastWalls, RaycastWallsImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("n_rays")
                                  .Attr<float>("ray_span")
                                  .Attr<float>("ray_length")
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

// ---- the same reduction through per-cell wall lists (fgx_raycast_walls_grid: bit-identical, ~30x faster at 1,000
// walls).  XLA owns the scratch, so the wall lists are rebuilt per call (reuse_grid = 0); a caller that keeps the
// walls fixed can carry the scratch as an explicit operand instead and pass reuse_grid = 1.
static ffi::Error RaycastWallsGridImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> y,
                                       ffi::Buffer<ffi::F32> ang, ffi::Buffer<ffi::F32> active,
                                       ffi::Buffer<ffi::F32> wx1, ffi::Buffer<ffi::F32> wy1, ffi::Buffer<ffi::F32> wx2,
                                       ffi::Buffer<ffi::F32> wy2, ffi::ScratchAllocator scratch, int32_t n_rays,
                                       float ray_span, float ray_length, float x_min, float x_max, float y_min,
                                       float y_max, ffi::Result<ffi::Buffer<ffi::F32>> dist,
                                       ffi::Result<ffi::Buffer<ffi::S32>> hit) {
  const int64_t n_walls = static_cast<int64_t>(wx1.element_count());
  const size_t bytes = fgx_raycast_walls_grid_scratch_bytes(n_walls, x_min, x_max, y_min, y_max, ray_length);
  auto buf = scratch.Allocate(bytes, 256);
  if (!buf.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "fgx_raycast_walls_grid: scratch");
  return Status(fgx_raycast_walls_grid(x.typed_data(), y.typed_data(), ang.typed_data(), active.typed_data(),
                                       static_cast<int64_t>(x.element_count()), n_rays, ray_span, ray_length,
                                       wx1.typed_data(), wy1.typed_data(), wx2.typed_data(), wy2.typed_data(), n_walls,
                                       x_min, x_max, y_min, y_max, dist->typed_data(), hit->typed_data(), *buf, bytes,
                                       /*reuse_grid=*/0, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxRaycastWallsGrid, RaycastWallsGridImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Attr<int32_t>("n_rays")
                                  .Attr<float>("ray_span")
                                  .Attr<float>("ray_length")
                                  .Attr<float>("x_min")
                                  .Attr<float>("x_max")
                                  .Attr<float>("y_min")
                                  .Attr<float>("y_max")
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

// ---- jax.random.split on a key batch (the K4 entry the others follow)
static ffi::Error SplitImpl(cudaStream_t stream, ffi::Buffer<ffi::U32> keys, int32_t num,
                            ffi::Result<ffi::Buffer<ffi::U32>> out) {
  return Status(fgx_threefry_split(keys.typed_data(), static_cast<int64_t>(keys.element_count() / 2), num,
                                   out->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxThreefrySplit, SplitImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Attr<int32_t>("num")
                                  .Ret<ffi::Buffer<ffi::U32>>());

// ---- vmapped Dice.step_agent (README.md:55-71); outputs alias the inputs (input_output_aliases)
static ffi::Error DiceStepImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> active, ffi::Buffer<ffi::S32> draw,
                               ffi::Buffer<ffi::U32> key, ffi::Buffer<ffi::F32> age, int32_t sides,
                               ffi::Result<ffi::Buffer<ffi::S32>> draw_out, ffi::Result<ffi::Buffer<ffi::U32>> key_out,
                               ffi::Result<ffi::Buffer<ffi::F32>> age_out) {
  const int64_t n = static_cast<int64_t>(active.element_count());
  if (draw_out->typed_data() != draw.typed_data() || key_out->typed_data() != key.typed_data() ||
      age_out->typed_data() != age.typed_data())
    return ffi::Error(ffi::ErrorCode::kInvalidArgument,
                      "fgx_dice_step updates in place: call with input_output_aliases={1:0, 2:1, 3:2}");
  return Status(fgx_dice_step(n, sides, active.typed_data(), draw_out->typed_data(), key_out->typed_data(),
                              age_out->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxDiceStep, DiceStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("sides")
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());
