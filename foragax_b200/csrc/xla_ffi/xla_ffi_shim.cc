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
#include <string>

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

// =====================================================================================================
// Round 2: the rest of the per-tick path.  In-place launchers are exposed the XLA way: every updated leaf is a
// result that must ALIAS its operand (ffi_call(..., input_output_aliases={operand index: result index})); the
// handlers verify the aliasing instead of copying.  Leaf lists ride in RemainingArgs / RemainingRets.
// =====================================================================================================

template <typename T>
static T* ArgPtr(ffi::RemainingArgs& a, size_t i) {
  auto b = a.get<ffi::AnyBuffer>(i);
  return b.has_value() ? static_cast<T*>(b->untyped_data()) : nullptr;
}
static void* RetPtr(ffi::RemainingRets& r, size_t i) {
  auto b = r.get<ffi::AnyBuffer>(i);
  return b.has_value() ? (*b)->untyped_data() : nullptr;
}
static ffi::Error NotAliased(const char* op) {
  return ffi::Error(ffi::ErrorCode::kInvalidArgument,
                    std::string(op) + " updates in place: pass input_output_aliases for every result");
}

// ---- jax.random.uniform / randint on a key batch
static ffi::Error UniformImpl(cudaStream_t stream, ffi::Buffer<ffi::U32> keys, int32_t n, float minval, float maxval,
                              ffi::Result<ffi::Buffer<ffi::F32>> out) {
  return Status(fgx_threefry_uniform(keys.typed_data(), static_cast<int64_t>(keys.element_count() / 2), n, minval,
                                     maxval, out->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxThreefryUniform, UniformImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Attr<int32_t>("n")
                                  .Attr<float>("minval")
                                  .Attr<float>("maxval")
                                  .Ret<ffi::Buffer<ffi::F32>>());

static ffi::Error RandintImpl(cudaStream_t stream, ffi::Buffer<ffi::U32> keys, int32_t n, int32_t minval,
                              int32_t maxval, ffi::Result<ffi::Buffer<ffi::S32>> out) {
  return Status(fgx_threefry_randint(keys.typed_data(), static_cast<int64_t>(keys.element_count() / 2), n, minval,
                                     maxval, out->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxThreefryRandint, RandintImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Attr<int32_t>("n")
                                  .Attr<int32_t>("minval")
                                  .Attr<int32_t>("maxval")
                                  .Ret<ffi::Buffer<ffi::S32>>());

// ---- jnp.sum(active_state, dtype=int32) (agent_methods.py:27)
static ffi::Error CountActiveImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> active,
                                  ffi::Result<ffi::Buffer<ffi::S32>> out) {
  return Status(fgx_count_active(active.typed_data(), static_cast<int64_t>(active.element_count()), out->typed_data(),
                                 stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxCountActive, CountActiveImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());

// ---- remove_agents / add_agents with the Dice callbacks (README.md:73-90); results alias draw, (key,) active
static ffi::Error DiceRemoveImpl(cudaStream_t stream, ffi::Buffer<ffi::S32> remove_ids, ffi::Buffer<ffi::S32> k,
                                 ffi::Buffer<ffi::S32> draw, ffi::Buffer<ffi::F32> active,
                                 ffi::Result<ffi::Buffer<ffi::S32>> draw_out,
                                 ffi::Result<ffi::Buffer<ffi::F32>> active_out) {
  if (draw_out->typed_data() != draw.typed_data() || active_out->typed_data() != active.typed_data())
    return NotAliased("fgx_dice_remove");
  return Status(fgx_dice_remove(static_cast<int64_t>(active.element_count()), remove_ids.typed_data(), k.typed_data(),
                                draw_out->typed_data(), active_out->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxDiceRemove, DiceRemoveImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

static ffi::Error DiceAddImpl(cudaStream_t stream, ffi::Buffer<ffi::S32> n_active, ffi::Buffer<ffi::S32> k,
                              ffi::Buffer<ffi::S32> draw, ffi::Buffer<ffi::U32> key, ffi::Buffer<ffi::F32> active,
                              int32_t sides, ffi::Result<ffi::Buffer<ffi::S32>> draw_out,
                              ffi::Result<ffi::Buffer<ffi::U32>> key_out,
                              ffi::Result<ffi::Buffer<ffi::F32>> active_out) {
  if (draw_out->typed_data() != draw.typed_data() || key_out->typed_data() != key.typed_data() ||
      active_out->typed_data() != active.typed_data())
    return NotAliased("fgx_dice_add");
  return Status(fgx_dice_add(static_cast<int64_t>(active.element_count()), sides, n_active.typed_data(), k.typed_data(),
                             draw_out->typed_data(), key_out->typed_data(), active_out->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxDiceAdd, DiceAddImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("sides")
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .Ret<ffi::Buffer<ffi::U32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>());

// ---- Forager leaves.  RemainingArgs layout (18 operands): the 11 state leaves in the order of
// fgx_forager_state_t, then J, tau, E, B, D, Z, then age.  RemainingRets (13 results): the 11 state leaves, Z, age
// -- each aliased to its operand.
static constexpr size_t kForagerArgs = 18, kForagerRets = 13;

static ffi::Error ForagerLeaves(const char* op, ffi::RemainingArgs& args, ffi::RemainingRets& rets,
                                fgx_forager_state_t* st, fgx_wc_policy_t* pol, float** age) {
  if (args.size() != kForagerArgs || rets.size() != kForagerRets)
    return ffi::Error(ffi::ErrorCode::kInvalidArgument,
                      std::string(op) + ": expected 18 leaf operands (11 state, J, tau, E, B, D, Z, age) and 13 results");
  float** s = reinterpret_cast<float**>(st);
  for (size_t i = 0; i < 11; ++i) {
    s[i] = ArgPtr<float>(args, i);
    if (RetPtr(rets, i) != s[i]) return NotAliased(op);
  }
  pol->J = ArgPtr<float>(args, 11);
  pol->tau = ArgPtr<float>(args, 12);
  pol->E = ArgPtr<float>(args, 13);
  pol->B = ArgPtr<float>(args, 14);
  pol->D = ArgPtr<float>(args, 15);
  pol->Z = ArgPtr<float>(args, 16);
  *age = ArgPtr<float>(args, 17);
  if (RetPtr(rets, 11) != pol->Z || RetPtr(rets, 12) != *age) return NotAliased(op);
  return ffi::Error::Success();
}

// ---- step_agents + vmapped Forager.step_agent with WilsonCowan.step_policy (sim.py:74-136, wilson_cowan.py:34-51)
static ffi::Error ForagerStepImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> active, ffi::Buffer<ffi::F32> obs,
                                  ffi::Buffer<ffi::F32> energy_in, ffi::RemainingArgs args, int32_t n_neurons,
                                  int32_t n_obs, int32_t n_act, float dt, float policy_dt, float action_scale,
                                  float x_min, float x_max, float y_min, float y_max, float repr_thresh,
                                  float die_thresh, float energy_min, float energy_max, float energy_cost_dt,
                                  int32_t clip_min, ffi::RemainingRets rets) {
  fgx_forager_params_t p{n_neurons, n_obs,       n_act,      dt,         policy_dt,  action_scale,   x_min,   x_max,
                         y_min,     y_max,       repr_thresh, die_thresh, energy_min, energy_max, energy_cost_dt, clip_min};
  fgx_forager_state_t st{};
  fgx_wc_policy_t pol{};
  float* age = nullptr;
  ffi::Error e = ForagerLeaves("fgx_forager_step", args, rets, &st, &pol, &age);
  if (!e.success()) return e;
  return Status(fgx_forager_step(&p, static_cast<int64_t>(active.element_count()), active.typed_data(),
                                 obs.typed_data(), energy_in.typed_data(), &st, &pol, age, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxForagerStep, ForagerStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .RemainingArgs()
                                  .Attr<int32_t>("n_neurons")
                                  .Attr<int32_t>("n_obs")
                                  .Attr<int32_t>("n_act")
                                  .Attr<float>("dt")
                                  .Attr<float>("policy_dt")
                                  .Attr<float>("action_scale")
                                  .Attr<float>("x_min")
                                  .Attr<float>("x_max")
                                  .Attr<float>("y_min")
                                  .Attr<float>("y_max")
                                  .Attr<float>("repr_thresh")
                                  .Attr<float>("die_thresh")
                                  .Attr<float>("energy_min")
                                  .Attr<float>("energy_max")
                                  .Attr<float>("energy_cost_dt")
                                  .Attr<int32_t>("clip_min")
                                  .RemainingRets());

// ---- the tick as one kernel: sensor (wall rays) -> step_agents (sim.py:652-660) = fgx_forager_sense_step.  Operands:
// active, energy_in, the four wall arrays, then the 18 Forager leaves; results: dist, hit, then the 13 aliased leaves.
// XLA owns the scratch, so the wall lists are rebuilt per call (reuse_grid = 0).
static ffi::Error ForagerSenseStepImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> active, ffi::Buffer<ffi::F32> energy_in,
                                       ffi::Buffer<ffi::F32> wx1, ffi::Buffer<ffi::F32> wy1, ffi::Buffer<ffi::F32> wx2,
                                       ffi::Buffer<ffi::F32> wy2, ffi::RemainingArgs args, ffi::ScratchAllocator scratch,
                                       int32_t n_neurons, int32_t n_obs, int32_t n_act, float dt, float policy_dt,
                                       float action_scale, float x_min, float x_max, float y_min, float y_max,
                                       float repr_thresh, float die_thresh, float energy_min, float energy_max,
                                       float energy_cost_dt, int32_t clip_min, int32_t n_rays, float ray_span,
                                       float ray_length, ffi::Result<ffi::Buffer<ffi::F32>> dist,
                                       ffi::Result<ffi::Buffer<ffi::S32>> hit, ffi::RemainingRets rets) {
  fgx_forager_params_t p{n_neurons, n_obs,       n_act,      dt,         policy_dt,  action_scale,   x_min,   x_max,
                         y_min,     y_max,       repr_thresh, die_thresh, energy_min, energy_max, energy_cost_dt, clip_min};
  fgx_forager_state_t st{};
  fgx_wc_policy_t pol{};
  float* age = nullptr;
  ffi::Error e = ForagerLeaves("fgx_forager_sense_step", args, rets, &st, &pol, &age);
  if (!e.success()) return e;
  const int64_t n_walls = static_cast<int64_t>(wx1.element_count());
  const size_t bytes = fgx_raycast_walls_grid_scratch_bytes(n_walls, x_min, x_max, y_min, y_max, ray_length);
  auto buf = scratch.Allocate(bytes, 256);
  if (!buf.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "fgx_forager_sense_step: scratch");
  return Status(fgx_forager_sense_step(&p, static_cast<int64_t>(active.element_count()), active.typed_data(),
                                       energy_in.typed_data(), &st, &pol, age, n_rays, ray_span, ray_length,
                                       wx1.typed_data(), wy1.typed_data(), wx2.typed_data(), wy2.typed_data(), n_walls,
                                       x_min, x_max, y_min, y_max, *buf, bytes, /*reuse_grid=*/0, dist->typed_data(),
                                       hit->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxForagerSenseStep, ForagerSenseStepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .RemainingArgs()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Attr<int32_t>("n_neurons")
                                  .Attr<int32_t>("n_obs")
                                  .Attr<int32_t>("n_act")
                                  .Attr<float>("dt")
                                  .Attr<float>("policy_dt")
                                  .Attr<float>("action_scale")
                                  .Attr<float>("x_min")
                                  .Attr<float>("x_max")
                                  .Attr<float>("y_min")
                                  .Attr<float>("y_max")
                                  .Attr<float>("repr_thresh")
                                  .Attr<float>("die_thresh")
                                  .Attr<float>("energy_min")
                                  .Attr<float>("energy_max")
                                  .Attr<float>("energy_cost_dt")
                                  .Attr<int32_t>("clip_min")
                                  .Attr<int32_t>("n_rays")
                                  .Attr<float>("ray_span")
                                  .Attr<float>("ray_length")
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>()
                                  .RemainingRets());

// ---- remove_agents / set_agents with the Forager callbacks (sim.py:139-151,213-225)
static ffi::Error ForagerRemoveImpl(cudaStream_t stream, ffi::Buffer<ffi::S32> remove_ids, ffi::Buffer<ffi::S32> k,
                                    ffi::Buffer<ffi::F32> active, ffi::RemainingArgs args, int32_t n_neurons,
                                    ffi::Result<ffi::Buffer<ffi::F32>> active_out, ffi::RemainingRets rets) {
  fgx_forager_state_t st{};
  fgx_wc_policy_t pol{};
  float* age = nullptr;
  ffi::Error e = ForagerLeaves("fgx_forager_remove", args, rets, &st, &pol, &age);
  if (!e.success()) return e;
  if (active_out->typed_data() != active.typed_data()) return NotAliased("fgx_forager_remove");
  return Status(fgx_forager_remove(static_cast<int64_t>(active.element_count()), n_neurons, remove_ids.typed_data(),
                                   k.typed_data(), &st, pol.Z, active_out->typed_data(), age, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxForagerRemove, ForagerRemoveImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .RemainingArgs()
                                  .Attr<int32_t>("n_neurons")
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .RemainingRets());

static ffi::Error ForagerSetImpl(cudaStream_t stream, ffi::Buffer<ffi::S32> set_ids, ffi::Buffer<ffi::S32> k,
                                 ffi::Buffer<ffi::F32> active, ffi::RemainingArgs args, ffi::RemainingRets rets) {
  fgx_forager_state_t st{};
  fgx_wc_policy_t pol{};
  float* age = nullptr;
  ffi::Error e = ForagerLeaves("fgx_forager_set", args, rets, &st, &pol, &age);
  if (!e.success()) return e;
  return Status(fgx_forager_set(static_cast<int64_t>(active.element_count()), set_ids.typed_data(), k.typed_data(), &st,
                                stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxForagerSet, ForagerSetImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .RemainingArgs()
                                  .RemainingRets());

// ---- add_agents with Forager.add_agent (sim.py:154-210): the serial key chain runs in a one-thread pre-pass
// inside the launcher (chain_scratch uint32[n, 2]: the chain key of every added row, from the scratch allocator)
static ffi::Error ForagerAddImpl(cudaStream_t stream, ffi::Buffer<ffi::S32> good_ids, ffi::Buffer<ffi::S32> n_active,
                                 ffi::Buffer<ffi::S32> k, ffi::Buffer<ffi::U32> key_in, ffi::Buffer<ffi::F32> active,
                                 ffi::RemainingArgs args, ffi::ScratchAllocator scratch, int32_t n_neurons,
                                 int32_t n_obs, int32_t n_act, float noise_scale, int32_t adaptive_noise,
                                 ffi::Result<ffi::Buffer<ffi::U32>> key_out,
                                 ffi::Result<ffi::Buffer<ffi::F32>> active_out, ffi::RemainingRets rets) {
  fgx_forager_state_t st{};
  fgx_wc_policy_t pol{};
  float* age = nullptr;
  ffi::Error e = ForagerLeaves("fgx_forager_add", args, rets, &st, &pol, &age);
  if (!e.success()) return e;
  if (active_out->typed_data() != active.typed_data()) return NotAliased("fgx_forager_add");
  const int64_t n = static_cast<int64_t>(active.element_count());
  auto buf = scratch.Allocate(static_cast<size_t>(n) * 2 * sizeof(uint32_t) + 64, 256);  // chain_scratch uint32[n, 2]
  if (!buf.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "fgx_forager_add: scratch");
  return Status(fgx_forager_add(n, n_neurons, n_obs, n_act, good_ids.typed_data(), n_active.typed_data(),
                                k.typed_data(), key_in.typed_data(), key_out->typed_data(), noise_scale, adaptive_noise,
                                &st, &pol, active_out->typed_data(), age, static_cast<uint32_t*>(*buf), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxForagerAdd, ForagerAddImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::U32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .RemainingArgs()
                                  .Ctx<ffi::ScratchAllocator>()
                                  .Attr<int32_t>("n_neurons")
                                  .Attr<int32_t>("n_obs")
                                  .Attr<int32_t>("n_act")
                                  .Attr<float>("noise_scale")
                                  .Attr<int32_t>("adaptive_noise")
                                  .Ret<ffi::Buffer<ffi::U32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .RemainingRets());

// ---- sensor_model (EcoEvoJax/sim.py:388-468): foragers x rays x [foragers, resources, walls] -> obs [F, 3 R]
static ffi::Error SensorModelImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> fx, ffi::Buffer<ffi::F32> fy,
                                  ffi::Buffer<ffi::F32> fang, ffi::Buffer<ffi::F32> frad, ffi::Buffer<ffi::F32> fenergy,
                                  ffi::Buffer<ffi::F32> factive, ffi::Buffer<ffi::S32> ftype, ffi::Buffer<ffi::F32> rx,
                                  ffi::Buffer<ffi::F32> ry, ffi::Buffer<ffi::F32> rrad, ffi::Buffer<ffi::F32> renergy,
                                  ffi::Buffer<ffi::F32> ractive, ffi::Buffer<ffi::S32> rtype, ffi::Buffer<ffi::F32> wx1,
                                  ffi::Buffer<ffi::F32> wy1, ffi::Buffer<ffi::F32> wx2, ffi::Buffer<ffi::F32> wy2,
                                  int32_t n_rays, float ray_span, float ray_length,
                                  ffi::Result<ffi::Buffer<ffi::F32>> obs) {
  return Status(fgx_sensor_model(static_cast<int64_t>(fx.element_count()), fx.typed_data(), fy.typed_data(),
                                 fang.typed_data(), frad.typed_data(), fenergy.typed_data(), factive.typed_data(),
                                 ftype.typed_data(), static_cast<int64_t>(rx.element_count()), rx.typed_data(),
                                 ry.typed_data(), rrad.typed_data(), renergy.typed_data(), ractive.typed_data(),
                                 rtype.typed_data(), static_cast<int64_t>(wx1.element_count()), wx1.typed_data(),
                                 wy1.typed_data(), wx2.typed_data(), wy2.typed_data(), n_rays, ray_span, ray_length,
                                 obs->typed_data(), stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxSensorModel, SensorModelImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::S32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Attr<int32_t>("n_rays")
                                  .Attr<float>("ray_span")
                                  .Attr<float>("ray_length")
                                  .Ret<ffi::Buffer<ffi::F32>>());

// ---- ray_generator + ray_agent_sensor + min/argmin over target circles through a uniform grid
// (space_methods.py:71-94, sim.py:444-447); targets as separate x / y / rad / active arrays
static ffi::Error RaycastAgentsImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> x, ffi::Buffer<ffi::F32> y,
                                    ffi::Buffer<ffi::F32> ang, ffi::Buffer<ffi::F32> active, ffi::Buffer<ffi::F32> tx,
                                    ffi::Buffer<ffi::F32> ty, ffi::Buffer<ffi::F32> trad, ffi::Buffer<ffi::F32> tactive,
                                    ffi::ScratchAllocator scratch, int32_t n_rays, float ray_span, float ray_length,
                                    float x_min, float x_max, float y_min, float y_max, float max_rad,
                                    ffi::Result<ffi::Buffer<ffi::F32>> dist, ffi::Result<ffi::Buffer<ffi::S32>> hit) {
  const int64_t n = static_cast<int64_t>(x.element_count()), m = static_cast<int64_t>(tx.element_count());
  const size_t bytes = fgx_raycast_agents_scratch_bytes(n, m, x_min, x_max, y_min, y_max, ray_length, max_rad);
  auto buf = scratch.Allocate(bytes, 256);
  if (!buf.has_value()) return ffi::Error(ffi::ErrorCode::kResourceExhausted, "fgx_raycast_agents: scratch");
  return Status(fgx_raycast_agents(x.typed_data(), y.typed_data(), ang.typed_data(), active.typed_data(), n, n_rays,
                                   ray_span, ray_length, tx.typed_data(), ty.typed_data(), trad.typed_data(),
                                   tactive.typed_data(), /*target_stride=*/1, m, x_min, x_max, y_min, y_max, max_rad,
                                   /*merge_walls=*/0, dist->typed_data(), hit->typed_data(), *buf, bytes, stream));
}
XLA_FFI_DEFINE_HANDLER_SYMBOL(FgxRaycastAgents, RaycastAgentsImpl,
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
                                  .Attr<float>("max_rad")
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::S32>>());
