// K2 for the hello-world Dice agent (README.md:35-129 of the reference):
// fused create / step / remove / add over struct-of-arrays leaves.  HBM-bound:
// 32 B per active slot (key 8 + age 4 + active 4 read; draw 4 + key 8 + age 4
// written), 4 B (the active flag) per inactive slot.
#include "common.cuh"
#include "threefry.cuh"

namespace fgx {

// Dice.create_agent under vmap (README.md:37-53): key, subkey = split(key);
// draw = randint(subkey) for active slots else 0; state key = subkey; age = 0.
__global__ void __launch_bounds__(256) dice_create_kernel(int64_t n, int32_t sides, const uint2* __restrict__ ckeys,
                                                          const float* __restrict__ active, int32_t* __restrict__ draw,
                                                          uint2* __restrict__ key, float* __restrict__ age) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint2 ck = ckeys[i];
    Key k0, sub;
    split2(Key{ck.x, ck.y}, k0, sub);
    draw[i] = active[i] != 0.0f ? randint1(sub, 1, sides + 1) : 0;
    key[i] = make_uint2(sub.a, sub.b);
    age[i] = 0.0f;
  }
}

// Dice.step_agent under vmap (README.md:55-71); lax.cond(active_state) == (active != 0).
// The body is ~460 integer instructions of threefry per active slot behind three dependent loads (flag -> key, age):
// left alone, a quarter of every warp's time went to waiting for them (ncu: long-scoreboard stalls on exactly those
// three consumers).  The grid-stride loop is therefore software-pipelined: the flag is loaded two iterations ahead,
// key and age one iteration ahead, and the six blocks of the current slot hide both.
template <typename Index>
__global__ void __launch_bounds__(256) dice_step_kernel(int64_t n64, const RandintPlan plan,
                                                        const float* __restrict__ active, int32_t* __restrict__ draw,
                                                        uint2* __restrict__ key, float* __restrict__ age) {
  // Index = uint32_t whenever n + 2 * stride fits (every launch this library makes below 2^31 slots): the 64-bit
  // index arithmetic of the loop was 14 % of all executed instructions, paid by inactive slots too
  const Index n = (Index)n64;
  const Index stride = (Index)gridDim.x * blockDim.x;
  Index i = (Index)blockIdx.x * blockDim.x + threadIdx.x;
  // prologue: flag of iterations 0 and 1, key / age of iteration 0
  float a0 = i < n ? __ldg(active + i) : 0.0f;
  float a1 = i + stride < n ? __ldg(active + i + stride) : 0.0f;
  uint2 k0 = make_uint2(0u, 0u);
  float g0 = 0.0f;
  if (a0 != 0.0f) {
    k0 = key[i];
    g0 = age[i];
  }
  for (; i < n; i += stride) {
    const float a_cur = a0;
    const uint2 kk = k0;
    const float ag = g0;
    // advance the pipeline: key / age of the next slot (its flag is already here), flag of the one after
    a0 = a1;
    if (a0 != 0.0f) {
      k0 = key[i + stride];
      g0 = age[i + stride];
    }
    a1 = i + 2 * stride < n ? __ldg(active + i + 2 * stride) : 0.0f;
    if (a_cur != 0.0f) {
      Key k_unused, sub;
      split2(Key{kk.x, kk.y}, k_unused, sub);
      draw[i] = randint1_plan(sub, plan);  // randint(subkey, (1,), 1, sides + 1)
      key[i] = make_uint2(sub.a, sub.b);
      age[i] = __fadd_rn(ag, 1.0f);
    }
  }
}

// Dice.remove_agent over remove_ids[0..k) (README.md:83-90): draw = 0, active = False
__global__ void __launch_bounds__(256) dice_remove_kernel(int64_t n, const int32_t* __restrict__ ids,
                                                          const int32_t* __restrict__ kp, int32_t* __restrict__ draw,
                                                          float* __restrict__ active) {
  const int64_t k = min((int64_t)*kp, n);
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < k; j += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = ids[j];
    if (i < 0) i += n;            // negative indices wrap (jnp indexing)
    if (i < 0 || i >= n) continue;  // .at[i].set drops out-of-range updates
    draw[i] = 0;
    active[i] = 0.0f;
  }
}

// Dice.add_agent over slots [a, a+k) (README.md:73-81)
__global__ void __launch_bounds__(256) dice_add_kernel(int64_t n, int32_t sides, const int32_t* __restrict__ ap,
                                                       const int32_t* __restrict__ kp, int32_t* __restrict__ draw,
                                                       uint2* __restrict__ key, float* __restrict__ active) {
  const int64_t a = *ap;
  const int64_t k = *kp;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < k; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = a + j;
    if (i < 0 || i >= n) continue;
    const uint2 kk = key[i];
    Key k0, sub;
    split2(Key{kk.x, kk.y}, k0, sub);
    draw[i] = randint1(sub, 1, sides + 1);
    key[i] = make_uint2(sub.a, sub.b);
    active[i] = 1.0f;
  }
}

}  // namespace fgx

extern "C" {

int fgx_dice_create(int64_t n, int32_t sides, const uint32_t* create_keys, const float* active_state, int32_t* draw,
                    uint32_t* key, float* age, fgx_stream_t stream) {
  FGX_REQUIRE(n >= 0 && sides >= 1, "fgx_dice_create: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(create_keys && active_state && draw && key && age, "fgx_dice_create: null leaf");
  const int grid = fgx::grid_for(n, 256, 8);
  fgx::dice_create_kernel<<<grid, 256, 0, fgx::as_stream(stream)>>>(n, sides,
                                                                   reinterpret_cast<const uint2*>(create_keys),
                                                                   active_state, draw, reinterpret_cast<uint2*>(key),
                                                                   age);
  return fgx::check_launch("fgx_dice_create");
}

int fgx_dice_step(int64_t n, int32_t sides, const float* active_state, int32_t* draw, uint32_t* key, float* age,
                  fgx_stream_t stream) {
  FGX_REQUIRE(n >= 0 && sides >= 1, "fgx_dice_step: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(active_state && draw && key && age, "fgx_dice_step: null leaf");
  const int grid = fgx::grid_for(n, 256, 8);
  const fgx::RandintPlan plan = fgx::randint_plan(1, sides + 1);
  if (n + 2 * (int64_t)grid * 256 < (1ll << 32))
    fgx::dice_step_kernel<uint32_t><<<grid, 256, 0, fgx::as_stream(stream)>>>(n, plan, active_state, draw,
                                                                             reinterpret_cast<uint2*>(key), age);
  else
    fgx::dice_step_kernel<int64_t><<<grid, 256, 0, fgx::as_stream(stream)>>>(n, plan, active_state, draw,
                                                                            reinterpret_cast<uint2*>(key), age);
  return fgx::check_launch("fgx_dice_step");
}

int fgx_dice_remove(int64_t n, const int32_t* remove_ids, const int32_t* k, int32_t* draw, float* active_state,
                    fgx_stream_t stream) {
  FGX_REQUIRE(n >= 0, "fgx_dice_remove: bad n");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(remove_ids && k && draw && active_state, "fgx_dice_remove: null pointer");
  const int grid = fgx::grid_for(n, 256, 4);
  fgx::dice_remove_kernel<<<grid, 256, 0, fgx::as_stream(stream)>>>(n, remove_ids, k, draw, active_state);
  return fgx::check_launch("fgx_dice_remove");
}

int fgx_dice_add(int64_t n, int32_t sides, const int32_t* n_active, const int32_t* k, int32_t* draw, uint32_t* key,
                 float* active_state, fgx_stream_t stream) {
  FGX_REQUIRE(n >= 0 && sides >= 1, "fgx_dice_add: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(n_active && k && draw && key && active_state, "fgx_dice_add: null pointer");
  const int grid = fgx::grid_for(n, 256, 4);
  fgx::dice_add_kernel<<<grid, 256, 0, fgx::as_stream(stream)>>>(n, sides, n_active, k, draw,
                                                                reinterpret_cast<uint2*>(key), active_state);
  return fgx::check_launch("fgx_dice_add");
}
}
