// K2 for the wolf-sheep-grass example (examples/basic/wolf_sheep_grass/sim.ipynb): Animal
// (Random_Policy grid walk, energy, reproduction flag) and Grass steps, their create / remove /
// add / set row ops, and the cell-equality interaction.  All state is int32 / uint32 / flags, so
// parity is bit-exact.  The reference's interaction builds W x S and S x G equality matrices
// (sim.ipynb:238-290); equality of grid cells is a histogram join, done here with per-cell
// counters (one atomic per animal) instead of the O(W*S) matrices.
#include "common.cuh"
#include "geometry.cuh"
#include "threefry.cuh"

namespace fgx {

__device__ __forceinline__ int64_t ws_wrap(int64_t i, int64_t n) { return i < 0 ? i + n : i; }
__device__ __forceinline__ int64_t ws_clamp(int64_t i, int64_t n) {
  i = ws_wrap(i, n);
  return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

// vmapped Animal.create_agent (sim.ipynb:40-70)
__global__ void __launch_bounds__(256)
    animal_create_kernel(int64_t n, int32_t x_max, int32_t y_max, int32_t delta_energy, const uint2* __restrict__ ckeys,
                         const float* __restrict__ active, const fgx_animal_state_t st, float* __restrict__ age) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint2 ck = ckeys[i];
    Key key, pkey;
    split2(Key{ck.x, ck.y}, key, pkey);
    reinterpret_cast<uint2*>(st.policy_key)[i] = make_uint2(pkey.a, pkey.b);
    Key skey = key;
    if (active[i] != 0.0f) {
      skey = split_child(key, 4, 0);
      st.X_pos[i] = randint1(split_child(key, 4, 1), 0, x_max - 1);
      st.Y_pos[i] = randint1(split_child(key, 4, 2), 0, y_max - 1);
      st.energy[i] = randint1(split_child(key, 4, 3), 1, 2 * delta_energy);
    } else {
      st.X_pos[i] = -1;
      st.Y_pos[i] = -1;
      st.energy[i] = -1;
    }
    st.reproduce[i] = 0;
    reinterpret_cast<uint2*>(st.key)[i] = make_uint2(skey.a, skey.b);
    age[i] = 0.0f;
  }
}

// vmapped Animal.step_agent (sim.ipynb:72-122) with Random_Policy.step_policy (random_policy.py:17-23)
__global__ void __launch_bounds__(256)
    animal_step_kernel(const fgx_animal_params_t p, int64_t n, const float* __restrict__ active,
                       const int32_t* __restrict__ energy_in, const fgx_animal_state_t st, float* __restrict__ age) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (active[i] == 0.0f) continue;
    const uint2 pk = reinterpret_cast<const uint2*>(st.policy_key)[i];
    Key k0, k1;
    split2(Key{pk.x, pk.y}, k0, k1);           // key, subkey = split(key)
    const float action = uniform1(k0, 0.0f, 1.0f);  // drawn with `key`; `subkey` is kept
    reinterpret_cast<uint2*>(st.policy_key)[i] = make_uint2(k1.a, k1.b);
    const int32_t X = st.X_pos[i], Y = st.Y_pos[i];
    int32_t Xn = action < 0.25f ? X + 1 : X;
    Xn = (action >= 0.25f && action < 0.5f) ? X - 1 : Xn;
    int32_t Yn = (action >= 0.5f && action < 0.75f) ? Y + 1 : Y;
    Yn = action >= 0.75f ? Y - 1 : Yn;
    if (p.torous && Xn < p.x_min) Xn = p.x_max - 1;
    if (p.torous && Xn > p.x_max) Xn = p.x_min;
    if (p.torous && Yn < p.y_min) Yn = p.y_max - 1;
    if (p.torous && Yn > p.y_max) Yn = p.y_min;
    Xn = min(max(Xn, p.x_min), p.x_max - 1);
    Yn = min(max(Yn, p.y_min), p.y_max - 1);
    st.X_pos[i] = Xn;
    st.Y_pos[i] = Yn;
    st.energy[i] = st.energy[i] - 1 + p.delta_energy * energy_in[i];
    const uint2 sk = reinterpret_cast<const uint2*>(st.key)[i];
    Key nk, hk;
    split2(Key{sk.x, sk.y}, nk, hk);
    if (uniform1(hk, 0.0f, 1.0f) < p.reproduction_rate) st.reproduce[i] = 1;
    reinterpret_cast<uint2*>(st.key)[i] = make_uint2(nk.a, nk.b);
    age[i] = f_add(age[i], 1.0f);
  }
}

// Animal.remove_agent (sim.ipynb:124-133) over ids[0..k)
__global__ void __launch_bounds__(256)
    animal_remove_kernel(int64_t n, const int32_t* __restrict__ ids, const int32_t* __restrict__ kp,
                         const fgx_animal_state_t st, float* __restrict__ active, float* __restrict__ age) {
  const int64_t k = min((int64_t)*kp, n);
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < k; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = ws_wrap(ids[j], n);
    if (i < 0 || i >= n) continue;
    st.X_pos[i] = -1;
    st.Y_pos[i] = -1;
    st.energy[i] = -1;
    st.reproduce[i] = 0;
    active[i] = 0.0f;
    age[i] = 0.0f;
  }
}

// float -> int32 scatter cast of `energy / 2` (sim.ipynb:145,159): true division, truncation toward zero
__device__ __forceinline__ int32_t half_energy(int32_t e) { return (int32_t)__fdiv_rn((float)e, 2.0f); }

// Animal.add_agent (sim.ipynb:135-156) over rows [a, a+k)
__global__ void __launch_bounds__(256)
    animal_add_kernel(int64_t n, const int32_t* __restrict__ copy_ids, const int32_t* __restrict__ ap,
                      const int32_t* __restrict__ kp, const fgx_animal_state_t st, float* __restrict__ active,
                      float* __restrict__ age) {
  const int64_t a = *ap, k = *kp;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < k; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t dst = a + j;
    if (dst < 0 || dst >= n) continue;
    const int64_t src = ws_clamp(copy_ids[ws_clamp(j, n)], n);
    st.X_pos[dst] = st.X_pos[src];
    st.Y_pos[dst] = st.Y_pos[src];
    st.energy[dst] = half_energy(st.energy[src]);
    st.reproduce[dst] = 0;
    active[dst] = 1.0f;
    age[dst] = 0.0f;
  }
}

// Animal.set_agent (sim.ipynb:158-170) over ids[0..k)
__global__ void __launch_bounds__(256)
    animal_set_kernel(int64_t n, const int32_t* __restrict__ ids, const int32_t* __restrict__ kp,
                      const fgx_animal_state_t st) {
  const int64_t k = min((int64_t)*kp, n);
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < k; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = ws_wrap(ids[j], n);
    if (i < 0 || i >= n) continue;
    st.energy[i] = half_energy(st.energy[i]);
    st.reproduce[i] = 0;
  }
}

// vmapped Grass.create_agent (sim.ipynb:177-199)
__global__ void __launch_bounds__(256)
    grass_create_kernel(int64_t n, int32_t x_max, int32_t regrowth, const uint2* __restrict__ ckeys,
                        const int32_t* __restrict__ uid, int32_t* __restrict__ X, int32_t* __restrict__ Y,
                        uint8_t* __restrict__ full, int32_t* __restrict__ count_down, float* __restrict__ age) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint2 ck = ckeys[i];
    const Key key{ck.x, ck.y};
    const int32_t u = uid[i];
    X[i] = u % x_max;   // jnp.mod / floor_divide on non-negative ids
    Y[i] = u / x_max;
    const bool fg = !(uniform1(split_child(key, 3, 2), 0.0f, 1.0f) < 0.5f);
    full[i] = fg ? 1 : 0;
    count_down[i] = fg ? 0 : randint1(split_child(key, 3, 1), 1, regrowth);
    age[i] = 0.0f;
  }
}

// vmapped Grass.step_agent (sim.ipynb:202-218); no active-state check, as in the reference
__global__ void __launch_bounds__(256)
    grass_step_kernel(int64_t n, int32_t regrowth, const float* __restrict__ energy_out, uint8_t* __restrict__ full,
                      int32_t* __restrict__ count_down) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const bool eaten = energy_out[i] != 0.0f;
    int32_t cd = eaten ? regrowth : count_down[i];
    bool fg = eaten ? false : full[i] != 0;
    if (!fg) cd -= 1;
    if (cd <= 0) fg = true;
    full[i] = fg ? 1 : 0;
    count_down[i] = cd;
  }
}

// ---- interaction (sim.ipynb:238-290) as a histogram join over grid cells -----------------
// cells cover X in [-1, x_max], Y in [-1, y_max] (inactive animals sit at (-1,-1) and, as in the
// reference, match each other there); positions outside never match anything.
struct CellGrid {
  int32_t w, h;  // x_max + 2, y_max + 2
};
__device__ __forceinline__ int cell_index(const CellGrid& g, int32_t x, int32_t y) {
  const int32_t cx = x + 1, cy = y + 1;
  return (cx < 0 || cy < 0 || cx >= g.w || cy >= g.h) ? -1 : cy * g.w + cx;
}
__global__ void __launch_bounds__(256) cell_count_kernel(CellGrid g, int64_t n, const int32_t* __restrict__ X,
                                                         const int32_t* __restrict__ Y,
                                                         const uint8_t* __restrict__ flag,
                                                         uint32_t* __restrict__ count) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (flag && !flag[i]) continue;
    const int c = cell_index(g, X[i], Y[i]);
    if (c >= 0) atomicAdd(count + c, 1u);
  }
}
__global__ void __launch_bounds__(256)
    interaction_out_kernel(CellGrid g, int64_t nw, const int32_t* __restrict__ wX, const int32_t* __restrict__ wY,
                           int64_t ns, const int32_t* __restrict__ sX, const int32_t* __restrict__ sY, int64_t ng,
                           const int32_t* __restrict__ gX, const int32_t* __restrict__ gY,
                           const uint8_t* __restrict__ gfull, const uint32_t* __restrict__ wolf_cnt,
                           const uint32_t* __restrict__ sheep_cnt, const uint32_t* __restrict__ grass_cnt,
                           int32_t* __restrict__ wolves_in, int32_t* __restrict__ sheeps_in,
                           float* __restrict__ sheeps_eaten, float* __restrict__ grasses_eaten) {
  const int64_t total = nw + ns + ng;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    if (t < nw) {
      const int c = cell_index(g, wX[t], wY[t]);
      wolves_in[t] = c >= 0 ? (int32_t)sheep_cnt[c] : 0;
    } else if (t < nw + ns) {
      const int64_t s = t - nw;
      const int c = cell_index(g, sX[s], sY[s]);
      sheeps_in[s] = c >= 0 ? (int32_t)grass_cnt[c] : 0;
      sheeps_eaten[s] = (c >= 0 && wolf_cnt[c] > 0) ? 1.0f : 0.0f;
    } else {
      const int64_t q = t - nw - ns;
      const int c = cell_index(g, gX[q], gY[q]);
      grasses_eaten[q] = (gfull[q] && c >= 0 && sheep_cnt[c] > 0) ? 1.0f : 0.0f;
    }
  }
}

}  // namespace fgx

extern "C" {

int fgx_animal_create(int64_t n, int32_t x_max, int32_t y_max, int32_t delta_energy, const uint32_t* create_keys,
                      const float* active_state, const fgx_animal_state_t* state, float* age, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(state && n >= 0 && x_max > 1 && y_max > 1 && delta_energy >= 1, "fgx_animal_create: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(create_keys && active_state && age, "fgx_animal_create: null pointer");
  animal_create_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(
      n, x_max, y_max, delta_energy, reinterpret_cast<const uint2*>(create_keys), active_state, *state, age);
  return check_launch("fgx_animal_create");
}

int fgx_animal_step(const fgx_animal_params_t* p, int64_t n, const float* active_state, const int32_t* energy_in,
                    const fgx_animal_state_t* state, float* age, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(p && state && n >= 0, "fgx_animal_step: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(active_state && energy_in && age, "fgx_animal_step: null pointer");
  animal_step_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(*p, n, active_state, energy_in, *state, age);
  return check_launch("fgx_animal_step");
}

int fgx_animal_remove(int64_t n, const int32_t* remove_ids, const int32_t* k, const fgx_animal_state_t* state,
                      float* active_state, float* age, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(state && n >= 0, "fgx_animal_remove: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(remove_ids && k && active_state && age, "fgx_animal_remove: null pointer");
  animal_remove_kernel<<<grid_for(n, 256, 4), 256, 0, as_stream(stream)>>>(n, remove_ids, k, *state, active_state,
                                                                          age);
  return check_launch("fgx_animal_remove");
}

int fgx_animal_add(int64_t n, const int32_t* copy_ids, const int32_t* n_active, const int32_t* k,
                   const fgx_animal_state_t* state, float* active_state, float* age, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(state && n >= 0, "fgx_animal_add: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(copy_ids && n_active && k && active_state && age, "fgx_animal_add: null pointer");
  animal_add_kernel<<<grid_for(n, 256, 4), 256, 0, as_stream(stream)>>>(n, copy_ids, n_active, k, *state,
                                                                       active_state, age);
  return check_launch("fgx_animal_add");
}

int fgx_animal_set(int64_t n, const int32_t* set_ids, const int32_t* k, const fgx_animal_state_t* state,
                   fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(state && n >= 0, "fgx_animal_set: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(set_ids && k, "fgx_animal_set: null pointer");
  animal_set_kernel<<<grid_for(n, 256, 4), 256, 0, as_stream(stream)>>>(n, set_ids, k, *state);
  return check_launch("fgx_animal_set");
}

int fgx_grass_create(int64_t n, int32_t x_max, int32_t regrowth_time, const uint32_t* create_keys,
                     const int32_t* unique_id, int32_t* X_pos, int32_t* Y_pos, uint8_t* fully_grown,
                     int32_t* count_down, float* age, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n >= 0 && x_max >= 1 && regrowth_time >= 1, "fgx_grass_create: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(create_keys && unique_id && X_pos && Y_pos && fully_grown && count_down && age,
              "fgx_grass_create: null pointer");
  grass_create_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(
      n, x_max, regrowth_time, reinterpret_cast<const uint2*>(create_keys), unique_id, X_pos, Y_pos, fully_grown,
      count_down, age);
  return check_launch("fgx_grass_create");
}

int fgx_grass_step(int64_t n, int32_t regrowth_time, const float* energy_out, uint8_t* fully_grown,
                   int32_t* count_down, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(n >= 0, "fgx_grass_step: bad n");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(energy_out && fully_grown && count_down, "fgx_grass_step: null pointer");
  grass_step_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(n, regrowth_time, energy_out, fully_grown,
                                                                       count_down);
  return check_launch("fgx_grass_step");
}

size_t fgx_wolf_sheep_interaction_scratch_bytes(int32_t x_max, int32_t y_max) {
  if (x_max < 0 || y_max < 0) return 0;
  return (size_t)3 * (size_t)(x_max + 2) * (size_t)(y_max + 2) * sizeof(uint32_t);
}

int fgx_wolf_sheep_interaction(int32_t x_max, int32_t y_max, int64_t n_wolves, const int32_t* wX, const int32_t* wY,
                               int64_t n_sheeps, const int32_t* sX, const int32_t* sY, int64_t n_grasses,
                               const int32_t* gX, const int32_t* gY, const uint8_t* g_fully_grown,
                               int32_t* wolves_energy_in, int32_t* sheeps_energy_in, float* sheeps_eaten,
                               float* grasses_eaten, void* scratch, size_t scratch_bytes, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(x_max >= 0 && y_max >= 0 && n_wolves >= 0 && n_sheeps >= 0 && n_grasses >= 0,
              "fgx_wolf_sheep_interaction: bad sizes");
  const size_t need = fgx_wolf_sheep_interaction_scratch_bytes(x_max, y_max);
  FGX_REQUIRE(scratch && scratch_bytes >= need, "fgx_wolf_sheep_interaction: scratch too small");
  FGX_REQUIRE(n_wolves == 0 || (wX && wY && wolves_energy_in), "fgx_wolf_sheep_interaction: null wolf array");
  FGX_REQUIRE(n_sheeps == 0 || (sX && sY && sheeps_energy_in && sheeps_eaten),
              "fgx_wolf_sheep_interaction: null sheep array");
  FGX_REQUIRE(n_grasses == 0 || (gX && gY && g_fully_grown && grasses_eaten),
              "fgx_wolf_sheep_interaction: null grass array");
  cudaStream_t s = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(scratch, 0, need, s);
  if (e != cudaSuccess) return fail(FGX_ERR_CUDA, "fgx_wolf_sheep_interaction: %s", cudaGetErrorString(e));
  CellGrid g{x_max + 2, y_max + 2};
  const size_t cells = (size_t)g.w * g.h;
  uint32_t* wolf_cnt = reinterpret_cast<uint32_t*>(scratch);
  uint32_t* sheep_cnt = wolf_cnt + cells;
  uint32_t* grass_cnt = sheep_cnt + cells;
  if (n_wolves) cell_count_kernel<<<grid_for(n_wolves, 256, 8), 256, 0, s>>>(g, n_wolves, wX, wY, nullptr, wolf_cnt);
  if (n_sheeps) cell_count_kernel<<<grid_for(n_sheeps, 256, 8), 256, 0, s>>>(g, n_sheeps, sX, sY, nullptr, sheep_cnt);
  if (n_grasses)
    cell_count_kernel<<<grid_for(n_grasses, 256, 8), 256, 0, s>>>(g, n_grasses, gX, gY, g_fully_grown, grass_cnt);
  const int64_t total = n_wolves + n_sheeps + n_grasses;
  if (total)
    interaction_out_kernel<<<grid_for(total, 256, 8), 256, 0, s>>>(g, n_wolves, wX, wY, n_sheeps, sX, sY, n_grasses,
                                                                  gX, gY, g_fully_grown, wolf_cnt, sheep_cnt,
                                                                  grass_cnt, wolves_energy_in, sheeps_energy_in,
                                                                  sheeps_eaten, grasses_eaten);
  return check_launch("fgx_wolf_sheep_interaction");
}
}
