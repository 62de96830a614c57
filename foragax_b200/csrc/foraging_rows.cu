// Row operations of the foraging example (EcoEvoJax/sim.py): create / remove / add / set for
// Forager and Resource -- the callbacks the reference runs inside serial lax.fori_loops
// (agent_methods.py:29-35,43-49,55-61).  The rows touched by one call are distinct, so the data
// part runs in parallel; the PRNG key that add_agents threads through its loop
// (agent_methods.py:29-35) is a serial hash chain, resolved by a one-thread pre-pass whose
// per-row keys feed the parallel kernels.
#include "common.cuh"
#include "geometry.cuh"
#include "threefry.cuh"

namespace fgx {

__device__ __forceinline__ float clipf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ int64_t wrap_row(int64_t i, int64_t n) { return i < 0 ? i + n : i; }
__device__ __forceinline__ int64_t clamp_row(int64_t i, int64_t n) {
  i = wrap_row(i, n);
  return i < 0 ? 0 : (i >= n ? n - 1 : i);
}

// ---- Forager.create_agent under vmap (sim.py:27-71) ---------------------------------
// scalar state + the five policy init keys (wilson_cowan.py:19) written leaf-major to `ik`
__global__ void __launch_bounds__(256)
    forager_create_kernel(const fgx_forager_params_t p, int64_t n, const uint2* __restrict__ ckeys,
                          const float* __restrict__ active, const fgx_forager_state_t st, float* __restrict__ Z,
                          uint2* __restrict__ ik, float* __restrict__ age) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint2 ck = ckeys[i];
    Key key, pkey;
    split2(Key{ck.x, ck.y}, key, pkey);
#pragma unroll
    for (int l = 0; l < 5; ++l) {
      const Key c = split_child(pkey, 6, 1 + l);
      ik[(int64_t)l * n + i] = make_uint2(c.a, c.b);
    }
    for (int j = 0; j < p.n_neurons; ++j) Z[i * p.n_neurons + j] = 0.0f;
    const bool act = active[i] != 0.0f;
    if (act) {
      // key, *keys = split(key, 8); bounds as sim.py:39,43 (float64 on the host, rounded once)
      const float Xv = uniform1(split_child(key, 8, 2), -1.0f, 1.0f);
      const float Yv = uniform1(split_child(key, 8, 5), -1.0f, 1.0f);
      const float en = uniform1(split_child(key, 8, 7), 5.0f, 10.0f);
      st.X[i] = uniform1(split_child(key, 8, 1), p.x_min, p.x_max);  // create bounds are passed in x_min/x_max
      st.X_vel[i] = Xv;
      st.X_acc[i] = uniform1(split_child(key, 8, 3), -1.0f, 1.0f);
      st.Y[i] = uniform1(split_child(key, 8, 4), p.y_min, p.y_max);
      st.Y_vel[i] = Yv;
      st.Y_acc[i] = uniform1(split_child(key, 8, 6), -1.0f, 1.0f);
      st.ang[i] = atan2f(Yv, Xv);
      st.energy[i] = en;
      st.rad[i] = en;
      st.time_above_rep_thresh[i] = 0.0f;
      st.time_below_die_thresh[i] = 0.0f;
    } else {
      st.X[i] = -1.0f; st.X_vel[i] = 0.0f; st.X_acc[i] = 0.0f;
      st.Y[i] = -1.0f; st.Y_vel[i] = 0.0f; st.Y_acc[i] = 0.0f;
      st.ang[i] = 0.0f; st.energy[i] = -1.0f; st.rad[i] = 0.0f;
      st.time_above_rep_thresh[i] = -1.0f; st.time_below_die_thresh[i] = -1.0f;
    }
    age[i] = 0.0f;
  }
}

// ---- Forager.remove_agent (sim.py:139-151) over ids[0..k) ------------------------------
__global__ void __launch_bounds__(256)
    forager_remove_kernel(int64_t n, int n_neurons, const int32_t* __restrict__ ids, const int32_t* __restrict__ kp,
                          const fgx_forager_state_t st, float* __restrict__ Z, float* __restrict__ active,
                          float* __restrict__ age) {
  const int64_t k = min((int64_t)*kp, n);
  // one warp per removed row: lanes clear Z
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t j = warp; j < k; j += nwarps) {
    const int64_t i = wrap_row(ids[j], n);
    if (i < 0 || i >= n) continue;
    for (int t = lane; t < n_neurons; t += 32) Z[i * n_neurons + t] = 0.0f;
    if (lane == 0) {
      st.X[i] = -1.0f; st.X_vel[i] = 0.0f; st.X_acc[i] = 0.0f;
      st.Y[i] = -1.0f; st.Y_vel[i] = 0.0f; st.Y_acc[i] = 0.0f;
      st.ang[i] = -1.0f; st.energy[i] = -1.0f; st.rad[i] = 0.0f;
      st.time_above_rep_thresh[i] = -1.0f; st.time_below_die_thresh[i] = -1.0f;
      active[i] = 0.0f;
      age[i] = 0.0f;
    }
  }
}

// ---- the serial key chain of add_agents (agent_methods.py:29-35) -----------------------
// chain[j] = the key handed to add_func for the j-th added row; *key_out = the key returned.
// MODE 0: Forager.add_agent  key, *ks = split(key, 6)            (sim.py:181)
// MODE 1: Resource.add_agent key, *ks = split(key, 4); key, pk = split(key)  (sim.py:322,335)
template <int MODE>
__global__ void key_chain_kernel(const uint32_t* __restrict__ key_in, const int32_t* __restrict__ kp, int64_t max_k,
                                 uint2* __restrict__ chain, uint32_t* __restrict__ key_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Key key{key_in[0], key_in[1]};
  const int64_t k = min((int64_t)*kp, max_k);
  for (int64_t j = 0; j < k; ++j) {
    chain[j] = make_uint2(key.a, key.b);
    if (MODE == 0) {
      key = split_child(key, 6, 0);
    } else {
      key = split_child(key, 4, 0);
      key = split_child(key, 2, 0);
    }
  }
  key_out[0] = key.a;
  key_out[1] = key.b;
}

// ---- Forager.add_agent (sim.py:154-210): scalar part, one thread per added row ---------------
__global__ void __launch_bounds__(256)
    forager_add_state_kernel(int64_t n, int n_neurons, const int32_t* __restrict__ good_ids,
                             const int32_t* __restrict__ ap, const int32_t* __restrict__ kp,
                             const fgx_forager_state_t st, float* __restrict__ active, float* __restrict__ age) {
  const int64_t a = *ap, k = *kp;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < k; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t dst = a + j;
    if (dst < 0 || dst >= n) continue;
    const int64_t src = clamp_row(good_ids[clamp_row(j, n)], n);
    const float en = __fdiv_rn(st.energy[src], 2.0f);
    st.X[dst] = st.X[src]; st.X_vel[dst] = 0.0f; st.X_acc[dst] = 0.0f;
    st.Y[dst] = st.Y[src]; st.Y_vel[dst] = 0.0f; st.Y_acc[dst] = 0.0f;
    st.ang[dst] = 0.0f; st.energy[dst] = en; st.rad[dst] = en;
    st.time_above_rep_thresh[dst] = 0.0f; st.time_below_die_thresh[dst] = 0.0f;
    active[dst] = 1.0f;
  }
}
// policy leaves: new = parent + noise_scale * normal(noise_key_l)  (sim.py:181-200); grid.y = leaf.
// age[dst] is reset by this kernel's leaf-0 lanes AFTER noise_scale has been read from the parent.
struct MutateArgs {
  float* leaf[5];
  int size[5];
};
__global__ void __launch_bounds__(256)
    forager_add_policy_kernel(int64_t n, const MutateArgs m, const int32_t* __restrict__ good_ids,
                              const int32_t* __restrict__ ap, const int32_t* __restrict__ kp,
                              const uint2* __restrict__ chain, const float* __restrict__ age, float noise_const,
                              int adaptive) {
  const int l = blockIdx.y;
  const int sz = m.size[l];
  float* leaf = m.leaf[l];
  const int64_t a = *ap, k = *kp;
  const int64_t total = k * sz;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = t / sz;
    const int e = (int)(t - j * sz);
    const int64_t dst = a + j;
    if (dst < 0 || dst >= n) continue;
    const int64_t src = clamp_row(good_ids[clamp_row(j, n)], n);
    float ns = noise_const;
    if (adaptive) ns = clipf(__fdiv_rn(0.1f, f_add(age[src], 0.01f)), 0.001f, 0.1f);  // notebook variant
    const uint2 ck = chain[j];
    const Key nk = split_child(Key{ck.x, ck.y}, 6, 1 + l);
    const float noise = bits_to_normal(bits_elem(nk, (uint32_t)sz, (uint32_t)e));
    leaf[dst * sz + e] = f_add(leaf[src * sz + e], f_mul(ns, noise));
  }
}
__global__ void __launch_bounds__(256) reset_age_rows_kernel(int64_t n, const int32_t* __restrict__ ap,
                                                             const int32_t* __restrict__ kp, float* __restrict__ age) {
  const int64_t a = *ap, k = *kp;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < k; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t dst = a + j;
    if (dst >= 0 && dst < n) age[dst] = 0.0f;
  }
}

// ---- Forager.set_agent (sim.py:213-225) over ids[0..k) -----------------------------------
__global__ void __launch_bounds__(256)
    forager_set_kernel(int64_t n, const int32_t* __restrict__ ids, const int32_t* __restrict__ kp,
                       const fgx_forager_state_t st) {
  const int64_t k = min((int64_t)*kp, n);
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < k; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = wrap_row(ids[j], n);
    if (i < 0 || i >= n) continue;
    const float en = __fdiv_rn(st.energy[i], 2.0f);
    st.energy[i] = en;
    st.rad[i] = en;
    st.time_above_rep_thresh[i] = 0.0f;
  }
}

// ---- Resource (sim.py:228-355) ---------------------------------------------------------
__global__ void __launch_bounds__(256)
    resource_create_kernel(int64_t n, float xlo, float xhi, float ylo, float yhi, const uint2* __restrict__ ckeys,
                           const float* __restrict__ active, const fgx_resource_state_t st,
                           float* __restrict__ growth, float* __restrict__ decay, float* __restrict__ age) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint2 ck = ckeys[i];
    Key key, pkey;
    split2(Key{ck.x, ck.y}, key, pkey);
    const float gr = uniform1(pkey, 0.01f, 1.0f);
    growth[i] = gr;
    decay[i] = f_mul(0.2f, gr);
    Key skey = key;
    if (active[i] != 0.0f) {
      skey = split_child(key, 4, 0);
      const float en = uniform1(split_child(key, 4, 3), 3.0f, 4.0f);
      st.X[i] = uniform1(split_child(key, 4, 1), xlo, xhi);
      st.Y[i] = uniform1(split_child(key, 4, 2), ylo, yhi);
      st.energy[i] = en;
      st.rad[i] = en;
      st.time_above_blossom[i] = 0.0f;
    } else {
      st.X[i] = -1.0f; st.Y[i] = -1.0f; st.energy[i] = 0.0f; st.rad[i] = 0.0f;
      st.time_above_blossom[i] = -1.0f;
    }
    st.blossom[i] = 0;
    reinterpret_cast<uint2*>(st.key)[i] = make_uint2(skey.a, skey.b);
    age[i] = 0.0f;
  }
}

__global__ void __launch_bounds__(256)
    resource_remove_kernel(int64_t n, const int32_t* __restrict__ ids, const int32_t* __restrict__ kp,
                           const fgx_resource_state_t st, float* __restrict__ active, float* __restrict__ age) {
  const int64_t k = min((int64_t)*kp, n);
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < k; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = wrap_row(ids[j], n);
    if (i < 0 || i >= n) continue;
    st.X[i] = -1.0f; st.Y[i] = -1.0f; st.energy[i] = 0.0f; st.rad[i] = 0.0f;
    st.blossom[i] = 0;
    st.time_above_blossom[i] = -1.0f;
    active[i] = 0.0f;
    age[i] = 0.0f;
  }
}

__global__ void __launch_bounds__(256)
    resource_add_kernel(int64_t n, float x_min, float x_max, float y_min, float y_max,
                        const int32_t* __restrict__ good_ids, const int32_t* __restrict__ ap,
                        const int32_t* __restrict__ kp, const uint2* __restrict__ chain,
                        const fgx_resource_state_t st, float* __restrict__ growth, float* __restrict__ decay,
                        float* __restrict__ active, float* __restrict__ age) {
  const int64_t a = *ap, k = *kp;
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < k; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t dst = a + j;
    if (dst < 0 || dst >= n) continue;
    const int64_t src = clamp_row(good_ids[clamp_row(j, n)], n);
    const uint2 ck = chain[j];
    const Key key{ck.x, ck.y};
    const Key k2 = split_child(key, 4, 0);
    const float rad = st.rad[src];
    const float lo = f_mul(-2.0f, rad), hi = f_mul(2.0f, rad);
    st.X[dst] = clipf(f_add(st.X[src], uniform1(split_child(key, 4, 1), lo, hi)), x_min, x_max);
    st.Y[dst] = clipf(f_add(st.Y[src], uniform1(split_child(key, 4, 2), lo, hi)), y_min, y_max);
    const float en = uniform1(split_child(key, 4, 3), 3.0f, 4.0f);
    st.energy[dst] = en;
    st.rad[dst] = en;
    st.blossom[dst] = 0;
    st.time_above_blossom[dst] = 0.0f;
    const float gr = clipf(f_add(growth[src], uniform1(split_child(k2, 2, 1), -0.01f, 0.01f)), 0.001f, 1.0f);
    growth[dst] = gr;
    decay[dst] = f_mul(0.2f, gr);
    active[dst] = 1.0f;
    age[dst] = 0.0f;
  }
}

__global__ void __launch_bounds__(256)
    resource_set_kernel(int64_t n, const int32_t* __restrict__ ids, const int32_t* __restrict__ kp,
                        const fgx_resource_state_t st) {
  const int64_t k = min((int64_t)*kp, n);
  for (int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; j < k; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = wrap_row(ids[j], n);
    if (i < 0 || i >= n) continue;
    const float en = __fdiv_rn(st.energy[i], 2.0f);
    st.energy[i] = en;
    st.rad[i] = en;
    st.blossom[i] = 0;
    st.time_above_blossom[i] = 0.0f;
  }
}

}  // namespace fgx

// the key add_agents' serial chain holds after `steps` added rows (no rows are touched): lets a rank of a sharded
// set start the chain at the global index of its first new row, and every rank compute the final key
__global__ void key_chain_advance_kernel(int mode, const uint32_t* __restrict__ key_in, const int32_t* __restrict__ steps,
                                         uint32_t* __restrict__ key_out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  fgx::Key key{key_in[0], key_in[1]};
  const int64_t k = *steps;
  for (int64_t j = 0; j < k; ++j) {
    if (mode == 0) {
      key = fgx::split_child(key, 6, 0);
    } else {
      key = fgx::split_child(key, 4, 0);
      key = fgx::split_child(key, 2, 0);
    }
  }
  key_out[0] = key.a;
  key_out[1] = key.b;
}

extern "C" {

int fgx_forager_create(const fgx_forager_params_t* p, int64_t n, const uint32_t* create_keys,
                       const float* active_state, const fgx_forager_state_t* state, float* Z, uint32_t* init_keys,
                       float* age, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(p && state && n >= 0, "fgx_forager_create: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(create_keys && active_state && Z && init_keys && age, "fgx_forager_create: null pointer");
  forager_create_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(
      *p, n, reinterpret_cast<const uint2*>(create_keys), active_state, *state, Z, reinterpret_cast<uint2*>(init_keys),
      age);
  return check_launch("fgx_forager_create");
}

int fgx_forager_remove(int64_t n, int32_t n_neurons, const int32_t* remove_ids, const int32_t* k,
                       const fgx_forager_state_t* state, float* Z, float* active_state, float* age,
                       fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(state && n >= 0 && n_neurons > 0, "fgx_forager_remove: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(remove_ids && k && Z && active_state && age, "fgx_forager_remove: null pointer");
  forager_remove_kernel<<<grid_for(n * 32, 256, 4), 256, 0, as_stream(stream)>>>(n, n_neurons, remove_ids, k, *state,
                                                                                Z, active_state, age);
  return check_launch("fgx_forager_remove");
}

int fgx_forager_add(int64_t n, int32_t n_neurons, int32_t n_obs, int32_t n_act, const int32_t* good_ids,
                    const int32_t* n_active, const int32_t* k, const uint32_t* key_in, uint32_t* key_out,
                    float noise_scale, int32_t adaptive_noise, const fgx_forager_state_t* state,
                    const fgx_wc_policy_t* policy, float* active_state, float* age, uint32_t* chain_scratch,
                    fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(state && policy && n >= 0, "fgx_forager_add: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(good_ids && n_active && k && key_in && key_out && active_state && age && chain_scratch,
              "fgx_forager_add: null pointer");
  cudaStream_t s = as_stream(stream);
  uint2* chain = reinterpret_cast<uint2*>(chain_scratch);
  key_chain_kernel<0><<<1, 32, 0, s>>>(key_in, k, n, chain, key_out);
  MutateArgs m;
  m.leaf[0] = policy->J;   m.size[0] = n_neurons * n_neurons;
  m.leaf[1] = policy->tau; m.size[1] = n_neurons;
  m.leaf[2] = policy->E;   m.size[2] = n_neurons * n_obs;
  m.leaf[3] = policy->B;   m.size[3] = n_neurons;
  m.leaf[4] = policy->D;   m.size[4] = n_act * n_neurons;
  dim3 grid(grid_for((int64_t)n * 64, 256, 8), 5);
  forager_add_policy_kernel<<<grid, 256, 0, s>>>(n, m, good_ids, n_active, k, chain, age, noise_scale,
                                                 adaptive_noise);
  forager_add_state_kernel<<<grid_for(n, 256, 4), 256, 0, s>>>(n, n_neurons, good_ids, n_active, k, *state,
                                                              active_state, age);
  reset_age_rows_kernel<<<grid_for(n, 256, 4), 256, 0, s>>>(n, n_active, k, age);
  return check_launch("fgx_forager_add");
}

int fgx_key_chain_advance(int32_t mode, const uint32_t* key_in, const int32_t* steps, uint32_t* key_out,
                          fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(mode == 0 || mode == 1, "fgx_key_chain_advance: mode must be 0 (Forager) or 1 (Resource)");
  FGX_REQUIRE(key_in && steps && key_out, "fgx_key_chain_advance: null pointer");
  key_chain_advance_kernel<<<1, 32, 0, as_stream(stream)>>>(mode, key_in, steps, key_out);
  return check_launch("fgx_key_chain_advance");
}

int fgx_forager_set(int64_t n, const int32_t* set_ids, const int32_t* k, const fgx_forager_state_t* state,
                    fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(state && n >= 0, "fgx_forager_set: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(set_ids && k, "fgx_forager_set: null pointer");
  forager_set_kernel<<<grid_for(n, 256, 4), 256, 0, as_stream(stream)>>>(n, set_ids, k, *state);
  return check_launch("fgx_forager_set");
}

int fgx_resource_create(int64_t n, float x_lo, float x_hi, float y_lo, float y_hi, const uint32_t* create_keys,
                        const float* active_state, const fgx_resource_state_t* state, float* growth_rate,
                        float* decay_rate, float* age, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(state && n >= 0, "fgx_resource_create: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(create_keys && active_state && growth_rate && decay_rate && age, "fgx_resource_create: null pointer");
  resource_create_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(
      n, x_lo, x_hi, y_lo, y_hi, reinterpret_cast<const uint2*>(create_keys), active_state, *state, growth_rate,
      decay_rate, age);
  return check_launch("fgx_resource_create");
}

int fgx_resource_remove(int64_t n, const int32_t* remove_ids, const int32_t* k, const fgx_resource_state_t* state,
                        float* active_state, float* age, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(state && n >= 0, "fgx_resource_remove: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(remove_ids && k && active_state && age, "fgx_resource_remove: null pointer");
  resource_remove_kernel<<<grid_for(n, 256, 4), 256, 0, as_stream(stream)>>>(n, remove_ids, k, *state, active_state,
                                                                            age);
  return check_launch("fgx_resource_remove");
}

int fgx_resource_add(int64_t n, float x_min, float x_max, float y_min, float y_max, const int32_t* good_ids,
                     const int32_t* n_active, const int32_t* k, const uint32_t* key_in, uint32_t* key_out,
                     const fgx_resource_state_t* state, float* growth_rate, float* decay_rate, float* active_state,
                     float* age, uint32_t* chain_scratch, fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(state && n >= 0, "fgx_resource_add: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(good_ids && n_active && k && key_in && key_out && growth_rate && decay_rate && active_state && age &&
                  chain_scratch,
              "fgx_resource_add: null pointer");
  cudaStream_t s = as_stream(stream);
  uint2* chain = reinterpret_cast<uint2*>(chain_scratch);
  key_chain_kernel<1><<<1, 32, 0, s>>>(key_in, k, n, chain, key_out);
  resource_add_kernel<<<grid_for(n, 256, 4), 256, 0, s>>>(n, x_min, x_max, y_min, y_max, good_ids, n_active, k, chain,
                                                         *state, growth_rate, decay_rate, active_state, age);
  return check_launch("fgx_resource_add");
}

int fgx_resource_set(int64_t n, const int32_t* set_ids, const int32_t* k, const fgx_resource_state_t* state,
                     fgx_stream_t stream) {
  using namespace fgx;
  FGX_REQUIRE(state && n >= 0, "fgx_resource_set: bad arguments");
  if (n == 0) return FGX_OK;
  FGX_REQUIRE(set_ids && k, "fgx_resource_set: null pointer");
  resource_set_kernel<<<grid_for(n, 256, 4), 256, 0, as_stream(stream)>>>(n, set_ids, k, *state);
  return check_launch("fgx_resource_set");
}
}
