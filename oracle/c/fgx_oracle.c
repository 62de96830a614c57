/* C restatement of the reference's hot-path arithmetic, used as the timed CPU baseline
 * ("port") and as a second checker.  TEST INFRASTRUCTURE: never linked into the product.
 *
 * Follows, operation by operation in float32 (compiled with -ffp-contract=off):
 *   ray_generator      foragax/base/space_methods.py:56-69
 *   ray_wall_sensor    foragax/base/space_methods.py:96-143
 *   ray_agent_sensor   foragax/base/space_methods.py:71-94
 *   min / first argmin / -1  examples/many_agent_systems/foraging_agents/EcoEvoJax/sim.py:444-447
 *   Dice.step_agent    README.md:55-71 with jax.random (legacy threefry2x32, see oracle/jax_random.py)
 * Like XLA's vmap of lax.cond, every candidate of ray_wall_sensor is evaluated for every
 * (ray, wall) pair and selected afterwards.  OpenMP over agents: uses every host core.
 * PARITY UNPINNED for the geometry (the reference has no tests for it); it is checked
 * against oracle/space_methods.py, which is restated from the same source lines.
 */
#include <math.h>
#include <stdint.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static inline int orient(float px, float py, float qx, float qy, float rx, float ry) {
  float a = (qy - py) * (rx - qx);
  float b = (qx - px) * (ry - qy);
  float v = a - b;
  return v == 0.0f ? 0 : (v > 0.0f ? 1 : 2);
}
static inline int on_seg(float px, float py, float qx, float qy, float rx, float ry) {
  return qx <= fmaxf(px, rx) && qx >= fminf(px, rx) && qy <= fmaxf(py, ry) && qy >= fminf(py, ry);
}
static inline float norm2(float dx, float dy) {
  float a = dx * dx;
  float b = dy * dy;
  return sqrtf(a + b);
}
static inline float min_nan(float a, float b) { return (a != a || b != b) ? NAN : fminf(a, b); }

static float ray_wall(float p1x, float p1y, float q1x, float q1y, float L, float p2x, float p2y, float q2x,
                      float q2y) {
  int o1 = orient(p1x, p1y, q1x, q1y, p2x, p2y);
  int o2 = orient(p1x, p1y, q1x, q1y, q2x, q2y);
  int o3 = orient(p2x, p2y, q2x, q2y, p1x, p1y);
  int o4 = orient(p2x, p2y, q2x, q2y, q1x, q1y);
  /* line_intercept, evaluated unconditionally (vmap of cond = select) */
  float ax = q1x - p1x, ay = q1y - p1y;
  float m1 = ax * (p2y - p1y), m2 = ay * (p2x - p1x);
  float m3 = ax * (q2y - p1y), m4 = ay * (q2x - p1x);
  float c1 = m1 - m2, c2 = m3 - m4;
  float r = c1 / ((c1 - c2) + 1e-6f);
  float t1 = r * (q2x - p2x), t2 = r * (q2y - p2y);
  float ix = p2x + t1, iy = p2y + t2;
  float e1 = (o1 != o2 && o3 != o4) ? norm2(ix - p1x, iy - p1y) : L;
  float e2 = (o1 == 0 && on_seg(p1x, p1y, p2x, p2y, q1x, q1y)) ? norm2(p1x - p2x, p1y - p2y) : L;
  float e3 = (o2 == 0 && on_seg(p1x, p1y, q2x, q2y, q1x, q1y)) ? norm2(p1x - q2x, p1y - q2y) : L;
  float e4 = (o3 == 0 && on_seg(p2x, p2y, p1x, p1y, q2x, q2y)) ? 0.0f : L;
  float e5 = (o4 == 0 && on_seg(p2x, p2y, q1x, q1y, q2x, q2y)) ? norm2(p1x - q1x, p1y - q1y) : L;
  return min_nan(e1, min_nan(e2, min_nan(e3, min_nan(e4, e5))));
}

static inline float linspace_elem(float a, float b, int n, int j) {
  if (n == 1) return a;
  if (j == n - 1) return b;
  float s = (float)j / (float)(n - 1);
  float u = a * (1.0f - s);
  float v = b * s;
  return u + v;
}

/* agents x rays x walls -> dist[n,R], hit[n,R] (hit may be NULL) */
void fgx_oracle_raycast_walls(const float* x, const float* y, const float* ang, const float* active, int64_t n,
                              int n_rays, float span, float L, const float* wx1, const float* wy1, const float* wx2,
                              const float* wy2, int64_t n_walls, float* dist, int32_t* hit) {
#pragma omp parallel for schedule(static)
  for (int64_t a = 0; a < n; ++a) {
    int act = active == NULL || active[a] != 0.0f;
    for (int j = 0; j < n_rays; ++j) {
      int64_t o = a * n_rays + j;
      if (!act) {
        dist[o] = 0.0f;
        if (hit) hit[o] = -1;
        continue;
      }
      float g = linspace_elem(ang[a] - span, ang[a] + span, n_rays, j);
      float c = (float)cos((double)g), s = (float)sin((double)g);
      float cx = c * L, sy = s * L;
      float q1x = x[a] + cx, q1y = y[a] + sy;
      float best = L;
      int32_t bi = -1;
      for (int64_t w = 0; w < n_walls; ++w) {
        float d = ray_wall(x[a], y[a], q1x, q1y, L, wx1[w], wy1[w], wx2[w], wy2[w]);
        if (d < best) {
          best = d;
          bi = (int32_t)w;
        }
      }
      dist[o] = best;
      if (hit) hit[o] = bi;
    }
  }
}

/* ---- threefry2x32 / jax.random (legacy layout), for the Dice step baseline ---- */
static inline uint32_t rotl(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }
static void tf(uint32_t k0, uint32_t k1, uint32_t* x0p, uint32_t* x1p) {
  static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
  uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  uint32_t x0 = *x0p + ks[0], x1 = *x1p + ks[1];
  for (int g = 0; g < 5; ++g) {
    for (int i = 0; i < 4; ++i) {
      x0 += x1;
      x1 = rotl(x1, R[g & 1][i]);
      x1 ^= x0;
    }
    x0 += ks[(g + 1) % 3];
    x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
  }
  *x0p = x0;
  *x1p = x1;
}
static void split2(const uint32_t k[2], uint32_t c0[2], uint32_t c1[2]) {
  uint32_t a0 = 0, a1 = 2, b0 = 1, b1 = 3;
  tf(k[0], k[1], &a0, &a1);
  tf(k[0], k[1], &b0, &b1);
  c0[0] = a0; c0[1] = b0; c1[0] = a1; c1[1] = b1;
}
static uint32_t bits1(const uint32_t k[2]) {
  uint32_t a = 0, b = 0;
  tf(k[0], k[1], &a, &b);
  return a;
}
static int32_t randint1(const uint32_t k[2], int32_t lo, int32_t hi) {
  uint32_t k1[2], k2[2];
  split2(k, k1, k2);
  uint32_t hb = bits1(k1), lb = bits1(k2);
  uint32_t span = hi > lo ? (uint32_t)hi - (uint32_t)lo : 1u;
  uint32_t m = 65536u % span;
  m = (m * m) % span;
  uint32_t off = ((hb % span) * m + (lb % span)) % span;
  return (int32_t)((uint32_t)lo + off);
}

/* vmapped Dice.step_agent, in place */
void fgx_oracle_dice_step(int64_t n, int sides, const float* active, int32_t* draw, uint32_t* key, float* age) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    if (active[i] != 0.0f) {
      uint32_t c0[2], sub[2];
      split2(key + 2 * i, c0, sub);
      draw[i] = randint1(sub, 1, sides + 1);
      key[2 * i] = sub[0];
      key[2 * i + 1] = sub[1];
      age[i] += 1.0f;
    }
  }
}

void fgx_oracle_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int fgx_oracle_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ---- vmapped Forager.step_agent with WilsonCowan.step_policy ----------------------------
 * examples/.../EcoEvoJax/sim.py:74-136, foragax/policy/wilson_cowan.py:34-51 (sim.py variant:
 * energy cost 0.3*dt*sum|a|, clip(energy_min, energy_max)).  state = 11 arrays float32[n] in the
 * order X, X_vel, X_acc, Y, Y_vel, Y_acc, ang, energy, rad, time_above_rep, time_below_die.
 * obs float32[n, n_obs-1]; in place.                                                      */
void fgx_oracle_forager_step(int64_t n, int nn, int nobs, int nact, float dt, float action_scale, float x_min,
                             float x_max, float y_min, float y_max, float repr_thresh, float die_thresh,
                             float energy_min, float energy_max, const float* active, const float* obs,
                             const float* energy_in, float** st, const float* J, const float* tau, const float* E,
                             const float* B, const float* D, float* Z, float* age) {
#pragma omp parallel for schedule(static)
  for (int64_t a = 0; a < n; ++a) {
    if (active[a] == 0.0f) continue;
    float tz[256], nz[256], ob[512], act[8];
    const float* Ja = J + a * nn * nn;
    const float* Ea = E + a * nn * nobs;
    const float* Da = D + a * nact * nn;
    float* Za = Z + a * nn;
    for (int k = 0; k < nobs - 1; ++k) ob[k] = obs[a * (nobs - 1) + k];
    ob[nobs - 1] = st[7][a];
    for (int j = 0; j < nn; ++j) tz[j] = tanhf(Za[j]);
    for (int r = 0; r < nn; ++r) {
      float sJ = 0.0f, sE = 0.0f;
      for (int j = 0; j < nn; ++j) sJ += Ja[r * nn + j] * tz[j];
      for (int k = 0; k < nobs; ++k) sE += Ea[r * nobs + k] * ob[k];
      float dz = ((sJ + sE) + B[a * nn + r]) - Za[r];
      dz = dz * (10.0f * (1.0f / (1.0f + expf(-tau[a * nn + r]))));
      nz[r] = Za[r] + dt * dz;
    }
    for (int r = 0; r < nn; ++r) { Za[r] = nz[r]; nz[r] = tanhf(nz[r]); }
    for (int k = 0; k < nact && k < 8; ++k) {
      float s = 0.0f;
      for (int j = 0; j < nn; ++j) s += Da[k * nn + j] * nz[j];
      act[k] = action_scale * tanhf(s);
    }
    float X = st[0][a], Y = st[3][a], Xv = st[1][a], Yv = st[4][a];
    float Xn = X + Xv * dt, Yn = Y + Yv * dt, Xvn = Xv + act[0] * dt, Yvn = Yv + act[1] * dt;
    if (Xn > x_max || Xn < x_min) { Xn = X; Xvn = -Xv; }
    if (Yn > y_max || Yn < y_min) { Yn = Y; Yvn = -Yv; }
    float en = (st[7][a] + energy_in[a]) - (0.3f * dt) * (fabsf(act[0]) + fabsf(act[1]));
    en = fminf(fmaxf(en, energy_min), energy_max);
    st[0][a] = Xn; st[3][a] = Yn; st[1][a] = Xvn; st[4][a] = Yvn; st[2][a] = act[0]; st[5][a] = act[1];
    st[6][a] = atan2f(Yvn, Xvn);
    st[7][a] = en; st[8][a] = en;
    st[9][a] = en > repr_thresh ? st[9][a] + dt : 0.0f;
    st[10][a] = en < die_thresh ? st[10][a] + dt : 0.0f;
    age[a] += dt;
  }
}
