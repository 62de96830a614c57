/* libforagax_b200 -- C ABI of the B200-native Foragax agent hot path.
 *
 * The reference (i-m-iron-man/Foragax v0.0.5) has no FFI: its boundary is the
 * Python API of foragax/base/agent_methods.py and foragax/base/space_methods.py,
 * whose arithmetic runs inside jax/XLA.  Every entry point below states the
 * reference function (file:line, relative to the reference root) it replaces.
 * The same launchers are what an XLA-FFI handler (jax.ffi) binds; see
 * INTEGRATION.md.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless its name ends in _host;
 *  - the caller owns all buffers; the library never allocates device memory,
 *    never retains a pointer across calls and never synchronises the device;
 *  - all work is enqueued asynchronously on `stream` (a cudaStream_t);
 *  - traced counts (k of remove/add/set, counts returned by select) live in
 *    device memory so that select -> remove -> sort -> add chains on-stream
 *    with no host round trip;
 *  - return value: 0 = OK, negative = error; text via fgx_last_error()
 *    (thread-local).  Arguments are validated on the host before launch;
 *  - scratch: ops that need temporary device memory take (scratch,
 *    scratch_bytes); the required size comes from fgx_<op>_scratch_bytes().
 *  - in-place: where an op has `in` and `out` arrays they may alias.
 */
#ifndef FORAGAX_B200_H
#define FORAGAX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fgx_stream_t; /* cudaStream_t */

/* element types of keys / masks / predicate leaves */
enum { FGX_U8 = 0, FGX_I32 = 1, FGX_F32 = 2, FGX_U32 = 3 };

/* ---- library ---------------------------------------------------------- */
const char* fgx_last_error(void);
int fgx_version(void);
/* SM count and compute capability of the current device (host query). */
int fgx_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- K4: threefry2x32, jax.random bit-for-bit (legacy counter layout) ---
 * Replaces the third-party jax.random calls of agent_methods.py:11,
 * random_policy.py:19-20, wilson_cowan.py:19-26, README.md:39,42,59,61,75-76.
 * keys: uint32[n_keys,2].  Each op is "vmapped" over the key batch.          */
/* jax.random.split(key, num): out uint32[n_keys, num, 2] */
int fgx_threefry_split(const uint32_t* keys, int64_t n_keys, int32_t num, uint32_t* out, fgx_stream_t stream);
/* jax.random.bits(key, (n,)): out uint32[n_keys, n] */
int fgx_threefry_random_bits(const uint32_t* keys, int64_t n_keys, int32_t n, uint32_t* out, fgx_stream_t stream);
/* jax.random.uniform(key, (n,), minval, maxval): out float32[n_keys, n] */
int fgx_threefry_uniform(const uint32_t* keys, int64_t n_keys, int32_t n, float minval, float maxval, float* out,
                         fgx_stream_t stream);
/* jax.random.randint(key, (n,), minval, maxval): out int32[n_keys, n] */
int fgx_threefry_randint(const uint32_t* keys, int64_t n_keys, int32_t n, int32_t minval, int32_t maxval,
                         int32_t* out, fgx_stream_t stream);
/* jax.random.normal(key, (n,)): out float32[n_keys, n] (erf_inv polynomial; tolerance parity) */
int fgx_threefry_normal(const uint32_t* keys, int64_t n_keys, int32_t n, float* out, fgx_stream_t stream);

/* ---- K3: select / sort / row movement ----------------------------------- */
/* One comparison of a predicate: (leaf[i*stride] <op> imm).  `imm_bits` holds
 * the int32 or float32 constant bit pattern.  FGX_OP_NONZERO ignores imm.   */
enum { FGX_OP_EQ = 0, FGX_OP_NE = 1, FGX_OP_LT = 2, FGX_OP_LE = 3, FGX_OP_GT = 4, FGX_OP_GE = 5, FGX_OP_NONZERO = 6 };
typedef struct {
  const void* data; /* device leaf, element i at data[i*stride] */
  int32_t dtype;    /* FGX_U8 / FGX_I32 / FGX_F32 */
  int32_t op;       /* FGX_OP_* */
  uint32_t imm_bits;
  int32_t stride; /* in elements */
} fgx_pred_term_t;
/* A predicate = up to 4 terms combined by an arbitrary boolean function given
 * as a 16-entry truth table: selected(i) = (lut >> (t0 | t1<<1 | t2<<2 | t3<<3)) & 1 */
typedef struct {
  fgx_pred_term_t term[4];
  int32_t n_terms;
  uint32_t lut;
} fgx_predicate_t;

size_t fgx_select_scratch_bytes(int64_t n);
/* select_agents (agent_methods.py:66-71): stable 2-way partition of the slot
 * indices.  perm int32[n] = selected ascending, then unselected ascending
 * (== argsort(-1*where(mask,1,0))); count int32[1] = number selected.       */
int fgx_select(const fgx_predicate_t* pred_host, int64_t n, int32_t* perm, int32_t* count, void* scratch,
               size_t scratch_bytes, fgx_stream_t stream);

size_t fgx_argsort_scratch_bytes(int64_t n);
/* the argsort of sort_agents (agent_methods.py:74-76): stable; descending !=0
 * sorts the NEGATED key ascending (ties keep ascending slot order); float keys
 * use lax.sort's total order (-0.0 == +0.0, NaN last).  key_dtype FGX_I32/F32. */
int fgx_argsort(const void* keys, int32_t key_dtype, int64_t n, int32_t descending, int32_t* perm, void* scratch,
                size_t scratch_bytes, fgx_stream_t stream);

/* one leaf of a struct-of-arrays agent pytree */
typedef struct {
  const void* src;
  void* dst;
  int64_t row_bytes; /* bytes per slot of this leaf */
} fgx_leaf_t;
#define FGX_MAX_LEAVES 32
/* the payload move of sort_agents (agent_methods.py:77): dst[i] = src[perm[i]]
 * for every leaf (jnp.take, indices clamped).  src and dst must not alias.  */
int fgx_gather_rows(const fgx_leaf_t* leaves_host, int32_t n_leaves, const int32_t* perm, int64_t n,
                    fgx_stream_t stream);
/* jnp.sum(active_state, dtype=int32) (agent_methods.py:27): out int32[1] */
int fgx_count_active(const float* active_state, int64_t n, int32_t* out, fgx_stream_t stream);
/* the argument construction of create_agents (agent_methods.py:10-15), fused:
 * unique_id = arange(n); create_keys = split(key, n+1)[1:]; active_state =
 * concat(ones(n_active), zeros(n-n_active)) float32; agent_types = tile(agent_type, n).
 * key_host is the HOST uint32[2] key passed to create_agents.                */
int fgx_create_base(const uint32_t key_host[2], int64_t n_total, int64_t n_active, int32_t agent_type,
                    int32_t* unique_id, uint32_t* create_keys, float* active_state, int32_t* agent_types,
                    fgx_stream_t stream);

/* ---- K1: ray casting (space_methods.py) ---------------------------------- */
/* Walls are struct-of-arrays float32[n_walls]: (wx1,wy1) = Wall.p1, (wx2,wy2) = Wall.p2
 * (space_classes.py:15-20, wall_generator space_methods.py:45-53).               */
/* The agents x rays x walls hot loop: ray_generator (space_methods.py:56-69) +
 * ray_wall_sensor (space_methods.py:96-143) for every (agent, ray, wall), reduced per ray
 * as the example-level composition does (EcoEvoJax/sim.py:444-447): dist float32[n,R] =
 * min over walls (ray_length if none), hit int32[n,R] = first wall index attaining it,
 * -1 if dist >= ray_length (hit may be NULL).  x, y, ang float32[n]; active float32[n] or
 * NULL: inactive agents get dist 0 and hit -1 (sim.py:459-460).                    */
int fgx_raycast_walls(const float* x, const float* y, const float* ang, const float* active, int64_t n_agents,
                      int32_t n_rays, float ray_span, float ray_length, const float* wx1, const float* wy1,
                      const float* wx2, const float* wy2, int64_t n_walls, float* dist, int32_t* hit,
                      fgx_stream_t stream);
/* ray_generator (space_methods.py:56-69), directions only: dir_x, dir_y float32[n, n_rays]
 * = cos / sin of linspace(ang - span, ang + span, n_rays).                          */
int fgx_ray_generator(const float* ang, int64_t n, int32_t n_rays, float ray_span, float* dir_x, float* dir_y,
                      fgx_stream_t stream);
/* jit_ray_wall_sensor (space_methods.py:96-143) vmapped over rays x walls:
 * out float32[n_rays_total, n_walls]; ray i = origin (ox,oy)[i], direction (dx,dy)[i].   */
int fgx_ray_wall_sensor(const float* ox, const float* oy, const float* dx, const float* dy, int64_t n_rays_total,
                        float ray_length, const float* wx1, const float* wy1, const float* wx2, const float* wy2,
                        int64_t n_walls, float* out, fgx_stream_t stream);
/* jit_ray_agent_sensor (space_methods.py:71-94) vmapped over rays x entities
 * (ex, ey, erad, eactive float32[n_entities]): out float32[n_rays_total, n_entities].    */
int fgx_ray_agent_sensor(const float* ox, const float* oy, const float* dx, const float* dy, int64_t n_rays_total,
                         float ray_length, const float* ex, const float* ey, const float* erad, const float* eactive,
                         int64_t n_entities, float* out, fgx_stream_t stream);
/* jit_check_collision (space_methods.py:10-43) vmapped over n moves old->new:
 * out uint8[n] = any wall intersects the move segment.                               */
int fgx_check_collision(const float* wx1, const float* wy1, const float* wx2, const float* wy2, int64_t n_walls,
                        const float* old_x, const float* old_y, const float* new_x, const float* new_y, int64_t n,
                        uint8_t* out, fgx_stream_t stream);

/* ---- K2: hello-world Dice agent (README.md:35-129) ---------------------- */
/* Leaves: draw int32[n] (state.content['draw'], shape [n,1]), key uint32[n,2],
 * unique_id int32[n], agent_type int32[n], active_state float32[n], age float32[n].
 * `sides` generalises randint(key,(1,),1,7) to randint(key,(1,),1,sides+1).  */
/* vmapped Dice.create_agent (README.md:37-53) over the batched arguments that
 * create_agents builds: create_keys uint32[n,2], active_state float32[n].     */
int fgx_dice_create(int64_t n, int32_t sides, const uint32_t* create_keys, const float* active_state, int32_t* draw,
                    uint32_t* key, float* age, fgx_stream_t stream);
/* step_agents + vmapped Dice.step_agent (agent_methods.py:20-24, README.md:55-71); in place */
int fgx_dice_step(int64_t n, int32_t sides, const float* active_state, int32_t* draw, uint32_t* key, float* age,
                  fgx_stream_t stream);
/* remove_agents with Dice.remove_agent (agent_methods.py:41-50, README.md:83-90):
 * rows remove_ids[0..*k) get draw=0, active_state=0; in place                 */
int fgx_dice_remove(int64_t n, const int32_t* remove_ids, const int32_t* k, int32_t* draw, float* active_state,
                    fgx_stream_t stream);
/* add_agents with Dice.add_agent (agent_methods.py:26-36, README.md:73-81): rows
 * [*n_active, *n_active + *k) are activated with a fresh draw; in place        */
int fgx_dice_add(int64_t n, int32_t sides, const int32_t* n_active, const int32_t* k, int32_t* draw, uint32_t* key,
                 float* active_state, fgx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FORAGAX_B200_H */
