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
/* the same with n output rows taken from leaves of n_src rows (x[ids] for a shorter id list) */
int fgx_gather_rows_from(const fgx_leaf_t* leaves_host, int32_t n_leaves, const int32_t* perm, int64_t n,
                         int64_t n_src, fgx_stream_t stream);
/* sort_agents (agent_methods.py:74-79) as ONE call: perm = the stable argsort of fgx_argsort AND
 * dst[i] = src[perm[i]] for every leaf (agent_methods.py:77).  When the key is two-valued (the sort on
 * active_state, or on a 0/1 flag) the keys are read once and the narrow payload rows are scattered to their
 * sorted positions by the same cooperative kernel -- no permutation round trip, no gather pass; otherwise the
 * radix sort runs and the rows are gathered by it.  scratch: fgx_argsort_scratch_bytes(n), 256-byte aligned. */
int fgx_sort_rows(const void* keys, int32_t key_dtype, int64_t n, int32_t descending, int32_t* perm,
                  const fgx_leaf_t* leaves_host, int32_t n_leaves, void* scratch, size_t scratch_bytes,
                  fgx_stream_t stream);
/* sort_agents(active_state, descending) (agent_methods.py:74-79) over a set sharded in contiguous blocks
 * across the GPUs of one node, compaction and exchange in ONE kernel: every rank scatters its rows into
 * the peers' destination buffers (peer-mapped device memory, stores over NVLink) at their slots of the
 * global stable order.  perm / n_front: this rank's stable partition from fgx_select (front = active).
 * front_counts: device int64[world], the all-gathered n_front.  bounds_host: int64[world+1] shard
 * bounds.  dst[d] of a leaf = that leaf's destination buffer on rank d (dst[rank] is local).  The caller
 * must order the ranks (a collective after this call) before anyone reads a destination buffer.       */
#define FGX_MAX_PEERS 8
#define FGX_MAX_PEER_LEAVES 16
typedef struct {
  const void* src;
  void* dst[FGX_MAX_PEERS];
  int64_t row_bytes;
} fgx_peer_leaf_t;
int fgx_pack_rows_peer(const fgx_peer_leaf_t* leaves_host, int32_t n_leaves, const int32_t* perm,
                       const int32_t* n_front, int64_t n_local, const int64_t* front_counts,
                       const int64_t* bounds_host, int32_t rank, int32_t world, fgx_stream_t stream);
/* Device-side exchange of a few counts between the ranks of one node, with NO host involvement and no NCCL
 * call: SURVEY.md 8e(2) ("all-gather of one int32 count per rank -> exclusive prefix") as a one-warp kernel over
 * peer-mapped memory.  Every rank owns a MAILBOX (fgx_peer_mailbox_bytes() of symmetric memory, zeroed once);
 * mailbox[d] is rank d's mailbox as mapped into this process.  One call = one epoch: the kernel stores this
 * rank's values into its row of every peer's mailbox, releases a flag there (st.release.sys after a system
 * fence -- which also publishes the peer stores of every kernel enqueued earlier on this stream, e.g. the row
 * scatter of fgx_pack_rows_peer), spins on the flags the peers set in its own mailbox (bounded: on a timeout of
 * ~5 s it raises *error and returns, it never hangs) and copies the gathered values out.  Every rank must call
 * the same sequence of exchanges.  epoch: device uint64 (zeroed once), error: device int32 (zeroed once).
 * With v0 == v1 == NULL the call is a pure barrier between the ranks' streams.
 * out_all (may be NULL): int64[2][FGX_MAX_PEERS], out_all[j][s] = value j of rank s (0 when not supplied).      */
typedef struct {
  void* mailbox[FGX_MAX_PEERS];
  uint64_t* epoch;
  int32_t* error;
  int32_t rank, world;
} fgx_peer_sync_t;
size_t fgx_peer_mailbox_bytes(void);
int fgx_peer_exchange(const fgx_peer_sync_t* sync, const int32_t* v0, const int32_t* v1, int64_t* out_all,
                      fgx_stream_t stream);
/* The position exchange of agent-agent sensing (all-gather of float4 (x, y, rad, active), 16 B per
 * agent) without NCCL: every rank packs its rows into its own STAGING buffer in peer-mapped memory (stage[d] = rank
 * d's staging buffer as mapped here, 2 * n_max float4: two parities), the ranks meet in a mailbox barrier
 * (fgx_peer_exchange semantics, same sync state rules) and every rank then pulls all rows over NVLink into
 * table[n_total] IN GLOBAL SLOT ORDER.  layout 0: contiguous blocks (bounds_host int64[world + 1]); layout 1:
 * round-robin (slot i on rank i mod world at row i / world; bounds_host ignored).  All launches are enqueued on
 * `stream`; no host synchronisation, capturable in a CUDA graph.                                             */
int fgx_peer_gather_positions(const fgx_peer_sync_t* sync, void* const* stage, int64_t n_max, const float* x,
                              const float* y, const float* rad, const float* active, int64_t n_local, int64_t n_total,
                              int32_t layout, const int64_t* bounds_host, float* table, fgx_stream_t stream);
/* add_agents over a globally packed, block-sharded set (agent_methods.py:26-38 + the caller's clip
 * min(n_selected, N - n_active), README.md:124-125) in one launch: exchanges (n_selected_local, n_active_local)
 * as above and writes out3 = { k = min(sum selected, n_total - sum active),
 *                              local_count = rows of the global new range [A, A + k) inside this rank's block
 *                                            [lo, lo + n_local),
 *                              j0 = global add index of this rank's first new row }.                       */
int fgx_peer_add_range(const fgx_peer_sync_t* sync, const int32_t* n_selected_local, const int32_t* n_active_local,
                       int64_t n_local, int64_t lo, int64_t n_total, int32_t* out3, fgx_stream_t stream);
/* sort_agents on a GENERAL key over the same block-sharded set: images[p] = order-preserving uint32 image of
 * keys[perm[p]] (perm = this rank's fgx_argsort); after an all-gather of every rank's sorted images
 * (images_all: [world][n_pad], rows padded), fgx_sort_rows_peer finds each local row's position in the
 * global stable order by binary searches (ties: rank order, then local order) and writes the row to its
 * slot on the owning GPU.  gpos_scratch: device int64[n_local].  Same ordering rule as fgx_pack_rows_peer. */
int fgx_sort_key_images(const void* keys, int32_t key_dtype, int64_t n, int32_t descending, const int32_t* perm,
                        uint32_t* images, fgx_stream_t stream);
int fgx_sort_rows_peer(const fgx_peer_leaf_t* leaves_host, int32_t n_leaves, const int32_t* perm, int64_t n_local,
                       const uint32_t* images_all, int64_t n_pad, const int64_t* bounds_host, int32_t rank,
                       int32_t world, int64_t* gpos_scratch, fgx_stream_t stream);
/* The recording ring of the example drivers (EcoEvoJax/sim.py:555-595 Rec_Data / record_data;
 * wolf_sheep_grass/sim.ipynb:301-362): for every leaf, ring[frame mod frames] = leaf, i.e. the n rows at src are
 * copied to dst + (frame mod frames) * n * row_bytes (dst = base of the [frames, n, ...] ring).  One launch for
 * all leaves; *frame is a device int32 so that a recorded tick can be captured in a CUDA graph.               */
int fgx_record_frame(const fgx_leaf_t* leaves_host, int32_t n_leaves, int64_t n, int32_t frames, const int32_t* frame,
                     fgx_stream_t stream);
/* jnp.sum(x) of a float32 leaf (EcoEvoJax/sim.py:486,491: the age sums around remove_agents) in one launch
 * with a fixed summation order (deterministic run to run; float32 accumulation like jnp.sum).  out float32[1]. */
size_t fgx_sum_f32_scratch_bytes(void);
int fgx_sum_f32(const float* x, int64_t n, float* out, void* scratch, size_t scratch_bytes, fgx_stream_t stream);
/* The life-expectancy update of Neuroevolution (EcoEvoJax/sim.py:486-493) on device scalars:
 * le = (age_before - age_after) / (n_removed + 0.01); out = (le == 0) ? old_value : le                        */
int fgx_life_expectancy(const float* age_before, const float* age_after, const int32_t* n_removed, const float* old_value,
                        float* out, fgx_stream_t stream);
/* jnp.sum(active_state, dtype=int32) (agent_methods.py:27): out int32[1] */
int fgx_count_active(const float* active_state, int64_t n, int32_t* out, fgx_stream_t stream);
/* the same count plus the caller's clip of hello_world.ipynb:138-139 / README.md:124-125 in one launch:
 * out3[0] = sum(active_state), out3[1] = min(*n_add, n - out3[0]) (new agents that still fit), out3[2] = scratch */
int fgx_count_active_clip(const float* active_state, int64_t n, const int32_t* n_add, int32_t* out3,
                          fgx_stream_t stream);
/* the argument construction of create_agents (agent_methods.py:10-15), fused:
 * unique_id = arange(n); create_keys = split(key, n+1)[1:]; active_state =
 * concat(ones(n_active), zeros(n-n_active)) float32; agent_types = tile(agent_type, n).
 * key_host is the HOST uint32[2] key passed to create_agents.                */
int fgx_create_base(const uint32_t key_host[2], int64_t n_total, int64_t n_active, int32_t agent_type,
                    int32_t* unique_id, uint32_t* create_keys, float* active_state, int32_t* agent_types,
                    fgx_stream_t stream);

/* the same for the slots first, first+stride, ... (n_local of them) of the set only: a shard of a
 * multi-GPU run produces exactly its rows of the single-GPU result (arrays have n_local entries);
 * stride 1 = a contiguous block, stride G with first = rank = round-robin sharding              */
int fgx_create_base_shard(const uint32_t key_host[2], int64_t n_total, int64_t n_active, int64_t first,
                          int64_t stride, int64_t n_local, int32_t agent_type, int32_t* unique_id,
                          uint32_t* create_keys, float* active_state, int32_t* agent_types, fgx_stream_t stream);

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
/* The same reduction with a spatial cull, BIT-IDENTICAL to fgx_raycast_walls: a uniform grid over
 * [x_min,x_max] x [y_min,y_max] lists, per cell, the walls that can yield a distance below ray_length for an
 * agent in that cell -- walls within reach of the cell, plus walls whose infinite line passes through it
 * (float32 sign noise of the orientation tests, space_methods.py:112-125, can make a far collinear wall
 * "hit"; the all-pairs evaluation reproduces that and so does this).  Agents outside the bounds and walls
 * with coordinates beyond twice the bounds fall back to the all-walls loop (exact).  reuse_grid != 0: the
 * scratch still holds the grid a previous call built from the SAME walls, bounds and ray_length (walls are
 * static in the reference: create_space, space_methods.py:146-156) and only the sensing kernel runs.     */
size_t fgx_raycast_walls_grid_scratch_bytes(int64_t n_walls, float x_min, float x_max, float y_min, float y_max,
                                            float ray_length);
int fgx_raycast_walls_grid(const float* x, const float* y, const float* ang, const float* active, int64_t n_agents,
                           int32_t n_rays, float ray_span, float ray_length, const float* wx1, const float* wy1,
                           const float* wx2, const float* wy2, int64_t n_walls, float x_min, float x_max, float y_min,
                           float y_max, float* dist, int32_t* hit, void* scratch, size_t scratch_bytes,
                           int32_t reuse_grid, fgx_stream_t stream);
/* Statistics of that cull (measurement aid for the instruction floor of the kernel): out2[0] = sum over the
 * ACTIVE agents of the number of walls listed for their cell (n_walls for agents that take the all-walls
 * loop), out2[1] = number of active agents.  Same grid arguments / scratch / reuse_grid as above.         */
int fgx_raycast_walls_grid_listed(const float* x, const float* y, const float* active, int64_t n_agents, const float* wx1,
                                  const float* wy1, const float* wx2, const float* wy2, int64_t n_walls, float x_min,
                                  float x_max, float y_min, float y_max, float ray_length, void* scratch,
                                  size_t scratch_bytes, int32_t reuse_grid, int64_t* out2, fgx_stream_t stream);
/* The agents x rays x agent-targets loop: ray_generator + ray_agent_sensor (space_methods.py:71-94)
 * for every ray of every sensing agent against every target circle (tx, ty, trad, tactive:
 * float32, element i at [i*target_stride] -- stride 4 for the interleaved float4 rows of a position
 * all-gather), reduced per ray to (min distance, FIRST target index attaining it, -1 if none), as
 * the all-pairs reference does (EcoEvoJax/sim.py:444-447).  A uniform grid over [x_min,x_max] x
 * [y_min,y_max] with cell >= ray_length + max_rad limits the search to 3x3 cells; the cull is exact
 * provided no target radius exceeds max_rad.  merge_walls != 0: dist / hit hold the result of
 * fgx_raycast_walls on entry and the merged result in entity order [targets, walls] on exit (wall w
 * becomes index n_targets + w).  Inactive sensing agents: dist 0, hit -1.                     */
size_t fgx_raycast_agents_scratch_bytes(int64_t n_agents, int64_t n_targets, float x_min, float x_max, float y_min,
                                        float y_max, float ray_length, float max_rad);
int fgx_raycast_agents(const float* x, const float* y, const float* ang, const float* active, int64_t n_agents,
                       int32_t n_rays, float ray_span, float ray_length, const float* tx, const float* ty,
                       const float* trad, const float* tactive, int64_t target_stride, int64_t n_targets,
                       float x_min, float x_max, float y_min, float y_max, float max_rad, int32_t merge_walls,
                       float* dist, int32_t* hit, void* scratch, size_t scratch_bytes, fgx_stream_t stream);
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

/* ---- K2: foraging example (examples/many_agent_systems/foraging_agents/EcoEvoJax/sim.py) ---
 * Forager = WilsonCowan policy (foragax/policy/wilson_cowan.py) + point-mass physics + energy;
 * Resource = logistic growth + blossom.  All leaves are struct-of-arrays with the slot axis
 * first; [n,1] leaves are passed as float32[n].                                          */
typedef struct {
  int32_t n_neurons, n_obs /* incl. the agent's own energy */, n_act;
  float dt;             /* forager.params.content['dt'] (sim.py:94) */
  float policy_dt;      /* policy.params.content['dt'] (wilson_cowan.py:44) */
  float action_scale;   /* wilson_cowan.py:45 */
  float x_min, x_max, y_min, y_max; /* space bounds (sim.py:100-104) */
  float repr_thresh, die_thresh, energy_min, energy_max; /* step params (sim.py:659) */
  float energy_cost_dt; /* float32(0.3)*dt in sim.py:109; dt in the notebook variant */
  int32_t clip_min;     /* 1: clip(e, energy_min, energy_max) (sim.py:111); 0: minimum(e, energy_max) */
} fgx_forager_params_t;
typedef struct {
  float *X, *X_vel, *X_acc, *Y, *Y_vel, *Y_acc, *ang, *energy, *rad, *time_above_rep_thresh, *time_below_die_thresh;
} fgx_forager_state_t; /* each float32[n] */
typedef struct {
  float *J;   /* [n, neurons, neurons] */
  float *tau; /* [n, neurons] */
  float *E;   /* [n, neurons, n_obs] */
  float *B;   /* [n, neurons] */
  float *D;   /* [n, n_act, neurons] */
  float *Z;   /* [n, neurons] policy state */
} fgx_wc_policy_t;
/* step_agents + vmapped Forager.step_agent (agent_methods.py:20-24, sim.py:74-136) with
 * WilsonCowan.step_policy (wilson_cowan.py:34-51) fused; in place.  obs float32[n, n_obs-1]
 * (sensor_model output), energy_in float32[n].  Inactive slots pass through.              */
int fgx_forager_step(const fgx_forager_params_t* p, int64_t n, const float* active_state, const float* obs,
                     const float* energy_in, const fgx_forager_state_t* state, const fgx_wc_policy_t* policy,
                     float* age, fgx_stream_t stream);

/* Sensing fused into the step (the composition EcoEvoJax/sim.py:652-660 makes per tick: sensor_model ->
 * step_agents), for a Forager whose observation is its n_rays wall distances + own energy: the warp that
 * steps an agent first casts its rays against the wall grid of fgx_raycast_walls_grid (same arguments,
 * same bits) while the agent's weights are in flight, feeds the distances to the policy from shared memory
 * and writes them to dist / hit (float32 / int32 [n, n_rays]; hit may be NULL).  Equals
 * fgx_raycast_walls_grid followed by fgx_forager_step with obs = dist.                              */
int fgx_forager_sense_step(const fgx_forager_params_t* p, int64_t n, const float* active_state, const float* energy_in,
                           const fgx_forager_state_t* state, const fgx_wc_policy_t* policy, float* age,
                           int32_t n_rays, float ray_span, float ray_length, const float* wx1, const float* wy1,
                           const float* wx2, const float* wy2, int64_t n_walls, float x_min, float x_max, float y_min,
                           float y_max, void* wall_scratch, size_t wall_scratch_bytes, int32_t reuse_grid, float* dist,
                           int32_t* hit, fgx_stream_t stream);
/* WilsonCowan.step_policy (wilson_cowan.py:34-51) on its own, vmapped: obs float32[n, n_obs],
 * actions float32[n, n_act]; Z updated in place; inactive slots (active_state == 0) are skipped. */
int fgx_wilson_cowan_step(int64_t n, int32_t n_neurons, int32_t n_obs, int32_t n_act, float policy_dt,
                          float action_scale, const float* active_state, const float* obs,
                          const fgx_wc_policy_t* policy, float* actions, fgx_stream_t stream);

typedef struct {
  float dt, x_min, x_max, blossom_thresh; /* sim.py:266,278-279,660 */
} fgx_resource_params_t;
typedef struct {
  float *X, *Y, *energy, *rad, *time_above_blossom; /* float32[n] */
  uint8_t* blossom;                                 /* bool[n] */
  uint32_t* key;                                    /* uint32[n,2] */
} fgx_resource_state_t;
/* step_agents + vmapped Resource.step_agent (sim.py:264-303); in place; energy_out float32[n] */
int fgx_resource_step(const fgx_resource_params_t* p, int64_t n, const float* active_state, const float* energy_out,
                      const fgx_resource_state_t* state, const float* growth_rate, const float* decay_rate,
                      float* age, fgx_stream_t stream);
/* forager_resource_interaction (sim.py:358-375): energy_in float32[n_foragers],
 * energy_out float32[n_resources]                                                        */
int fgx_forager_resource_interaction(int64_t n_foragers, const float* fx, const float* fy, int64_t n_resources,
                                     const float* rx, const float* ry, const float* rrad, const float* renergy,
                                     float* energy_in, float* energy_out, fgx_stream_t stream);
/* sensor_model (sim.py:388-468): rays of every active forager against the entity table
 * [foragers, resources, walls] with the reference's distance prefilter (sim.py:435) and
 * min / first-argmin reduction; obs float32[n_foragers, 3*n_rays] = per ray (energy, type,
 * distance) of the hit entity, zeros for inactive foragers.                               */
int fgx_sensor_model(int64_t n_foragers, const float* fx, const float* fy, const float* fang, const float* frad,
                     const float* fenergy, const float* factive, const int32_t* ftype, int64_t n_resources,
                     const float* rx, const float* ry, const float* rrad, const float* renergy, const float* ractive,
                     const int32_t* rtype, int64_t n_walls, const float* wx1, const float* wy1, const float* wx2,
                     const float* wy2, int32_t n_rays, float ray_span, float ray_length, float* obs,
                     fgx_stream_t stream);

/* Row operations of the foraging example.  The reference runs these callbacks inside serial
 * lax.fori_loops (agent_methods.py:29-35,43-49,55-61); here all k rows are processed in one
 * call, k and n_active being int32 DEVICE scalars.  add_agents threads a PRNG key through its
 * loop: key_in / key_out are uint32[2] device arrays and chain_scratch is uint32[n,2].        */
/* vmapped Forager.create_agent (sim.py:27-71).  p->x_min..y_max hold the CREATE bounds of X, Y
 * (sim.py:39,43).  init_keys uint32[5,n,2] receives the five WilsonCowan.create_policy keys
 * (wilson_cowan.py:19-26), leaf-major; the caller draws J,tau,E,B,D with fgx_threefry_uniform. */
int fgx_forager_create(const fgx_forager_params_t* p, int64_t n, const uint32_t* create_keys,
                       const float* active_state, const fgx_forager_state_t* state, float* Z, uint32_t* init_keys,
                       float* age, fgx_stream_t stream);
/* remove_agents with Forager.remove_agent (sim.py:139-151) */
int fgx_forager_remove(int64_t n, int32_t n_neurons, const int32_t* remove_ids, const int32_t* k,
                       const fgx_forager_state_t* state, float* Z, float* active_state, float* age,
                       fgx_stream_t stream);
/* add_agents with Forager.add_agent (sim.py:154-210): rows [*n_active, *n_active + *k) become
 * children of rows good_ids[0..*k): parent weights + noise_scale * normal() (noise_scale constant,
 * or clip(0.1/(age+0.01), 0.001, 0.1) when adaptive_noise, the notebook variant)              */
int fgx_forager_add(int64_t n, int32_t n_neurons, int32_t n_obs, int32_t n_act, const int32_t* good_ids,
                    const int32_t* n_active, const int32_t* k, const uint32_t* key_in, uint32_t* key_out,
                    float noise_scale, int32_t adaptive_noise, const fgx_forager_state_t* state,
                    const fgx_wc_policy_t* policy, float* active_state, float* age, uint32_t* chain_scratch,
                    fgx_stream_t stream);
/* set_agents with Forager.set_agent (sim.py:213-225) */
int fgx_forager_set(int64_t n, const int32_t* set_ids, const int32_t* k, const fgx_forager_state_t* state,
                    fgx_stream_t stream);
/* The key held by add_agents' serial chain (agent_methods.py:29-35) after `steps` (device int32) added rows, without
 * touching any row: mode 0 = Forager.add_agent's `key, *ks = split(key, 6)` (sim.py:181), mode 1 =
 * Resource.add_agent's two splits (sim.py:322,335).  Lets a rank of a sharded set start the chain at the global
 * index of its first new row.                                                                          */
int fgx_key_chain_advance(int32_t mode, const uint32_t* key_in, const int32_t* steps, uint32_t* key_out,
                          fgx_stream_t stream);
/* vmapped Resource.create_agent (sim.py:231-261); x_lo..y_hi = create bounds of X, Y (sim.py:248-249) */
int fgx_resource_create(int64_t n, float x_lo, float x_hi, float y_lo, float y_hi, const uint32_t* create_keys,
                        const float* active_state, const fgx_resource_state_t* state, float* growth_rate,
                        float* decay_rate, float* age, fgx_stream_t stream);
/* remove_agents with Resource.remove_agent (sim.py:306-313) */
int fgx_resource_remove(int64_t n, const int32_t* remove_ids, const int32_t* k, const fgx_resource_state_t* state,
                        float* active_state, float* age, fgx_stream_t stream);
/* add_agents with Resource.add_agent (sim.py:316-344); x_min..y_max = the clip bounds */
int fgx_resource_add(int64_t n, float x_min, float x_max, float y_min, float y_max, const int32_t* good_ids,
                     const int32_t* n_active, const int32_t* k, const uint32_t* key_in, uint32_t* key_out,
                     const fgx_resource_state_t* state, float* growth_rate, float* decay_rate, float* active_state,
                     float* age, uint32_t* chain_scratch, fgx_stream_t stream);
/* set_agents with Resource.set_agent (sim.py:347-355) */
int fgx_resource_set(int64_t n, const int32_t* set_ids, const int32_t* k, const fgx_resource_state_t* state,
                     fgx_stream_t stream);

/* ---- K2: wolf-sheep-grass example (examples/basic/wolf_sheep_grass/sim.ipynb) --------------
 * Animal state leaves ([n,1] leaves as int32[n]); policy_key = Random_Policy state (random_policy.py). */
typedef struct {
  int32_t *X_pos, *Y_pos, *energy, *reproduce;
  uint32_t *key, *policy_key; /* uint32[n,2] */
} fgx_animal_state_t;
typedef struct {
  int32_t x_min, x_max, y_min, y_max, torous, delta_energy; /* space + animal.params (sim.ipynb:44-45,73) */
  float reproduction_rate;
} fgx_animal_params_t;
/* vmapped Animal.create_agent (sim.ipynb:40-70) */
int fgx_animal_create(int64_t n, int32_t x_max, int32_t y_max, int32_t delta_energy, const uint32_t* create_keys,
                      const float* active_state, const fgx_animal_state_t* state, float* age, fgx_stream_t stream);
/* step_agents + vmapped Animal.step_agent (sim.ipynb:72-122) with Random_Policy.step_policy
 * (random_policy.py:17-23); in place; energy_in int32[n] */
int fgx_animal_step(const fgx_animal_params_t* p, int64_t n, const float* active_state, const int32_t* energy_in,
                    const fgx_animal_state_t* state, float* age, fgx_stream_t stream);
/* remove / add / set with Animal.remove_agent, add_agent, set_agent (sim.ipynb:124-170); k, n_active on device */
int fgx_animal_remove(int64_t n, const int32_t* remove_ids, const int32_t* k, const fgx_animal_state_t* state,
                      float* active_state, float* age, fgx_stream_t stream);
int fgx_animal_add(int64_t n, const int32_t* copy_ids, const int32_t* n_active, const int32_t* k,
                   const fgx_animal_state_t* state, float* active_state, float* age, fgx_stream_t stream);
int fgx_animal_set(int64_t n, const int32_t* set_ids, const int32_t* k, const fgx_animal_state_t* state,
                   fgx_stream_t stream);
/* vmapped Grass.create_agent / step_agent (sim.ipynb:177-218); fully_grown bool[n], energy_out float32[n] */
int fgx_grass_create(int64_t n, int32_t x_max, int32_t regrowth_time, const uint32_t* create_keys,
                     const int32_t* unique_id, int32_t* X_pos, int32_t* Y_pos, uint8_t* fully_grown,
                     int32_t* count_down, float* age, fgx_stream_t stream);
int fgx_grass_step(int64_t n, int32_t regrowth_time, const float* energy_out, uint8_t* fully_grown,
                   int32_t* count_down, fgx_stream_t stream);
/* interaction (sim.ipynb:238-290): wolves_energy_in int32[W] = sheep on the wolf's cell,
 * sheeps_energy_in int32[S] = fully grown grass on the sheep's cell, sheeps_eaten float32[S],
 * grasses_eaten float32[G]; cell equality evaluated as a histogram join over the grid
 * [-1, x_max] x [-1, y_max] (no active-state check, as in the reference)                     */
size_t fgx_wolf_sheep_interaction_scratch_bytes(int32_t x_max, int32_t y_max);
int fgx_wolf_sheep_interaction(int32_t x_max, int32_t y_max, int64_t n_wolves, const int32_t* wX, const int32_t* wY,
                               int64_t n_sheeps, const int32_t* sX, const int32_t* sY, int64_t n_grasses,
                               const int32_t* gX, const int32_t* gY, const uint8_t* g_fully_grown,
                               int32_t* wolves_energy_in, int32_t* sheeps_energy_in, float* sheeps_eaten,
                               float* grasses_eaten, void* scratch, size_t scratch_bytes, fgx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FORAGAX_B200_H */
