/*
 * xp_physics_cuda.h -- C ABI of the B200-native (sm_100a) replacement for the xp_physics
 * swept-sphere vs triangle collision path of bjrnmrtns/xp-game-engine.
 *
 * The reference has NO FFI today: its boundary is the Rust crate's public fn API
 * (xp_physics/src/lib.rs:9-15,17,29).  Each entry point below names the reference item it
 * replaces; INTEGRATION.md shows the Rust `extern "C"` block a maintainer would add.
 *
 * Conventions
 *   - every function returns XP_OK (0) or a negative XP_E* code; nothing unwinds across the ABI;
 *   - the reference's panics become per-sphere status bits (XP_ST_*);
 *   - plain pointers + sizes only; `*_dev` variants take device pointers on the scene's device
 *     and are asynchronous on the scene's stream, the others take host pointers, copy, and
 *     return after the results are back in the caller's buffers;
 *   - one xp_scene per device; calls on one scene are serialised by the caller (the reference
 *     is single-threaded); different scenes are independent;
 *   - there is no CPU fallback: without a usable CUDA device every call fails with XP_ECUDA;
 *   - inputs: triangle vertices must be finite; sphere centres / movements may be anything except
 *     +-infinity, and a non-zero movement must be long enough that |m|^2 does not underflow to zero
 *     (|m| > ~1e-19); NaN, exactly zero and huge finite values behave exactly like the reference.
 *     (For the two excluded inputs the reference reports "hits" at t = +inf; this library reports none.)
 *
 * All arithmetic is IEEE binary32.  XP_ARITH_EXACT (default) evaluates the reference's
 * operations in the reference's order without fused multiply-add and is bit-identical to
 * the Rust code's results; XP_ARITH_FAST lets the compiler contract a*b+c (<= 1e-5 relative).
 */
#ifndef XP_PHYSICS_CUDA_H
#define XP_PHYSICS_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- value types: byte-for-byte the reference's structs with nalgebra Vec3 = 3 x f32 ---- */
typedef struct { float x, y, z; } xp_vec3;                              /* nalgebra_glm::Vec3            */
typedef struct { xp_vec3 c; float r; } xp_sphere;                       /* xp_physics/src/sphere.rs:4-7   16 B */
typedef struct { xp_vec3 v0, v1, v2; } xp_triangle;                     /* xp_physics/src/triangle.rs:4-8 36 B */
typedef struct { xp_sphere sphere; xp_vec3 movement; } xp_response;     /* xp_physics/src/response.rs:5-8 28 B */
typedef struct { float time_to, distance_to;
                 xp_vec3 position, intersection; } xp_collision;        /* xp_physics/src/collision.rs:5-10 32 B */

typedef struct xp_scene xp_scene;   /* opaque: device triangle records + LBVH + workspaces */

/* ---- error codes ---- */
#define XP_OK        0
#define XP_EINVAL   -1   /* bad argument (null pointer, n_vec3 % 3 != 0 like lib.rs:19-24, ...) */
#define XP_ECUDA    -2   /* CUDA runtime error; xp_last_cuda_error() has the text               */
#define XP_ENOMEM   -3   /* device or host allocation failed                                     */
#define XP_ESTATE   -4   /* scene not built / no triangles set                                   */
#define XP_ECAP     -5   /* output capacity too small (xp_broadphase_pairs)                      */

/* ---- per-sphere status bits (where the reference panics) ---- */
#define XP_ST_RADIUS_NE_1          1u  /* assert_eq!(sphere.r, 1.0)  collision_detect.rs:161   */
#define XP_ST_ASSERT_NEG_DISTANCE  2u  /* assert!(d >= 0.0)          collision_response.rs:22  */
#define XP_ST_ITER_CAP             4u  /* iteration cap reached; the reference loop (lib.rs:32) has none */

/* ---- horizon_mode: the broadphase query volume ---- */
#define XP_HORIZON_INF   0u  /* ray t in [0,+inf): reference-faithful (hits at any t>0, collision_detect.rs:201-202) */
#define XP_HORIZON_UNIT  1u  /* ray t in [0,1]: Fauerby-correct, NOT what the reference computes */

/* ---- scene options (xp_scene_set_option) ---- */
#define XP_OPT_ARITH      1  /* XP_ARITH_*                                      */
#define XP_OPT_METHOD     2  /* XP_METHOD_*                                     */
#define XP_OPT_TIMING     3  /* 1: record CUDA events around every kernel stage (no extra synchronisation)
                                and count narrowphase tests per code path */
#define XP_OPT_SORT_SPHERES 4 /* 1: process spheres in Morton order of their centre (default 1) */
#define XP_OPT_MAX_PAIRS  5  /* capacity (pairs) of the two-stage candidate buffer per chunk */
#define XP_OPT_PRUNE      6  /* 0: evaluate detect on EVERY pair passing P, so narrowphase_tests == |P| and the
                                tested set is the bit-exact candidate set.
                                1: find the same closest collision with fewer tests: the fused methods skip subtrees
                                entered after the best t so far; LBVH_PAIRS sweeps growing distance windows
                                [0,4) [4,16) ... [1024,inf) and drops spheres already hit before the window.  Pays off
                                when a sphere has hundreds of candidates (volumetric scenes: 8.5x on config 3), costs
                                ~15 % on terrain.  tri_index may differ among exactly equal times.
                                2 (default): auto for LBVH_PAIRS: like 0 until a sweep saw > 96 candidates per sphere
                                (volumetric scenes), like 1 from then on; terrain-like scenes therefore run as 0 */
#define XP_OPT_FAR_EXACT  7  /* 1: also reproduce the reference's far-range grazing noise.  Its vertex / edge quadratics
                                lose ~3e-7 * D^2 of miss distance at range D, so beyond D ~ 18 it can report a hit on
                                a triangle whose (1 + 1e-4)-inflated box the path never enters; with this option a
                                second pass re-walks the spheres whose best hit lies more than 10 units away (or that
                                have none) with every box inflated by 8e-7 * D^2 and tests what that adds.  Default 0:
                                hit/miss can differ from the brute-force reference for exactly those contacts. */

#define XP_OPT_GRID_WALK  8  /* 1 (default): a scene set with xp_scene_set_heightmap is swept by walking the height grid
                                itself (computed cell addresses, no tree; csrc/xp_gridwalk.cuh) instead of the LBVH.
                                Same candidate set, bit for bit; applies to XP_METHOD_LBVH_PAIRS without distance
                                windows.  0: always the LBVH. */
#define XP_OPT_CONTACT_RECORD 9 /* what the contact entry points (xp_detect_contacts_*) write per sphere:
                                * 0 (default) xp_contact, 40 B;  1 xp_contact16, 16 B = the sphere's centre at the contact and
                                * the triangle -- what a PEER can use without the sphere's own Response (time_to, distance_to
                                * and intersection are functions of the Response, collision_detect.rs:188-194, which only the
                                * owning rank holds).  128 M spheres on 8 GPUs: 2.1 GB instead of 5.4 GB landed per GPU. */

#define XP_OPT_LEAF_RUNS 10  /* 1 (default): the LBVH walk of XP_METHOD_LBVH_PAIRS does not visit an internal node over exactly two
                                leaves (Morton neighbours; usually the two triangles of a quad): when the ray hits their union
                                box it emits both, and the narrowphase kernel applies P to each triangle's own box before testing
                                it.  The tested set stays P bit for bit (narrowphase_tests == |P|; broadphase_pairs counts what
                                was emitted, a few per cent more); a fifth of the node visits disappear.  In the windowed passes of
                                XP_OPT_PRUNE a run belongs to the pass that holds its union box's entry time (never later than
                                either triangle's own pass: the closest hit cannot change) and, with the infinite horizon, both
                                triangles are tested unfiltered.  xp_broadphase_pairs always walks without runs.  0: off. */

#define XP_ARITH_EXACT    0  /* no FMA, reference operation order: bit-identical to the reference's results         */
#define XP_ARITH_FAST     1  /* tolerance mode: FMA, hardware reciprocal / square root, reduced quadratics (<= 1e-5) */
#define XP_ARITH_EXACT_LITERAL 2 /* same results as 0 through the compiler's IEEE div / sqrt (the round-1 kernel; A/B)  */
#define XP_ARITH_FAST_LITERAL  3 /* the literal formulas with FMA contraction only (the round-1 fast mode; A/B)        */

#define XP_METHOD_WIDE_FUSED  0  /* 32-wide implicit tree over the Morton-sorted triangles; one WARP per
                                   sphere: the 32 lanes test a node's 32 child boxes (six coalesced 128 B loads),
                                   leaf hits go to a warp-shared queue in shared memory that is drained 32 pairs at
                                   a time through the narrowphase */
#define XP_METHOD_LBVH_FUSED  1  /* binary LBVH, one LANE per sphere with a private stack, same warp queue */
#define XP_METHOD_LBVH_PAIRS  2  /* default. Persistent-warp binary LBVH traversal (4 blocks of 256 per SM, 60 registers) emits every pair passing P to HBM; narrowphase over the pairs */
#define XP_METHOD_BRUTE       3  /* the reference's literal algorithm: every triangle for every sphere (lib.rs:33-35) */
#define XP_METHOD_WIDE_PAIRS  4  /* 32-wide tree traversal emits the pairs to HBM; narrowphase over the pairs */
#define XP_METHOD_Q4_PAIRS    5  /* quantised 4-wide collapse of the LBVH (64 B nodes, 8-bit boxes rounded outward):
                                   stage 1 emits a superset, stage 2 re-checks the exact P before the narrowphase */

/* ---- statistics of the last detect / response call on a scene ---- */
#define XP_STAGE_COUNT 12
typedef struct {
    uint64_t narrowphase_tests;   /* (sphere,triangle) pairs on which detect was evaluated          */
    uint64_t broadphase_pairs;    /* pairs passing P (BVH_PAIRS) */
    uint64_t node_visits;         /* internal LBVH nodes fetched */
    uint32_t kernel_launches;     /* kernels of this library launched by the call                    */
    uint32_t iterations_run;      /* outer response iterations executed (max over spheres)           */
    float    stage_ms[XP_STAGE_COUNT]; /* XP_OPT_TIMING=1: CUDA-event ms per stage, see XP_STAGE_*   */
    uint64_t path_tests[5];       /* XP_OPT_TIMING=1: tests per code path: back-face cull, parallel-far,
                                     face hit, full vertex+edge path, full path after a failed face test */
} xp_stats;

#define XP_STAGE_BUILD_BOUNDS   0
#define XP_STAGE_BUILD_MORTON   1
#define XP_STAGE_BUILD_SORT     2
#define XP_STAGE_BUILD_PACK     3
#define XP_STAGE_BUILD_TREE     4
#define XP_STAGE_BUILD_REFIT    5
#define XP_STAGE_RAY_SETUP      6
#define XP_STAGE_TRAVERSE       7   /* broadphase traversal (BVH_PAIRS) or fused traversal+narrowphase */
#define XP_STAGE_NARROWPHASE    8   /* narrowphase over pairs / brute force                            */
#define XP_STAGE_RESPONSE       9
#define XP_STAGE_FINALIZE      10
#define XP_STAGE_SPHERE_SORT   11

/* ---- scene lifetime ---- */
int  xp_scene_create(int device, xp_scene **out);
void xp_scene_destroy(xp_scene *s);
/* cudaStream_t (as void*) every later call on this scene is enqueued on; NULL = default stream. */
int  xp_scene_set_stream(xp_scene *s, void *cuda_stream);
int  xp_scene_set_option(xp_scene *s, int option, int64_t value);
int  xp_scene_get_stats(const xp_scene *s, xp_stats *out);

/* The `&[Triangle]` argument of collision_response (lib.rs:29): uploaded ONCE instead of
 * being walked on every call.  Copies; the caller keeps ownership. */
int  xp_scene_set_triangles(xp_scene *s, const xp_triangle *aos, uint64_t n);
int  xp_scene_set_triangles_dev(xp_scene *s, const xp_triangle *d_aos, uint64_t n);
/* The `&[Vec3]` argument of collision_response_non_trianulated (lib.rs:17-25): every three
 * consecutive Vec3 form one triangle; n_vec3 % 3 != 0 -> XP_EINVAL (the reference panics). */
int  xp_scene_set_positions(xp_scene *s, const xp_vec3 *soup, uint64_t n_vec3);
/* Row f2: heightmap terrain emitted as triangles ON THE DEVICE.  heights is a host
 * array of (nx+1) x (nz+1) floats, x-major (heights[ix*(nz+1)+iz]); quad (ix,iz) yields the triangles
 * (p00,p01,p11), (p00,p11,p10) with p_ab = (x0+spacing*(ix+a), h, z0+spacing*(iz+b)), x outer / z inner --
 * the order of low_poly_nice_graphics/src/world/heightmap_loader.rs:19-63.  Equivalent to
 * xp_scene_set_triangles on that soup, with 4 B per vertex uploaded instead of 36 B per triangle. */
int  xp_scene_set_heightmap(xp_scene *s, const float *heights, uint32_t nx, uint32_t nz, float x0, float z0,
                            float spacing);
/* Row f2, clipmap half: the triangles of the geometry clipmap (game/src/graphics/clipmap.rs:566-625 parts,
 * vertices placed like game/src/shaders/shader_clipmap.vert:43-65) emitted on the device from the per-level
 * toroidal height textures: heights[levels][size][size] (host), vertex (ix, iz) of level L =
 * (ix * unit0 * 2^L, heights[L][iz mod size][ix mod size], iz * unit0 * 2^L).  parts holds n_parts rows of 8
 * int32 {level, first vertex x, first vertex z, vertices per side x, z, first triangle, triangles, 0}; a part
 * emits its quads x fastest, all [i0,i2,i1] triangles first, then all [i3,i1,i2] (create_grid,
 * clipmap.rs:752-786).  xp_physics_cuda/clipmap.py builds both arrays (the host mirror of `Clipmap`). */
int  xp_scene_set_clipmap(xp_scene *s, const float *heights, uint32_t levels, uint32_t size,
                          const int32_t *parts, uint32_t n_parts, float unit0);
/* (Re)build the device LBVH: per-triangle normal / plane constant / inflated AABB, 30-bit
 * Morton codes, radix sort, hierarchy and boxes in one bottom-up pass (the binary radix tree of
 * Karras 2012, built the way of Apetrei 2014).  New: the reference has no
 * broadphase (lib.rs:33-35 walks every triangle). */
int  xp_scene_build(xp_scene *s);
uint64_t xp_scene_triangle_count(const xp_scene *s);

/* ---- a11 + the min_by of lib.rs:33-42: closest collision of each sphere against the scene.
 * out[i] is written only where hit[i] != 0; tri_index (nullable) is the index, in upload
 * order, of the triangle the reference's min_by would return (ties -> later triangle), or
 * 0xFFFFFFFF; status (nullable) carries XP_ST_RADIUS_NE_1. */
int  xp_detect_batch(xp_scene *s, const xp_response *in, uint64_t n, uint32_t horizon_mode,
                     xp_collision *out, uint8_t *hit, uint32_t *tri_index, uint32_t *status);
int  xp_detect_batch_dev(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t horizon_mode,
                         xp_collision *d_out, uint8_t *d_hit, uint32_t *d_tri_index,
                         uint32_t *d_status);
/* The same sweep for a STREAM of batches (new; the reference's call is synchronous, lib.rs:29):
 * submit queues the H2D copy, the sweep and the D2H copies of one batch and returns a ticket;
 * wait blocks until that batch's results are in the caller's buffers and fills xp_scene_get_stats
 * for it.  Up to XP_ASYNC_SLOTS batches may be in flight, so one batch's copies run under another
 * batch's kernels.  The caller's buffers must stay valid and untouched until wait returns, and
 * should be page-locked (pageable memory makes the copies synchronous).  submit returns XP_ESTATE
 * when every slot is in flight, wait returns XP_ESTATE for a ticket that is not in flight. */
#define XP_ASYNC_SLOTS 2
int  xp_detect_batch_submit(xp_scene *s, const xp_response *in, uint64_t n, uint32_t horizon_mode,
                            xp_collision *out, uint8_t *hit, uint32_t *tri_index, uint32_t *status,
                            int *ticket);
int  xp_detect_batch_wait(xp_scene *s, int ticket);

/* ---- compact results: 8 B per sphere over the host link instead of 41 B ----
 * All four fields of the reference's Collision are functions of (c, m, t) alone (collision_detect.rs:188-194:
 * position = c + m*t, distance_to = |m|*t, intersection = position + m/|m|*r), so a caller that keeps its own
 * Responses only needs t and the triangle: out[i] = { time_to, tri_index } with tri_index = 0xFFFFFFFF for "no
 * collision" (time_to is then 0).  cpp/xp_physics.hpp (xp::collision_from_hit) and xp_physics_cuda.api
 * (collision_from_hit) rebuild the full Collision from it, bit for bit.  status (nullable) as in xp_detect_batch.
 * The _submit form shares tickets, slots and xp_detect_batch_wait with xp_detect_batch_submit. */
typedef struct { float time_to; uint32_t tri_index; } xp_hit;
int  xp_detect_batch_compact(xp_scene *s, const xp_response *in, uint64_t n, uint32_t horizon_mode,
                             xp_hit *out, uint32_t *status);
int  xp_detect_batch_compact_dev(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t horizon_mode,
                                 xp_hit *d_out, uint32_t *d_status);
int  xp_detect_batch_compact_submit(xp_scene *s, const xp_response *in, uint64_t n, uint32_t horizon_mode,
                                    xp_hit *out, uint32_t *status, int *ticket);

/* ---- a16: collision_response (lib.rs:29-62) for a batch of independent spheres.
 * max_iter == 0: unbounded like the reference (an internal guard of 65536 iterations sets
 * XP_ST_ITER_CAP).  out[i] = the Response the reference returns; where a panic site is hit
 * the sphere stops, the status bit is set and out[i] is the Response current at that point.
 * iterations / status / last_slide_normal (collision_response.rs:16) are nullable. */
int  xp_collision_response_batch(xp_scene *s, const xp_response *in, uint64_t n,
                                 uint32_t max_iter, uint32_t horizon_mode, xp_response *out,
                                 uint32_t *iterations, uint32_t *status,
                                 xp_vec3 *last_slide_normal);
int  xp_collision_response_batch_dev(xp_scene *s, const xp_response *d_in, uint64_t n,
                                     uint32_t max_iter, uint32_t horizon_mode,
                                     xp_response *d_out, uint32_t *d_iterations,
                                     uint32_t *d_status, xp_vec3 *d_last_slide_normal);

/* ---- row f4 of SURVEY.md section 8 (beyond the reference): Fauerby's collide-and-slide on the device ----
 * What the author's note at collision_response.rs:9-14 points at, for a batch: every sphere moves by its velocity;
 * only contacts inside the step count (t <= 1: the sweep runs with XP_HORIZON_UNIT); the sphere stops 0.005 short of
 * the first contact and continues with its velocity projected onto the sliding plane through the TRUE contact point
 * (the closest point of the hit triangle to the sphere centre at the time of impact), up to max_iter times
 * (0 = 5).  radii (nullable): an ellipsoid's radii -- positions and velocities are taken to its space, in which it
 * is the unit sphere the sweep requires (collision_detect.rs:161), and back; the scene's triangles must already be
 * in that space.  out_velocity = what is left of the step (zero for a sphere that came to rest or ran free),
 * handled (nullable) = contacts handled per sphere.  XP_METHOD_LBVH_PAIRS only (the default method). */
int  xp_collide_and_slide_batch(xp_scene *s, const xp_vec3 *position, const xp_vec3 *velocity, uint64_t n,
                                uint32_t max_iter, const float radii[3], xp_vec3 *out_position,
                                xp_vec3 *out_velocity, uint32_t *handled);
int  xp_collide_and_slide_batch_dev(xp_scene *s, const xp_vec3 *d_position, const xp_vec3 *d_velocity, uint64_t n,
                                    uint32_t max_iter, const float radii[3], xp_vec3 *d_out_position,
                                    xp_vec3 *d_out_velocity, uint32_t *d_handled);

/* ---- the 1-sphere calls with the reference's exact shapes (host pointers) ----
 * sphere_triangle_detect_collision (collision_detect.rs:155): returns 1 = Some, 0 = None,
 * XP_EINVAL where the reference panics on r != 1.  Evaluated on the device. */
int  xp_sphere_triangle_detect_collision(int device, const xp_response *response,
                                         const xp_triangle *triangle, xp_collision *out);
/* sphere_triangle_calculate_response (collision_response.rs:6): returns 0, or 1 where
 * assert!(d >= 0) would fire.  Evaluated on the device. */
int  xp_sphere_triangle_calculate_response(int device, const xp_response *response,
                                           const xp_collision *collision, xp_response *out);

/* ---- broadphase candidate set, for the bit-exact check against predicate P ----
 * Writes the (sphere, triangle-upload-index) pairs whose triangle AABB, inflated by
 * 1 + 1e-4, is hit by the ray c + t*m over the horizon.  Order unspecified.  If more than
 * cap pairs exist, *count still receives the total and XP_ECAP is returned. */
int  xp_broadphase_pairs(xp_scene *s, const xp_response *in, uint64_t n, uint32_t horizon_mode,
                         uint64_t cap, uint32_t *sphere_idx, uint32_t *tri_idx, uint64_t *count);

/* ---- LBVH introspection for tests (host buffers, nullable) ----
 * sorted_tri[j]  = upload index of the triangle stored in leaf j (Morton order)
 * morton[j]      = its 30-bit code
 * node_child[2i], node_child[2i+1] = children of internal node i (>= 0 internal, < 0 leaf ~c)
 * node_lo/hi[i]  = AABB of internal node i. */
int  xp_scene_debug_tree(xp_scene *s, uint32_t *sorted_tri, uint32_t *morton, int32_t *node_child,
                         xp_vec3 *node_lo, xp_vec3 *node_hi);
/* node_run[2i], node_run[2i+1] = leaf run of the left / right child of internal node i (XP_OPT_LEAF_RUNS): l + 1 when that
 * child is an internal node over exactly the leaves l and l + 1, else 0. */
int  xp_scene_debug_leaf_runs(xp_scene *s, int32_t *node_run);

/* ---- contact record for the multi-GPU allgather ----
 * 40 B per sphere: the 32 B collision + tri_index + (hit | status << 8 | iterations << 16). */
typedef struct { xp_collision col; uint32_t tri_index; uint32_t flags; } xp_contact;
/* XP_OPT_CONTACT_RECORD = 1: position = c + m * time_to (all zero without a hit), tri_index = 0xFFFFFFFF without a hit.
 * Every buffer / slot argument of the contact entry points then counts records of this size. */
typedef struct { xp_vec3 position; uint32_t tri_index; } xp_contact16;
/* Closest collision per sphere written as xp_contact records straight into a (send) buffer,
 * e.g. this rank's slot of the NCCL allgather buffer.  Device pointers. */
int  xp_detect_contacts_dev(xp_scene *s, const xp_response *d_in, uint64_t n,
                            uint32_t horizon_mode, xp_contact *d_out);

/* ---- fused contact build + all-gather over NVLink peer memory (one process per GPU) ----
 * Each rank allocates its gather buffer with xp_exchange_alloc, the 64-byte IPC handles are exchanged by
 * the host (any transport), every rank maps its peers' buffers with xp_exchange_open, and
 * xp_detect_contacts_exchange_dev then writes each sphere's 40 B contact record into ALL buffers (its own
 * included, bufs[] holds n_bufs device pointers) at record index slot_first + i -- the kernel that builds
 * the records performs the all-gather with P2P stores.  slot_first must be even (8-byte aligned spans).
 * The caller synchronises the ranks (a barrier) before reading the buffers. */
int  xp_exchange_alloc(xp_scene *s, uint64_t bytes, void **dev_ptr, uint8_t ipc_handle[64]);
int  xp_exchange_free(xp_scene *s, void *dev_ptr);
int  xp_exchange_open(xp_scene *s, const uint8_t ipc_handle[64], void **peer_ptr);
int  xp_exchange_close(xp_scene *s, void *peer_ptr);
int  xp_detect_contacts_exchange_dev(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t horizon_mode,
                                     void *const *bufs, int n_bufs, uint64_t slot_first);
/* The same with ONE store per record: multicast_buf is a multicast address of the NVSwitch fabric bound to
 * every rank's gather buffer (cuMulticast* / torch symmetric memory set it up; 16-byte aligned), so a span
 * leaves this GPU once and the switch replicates it into all buffers, this rank's included
 * (multimem.st).  slot_first must be even.  The caller synchronises the ranks before reading. */
int  xp_detect_contacts_multicast_dev(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t horizon_mode,
                                      void *multicast_buf, uint64_t slot_first);
/* A LOOP of steps: submit queues one step and returns; the sweep runs on the scene's stream, the record
 * build + exchange behind it on the stream given to xp_scene_set_exchange_stream, so the NEXT step's sweep
 * overlaps this step's exchange (two steps may be in flight; each owns the buffers its exchange reads).
 * The caller orders its cross-rank barrier after the exchange on that second stream.  wait blocks until
 * the step's records have left this GPU, fills xp_scene_get_stats for it, and returns XP_ECAP in the rare
 * case that the optimistic candidate buffer overflowed: redo that step with the synchronous call.
 * self_index >= 0 names this rank's own buffer in bufs[]: the kernel then fills only that buffer and the
 * copy engines carry the span to the others (peer stores issued by SMs stall a concurrent sweep while
 * NVLink is saturated; DMA copies do not); self_index = -1 keeps the kernel's P2P stores.
 * n <= 2^24 per step; d_in must stay untouched until wait returns. */
int  xp_scene_set_exchange_stream(xp_scene *s, void *cuda_stream);
int  xp_detect_contacts_exchange_submit(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t horizon_mode,
                                        void *const *bufs, int n_bufs, int self_index, uint64_t slot_first,
                                        int *ticket);
int  xp_detect_contacts_exchange_wait(xp_scene *s, int ticket);

/* ---- device self-check of the narrowphase's arithmetic building blocks (csrc/xp_debug.cu) ----
 * The exact mode's division / square-root sequences against the compiler's IEEE intrinsics on every square-root
 * input of the guarded range and on n_random hashed + structured operand pairs, and the operand-side decisions
 * (edge parameter, face interval, shared-denominator minimum) against the quotient-side ones of the reference.
 * counters[0..5] = mismatches per check (sqrt, div random, div per mantissa, edge parameter, face interval,
 * vertex minimum); all zero = the hand-rolled forms are exact. */
int  xp_debug_check_arith(int device, uint64_t n_random, uint64_t seed, uint64_t counters[8]);

/* ---- measurement helpers ---- */
/* Dependent-free FFMA loop on every SM: the fp32 CUDA-core peak used as roofline denominator. */
int  xp_probe_fp32_peak(int device, double *tflops);
int  xp_device_count(int *count);
const char *xp_strerror(int code);
const char *xp_last_cuda_error(void);
const char *xp_version(void);

#ifdef __cplusplus
}
#endif
#endif /* XP_PHYSICS_CUDA_H */
