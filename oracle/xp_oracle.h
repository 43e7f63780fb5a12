/*
 * xp_oracle.h -- CPU ORACLE for the xp_physics swept-sphere vs triangle path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may build, load or call it.  The
 * product (xp-game-engine_b200/csrc -> libxp_physics_cuda.so) never links or calls it
 * and has no CPU fallback.
 *
 * What it is: a plain-C, operation-for-operation restatement of the reference's
 * Rust code (all citations relative to /root/reference):
 *   xp_physics/src/collision_detect.rs:6-213   detect (gate, face, vertices, edges, min)
 *   xp_physics/src/collision_response.rs:6-35  response step
 *   xp_physics/src/triangle.rs:11-54           point_in_triangle, plane_constant, normal
 *   xp_physics/src/lib.rs:17-62                collision_response loop, non_trianulated
 *   xp_physics/src/collision.rs:3              DISTANCE_EPSILON
 *   xp_math/src/lib.rs:3-12                    get_roots
 * and of the nalgebra-glm 0.8.0 (nalgebra 0.22) primitives those files call -- a
 * crates.io dependency that is NOT under /root/reference (xp_physics/Cargo.toml:10,
 * no Cargo.lock), restated from its published algorithm:
 *   dot(a,b)  = (a0*b0 + a1*b1) + a2*b2        length(a) = sqrt(dot(a,a))
 *   normalize(a) = a / length(a) component-wise division
 *   cross(a,b) = (a1*b2 - a2*b1, a2*b0 - a0*b2, a0*b1 - a1*b0)
 *   triangle_normal(p1,p2,p3) = normalize(cross(p2-p1, p3-p1))
 * All arithmetic is IEEE-754 binary32, round-to-nearest, NO fused multiply-add
 * (build with -ffp-contract=off, see oracle/Makefile).
 *
 * PARITY PINNING: the Rust reference cannot be built here (no cargo/rustc in the image),
 * so there is no oracle/_ref.  The oracle is pinned against every known-answer test the
 * reference holds for this path (tests/test_oracle_kat.py):
 *   collision_detect.rs:222-232  face   time_to == 0.5
 *   collision_detect.rs:235-246  vertex time_to == 0.375
 *   collision_detect.rs:249-260  edge   time_to == 0.5
 *   triangle.rs:62-76            point_in_triangle (6 assertions)
 * The reference has NO test for collision_response / sphere_triangle_calculate_response /
 * get_roots / misses / position,intersection,distance_to: for those this file is
 * "parity unpinned" -- its fidelity rests on line-by-line review against the citations.
 *
 * Panics in the reference become status bits here (XPO_ST_*), see each function.
 */
#ifndef XP_ORACLE_H
#define XP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float x, y, z; } xpo_vec3;
typedef struct { xpo_vec3 c; float r; } xpo_sphere;              /* sphere.rs:4-7    */
typedef struct { xpo_vec3 v0, v1, v2; } xpo_triangle;            /* triangle.rs:4-8  */
typedef struct { xpo_sphere sphere; xpo_vec3 movement; } xpo_response; /* response.rs:5-8 */
typedef struct {                                                 /* collision.rs:5-10 */
    float time_to, distance_to;
    xpo_vec3 position, intersection;
} xpo_collision;

#define XPO_DISTANCE_EPSILON 0.001f   /* collision.rs:3 */

/* status bits: where the reference would panic */
#define XPO_ST_RADIUS_NE_1          1u  /* assert_eq!(sphere.r, 1.0) collision_detect.rs:161 */
#define XPO_ST_ASSERT_NEG_DISTANCE  2u  /* assert!(d >= 0.0)        collision_response.rs:22 */
#define XPO_ST_ITER_CAP             4u  /* max_iter reached (reference has no cap, lib.rs:32) */

/* xp_math/src/lib.rs:3-12.  Returns 1 and writes roots, or 0 when det < 0. */
int xpo_get_roots(float a, float b, float c, float *r0, float *r1);

/* triangle.rs:11-32 */
int xpo_point_in_triangle(const xpo_vec3 *p0, const xpo_vec3 *p1, const xpo_vec3 *p2,
                          const xpo_vec3 *p);
/* triangle.rs:44-46 and :47-50 */
xpo_vec3 xpo_triangle_normal(const xpo_triangle *t);
float xpo_triangle_plane_constant(const xpo_triangle *t);

/* collision_detect.rs:155-213.  Returns 1 = Some(collision) (written to *out),
 * 0 = None, -1 = the reference would panic on r != 1.0. */
int xpo_detect(const xpo_response *resp, const xpo_triangle *tri, xpo_collision *out);

/* collision_response.rs:6-35.  Returns 0, or -1 where assert!(d>=0) would fire (then
 * *out is not written).  slide_normal (nullable) receives slide_plane_normal (:16). */
int xpo_respond(const xpo_response *resp, const xpo_collision *col, xpo_response *out,
                xpo_vec3 *slide_normal);

/* lib.rs:33-42: closest = min_by(time_to) over every Some(detect); ties go to the LATER
 * triangle.  Returns 1/0/-1 like xpo_detect; tri_index nullable. */
int xpo_closest(const xpo_response *resp, const xpo_triangle *tris, uint64_t n,
                xpo_collision *out, uint32_t *tri_index);

/* lib.rs:29-62.  max_iter == 0 means unbounded like the reference.  *out is the
 * Response the reference would return; when a panic site is hit the loop stops, the
 * status bit is set and *out is the Response that was current at the panic.
 * last_slide_normal (nullable) = slide_plane_normal of the last response step (0 if none). */
void xpo_collision_response(const xpo_response *in, const xpo_triangle *tris, uint64_t n,
                            uint32_t max_iter, xpo_response *out, uint32_t *iterations,
                            uint32_t *status, xpo_vec3 *last_slide_normal);

/* lib.rs:17-27: every 3 consecutive Vec3 form a triangle.  Returns -1 if n_vec3 % 3 != 0
 * (the reference panics on the out-of-bounds chunk index), else 0. */
int xpo_collision_response_non_trianulated(const xpo_response *in, const xpo_vec3 *soup,
                                           uint64_t n_vec3, uint32_t max_iter,
                                           xpo_response *out, uint32_t *iterations,
                                           uint32_t *status);

/* Batches (pthreads over spheres; threads <= 0 -> all cores).  tests (nullable) receives the
 * number of xpo_detect evaluations performed. */
void xpo_detect_batch(const xpo_response *in, uint64_t n_spheres, const xpo_triangle *tris,
                      uint64_t n_tris, int threads, xpo_collision *out, uint8_t *hit,
                      uint32_t *tri_index, uint32_t *status, uint64_t *tests);
void xpo_collision_response_batch(const xpo_response *in, uint64_t n_spheres,
                                  const xpo_triangle *tris, uint64_t n_tris, uint32_t max_iter,
                                  int threads, xpo_response *out, uint32_t *iterations,
                                  uint32_t *status, xpo_vec3 *last_slide_normal,
                                  uint64_t *tests);

/* ---- broadphase predicate P (NEW: the reference has no broadphase, lib.rs:33-35) ----
 * Ray c + t*m, t in [0, horizon], against AABB(tri) inflated by R = 1 + 1e-4 (fp32 add/sub).
 * Per axis k: m_k == 0 -> require lo_k <= c_k <= hi_k; else inv_k = 1/m_k (IEEE divide),
 * a = (lo_k - c_k)*inv_k, b = (hi_k - c_k)*inv_k, near/far chosen by the sign of inv_k,
 * tmin = near if near > tmin, tmax = far if far < tmax.  Accept iff tmin <= tmax.
 * horizon = +inf (clamped to FLT_MAX, so an entry time of +inf never passes) is the
 * reference-faithful query volume: hits at any t > 0 are accepted, collision_detect.rs:201-202.  Only sub, mul, one divide per axis per sphere and
 * comparisons: bit-identical on CPU and GPU. */
#define XPO_BROADPHASE_INFLATE 1.0001f
void xpo_triangle_aabb(const xpo_triangle *t, xpo_vec3 *lo, xpo_vec3 *hi);
int xpo_broadphase_accept(const xpo_response *resp, const xpo_triangle *tri, float horizon);
/* Writes up to cap pairs in (sphere-major, triangle-minor) order; *count = total found. */
void xpo_broadphase_pairs(const xpo_response *in, uint64_t n_spheres, const xpo_triangle *tris,
                          uint64_t n_tris, float horizon, uint64_t cap, uint32_t *sphere_idx,
                          uint32_t *tri_idx, uint64_t *count);

/* ---- float64 "grazing" shadow ----
 * Re-runs detect in double precision and returns the smallest normalised margin
 * |operand - threshold| / max(1, |operand|, |threshold|) over every branch it takes
 * (collision_detect.rs:166,175,184,33; triangle.rs:31; xp_math lib.rs:5;
 * collision_detect.rs:107,202).  hit/t (nullable) receive the double-precision outcome. */
double xpo_detect_margin_f64(const xpo_response *resp, const xpo_triangle *tri, int *hit,
                             double *t);
/* Per sphere: min margin over all triangles that pass P(horizon); used to excuse
 * hit/miss disagreements of the contracted ("fast") GPU mode. */
void xpo_grazing_batch(const xpo_response *in, uint64_t n_spheres, const xpo_triangle *tris,
                       uint64_t n_tris, float horizon, int threads, double *min_margin);

int xpo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
