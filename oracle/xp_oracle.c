/*
 * xp_oracle.c -- CPU ORACLE (test infrastructure, never shipped).  See xp_oracle.h for the
 * scope, the reference files restated and the parity-pinning statement.
 *
 * Style rule: every function below follows the cited Rust lines statement by statement,
 * keeps the reference's redundant work (normal computed twice, normalised twice, length
 * squared as sqrt*sqrt) and its quirks, and uses only unfused binary32 operations.
 * Build flags (oracle/Makefile): -O2 -ffp-contract=off -fno-fast-math.
 */
#include "xp_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <stdatomic.h>
#include <unistd.h>

/* ---------- nalgebra-glm 0.8.0 primitives (restated; see header) ---------- */

static inline xpo_vec3 v3(float x, float y, float z) { xpo_vec3 r = {x, y, z}; return r; }
static inline xpo_vec3 vadd(xpo_vec3 a, xpo_vec3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline xpo_vec3 vsub(xpo_vec3 a, xpo_vec3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline xpo_vec3 vscale(xpo_vec3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
/* nalgebra dot for 3-vectors: (a0*b0 + a1*b1) + a2*b2 */
static inline float vdot(xpo_vec3 a, xpo_vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float vlen(xpo_vec3 a) { return sqrtf(vdot(a, a)); }
/* normalize = self / norm, component-wise division */
static inline xpo_vec3 vnormalize(xpo_vec3 a) {
    float l = vlen(a);
    return v3(a.x / l, a.y / l, a.z / l);
}
static inline xpo_vec3 vcross(xpo_vec3 a, xpo_vec3 b) {
    return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

/* ---------- xp_math/src/lib.rs:3-12 ---------- */
int xpo_get_roots(float a, float b, float c, float *r0, float *r1) {
    float det = b * b - 4.0f * a * c;               /* :4  (4*a)*c, left to right */
    if (det < 0.0f) return 0;                        /* :5 */
    float sqrt_det = sqrtf(det);                     /* :8 */
    *r0 = (-b - sqrt_det) / (2.0f * a);              /* :9 */
    *r1 = (-b + sqrt_det) / (2.0f * a);              /* :10 */
    return 1;
}

/* ---------- xp_physics/src/triangle.rs ---------- */
int xpo_point_in_triangle(const xpo_vec3 *p0, const xpo_vec3 *p1, const xpo_vec3 *p2,
                          const xpo_vec3 *p) {
    xpo_vec3 v0 = vsub(*p2, *p0);                    /* :13 */
    xpo_vec3 v1 = vsub(*p1, *p0);                    /* :14 */
    xpo_vec3 v2 = vsub(*p, *p0);                     /* :15 */
    float dot00 = vdot(v0, v0);
    float dot01 = vdot(v0, v1);
    float dot02 = vdot(v0, v2);
    float dot11 = vdot(v1, v1);
    float dot12 = vdot(v1, v2);
    float inv_denom = 1.0f / (dot00 * dot11 - dot01 * dot01);   /* :23 */
    float u = (dot11 * dot02 - dot01 * dot12) * inv_denom;      /* :24 */
    float v = (dot00 * dot12 - dot01 * dot02) * inv_denom;      /* :25 */
    return u >= 0.0f && v >= 0.0f && u + v <= 1.0f;             /* :31 */
}

/* triangle.rs:34-38 */
static inline float plane_constant(xpo_vec3 point_on_plane, xpo_vec3 n) {
    return -(n.x * point_on_plane.x + n.y * point_on_plane.y + n.z * point_on_plane.z);
}

xpo_vec3 xpo_triangle_normal(const xpo_triangle *t) {            /* :44-46 */
    return vnormalize(vcross(vsub(t->v1, t->v0), vsub(t->v2, t->v0)));
}

float xpo_triangle_plane_constant(const xpo_triangle *t) {       /* :47-50 */
    xpo_vec3 normal = vnormalize(xpo_triangle_normal(t));
    return plane_constant(t->v0, normal);
}

/* ---------- xp_physics/src/collision_detect.rs ---------- */

/* :6-8 */
static inline float signed_distance(xpo_vec3 p, float pc, xpo_vec3 n) { return vdot(n, p) + pc; }

/* The Vec<f32> the reference grows with push/extend (at most 6 vertex roots + 3 edge t). */
typedef struct { float v[9]; int n; } tlist;
static inline void tpush(tlist *l, float x) { l->v[l->n++] = x; }

/* :10-21  seeds with vals[0], replaces on strict <  (a leading NaN stays) */
static int min_vals(const float *vals, int n, float *out) {
    if (n == 0) return 0;
    float m = vals[0];
    for (int i = 0; i < n; ++i)
        if (vals[i] < m) m = vals[i];
    *out = m;
    return 1;
}

/* :23-42 */
static int detect_triangle_collision(const xpo_sphere *sphere, const xpo_triangle *tri,
                                     xpo_vec3 movement, float sd, float pndm, float *t_out) {
    float t0 = (1.0f - sd) / pndm;
    float t1 = (-1.0f - sd) / pndm;
    if (!(t0 < t1)) { float tmp = t0; t0 = t1; t1 = tmp; }      /* :32 */
    if (t0 > 1.0f || t1 < 0.0f) return 0;                        /* :33 */
    t0 = (t0 > 0.0f) ? t0 : 0.0f;                                /* :36 */
    xpo_vec3 p = vadd(sphere->c, vscale(movement, t0));          /* :37 */
    if (xpo_point_in_triangle(&tri->v0, &tri->v1, &tri->v2, &p)) { *t_out = t0; return 1; }
    return 0;
}

/* :44-59 */
static void detect_vertex_collision(const xpo_sphere *sphere, xpo_vec3 v0, xpo_vec3 movement,
                                    float msl, tlist *ts) {
    float a = msl;
    float b = 2.0f * vdot(movement, vsub(sphere->c, v0));
    float c = vlen(vsub(v0, sphere->c));
    c = c * c - 1.0f;
    float r0, r1;
    if (xpo_get_roots(a, b, c, &r0, &r1)) { tpush(ts, r0); tpush(ts, r1); }
}

/* :83-112 */
static void detect_edge_collision(const xpo_sphere *sphere, xpo_vec3 v0, xpo_vec3 v1,
                                  xpo_vec3 movement, float msl, tlist *ts) {
    xpo_vec3 edge = vsub(v1, v0);
    xpo_vec3 base_to_vertex = vsub(v0, sphere->c);
    float base_to_vertex_length = vlen(base_to_vertex);
    float edge_length = vlen(edge);
    float edge_length_squared = edge_length * edge_length;
    float edge_dot_movement = vdot(edge, movement);
    float edge_dot_base_to_vertex = vdot(edge, base_to_vertex);

    float a = edge_length_squared * -msl + edge_dot_movement * edge_dot_movement;       /* :98 */
    float b = edge_length_squared * (2.0f * vdot(movement, base_to_vertex))
            - 2.0f * edge_dot_movement * edge_dot_base_to_vertex;                       /* :99-100 */
    float c = edge_length_squared * (1.0f - base_to_vertex_length * base_to_vertex_length)
            + edge_dot_base_to_vertex * edge_dot_base_to_vertex;                        /* :101-102 */
    float r0, r1;
    if (xpo_get_roots(a, b, c, &r0, &r1)) {
        float two[2] = {r0, r1};
        float t;
        min_vals(two, 2, &t);                                                           /* :105 */
        float f = (edge_dot_movement * t - edge_dot_base_to_vertex) / edge_length_squared;
        if (f >= 0.0f && f <= 1.0f) tpush(ts, t);                                       /* :107-109 */
    }
}

static void fill_collision(const xpo_sphere *sphere, xpo_vec3 movement, float movement_length,
                           xpo_vec3 normalized_movement, float t, xpo_collision *out) {
    xpo_vec3 position = vadd(sphere->c, vscale(movement, t));               /* :188 / :203 */
    out->time_to = t;
    out->distance_to = movement_length * t;
    out->position = position;
    out->intersection = vadd(position, vscale(normalized_movement, sphere->r));
}

/* :155-213 */
int xpo_detect(const xpo_response *resp, const xpo_triangle *tri, xpo_collision *out) {
    const xpo_sphere *sphere = &resp->sphere;
    xpo_vec3 movement = resp->movement;
    if (!(sphere->r == 1.0f)) return -1;                                     /* :161 */
    xpo_vec3 normal = vnormalize(xpo_triangle_normal(tri));                  /* :162 */
    xpo_vec3 normalized_movement = vnormalize(movement);                     /* :163 */

    if (vdot(normal, normalized_movement) > 0.0f) return 0;                  /* :166-168 */

    float pndm = vdot(normal, movement);                                     /* :170 */
    float sd = signed_distance(sphere->c, xpo_triangle_plane_constant(tri), normal); /* :171 */

    if (pndm == 0.0f && fabsf(sd) >= 1.0f) return 0;                         /* :175-177 */

    float movement_length = vlen(movement);                                  /* :179 */
    float msl = movement_length * movement_length;                           /* :180 */

    if (pndm != 0.0f && fabsf(sd) >= 1.0f) {                                 /* :184 */
        float t;
        if (detect_triangle_collision(sphere, tri, movement, sd, pndm, &t)) {
            fill_collision(sphere, movement, movement_length, normalized_movement, t, out);
            return 1;                                                        /* :188-194 */
        }
    }
    tlist ts; ts.n = 0;
    detect_vertex_collision(sphere, tri->v0, movement, msl, &ts);            /* :61-81 */
    detect_vertex_collision(sphere, tri->v1, movement, msl, &ts);
    detect_vertex_collision(sphere, tri->v2, movement, msl, &ts);
    detect_edge_collision(sphere, tri->v0, tri->v1, movement, msl, &ts);     /* :114-152 */
    detect_edge_collision(sphere, tri->v1, tri->v2, movement, msl, &ts);
    detect_edge_collision(sphere, tri->v2, tri->v0, movement, msl, &ts);
    float t;
    if (min_vals(ts.v, ts.n, &t)) {                                          /* :201 */
        if (t > 0.0f) {                                                      /* :202 */
            fill_collision(sphere, movement, movement_length, normalized_movement, t, out);
            return 1;
        }
    }
    return 0;
}

/* ---------- xp_physics/src/collision_response.rs:6-35 ---------- */
int xpo_respond(const xpo_response *resp, const xpo_collision *col, xpo_response *out,
                xpo_vec3 *slide_normal) {
    xpo_vec3 slide_plane_origin = col->intersection;                                   /* :15 */
    xpo_vec3 slide_plane_normal = vnormalize(vsub(col->position, slide_plane_origin)); /* :16 */
    xpo_vec3 original_destination = vadd(resp->sphere.c, resp->movement);              /* :18 */
    float d = vdot(original_destination, slide_plane_normal)
            + plane_constant(slide_plane_origin, slide_plane_normal);                  /* :20-21 */
    if (slide_normal) *slide_normal = slide_plane_normal;
    if (!(d >= 0.0f)) return -1;                                                       /* :22 */
    /* f32 * Vec3: each component is d * n_k */
    xpo_vec3 new_destination = vsub(original_destination,
                                    v3(d * slide_plane_normal.x, d * slide_plane_normal.y,
                                       d * slide_plane_normal.z));                     /* :24-25 */
    xpo_vec3 new_movement = vsub(new_destination, col->position);                      /* :26 */
    out->sphere.c = new_destination;
    out->sphere.r = 1.0f;                                                              /* :31 */
    out->movement = new_movement;
    return 0;
}

/* ---------- xp_physics/src/lib.rs ---------- */

/* :33-42  Iterator::min_by folds (acc, x): keeps acc only when the comparator says Less,
 * i.e. acc.time_to < x.time_to; anything else (ties, NaN) takes the later element. */
int xpo_closest(const xpo_response *resp, const xpo_triangle *tris, uint64_t n,
                xpo_collision *out, uint32_t *tri_index) {
    int have = 0;
    xpo_collision best;
    uint32_t best_i = 0;
    memset(&best, 0, sizeof best);
    for (uint64_t i = 0; i < n; ++i) {
        xpo_collision c;
        int r = xpo_detect(resp, &tris[i], &c);
        if (r < 0) return -1;
        if (r == 0) continue;
        if (!have) { best = c; best_i = (uint32_t)i; have = 1; }
        else if (best.time_to < c.time_to) { /* keep */ }
        else { best = c; best_i = (uint32_t)i; }
    }
    if (have) { *out = best; if (tri_index) *tri_index = best_i; }
    return have;
}

static uint64_t collision_response_counted(const xpo_response *in, const xpo_triangle *tris,
                                           uint64_t n, uint32_t max_iter, xpo_response *out,
                                           uint32_t *iterations, uint32_t *status,
                                           xpo_vec3 *last_slide_normal) {
    uint32_t it = 0, st = 0;
    uint64_t tests = 0;
    xpo_response response = *in;                                             /* :31 */
    xpo_vec3 last_n = v3(0.0f, 0.0f, 0.0f);
    for (;;) {                                                               /* :32 */
        if (max_iter != 0 && it >= max_iter) { st |= XPO_ST_ITER_CAP; break; }
        xpo_collision closest;
        int r = xpo_closest(&response, tris, n, &closest, NULL);             /* :33-42 */
        if (r < 0) { st |= XPO_ST_RADIUS_NE_1; break; }
        tests += n;
        if (r == 0) break;                                                   /* :48-53 */
        xpo_response next;
        xpo_vec3 sn;
        if (xpo_respond(&response, &closest, &next, &sn) != 0) {             /* :46 */
            st |= XPO_ST_ASSERT_NEG_DISTANCE; last_n = sn; break;
        }
        it += 1;                                                             /* :45 */
        last_n = sn;
        response = next;
        if (vlen(response.movement) < XPO_DISTANCE_EPSILON) break;           /* :55-60 */
    }
    *out = response;
    if (iterations) *iterations = it;
    if (status) *status = st;
    if (last_slide_normal) *last_slide_normal = last_n;
    return tests;
}

void xpo_collision_response(const xpo_response *in, const xpo_triangle *tris, uint64_t n,
                            uint32_t max_iter, xpo_response *out, uint32_t *iterations,
                            uint32_t *status, xpo_vec3 *last_slide_normal) {
    collision_response_counted(in, tris, n, max_iter, out, iterations, status, last_slide_normal);
}

int xpo_collision_response_non_trianulated(const xpo_response *in, const xpo_vec3 *soup,
                                           uint64_t n_vec3, uint32_t max_iter,
                                           xpo_response *out, uint32_t *iterations,
                                           uint32_t *status) {
    if (n_vec3 % 3 != 0) return -1;                                          /* :19-24 OOB panic */
    uint64_t n = n_vec3 / 3;
    xpo_triangle *tris = (xpo_triangle *)malloc((n ? n : 1) * sizeof(xpo_triangle));
    for (uint64_t i = 0; i < n; ++i) {                                       /* :18-25 */
        tris[i].v0 = soup[3 * i + 0];
        tris[i].v1 = soup[3 * i + 1];
        tris[i].v2 = soup[3 * i + 2];
    }
    xpo_collision_response(in, tris, n, max_iter, out, iterations, status, NULL);
    free(tris);
    return 0;
}

/* ---------- batches: a small pthread parallel-for over spheres ---------- */
int xpo_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

static int pick_threads(int threads) {
    int m = xpo_max_threads();
    if (threads <= 0 || threads > m) return m;
    return threads;
}

typedef void (*item_fn)(int64_t i, void *ctx, uint64_t *acc);
typedef struct { atomic_llong next; int64_t n; item_fn fn; void *ctx; } pf_shared;
typedef struct { pf_shared *sh; uint64_t acc; } pf_worker;

static void *pf_main(void *arg) {
    pf_worker *w = (pf_worker *)arg;
    for (;;) {
        int64_t i = atomic_fetch_add(&w->sh->next, 1);
        if (i >= w->sh->n) break;
        w->sh->fn(i, w->sh->ctx, &w->acc);
    }
    return NULL;
}

/* Runs fn(i) for i in [0,n) on `threads` threads (dynamic schedule); returns the sum of acc. */
static uint64_t parallel_for(int64_t n, int threads, item_fn fn, void *ctx) {
    pf_shared sh;
    atomic_init(&sh.next, 0);
    sh.n = n; sh.fn = fn; sh.ctx = ctx;
    int nt = pick_threads(threads);
    if (nt > n) nt = n > 0 ? (int)n : 1;
    pf_worker *ws = (pf_worker *)calloc((size_t)nt, sizeof(pf_worker));
    pthread_t *ths = (pthread_t *)calloc((size_t)nt, sizeof(pthread_t));
    for (int t = 0; t < nt; ++t) { ws[t].sh = &sh; ws[t].acc = 0; }
    for (int t = 1; t < nt; ++t) pthread_create(&ths[t], NULL, pf_main, &ws[t]);
    pf_main(&ws[0]);
    uint64_t total = ws[0].acc;
    for (int t = 1; t < nt; ++t) { pthread_join(ths[t], NULL); total += ws[t].acc; }
    free(ws); free(ths);
    return total;
}

typedef struct {
    const xpo_response *in; const xpo_triangle *tris; uint64_t n_tris; uint32_t max_iter;
    float horizon;
    xpo_collision *col; uint8_t *hit; uint32_t *tri_index; uint32_t *status;
    xpo_response *out; uint32_t *iterations; xpo_vec3 *last_slide_normal; double *min_margin;
} batch_ctx;

static void detect_item(int64_t s, void *vctx, uint64_t *acc) {
    batch_ctx *b = (batch_ctx *)vctx;
    xpo_collision c;
    uint32_t ti = 0xFFFFFFFFu;
    memset(&c, 0, sizeof c);
    int r = xpo_closest(&b->in[s], b->tris, b->n_tris, &c, &ti);
    if (b->col) b->col[s] = c;
    if (b->hit) b->hit[s] = (uint8_t)(r == 1);
    if (b->tri_index) b->tri_index[s] = (r == 1) ? ti : 0xFFFFFFFFu;
    if (b->status) b->status[s] = (r < 0) ? XPO_ST_RADIUS_NE_1 : 0u;
    *acc += b->n_tris;
}

void xpo_detect_batch(const xpo_response *in, uint64_t n_spheres, const xpo_triangle *tris,
                      uint64_t n_tris, int threads, xpo_collision *out, uint8_t *hit,
                      uint32_t *tri_index, uint32_t *status, uint64_t *tests) {
    batch_ctx b;
    memset(&b, 0, sizeof b);
    b.in = in; b.tris = tris; b.n_tris = n_tris;
    b.col = out; b.hit = hit; b.tri_index = tri_index; b.status = status;
    uint64_t total = parallel_for((int64_t)n_spheres, threads, detect_item, &b);
    if (tests) *tests = total;
}

static void response_item(int64_t s, void *vctx, uint64_t *acc) {
    batch_ctx *b = (batch_ctx *)vctx;
    uint32_t it = 0, st = 0;
    xpo_vec3 ln;
    *acc += collision_response_counted(&b->in[s], b->tris, b->n_tris, b->max_iter, &b->out[s],
                                       &it, &st, &ln);
    if (b->iterations) b->iterations[s] = it;
    if (b->status) b->status[s] = st;
    if (b->last_slide_normal) b->last_slide_normal[s] = ln;
}

void xpo_collision_response_batch(const xpo_response *in, uint64_t n_spheres,
                                  const xpo_triangle *tris, uint64_t n_tris, uint32_t max_iter,
                                  int threads, xpo_response *out, uint32_t *iterations,
                                  uint32_t *status, xpo_vec3 *last_slide_normal,
                                  uint64_t *tests) {
    batch_ctx b;
    memset(&b, 0, sizeof b);
    b.in = in; b.tris = tris; b.n_tris = n_tris; b.max_iter = max_iter;
    b.out = out; b.iterations = iterations; b.status = status;
    b.last_slide_normal = last_slide_normal;
    uint64_t total = parallel_for((int64_t)n_spheres, threads, response_item, &b);
    if (tests) *tests = total;
}

/* ---------- broadphase predicate P (new; see header) ---------- */
void xpo_triangle_aabb(const xpo_triangle *t, xpo_vec3 *lo, xpo_vec3 *hi) {
    const float R = XPO_BROADPHASE_INFLATE;
    float lx = fminf(fminf(t->v0.x, t->v1.x), t->v2.x), hx = fmaxf(fmaxf(t->v0.x, t->v1.x), t->v2.x);
    float ly = fminf(fminf(t->v0.y, t->v1.y), t->v2.y), hy = fmaxf(fmaxf(t->v0.y, t->v1.y), t->v2.y);
    float lz = fminf(fminf(t->v0.z, t->v1.z), t->v2.z), hz = fmaxf(fmaxf(t->v0.z, t->v1.z), t->v2.z);
    *lo = v3(lx - R, ly - R, lz - R);
    *hi = v3(hx + R, hy + R, hz + R);
}

static inline int slab_axis(float c, float m, float lo, float hi, float *tmin, float *tmax) {
    if (m == 0.0f) return !(c < lo || c > hi);
    float inv = 1.0f / m;
    float a = (lo - c) * inv;
    float b = (hi - c) * inv;
    float nr = (inv > 0.0f) ? a : b;
    float fr = (inv > 0.0f) ? b : a;
    if (nr > *tmin) *tmin = nr;
    if (fr < *tmax) *tmax = fr;
    return 1;
}

static int slab_accept(const xpo_response *resp, xpo_vec3 lo, xpo_vec3 hi, float horizon) {
    /* an infinite horizon is FLT_MAX, so that an entry time of +inf (overflow, or the branch-free
     * GPU form of an m_k == 0 axis that starts outside the slab) always rejects */
    float tmin = 0.0f, tmax = (horizon > FLT_MAX) ? FLT_MAX : horizon;
    const xpo_vec3 c = resp->sphere.c, m = resp->movement;
    if (!slab_axis(c.x, m.x, lo.x, hi.x, &tmin, &tmax)) return 0;
    if (!slab_axis(c.y, m.y, lo.y, hi.y, &tmin, &tmax)) return 0;
    if (!slab_axis(c.z, m.z, lo.z, hi.z, &tmin, &tmax)) return 0;
    return tmin <= tmax;
}

int xpo_broadphase_accept(const xpo_response *resp, const xpo_triangle *tri, float horizon) {
    xpo_vec3 lo, hi;
    xpo_triangle_aabb(tri, &lo, &hi);
    return slab_accept(resp, lo, hi, horizon);
}

void xpo_broadphase_pairs(const xpo_response *in, uint64_t n_spheres, const xpo_triangle *tris,
                          uint64_t n_tris, float horizon, uint64_t cap, uint32_t *sphere_idx,
                          uint32_t *tri_idx, uint64_t *count) {
    xpo_vec3 *lo = (xpo_vec3 *)malloc((n_tris ? n_tris : 1) * sizeof(xpo_vec3));
    xpo_vec3 *hi = (xpo_vec3 *)malloc((n_tris ? n_tris : 1) * sizeof(xpo_vec3));
    for (uint64_t j = 0; j < n_tris; ++j) xpo_triangle_aabb(&tris[j], &lo[j], &hi[j]);
    uint64_t k = 0;
    for (uint64_t i = 0; i < n_spheres; ++i)
        for (uint64_t j = 0; j < n_tris; ++j)
            if (slab_accept(&in[i], lo[j], hi[j], horizon)) {
                if (k < cap) { sphere_idx[k] = (uint32_t)i; tri_idx[k] = (uint32_t)j; }
                ++k;
            }
    *count = k;
    free(lo); free(hi);
}

/* ---------- float64 grazing shadow ---------- */
typedef struct { double x, y, z; } d3;
static inline d3 D3(xpo_vec3 a) { d3 r = {a.x, a.y, a.z}; return r; }
static inline d3 dsub(d3 a, d3 b) { d3 r = {a.x - b.x, a.y - b.y, a.z - b.z}; return r; }
static inline d3 dadd(d3 a, d3 b) { d3 r = {a.x + b.x, a.y + b.y, a.z + b.z}; return r; }
static inline d3 dscale(d3 a, double s) { d3 r = {a.x * s, a.y * s, a.z * s}; return r; }
static inline double ddot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline d3 dcross(d3 a, d3 b) {
    d3 r = {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
    return r;
}
static inline double margin(double a, double thr) {
    double s = fmax(1.0, fmax(fabs(a), fabs(thr)));
    return fabs(a - thr) / s;
}
#define TRACK(a, thr) do { double _m = margin((a), (thr)); if (_m < mg) mg = _m; } while (0)

static int droots(double a, double b, double c, double *r0, double *r1, double *mgp) {
    double mg = *mgp;
    double det = b * b - 4.0 * a * c;
    /* det's natural scale is b*b + |4ac|: normalise so the margin is relative */
    double sc = fmax(b * b, fabs(4.0 * a * c));
    if (sc > 0.0) TRACK(det / sc, 0.0);
    *mgp = mg;
    if (det < 0.0) return 0;
    double s = sqrt(det);
    *r0 = (-b - s) / (2.0 * a);
    *r1 = (-b + s) / (2.0 * a);
    return 1;
}

double xpo_detect_margin_f64(const xpo_response *resp, const xpo_triangle *tri, int *hit_out,
                             double *t_out) {
    double mg = INFINITY;
    int hit = 0;
    double t_hit = 0.0;
    d3 c = D3(resp->sphere.c), m = D3(resp->movement);
    d3 v0 = D3(tri->v0), v1 = D3(tri->v1), v2 = D3(tri->v2);
    d3 n = dcross(dsub(v1, v0), dsub(v2, v0));
    double nl = sqrt(ddot(n, n));
    n = dscale(n, 1.0 / nl);
    double ml = sqrt(ddot(m, m));
    d3 mh = dscale(m, 1.0 / ml);
    double facing = ddot(n, mh);
    TRACK(facing, 0.0);                                        /* collision_detect.rs:166 */
    if (facing > 0.0) goto done;
    {
        double pndm = ddot(n, m);
        double sd = ddot(n, c) - ddot(n, v0);
        TRACK(fabs(sd), 1.0);                                  /* :175, :184 */
        if (pndm == 0.0 && fabs(sd) >= 1.0) goto done;
        double msl = ml * ml;
        if (pndm != 0.0 && fabs(sd) >= 1.0) {
            double t0 = (1.0 - sd) / pndm, t1 = (-1.0 - sd) / pndm;
            if (!(t0 < t1)) { double tmp = t0; t0 = t1; t1 = tmp; }
            TRACK(t0, 1.0); TRACK(t1, 0.0);                    /* :33 */
            if (!(t0 > 1.0 || t1 < 0.0)) {
                if (!(t0 > 0.0)) t0 = 0.0;
                d3 p = dadd(c, dscale(m, t0));
                d3 a0 = dsub(v2, v0), a1 = dsub(v1, v0), a2 = dsub(p, v0);
                double d00 = ddot(a0, a0), d01 = ddot(a0, a1), d02 = ddot(a0, a2);
                double d11 = ddot(a1, a1), d12 = ddot(a1, a2);
                double inv = 1.0 / (d00 * d11 - d01 * d01);
                double u = (d11 * d02 - d01 * d12) * inv, v = (d00 * d12 - d01 * d02) * inv;
                TRACK(u, 0.0); TRACK(v, 0.0); TRACK(u + v, 1.0);   /* triangle.rs:31 */
                if (u >= 0.0 && v >= 0.0 && u + v <= 1.0) { hit = 1; t_hit = t0; goto done; }
            }
        }
        double ts[9]; int nts = 0;
        d3 vs[3] = {v0, v1, v2};
        for (int k = 0; k < 3; ++k) {
            d3 d = dsub(c, vs[k]);
            double r0, r1;
            if (droots(msl, 2.0 * ddot(m, d), ddot(d, d) - 1.0, &r0, &r1, &mg)) {
                ts[nts++] = r0; ts[nts++] = r1;
            }
        }
        for (int k = 0; k < 3; ++k) {
            d3 p = vs[k], q = vs[(k + 1) % 3];
            d3 e = dsub(q, p), btv = dsub(p, c);
            double els = ddot(e, e), edm = ddot(e, m), edb = ddot(e, btv);
            double a = els * -msl + edm * edm;
            double b = els * (2.0 * ddot(m, btv)) - 2.0 * edm * edb;
            double cc = els * (1.0 - ddot(btv, btv)) + edb * edb;
            double r0, r1;
            if (droots(a, b, cc, &r0, &r1, &mg)) {
                double t = (r1 < r0) ? r1 : r0;
                double f = (edm * t - edb) / els;
                TRACK(f, 0.0); TRACK(f, 1.0);                  /* :107 */
                if (f >= 0.0 && f <= 1.0) ts[nts++] = t;
            }
        }
        if (nts > 0) {
            double t = ts[0];
            for (int k = 0; k < nts; ++k) if (ts[k] < t) t = ts[k];
            TRACK(t, 0.0);                                     /* :202 */
            if (t > 0.0) { hit = 1; t_hit = t; }
        }
    }
done:
    if (hit_out) *hit_out = hit;
    if (t_out) *t_out = t_hit;
    return mg;
}

static void grazing_item(int64_t s, void *vctx, uint64_t *acc) {
    batch_ctx *b = (batch_ctx *)vctx;
    double mg = INFINITY;
    (void)acc;
    for (uint64_t j = 0; j < b->n_tris; ++j) {
        if (!xpo_broadphase_accept(&b->in[s], &b->tris[j], b->horizon)) continue;
        double x = xpo_detect_margin_f64(&b->in[s], &b->tris[j], NULL, NULL);
        if (x < mg) mg = x;
    }
    b->min_margin[s] = mg;
}

void xpo_grazing_batch(const xpo_response *in, uint64_t n_spheres, const xpo_triangle *tris,
                       uint64_t n_tris, float horizon, int threads, double *min_margin) {
    batch_ctx b;
    memset(&b, 0, sizeof b);
    b.in = in; b.tris = tris; b.n_tris = n_tris; b.horizon = horizon; b.min_margin = min_margin;
    parallel_for((int64_t)n_spheres, threads, grazing_item, &b);
}
