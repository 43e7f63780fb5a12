// TEST-ONLY host build of the product's device header (xp_narrowphase.cuh compiles as plain C++
// when __CUDACC__ is absent).  It lets the CPU test-suite check that the GPU's RE-ORGANISED
// formulas (cached normal / plane constant, shared vertex terms, running minimum, branch-free
// slab test) are bit-identical to the oracle's line-by-line restatement, before any GPU time is
// spent.  Nothing in the product links this; the product has no CPU path.
// Build: g++ -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see tests/host_emul/__init__.py)
#include <cstring>
#include "../../xp-game-engine_b200/csrc/xp_narrowphase.cuh"
#include "../../xp-game-engine_b200/csrc/xp_gridwalk.cuh"

using namespace xp;

static long g_bad = 0;       // tests of modes 2 / 3 that fell back to the literal form

extern "C" {

long emul_bad_count(int reset) { const long v = g_bad; if (reset) g_bad = 0; return v; }

// resp: 7 floats (c, r, m); tri: 9 floats; out: 8 floats (collision). returns 1/0/-1.  fast = mode (0 / 1)
int emul_detect(const float *resp, const float *tri, float *out, int fast) {
    if (!(resp[3] == 1.0f)) return -1;
    V3 c = {resp[0], resp[1], resp[2]}, m = {resp[4], resp[5], resp[6]};
    V3 v0 = {tri[0], tri[1], tri[2]}, v1 = {tri[3], tri[4], tri[5]}, v2 = {tri[6], tri[7], tri[8]};
    TriDerived d = xp_tri_derive(v0, v1, v2);
    Ray ray = xp_ray_setup<ExactOps>(c, m);
    float t;
    // mode 1: the two squared edge lengths come precomputed, as k_pack stores them in the triangle record
    const float els01[2] = {xp_edge_len2<ExactOps>(v0, v1), xp_edge_len2<ExactOps>(v1, v2)};
    if (fast == 2 || fast == 3) {
        // modes 2 / 3: the production forms xp_detect_x / xp_detect_f exactly as k_narrow_pairs calls them (per-ray
        // operand from xp_ray_r2a, +inf marker for an irregular triangle, literal form when `bad`); the host build
        // substitutes IEEE division / square root for the device sequences (see the header)
        const float r2a = xp_ray_r2a<ExactOps>(ray);
        const float e0 = xp_tri_regular(v0, v1, v2) ? els01[0] : XP_INF;
        bool bad = false;
        bool hit = fast == 2 ? xp_detect_x(ray, r2a, v0, v1, v2, d.n, d.pc, e0, els01[1], t, bad)
                             : xp_detect_f(ray, r2a, v0, v1, v2, d.n, d.pc, e0, els01[1], t, bad);
        g_bad += bad ? 1 : 0;
        if (bad) hit = xp_detect<ExactOps>(ray, v0, v1, v2, d.n, d.pc, t);
        if (!hit) return 0;
    } else
    if (!xp_detect<ExactOps>(ray, v0, v1, v2, d.n, d.pc, t, nullptr, fast == 1 ? els01 : nullptr)) return 0;
    Contact k = xp_make_collision<ExactOps>(ray, t);
    out[0] = k.time_to; out[1] = k.distance_to;
    out[2] = k.position.x; out[3] = k.position.y; out[4] = k.position.z;
    out[5] = k.intersection.x; out[6] = k.intersection.y; out[7] = k.intersection.z;
    return 1;
}

// returns 0 ok / 1 assert
int emul_respond(const float *resp, const float *col, float *out, float *normal) {
    V3 nc, nm, sn;
    bool ok = xp_respond<ExactOps>(V3{resp[0], resp[1], resp[2]}, V3{resp[4], resp[5], resp[6]},
                                   V3{col[2], col[3], col[4]}, V3{col[5], col[6], col[7]}, nc, nm, sn);
    normal[0] = sn.x; normal[1] = sn.y; normal[2] = sn.z;
    if (!ok) return 1;
    out[0] = nc.x; out[1] = nc.y; out[2] = nc.z; out[3] = 1.0f; out[4] = nm.x; out[5] = nm.y; out[6] = nm.z;
    return 0;
}

// the branch-free slab form used by the traversal kernel (inv = 1/m even when m == 0)
static bool slab_branchfree(V3 c, V3 inv, float horizon, V3 lo, V3 hi) {
    float tmin = 0.0f, tmax = horizon;
    const float cs[3] = {c.x, c.y, c.z}, is[3] = {inv.x, inv.y, inv.z};
    const float ls[3] = {lo.x, lo.y, lo.z}, hs[3] = {hi.x, hi.y, hi.z};
    for (int k = 0; k < 3; ++k) {
        const float a = (ls[k] - cs[k]) * is[k], b = (hs[k] - cs[k]) * is[k];
        const float nr = (is[k] > 0.0f) ? a : b, fr = (is[k] > 0.0f) ? b : a;
        if (nr > tmin) tmin = nr;
        if (fr < tmax) tmax = fr;
    }
    return tmin <= tmax;
}

// returns bit0 = header form (explicit m==0 branch), bit1 = branch-free kernel form
int emul_broadphase(const float *resp, const float *tri, float horizon) {
    V3 c = {resp[0], resp[1], resp[2]}, m = {resp[4], resp[5], resp[6]};
    V3 v0 = {tri[0], tri[1], tri[2]}, v1 = {tri[3], tri[4], tri[5]}, v2 = {tri[6], tri[7], tri[8]};
    V3 lo, hi;
    xp_tri_aabb(v0, v1, v2, lo, hi);
    SlabRay s = xp_slab_setup(c, m, horizon);
    float tn;
    int r = xp_slab_hit(s, lo, hi, tn) ? 1 : 0;
    if (slab_branchfree(c, s.inv, s.horizon, lo, hi)) r |= 2;
    return r;
}

// The height-grid walk of k_broadphase_grid as plain nested loops over the same functions (xp_gridwalk.cuh):
// every (column, cell) the ray enumerates, the two triangles' predicate P.  Writes the upload indices of the
// candidates; returns their number (cells tested in *cells).
long emul_grid_walk(const float *heights, unsigned nx, unsigned nz, float x0, float z0, float sp, float eps,
                    const float *resp, float horizon, unsigned *out_tris, long cap, long *cells) {
    GridDesc g;
    g.h = heights; g.nx = nx; g.nz = nz; g.x0 = x0; g.z0 = z0; g.sp = sp; g.inv_sp = 1.0f / sp; g.eps = eps; g.mm = nullptr;
    float ymin = heights[0], ymax = heights[0];
    for (unsigned long i = 0; i < (unsigned long)(nx + 1) * (nz + 1); ++i) { ymin = fminf(ymin, heights[i]); ymax = fmaxf(ymax, heights[i]); }
    const V3 c = {resp[0], resp[1], resp[2]}, m = {resp[4], resp[5], resp[6]};
    const SlabRay sr = xp_slab_setup(c, m, horizon);          // 1/m and the FLT_MAX-clamped horizon, as k_ray_setup stores them
    GridRay r;
    long n = 0;
    *cells = 0;
    if (!grid_ray_setup(g, ymin, ymax, c, m, sr.inv, sr.horizon, r)) return 0;
    const unsigned row = nz + 1;
    for (int ix = r.ix; ix <= r.ix_hi; ++ix) {
        GridCol col;
        if (!grid_col_setup(g, r, ix, col)) continue;
        for (int iz = col.iz; iz <= col.iz_hi; ++iz) {
            if ((iz - col.iz) % XP_GW_WINDOW == 0) {          // a new window of cells: k_height_windows' (min, max), then the cull
                float hmin = HUGE_VALF, hmax = -HUGE_VALF;
                for (unsigned j = (unsigned)iz; j <= ((unsigned)iz + XP_GW_WINDOW < nz ? (unsigned)iz + XP_GW_WINDOW : nz); ++j) {
                    const float a = heights[(unsigned long)ix * row + j], b = heights[(unsigned long)(ix + 1) * row + j];
                    hmin = fminf(hmin, fminf(a, b)); hmax = fmaxf(hmax, fmaxf(a, b));
                }
                if (!grid_window_test(r, col, hmin, hmax)) { iz += XP_GW_WINDOW - 1; continue; }
            }
            const float *p = heights + (unsigned long)ix * row + iz;
            const unsigned hit = grid_cell_test(r, col, grid_coord(z0, sp, iz), grid_coord(z0, sp, iz + 1), p[0], p[1], p[row], p[row + 1]);
            ++*cells;
            for (unsigned k = 0; k < 2; ++k)
                if ((hit >> k) & 1u) { if (n < cap) out_tris[n] = 2u * ((unsigned)ix * nz + (unsigned)iz) + k; ++n; }
        }
    }
    return n;
}

void emul_detect_many(const float *resp, const float *tris, long n_tris, float *out_t, int *out_hit, int mode) {
    for (long j = 0; j < n_tris; ++j) {
        float col[8];
        int r = emul_detect(resp, tris + 9 * j, col, mode);
        out_hit[j] = r;
        out_t[j] = (r == 1) ? col[0] : 0.0f;
    }
}

}
