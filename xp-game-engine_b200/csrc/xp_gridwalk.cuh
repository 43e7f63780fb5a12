// xp_gridwalk.cuh -- broadphase for heightmap terrain without a tree: the ray walks the height grid itself.
//
// A heightmap scene (xp_scene_set_heightmap: low_poly_nice_graphics/src/world/heightmap_loader.rs:19-63, the
// clipmap's per-level grids) is an implicit hierarchy: the two triangles of quad (ix, iz) have the SAME box in x
// and z -- the cell [X(ix), X(ix+1)] x [Z(iz), Z(iz+1)] inflated by R -- and differ only in their height range.
// So instead of chasing 64 B nodes of an LBVH the ray enumerates the columns ix its clipped path can touch, and
// in each column the run of cells iz it can touch, from two multiplications per bound; node addresses are
// computed, nothing is pointer-chased, and a candidate costs 8 B of heights instead of ~130 B of node records.
//
// The candidate set stays bit-exact: the DECISION for a triangle is the very slab test of predicate P (same
// operations on the same box values k_pack would derive from the emitted vertices), evaluated axis by axis --
// max / min are associative, so hoisting the x axis per column and the z axis per cell changes no value.  The
// ENUMERATION only has to be a superset; its positions are widened by `eps` (a distance set from the grid's
// coordinate magnitude at upload, ~32 ulp of the largest coordinate) against rounding.  Runs of GW_WINDOW cells
// are skipped wholesale when the ray misses the box of their height range (mm: sliding-window min / max, the
// implicit "mip" level of this hierarchy) -- again predicate P's own slab test on an enclosing box, so nothing
// that P accepts is ever skipped.  Compiles for the host too: tests/host_emul runs the same functions against
// the oracle's P on the CPU.
#pragma once

#include "xp_narrowphase.cuh"

namespace xp {

struct GridDesc {
    const float *h;            // (nx+1) x (nz+1) heights, x-major: h[ix * (nz+1) + iz]
    unsigned nx, nz;           // quads per axis
    float x0, z0, sp, inv_sp;  // vertex (ix, iz) = (x0 + sp*ix, h, z0 + sp*iz), as k_emit_heightmap forms it
    float eps;                 // slack (a distance) on every enumerated position
    const float *mm;           // nx x nz pairs (min, max) of the heights under cells (ix, iz .. iz + GW_WINDOW - 1): h[ix..ix+1][iz..iz+GW_WINDOW]
};
#ifndef XP_GW_WINDOW
#define XP_GW_WINDOW 8
#endif

// coordinate of grid line i exactly as the emitted triangles carry it (k_emit_heightmap)
XP_HD float grid_coord(float org, float sp, int i) { return ExactOps::add(org, ExactOps::mul(sp, (float)i)); }

// one axis of predicate P's slab test, the same operations as xp_query.cu's slab(): NaN operands are ignored
XP_HD void grid_slab_axis(float c, float inv, float lo, float hi, float &tmin, float &tmax) {
    const float a = ExactOps::mul(ExactOps::sub(lo, c), inv), b = ExactOps::mul(ExactOps::sub(hi, c), inv);
    tmin = fmaxf(tmin, (inv > 0.0f) ? a : b);
    tmax = fminf(tmax, (inv > 0.0f) ? b : a);
}

XP_HD int grid_floor_clamped(float u, int lo, int hi) {
    // floor with saturation, branch-free; NaN cannot occur (finite rays only; fmaxf would map it to lo), +-inf saturate
    return (int)fminf(fmaxf(floorf(u), (float)lo), (float)hi);
}

struct GridRay {
    V3 c, m, inv;
    float horizon;
    float t_in, t_out;         // the ray inside the scene's inflated bounding box (enumeration only, never a decision)
    int ix, ix_hi;             // columns still to visit: ix .. ix_hi
};

// false: the ray misses the box of the whole terrain, so it misses every triangle box inside it.
// Rays must be finite (c, m); 1/m may be +-inf.
XP_HD bool grid_ray_setup(const GridDesc &g, float ymin, float ymax, V3 c, V3 m, V3 inv, float horizon, GridRay &r) {
    r.c = c; r.m = m; r.inv = inv; r.horizon = horizon;
    float tmin = 0.0f, tmax = horizon;
    grid_slab_axis(c.x, inv.x, ExactOps::sub(grid_coord(g.x0, g.sp, 0), XP_INFLATE), ExactOps::add(grid_coord(g.x0, g.sp, (int)g.nx), XP_INFLATE), tmin, tmax);
    grid_slab_axis(c.y, inv.y, ExactOps::sub(ymin, XP_INFLATE), ExactOps::add(ymax, XP_INFLATE), tmin, tmax);
    grid_slab_axis(c.z, inv.z, ExactOps::sub(grid_coord(g.z0, g.sp, 0), XP_INFLATE), ExactOps::add(grid_coord(g.z0, g.sp, (int)g.nz), XP_INFLATE), tmin, tmax);
    r.t_in = tmin; r.t_out = tmax;
    r.ix = 0; r.ix_hi = -1;
    if (!(tmin <= tmax)) return false;
    const float xa = c.x + tmin * m.x, xb = c.x + tmax * m.x;
    const float xlo = fminf(xa, xb), xhi = fmaxf(xa, xb);
    // cell ix overlaps [xlo, xhi] iff X(ix+1) + R >= xlo and X(ix) - R <= xhi, i.e. u - 1 <= ix <= v with
    // u = (xlo - R - x0) / sp, v = (xhi + R - x0) / sp; floor(u - eps/sp) <= ceil(u - 1) for any eps > 0
    r.ix = grid_floor_clamped((xlo - (XP_INFLATE + g.eps) - g.x0) * g.inv_sp, -(1 << 28), (int)g.nx);
    r.ix_hi = grid_floor_clamped((xhi + (XP_INFLATE + g.eps) - g.x0) * g.inv_sp, -(1 << 28), (int)g.nx);
    if (r.ix < 0) r.ix = 0;
    if (r.ix_hi > (int)g.nx - 1) r.ix_hi = (int)g.nx - 1;
    return r.ix <= r.ix_hi;
}

struct GridCol {
    float tx_lo, tx_hi;        // P's running (tmin, tmax) after the x axis of this column's boxes
    int iz, iz_hi;             // cells still to visit: iz .. iz_hi
};

// false: no cell of column ix can pass
XP_HD bool grid_col_setup(const GridDesc &g, const GridRay &r, int ix, GridCol &col) {
    float tmin = 0.0f, tmax = r.horizon;
    grid_slab_axis(r.c.x, r.inv.x, ExactOps::sub(grid_coord(g.x0, g.sp, ix), XP_INFLATE),
                   ExactOps::add(grid_coord(g.x0, g.sp, ix + 1), XP_INFLATE), tmin, tmax);
    col.tx_lo = tmin; col.tx_hi = tmax;
    col.iz = 0; col.iz_hi = -1;
    if (!(tmin <= tmax)) return false;             // exact: the x axis alone already rejects every box of the column
    // where along z can the ray be while inside this x slab (and inside the terrain's box)?
    const float t0 = fmaxf(tmin, r.t_in), t1 = fminf(tmax, r.t_out);
    if (!(t0 <= t1)) return false;
    const float za = r.c.z + t0 * r.m.z, zb = r.c.z + t1 * r.m.z;
    const float zlo = fminf(za, zb), zhi = fmaxf(za, zb);
    col.iz = grid_floor_clamped((zlo - (XP_INFLATE + g.eps) - g.z0) * g.inv_sp, -(1 << 28), (int)g.nz);
    col.iz_hi = grid_floor_clamped((zhi + (XP_INFLATE + g.eps) - g.z0) * g.inv_sp, -(1 << 28), (int)g.nz);
    if (col.iz < 0) col.iz = 0;
    if (col.iz_hi > (int)g.nz - 1) col.iz_hi = (int)g.nz - 1;
    return col.iz <= col.iz_hi;
}

// Can any triangle of cells (ix, iz .. iz + GW_WINDOW - 1) pass?  P's x and y axes on the box of the window's height
// range (hmin, hmax from g.mm): every triangle box of the window lies inside it, and the slab test is monotone.
XP_HD bool grid_window_test(const GridRay &r, const GridCol &col, float hmin, float hmax) {
    float tmin = col.tx_lo, tmax = col.tx_hi;
    grid_slab_axis(r.c.y, r.inv.y, ExactOps::sub(hmin, XP_INFLATE), ExactOps::add(hmax, XP_INFLATE), tmin, tmax);
    return tmin <= tmax;
}

// Predicate P for the two triangles of cell (ix, iz): bit 0 = (p00, p01, p11), bit 1 = (p00, p11, p10).
// z_lo / z_hi: grid_coord(z0, sp, iz), grid_coord(z0, sp, iz + 1); (tx_lo, tx_hi): P's running (tmin, tmax) after the x
// axis of the column's boxes (GridCol).  The scalar form is what the cell-parallel walk calls: a lane that tests a
// cell holds only these values of the ray, not the ray.
XP_HD unsigned grid_cell_test_s(float cy, float inv_y, float cz, float inv_z, float tx_lo, float tx_hi, float z_lo, float z_hi,
                                float h00, float h01, float h10, float h11) {
    float tmin = tx_lo, tmax = tx_hi;
    grid_slab_axis(cz, inv_z, ExactOps::sub(z_lo, XP_INFLATE), ExactOps::add(z_hi, XP_INFLATE), tmin, tmax);
    if (!(tmin <= tmax)) return 0u;
    float amin = tmin, amax = tmax, bmin = tmin, bmax = tmax;
    grid_slab_axis(cy, inv_y, ExactOps::sub(x_min3(h00, h01, h11), XP_INFLATE), ExactOps::add(x_max3(h00, h01, h11), XP_INFLATE), amin, amax);
    grid_slab_axis(cy, inv_y, ExactOps::sub(x_min3(h00, h11, h10), XP_INFLATE), ExactOps::add(x_max3(h00, h11, h10), XP_INFLATE), bmin, bmax);
    return (amin <= amax ? 1u : 0u) | (bmin <= bmax ? 2u : 0u);
}
XP_HD unsigned grid_cell_test(const GridRay &r, const GridCol &col, float z_lo, float z_hi, float h00, float h01,
                              float h10, float h11) {
    return grid_cell_test_s(r.c.y, r.inv.y, r.c.z, r.inv.z, col.tx_lo, col.tx_hi, z_lo, z_hi, h00, h01, h10, h11);
}

}  // namespace xp
