// xp_rays.cuh -- the per-sphere, per-sweep ray record (RayRec, layout in xp_device.cuh), built wherever a ray is born:
// k_ray_setup, the sphere ordering's scatter (xp_build.cu) and the response kernels (xp_query.cu).
#pragma once

#include "xp_device.cuh"
#include "xp_narrowphase.cuh"

namespace xp {

// The 64 B record of one sphere for one sweep (layout in xp_device.cuh).  valid: the r == 1 assert held.
template <class O>
__device__ __forceinline__ RayRec make_ray_rec(V3 c, V3 m, bool valid) {
    uint32_t flags = valid ? RAY_VALID : 0u;
    {   // finite centre and movement: the height-grid walk can enumerate this ray's cells (anything else walks the LBVH)
        const float big = fmaxf(fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fabsf(c.z)), fmaxf(fmaxf(fabsf(m.x), fabsf(m.y)), fabsf(m.z)));
        if (big <= XP_FLT_MAX && c.x == c.x && c.y == c.y && c.z == c.z && m.x == m.x && m.y == m.y && m.z == m.z) flags |= RAY_GRID_OK;
    }
    const Ray ray = xp_ray_setup<O>(c, m);
    RayRec rr;
    rr.q0 = make_float4(c.x, c.y, c.z, ray.ml);
    rr.q1 = make_float4(m.x, m.y, m.z, ray.msl);
    // q2.w: the per-ray operand of the narrowphase's one vertex-root division (+inf = irregular ray, literal path)
    rr.q2 = make_float4(ray.mh.x, ray.mh.y, ray.mh.z, xp_ray_r2a<O>(ray));
    rr.q3 = make_float4(frcp_z(m.x), frcp_z(m.y), frcp_z(m.z), __uint_as_float(flags));
    return rr;
}

}  // namespace xp
