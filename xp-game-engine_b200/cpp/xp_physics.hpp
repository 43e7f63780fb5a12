// xp_physics.hpp -- header-only C++ host mirror of the xp_physics crate's public API
// (/root/reference/xp_physics/src/lib.rs:9-15,17,29) on top of the C ABI.
//
// The reference is compiled code (Rust) and this image has no Rust toolchain, so the host layer
// a maintainer would write in Rust is provided in C++ with the same names, argument meaning and
// error behaviour: where the reference panics these throw (std::logic_error), batch calls return
// per-sphere status instead.  All arithmetic happens on the GPU in libxp_physics_cuda.so.
#pragma once

#include <cmath>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/xp_physics_cuda.h"

namespace xp_physics {

using Vec3 = xp_vec3;
struct Sphere { Vec3 c; float r; };                       // sphere.rs:4-7
struct Triangle { Vec3 v0, v1, v2; };                     // triangle.rs:4-8
struct Response { Sphere sphere; Vec3 movement; };        // response.rs:5-8
struct Collision { float time_to, distance_to; Vec3 position, intersection; };   // collision.rs:5-10
static_assert(sizeof(Sphere) == sizeof(xp_sphere) && sizeof(Triangle) == sizeof(xp_triangle) &&
              sizeof(Response) == sizeof(xp_response) && sizeof(Collision) == sizeof(xp_collision),
              "layouts must match the C ABI");

inline void check(int rc, const char *what) {
    if (rc < 0) throw std::runtime_error(std::string(what) + ": " + xp_strerror(rc) + " " + xp_last_cuda_error());
}

// The triangle soup uploaded once + the device LBVH: what `&[Triangle]` becomes on a GPU.
class Scene {
  public:
    explicit Scene(int device = 0) { check(xp_scene_create(device, &s_), "xp_scene_create"); }
    Scene(int device, const std::vector<Triangle> &tris) : Scene(device) { set_triangles(tris); build(); }
    ~Scene() { xp_scene_destroy(s_); }
    Scene(const Scene &) = delete;
    Scene &operator=(const Scene &) = delete;

    void set_triangles(const std::vector<Triangle> &t) {
        check(xp_scene_set_triangles(s_, reinterpret_cast<const xp_triangle *>(t.data()), t.size()), "set_triangles");
    }
    // lib.rs:17-25: every three Vec3 are one triangle; a ragged slice panics in the reference
    void set_positions(const std::vector<Vec3> &soup) {
        int rc = xp_scene_set_positions(s_, soup.data(), soup.size());
        if (rc == XP_EINVAL) throw std::out_of_range("index out of bounds: chunks(3) (lib.rs:19-24)");
        check(rc, "set_positions");
    }
    void build() { check(xp_scene_build(s_), "xp_scene_build"); }

    struct Closest { std::vector<Collision> collision; std::vector<uint8_t> hit; std::vector<uint32_t> tri_index, status; };
    Closest detect_batch(const std::vector<Response> &in, uint32_t horizon = XP_HORIZON_INF) {
        Closest o;
        o.collision.resize(in.size()); o.hit.resize(in.size()); o.tri_index.resize(in.size()); o.status.resize(in.size());
        check(xp_detect_batch(s_, reinterpret_cast<const xp_response *>(in.data()), in.size(), horizon,
                              reinterpret_cast<xp_collision *>(o.collision.data()), o.hit.data(), o.tri_index.data(),
                              o.status.data()), "xp_detect_batch");
        return o;
    }
    // stream of batches: submit() returns a ticket, wait() blocks until that batch is in `out`; up to
    // XP_ASYNC_SLOTS in flight.  `in` and `out` must outlive wait() and should be page-locked.
    int detect_batch_submit(const Response *in, size_t n, Closest &out, uint32_t horizon = XP_HORIZON_INF) {
        out.collision.resize(n); out.hit.resize(n); out.tri_index.resize(n); out.status.resize(n);
        int ticket = -1;
        check(xp_detect_batch_submit(s_, reinterpret_cast<const xp_response *>(in), n, horizon,
                                     reinterpret_cast<xp_collision *>(out.collision.data()), out.hit.data(),
                                     out.tri_index.data(), out.status.data(), &ticket), "xp_detect_batch_submit");
        return ticket;
    }
    void detect_batch_wait(int ticket) { check(xp_detect_batch_wait(s_, ticket), "xp_detect_batch_wait"); }
    // the same sweep with compact results: 8 B (time_to, tri_index) per sphere cross the host link instead of 41 B;
    // collision_from_hit() below rebuilds the reference's Collision where a caller wants it
    std::vector<xp_hit> detect_batch_compact(const std::vector<Response> &in, uint32_t horizon = XP_HORIZON_INF) {
        std::vector<xp_hit> out(in.size());
        check(xp_detect_batch_compact(s_, reinterpret_cast<const xp_response *>(in.data()), in.size(), horizon, out.data(), nullptr),
              "xp_detect_batch_compact");
        return out;
    }
    int detect_batch_compact_submit(const Response *in, size_t n, xp_hit *out, uint32_t horizon = XP_HORIZON_INF) {
        int ticket = -1;
        check(xp_detect_batch_compact_submit(s_, reinterpret_cast<const xp_response *>(in), n, horizon, out, nullptr, &ticket),
              "xp_detect_batch_compact_submit");
        return ticket;
    }

    struct Resolved { std::vector<Response> response; std::vector<uint32_t> iterations, status; std::vector<Vec3> slide_normal; };
    Resolved collision_response_batch(const std::vector<Response> &in, uint32_t max_iter = 0,
                                      uint32_t horizon = XP_HORIZON_INF) {
        Resolved o;
        o.response.resize(in.size()); o.iterations.resize(in.size()); o.status.resize(in.size()); o.slide_normal.resize(in.size());
        check(xp_collision_response_batch(s_, reinterpret_cast<const xp_response *>(in.data()), in.size(), max_iter, horizon,
                                          reinterpret_cast<xp_response *>(o.response.data()), o.iterations.data(),
                                          o.status.data(), o.slide_normal.data()), "xp_collision_response_batch");
        return o;
    }
    xp_scene *handle() { return s_; }

  private:
    xp_scene *s_ = nullptr;
};

// The reference's Collision from the compact result (collision_detect.rs:188-194 / :203-209): every field is a function
// of (c, m, t).  Same operations in the same order as the device (and the reference): compile the including
// translation unit without floating-point contraction (-ffp-contract=off) and the fields are bit-identical to what
// xp_detect_batch returns.  nullopt = no collision.
inline std::optional<Collision> collision_from_hit(const Response &r, const xp_hit &h) {
    if (h.tri_index == 0xFFFFFFFFu) return std::nullopt;
    const Vec3 &c = r.sphere.c, &m = r.movement;
    const float t = h.time_to;
    const float ml = std::sqrt((m.x * m.x + m.y * m.y) + m.z * m.z);          // nalgebra's 3-vector dot order
    Collision k;
    k.time_to = t;
    k.distance_to = ml * t;
    const float px = m.x * t, py = m.y * t, pz = m.z * t;
    k.position = Vec3{c.x + px, c.y + py, c.z + pz};
    const float nx = m.x / ml, ny = m.y / ml, nz = m.z / ml;
    const float ix = nx * r.sphere.r, iy = ny * r.sphere.r, iz = nz * r.sphere.r;
    k.intersection = Vec3{k.position.x + ix, k.position.y + iy, k.position.z + iz};
    return k;
}

inline void raise_status(uint32_t st) {
    if (st & XP_ST_RADIUS_NE_1) throw std::logic_error("assertion failed: sphere.r == 1.0 (collision_detect.rs:161)");
    if (st & XP_ST_ASSERT_NEG_DISTANCE)
        throw std::logic_error("assertion failed: original_destination_to_plane_distance >= 0.0 (collision_response.rs:22)");
}

// collision_detect.rs:155
inline std::optional<Collision> sphere_triangle_detect_collision(const Response &response, const Triangle &triangle, int device = 0) {
    Collision c;
    int rc = xp_sphere_triangle_detect_collision(device, reinterpret_cast<const xp_response *>(&response),
                                                 reinterpret_cast<const xp_triangle *>(&triangle),
                                                 reinterpret_cast<xp_collision *>(&c));
    if (rc == XP_EINVAL) raise_status(XP_ST_RADIUS_NE_1);
    check(rc, "xp_sphere_triangle_detect_collision");
    return rc == 1 ? std::optional<Collision>(c) : std::nullopt;
}
// collision_response.rs:6
inline Response sphere_triangle_calculate_response(const Response &response, const Collision &collision, int device = 0) {
    Response out;
    int rc = xp_sphere_triangle_calculate_response(device, reinterpret_cast<const xp_response *>(&response),
                                                   reinterpret_cast<const xp_collision *>(&collision),
                                                   reinterpret_cast<xp_response *>(&out));
    check(rc, "xp_sphere_triangle_calculate_response");
    if (rc == 1) raise_status(XP_ST_ASSERT_NEG_DISTANCE);
    return out;
}
// lib.rs:29
inline Response collision_response(const Response &response, const std::vector<Triangle> &triangles, int device = 0) {
    Scene scene(device, triangles);
    auto r = scene.collision_response_batch({response});
    raise_status(r.status[0]);
    return r.response[0];
}
// lib.rs:17
inline Response collision_response_non_trianulated(const Response &response, const std::vector<Vec3> &triangles, int device = 0) {
    Scene scene(device);
    scene.set_positions(triangles);
    scene.build();
    auto r = scene.collision_response_batch({response});
    raise_status(r.status[0]);
    return r.response[0];
}

}  // namespace xp_physics
