// Use of the C++ host mirror from a native program (needs a CUDA device at run time): the reference's
// known answers through the 1-sphere calls, then a small scene through the batch, response and
// stream-of-batches calls, checked against each other.  Exit code 0 = everything consistent.
//   g++ -std=c++17 example.cpp -L../csrc -lxp_physics_cuda -Wl,-rpath,'$ORIGIN/../csrc' -o example
#include <cmath>
#include <cstdio>
#include <cstring>
#include "xp_physics.hpp"

using namespace xp_physics;

static int fail(const char *what) { std::printf("FAILED: %s\n", what); return 1; }

int main() {
    // collision_detect.rs:222-232 / :234-246
    Triangle tri{{-2, 2, -2}, {-2, 2, 2}, {2, 2, 0}};
    Response r{{{0, 4, 0}, 1.0f}, {0, -2, 0}};
    auto c = sphere_triangle_detect_collision(r, tri);
    std::printf("face time_to = %g (reference test expects 0.5)\n", c ? c->time_to : -1.0f);
    if (!c || c->time_to != 0.5f) return fail("face KAT");
    auto v = sphere_triangle_detect_collision(Response{{{0, 4, 0}, 1.0f}, {0, -8, 0}}, Triangle{{0, 0, 0}, {-2, -1, 0}, {2, -1, 0}});
    if (!v || v->time_to != 0.375f) return fail("vertex KAT");
    try {
        sphere_triangle_detect_collision(Response{{{0, 4, 0}, 2.0f}, {0, -2, 0}}, tri);
        return fail("r != 1 must throw like the reference panics");
    } catch (const std::logic_error &) {}

    // a 16 x 16 bumpy floor and 4096 spheres dropped onto it
    const int n = 16;
    std::vector<Triangle> floor;
    auto h = [](int x, int z) { return 0.3f * std::sin(0.9f * x) * std::cos(0.7f * z); };
    for (int x = 0; x < n; ++x)
        for (int z = 0; z < n; ++z) {
            Vec3 p00{(float)x, h(x, z), (float)z}, p01{(float)x, h(x, z + 1), (float)z + 1};
            Vec3 p11{(float)x + 1, h(x + 1, z + 1), (float)z + 1}, p10{(float)x + 1, h(x + 1, z), (float)z};
            floor.push_back({p00, p01, p11});
            floor.push_back({p00, p11, p10});
        }
    std::vector<Response> spheres;
    unsigned s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) * (1.0f / 16777216.0f); };
    for (int i = 0; i < 4096; ++i)
        spheres.push_back({{{rnd() * n, 1.5f + 2.0f * rnd(), rnd() * n}, 1.0f}, {0.4f * (rnd() - 0.5f), -0.2f - rnd(), 0.4f * (rnd() - 0.5f)}});

    Scene scene(0, floor);
    auto a = scene.detect_batch(spheres);
    size_t hits = 0;
    for (auto x : a.hit) hits += x;
    std::printf("%zu of %zu spheres hit the floor\n", hits, spheres.size());
    if (hits < spheres.size() / 2) return fail("too few hits");

    // the same batch twice through the asynchronous calls, both in flight
    Scene::Closest b0, b1;
    int t0 = scene.detect_batch_submit(spheres.data(), spheres.size(), b0);
    int t1 = scene.detect_batch_submit(spheres.data(), spheres.size(), b1);
    scene.detect_batch_wait(t0);
    scene.detect_batch_wait(t1);
    for (const Scene::Closest *b : {&b0, &b1})
        for (size_t i = 0; i < spheres.size(); ++i) {
            if (b->hit[i] != a.hit[i] || b->tri_index[i] != a.tri_index[i]) return fail("async hit / triangle differs");
            if (a.hit[i] && std::memcmp(&b->collision[i], &a.collision[i], sizeof(Collision)) != 0) return fail("async collision differs");
        }

    // compact results (8 B per sphere) + collision_from_hit rebuild the full records bit for bit
    auto compact = scene.detect_batch_compact(spheres);
    for (size_t i = 0; i < spheres.size(); ++i) {
        auto c = collision_from_hit(spheres[i], compact[i]);
        if (c.has_value() != (a.hit[i] != 0) || compact[i].tri_index != a.tri_index[i]) return fail("compact hit / triangle differs");
        if (c && std::memcmp(&*c, &a.collision[i], sizeof(Collision)) != 0) return fail("collision_from_hit differs from the device's record");
    }

    // the response loop agrees with the single-sphere reference-shaped call
    auto res = scene.collision_response_batch(spheres, 3);
    for (size_t i = 0; i < 8; ++i) {
        if (res.status[i] & (XP_ST_RADIUS_NE_1 | XP_ST_ASSERT_NEG_DISTANCE)) continue;
        Scene one(0, floor);
        auto single = one.collision_response_batch({spheres[i]}, 3);
        if (std::memcmp(&single.response[0], &res.response[i], sizeof(Response)) != 0) return fail("batch vs single response");
    }
    std::printf("ok\n");
    return 0;
}
