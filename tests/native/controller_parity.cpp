// TEST-ONLY native program: the game's simulation step (cpp/controller.hpp) driven once by the CUDA
// library and once by the CPU oracle (linked here because this is a test), frame by frame, for a crowd of
// players walking over a bumpy terrain; every position must agree bit for bit.  Exit code 0 = parity.
#include <cmath>
#include <cstdio>
#include <cstring>

#include "../../xp-game-engine_b200/cpp/controller.hpp"
extern "C" {
#include "../../oracle/xp_oracle.h"
}

using namespace xp_physics;

struct OracleSweep {
    const std::vector<Triangle> *tris;
    uint32_t max_iter;
    std::vector<Response> operator()(const std::vector<Response> &in) const {
        std::vector<Response> out(in.size());
        std::vector<uint32_t> it(in.size()), st(in.size());
        std::vector<Vec3> nrm(in.size());
        uint64_t tests = 0;
        xpo_collision_response_batch(reinterpret_cast<const xpo_response *>(in.data()), in.size(),
                                     reinterpret_cast<const xpo_triangle *>(tris->data()), tris->size(), max_iter,
                                     /*threads=*/0, reinterpret_cast<xpo_response *>(out.data()), it.data(), st.data(),
                                     reinterpret_cast<xpo_vec3 *>(nrm.data()), &tests);
        return out;
    }
};

int main() {
    const int n = 24;
    std::vector<Triangle> terrain;
    auto h = [](int x, int z) { return 0.25f * std::sin(0.8f * x) + 0.25f * std::cos(0.6f * z); };
    for (int x = 0; x < n; ++x)
        for (int z = 0; z < n; ++z) {
            Vec3 p00{(float)x, h(x, z), (float)z}, p01{(float)x, h(x, z + 1), (float)z + 1};
            Vec3 p11{(float)x + 1, h(x + 1, z + 1), (float)z + 1}, p10{(float)x + 1, h(x + 1, z), (float)z};
            terrain.push_back({p00, p01, p11});
            terrain.push_back({p00, p11, p10});
        }
    Scene scene(0, terrain);
    const uint32_t max_iter = 4;
    Client<SceneSweep> gpu(SceneSweep{&scene, max_iter}, 0.5f, Vec3{0, -2.0f, 0});
    Client<OracleSweep> cpu(OracleSweep{&terrain, max_iter}, 0.5f, Vec3{0, -2.0f, 0});

    // one player through the per-frame handler ...
    Player a{{{3.0f, 1.2f, 3.0f}, {0.9238795f, 0, 0.3826834f, 0}}, 3.0f}, b = a;
    // ... and a crowd through the batch form
    std::vector<Vec3> crowd_gpu, crowd_cpu;
    unsigned s = 99;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return (s >> 8) * (1.0f / 16777216.0f); };
    for (int i = 0; i < 512; ++i) crowd_gpu.push_back({2 + rnd() * (n - 4), 0.9f + rnd(), 2 + rnd() * (n - 4)});
    crowd_cpu = crowd_gpu;
    std::vector<Vec3> walk(crowd_gpu.size());
    for (auto &w : walk) w = {0.1f * (rnd() - 0.5f), 0, 0.1f * (rnd() - 0.5f)};

    for (uint64_t frame = 1; frame <= 40; ++frame) {
        InputState in;
        if (frame % 7 != 0) in.movement = Movement{frame % 3 ? 1.0f : -0.5f, frame % 5 ? 0.3f : -1.0f};
        gpu.handle(frame, in, a, 1.0f / 60.0f);
        cpu.handle(frame, in, b, 1.0f / 60.0f);
        if (std::memcmp(&a.pose.position, &b.pose.position, sizeof(Vec3)) != 0) {
            std::printf("FAILED: player differs at frame %llu\n", (unsigned long long)frame);
            return 1;
        }
        crowd_gpu = gpu.resolve(crowd_gpu, walk);
        crowd_cpu = cpu.resolve(crowd_cpu, walk);
        crowd_gpu = gpu.resolve(crowd_gpu, {Vec3{0, -2.0f / 60.0f, 0}});
        crowd_cpu = cpu.resolve(crowd_cpu, {Vec3{0, -2.0f / 60.0f, 0}});
        if (std::memcmp(crowd_gpu.data(), crowd_cpu.data(), crowd_gpu.size() * sizeof(Vec3)) != 0) {
            std::printf("FAILED: crowd differs at frame %llu\n", (unsigned long long)frame);
            return 1;
        }
    }
    std::printf("player at (%g, %g, %g) after 40 frames; 512-player crowd identical on GPU and oracle: ok\n",
                a.pose.position.x, a.pose.position.y, a.pose.position.z);
    return 0;
}
