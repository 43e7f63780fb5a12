// Minimal use of the C++ host mirror (needs a CUDA device at run time).
//   g++ -std=c++17 example.cpp -L../csrc -lxp_physics_cuda -Wl,-rpath,'$ORIGIN/../csrc' -o example
#include <cstdio>
#include "xp_physics.hpp"

int main() {
    using namespace xp_physics;
    Triangle tri{{-2, 2, -2}, {-2, 2, 2}, {2, 2, 0}};          // collision_detect.rs:223-227
    Response r{{{0, 4, 0}, 1.0f}, {0, -2, 0}};
    auto c = sphere_triangle_detect_collision(r, tri);
    std::printf("time_to = %g (reference test expects 0.5)\n", c ? c->time_to : -1.0f);
    return (c && c->time_to == 0.5f) ? 0 : 1;
}
