// controller.hpp -- row f1: the caller.  C++ mirror of the game's per-frame
// simulation step (/root/reference/game/src/simulation.rs:5-52) with its physics backend replaced by
// this library: movement = frame_time * max_velocity * {forward, right} rotated into the player's frame
// (game/src/transformation.rs:3-18), then -- where the reference calls physics.move_ball + physics.step
// (nphysics) -- one collision_response sweep of the player's sphere against the terrain scene in
// unit-sphere space (positions and movement divided by the player radius, collision_detect.rs:161),
// plus an optional gravity sweep (Fauerby's second pass).  The sweep is a template parameter so the same
// loop can be driven by the library (Scene) or, in tests, by the CPU oracle.
#pragma once

#include <cassert>
#include <optional>
#include <vector>

#include "xp_physics.hpp"

namespace xp_physics {

struct Quat { float w, x, y, z; };                                  // unit quaternion
struct Pose { Vec3 position; Quat orientation{1, 0, 0, 0}; };       // game/src/scene/entities.rs
struct Movement { float forward = 0, right = 0; };                  // window_input::input_state::Movement
struct InputState { std::optional<Movement> movement; };
struct Player { Pose pose; float max_velocity = 3.0f; };            // scene::Entity::Player, config.ron:3

inline Vec3 cross(Vec3 a, Vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// quat_to_mat4(q) * v for a unit quaternion: v + w*t + u x t with t = 2 (u x v)
inline Vec3 quat_rotate(Quat q, Vec3 v) {
    const Vec3 u{q.x, q.y, q.z};
    Vec3 t = cross(u, v);
    t = {2.0f * t.x, 2.0f * t.y, 2.0f * t.z};
    const Vec3 ut = cross(u, t);
    return {v.x + q.w * t.x + ut.x, v.y + q.w * t.y + ut.y, v.z + q.w * t.z + ut.z};
}

// transformation.rs:3-18: GLOBAL_DIR = (1, 1, -1)
inline Vec3 move_along_local_axis(Quat orientation, float forward, float right, float up) {
    return quat_rotate(orientation, Vec3{1.0f * right, 1.0f * up, -1.0f * forward});
}

// Sweep: std::vector<Response> operator()(const std::vector<Response>&) -- collision_response for each.
template <class Sweep>
class Client {                                                      // simulation.rs:15-52
  public:
    Client(Sweep sweep, float radius = 0.5f, Vec3 gravity = {0, 0, 0}) : sweep_(sweep), radius_(radius), gravity_(gravity) {}

    void handle(uint64_t frame, const InputState &input, Player &player, float frame_time) {
        assert(!last_frame_ || *last_frame_ < frame);               // simulation.rs:37
        last_frame_ = frame;
        std::vector<Vec3> pos{player.pose.position};
        if (input.movement) {                                       // simulation.rs:40-46
            const float forward = frame_time * player.max_velocity * input.movement->forward;
            const float right = frame_time * player.max_velocity * input.movement->right;
            pos = resolve(pos, {move_along_local_axis(player.pose.orientation, forward, right, 0.0f)});
        }
        if (gravity_.x != 0 || gravity_.y != 0 || gravity_.z != 0)  // physics.step(): gravity as a second sweep
            pos = resolve(pos, {Vec3{gravity_.x * frame_time, gravity_.y * frame_time, gravity_.z * frame_time}});
        player.pose.position = pos[0];                              // simulation.rs:49-51
    }

    // the same step for many spheres at once (what the batch API is for)
    std::vector<Vec3> resolve(const std::vector<Vec3> &positions, const std::vector<Vec3> &movement) {
        std::vector<Response> in(positions.size());
        for (size_t i = 0; i < in.size(); ++i) {
            const Vec3 p = positions[i], m = movement.size() == 1 ? movement[0] : movement[i];
            in[i] = {{{p.x / radius_, p.y / radius_, p.z / radius_}, 1.0f}, {m.x / radius_, m.y / radius_, m.z / radius_}};
        }
        const std::vector<Response> out = sweep_(in);
        std::vector<Vec3> res(out.size());
        for (size_t i = 0; i < out.size(); ++i)
            res[i] = {out[i].sphere.c.x * radius_, out[i].sphere.c.y * radius_, out[i].sphere.c.z * radius_};
        return res;
    }

  private:
    Sweep sweep_;
    float radius_;
    Vec3 gravity_;
    std::optional<uint64_t> last_frame_;
};

// the library as the sweep
struct SceneSweep {
    Scene *scene;
    uint32_t max_iter = 0;
    std::vector<Response> operator()(const std::vector<Response> &in) const {
        return scene->collision_response_batch(in, max_iter).response;
    }
};

}  // namespace xp_physics
