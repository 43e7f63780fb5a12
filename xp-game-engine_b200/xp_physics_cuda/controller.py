"""Row f1 of SURVEY.md section 8: the caller.  A mirror of the game's per-frame simulation step
(/root/reference/game/src/simulation.rs:5-52) with its physics backend replaced by this library.

The reference step is: movement = frame_time * max_velocity * {forward, right} rotated into the player's
frame (transformation.rs:7-18) -> physics.move_ball(movement); physics.step() (nphysics, simulation.rs:45-48)
-> pose.position = physics.get_position_ball().  Here `move_ball + step` becomes what INTEGRATION.md
section 3 describes: one collision_response sweep of the player's sphere against the terrain scene, in
unit-sphere space (positions and movement divided by the player radius, collision_detect.rs:161), plus an
optional gravity sweep (Fauerby's second pass).  Host-side glue only: the quaternion rotation is numpy,
all collision arithmetic runs on the GPU through Scene.collision_response_batch.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from .api import Scene, make_responses

GLOBAL_DIR = np.array([1.0, 1.0, -1.0], np.float32)        # transformation.rs:3-5


@dataclass
class Pose:                                   # game/src/scene/entities.rs
    position: np.ndarray
    orientation: np.ndarray = field(default_factory=lambda: np.array([1, 0, 0, 0], np.float32))   # w, x, y, z


@dataclass
class Movement:                               # window_input::input_state::Movement
    forward: float = 0.0
    right: float = 0.0


@dataclass
class InputState:
    movement: Optional[Movement] = None


@dataclass
class Player:                                 # scene::Entity::Player { pose, max_velocity }
    pose: Pose
    max_velocity: float = 3.0                 # config.ron:3


def quat_rotate(q: np.ndarray, v: np.ndarray) -> np.ndarray:
    """quat_to_mat4(q) * v for unit quaternions (w, x, y, z), float32."""
    q = np.asarray(q, np.float32).reshape(-1, 4)
    v = np.asarray(v, np.float32).reshape(-1, 3)
    w, u = q[:, :1], q[:, 1:]
    t = np.float32(2.0) * np.cross(u, v)
    return (v + w * t + np.cross(u, t)).astype(np.float32)


def move_along_local_axis(orientation, forward, right, up=0.0) -> np.ndarray:
    """transformation.rs:7-18."""
    local = np.stack([GLOBAL_DIR[0] * np.asarray(right, np.float32), GLOBAL_DIR[1] * np.asarray(up, np.float32)
                      * np.ones_like(np.asarray(right, np.float32)), GLOBAL_DIR[2] * np.asarray(forward, np.float32)], axis=-1)
    return quat_rotate(orientation, local)


class FrameInputHandler:
    """simulation.rs:5-13."""

    def handle(self, frame: int, command: InputState, player: Player, frame_time: float) -> None:
        raise NotImplementedError


class Client(FrameInputHandler):
    """simulation.rs:15-52 with xp_physics_cuda behind it.  `sweep` is injectable so tests can run the
    identical loop against the CPU oracle."""

    def __init__(self, scene: Optional[Scene], radius: float = 0.5, gravity: Sequence[float] = (0.0, 0.0, 0.0),
                 max_iter: int = 0, sweep=None):
        self.scene, self.radius, self.max_iter = scene, np.float32(radius), max_iter
        self.gravity = np.asarray(gravity, np.float32)
        self.last_frame: Optional[int] = None
        self._sweep = sweep or (lambda resp: scene.collision_response_batch(resp, max_iter=max_iter)[0])

    def _resolve(self, positions: np.ndarray, movement: np.ndarray) -> np.ndarray:
        r = self.radius
        resp = make_responses(positions / r, movement / r)             # into unit-sphere space
        out = self._sweep(resp)
        return (out["c"] * r).astype(np.float32)

    def handle(self, frame: int, input_state: InputState, player: Player, frame_time: float) -> None:
        assert self.last_frame is None or self.last_frame < frame      # simulation.rs:37
        self.last_frame = frame
        pos = np.asarray(player.pose.position, np.float32).reshape(1, 3)
        if input_state.movement is not None:                           # simulation.rs:40-46
            forward = np.float32(frame_time) * np.float32(player.max_velocity) * np.float32(input_state.movement.forward)
            right = np.float32(frame_time) * np.float32(player.max_velocity) * np.float32(input_state.movement.right)
            move = move_along_local_axis(player.pose.orientation, forward, right, 0.0)
            pos = self._resolve(pos, move)
        if np.any(self.gravity != 0):                                  # physics.step(): gravity as a second sweep
            pos = self._resolve(pos, (self.gravity * np.float32(frame_time)).reshape(1, 3))
        player.pose.position = pos[0]                                  # simulation.rs:49-51


class Crowd:
    """The same step for N players at once (what the batch API is for): positions [n,3], orientations [n,4]."""

    def __init__(self, scene: Optional[Scene], positions, orientations=None, max_velocity: float = 3.0,
                 radius: float = 0.5, gravity: Sequence[float] = (0.0, 0.0, 0.0), max_iter: int = 0, sweep=None):
        self.client = Client(scene, radius, gravity, max_iter, sweep)
        self.positions = np.ascontiguousarray(positions, np.float32).reshape(-1, 3)
        n = len(self.positions)
        self.orientations = (np.tile(np.array([1, 0, 0, 0], np.float32), (n, 1)) if orientations is None
                             else np.ascontiguousarray(orientations, np.float32).reshape(n, 4))
        self.max_velocity = np.float32(max_velocity)

    def step(self, forward, right, frame_time: float) -> np.ndarray:
        f = np.float32(frame_time) * self.max_velocity * np.asarray(forward, np.float32)
        r = np.float32(frame_time) * self.max_velocity * np.asarray(right, np.float32)
        move = move_along_local_axis(self.orientations, f, r, 0.0)
        self.positions = self.client._resolve(self.positions, move)
        if np.any(self.client.gravity != 0):
            g = np.tile(self.client.gravity * np.float32(frame_time), (len(self.positions), 1))
            self.positions = self.client._resolve(self.positions, g)
        return self.positions
