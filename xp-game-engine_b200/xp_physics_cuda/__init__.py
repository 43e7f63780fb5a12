"""xp_physics_cuda -- B200-native (sm_100a) swept-sphere vs triangle collision, a drop-in for the
xp_physics crate of bjrnmrtns/xp-game-engine behind the same collide / sweep API.

    from xp_physics_cuda import Scene, Response, Sphere, Triangle, collision_response

The compute lives in csrc/libxp_physics_cuda.so (hand-written CUDA, C ABI in include/xp_physics_cuda.h);
this package is the thin host-side mirror of the reference's Rust API.  There is no CPU fallback.
"""
from . import _lib
from ._lib import XpError, build
from .api import (COLLISION_DTYPE, CONTACT_DTYPE, CONTACT16_DTYPE, HIT_DTYPE, RESPONSE_DTYPE, Collision, collision_from_hit, Response, Scene, Sphere, Triangle,
                  as_triangles, collision_response, collision_response_non_trianulated, debug_check_arith, device_count,
                  make_responses, probe_fp32_peak, sphere_triangle_calculate_response,
                  sphere_triangle_detect_collision)

from . import clipmap, controller, distributed, fauerby, ingest, scenes   # noqa: E402,F401  (host-side glue: caller mirror, sharding, asset loaders, generators)

__all__ = ["Scene", "Sphere", "Triangle", "Response", "Collision", "collision_response",
           "collision_response_non_trianulated", "sphere_triangle_detect_collision",
           "sphere_triangle_calculate_response", "make_responses", "as_triangles", "device_count",
           "probe_fp32_peak", "debug_check_arith", "build", "XpError", "RESPONSE_DTYPE", "COLLISION_DTYPE", "CONTACT_DTYPE", "CONTACT16_DTYPE", "HIT_DTYPE", "collision_from_hit"]
