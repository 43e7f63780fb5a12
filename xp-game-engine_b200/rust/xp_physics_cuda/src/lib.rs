//! UNVERIFIED: written without a Rust toolchain (none in the build image); never compiled.
//!
//! Drop-in for the `xp_physics` crate (xp_physics/src/lib.rs:9-15,17,29): same type and function
//! names, same argument meaning, panics where the reference panics.  The work happens on the GPU
//! in libxp_physics_cuda.so; this file is only the FFI declarations and safe wrappers.
use nalgebra_glm::Vec3;
use std::os::raw::c_int;

// ---- #[repr(C)] mirrors (xp_math has no vector types; nalgebra's Vec3 is not repr(C)) ----------
#[repr(C)] #[derive(Clone, Copy, Debug, PartialEq)] pub struct Vec3f { pub x: f32, pub y: f32, pub z: f32 }
#[repr(C)] #[derive(Clone, Copy, Debug, PartialEq)] pub struct Sphere { pub c: Vec3f, pub r: f32 }
#[repr(C)] #[derive(Clone, Copy, Debug, PartialEq)] pub struct Triangle { pub v0: Vec3f, pub v1: Vec3f, pub v2: Vec3f }
#[repr(C)] #[derive(Clone, Copy, Debug)] pub struct Response { pub sphere: Sphere, pub movement: Vec3f }
#[repr(C)] #[derive(Clone, Copy, Debug)] pub struct Collision {
    pub time_to: f32, pub distance_to: f32, pub position: Vec3f, pub intersection: Vec3f,
}
/// xp_hit: time of impact + triangle (u32::MAX = no collision)
#[repr(C)] #[derive(Clone, Copy, Debug)] pub struct Hit { pub time_to: f32, pub tri_index: u32 }
impl Collision {
    /// collision_detect.rs:188-194: every field is a function of (c, m, t)
    pub fn from_hit(r: &Response, h: &Hit) -> Option<Collision> {
        if h.tri_index == u32::MAX { return None; }
        let (c, m, t) = (r.sphere.c, r.movement, h.time_to);
        let ml = ((m.x * m.x + m.y * m.y) + m.z * m.z).sqrt();
        let position = Vec3f { x: c.x + m.x * t, y: c.y + m.y * t, z: c.z + m.z * t };
        let intersection = Vec3f { x: position.x + m.x / ml * r.sphere.r, y: position.y + m.y / ml * r.sphere.r, z: position.z + m.z / ml * r.sphere.r };
        Some(Collision { time_to: t, distance_to: ml * t, position, intersection })
    }
}
impl From<Vec3> for Vec3f { fn from(v: Vec3) -> Self { Vec3f { x: v.x, y: v.y, z: v.z } } }
impl From<Vec3f> for Vec3 { fn from(v: Vec3f) -> Self { nalgebra_glm::vec3(v.x, v.y, v.z) } }
impl Sphere { pub fn new(c: Vec3, r: f32) -> Self { Sphere { c: c.into(), r } } }
impl Triangle { pub fn new(v0: Vec3, v1: Vec3, v2: Vec3) -> Self { Triangle { v0: v0.into(), v1: v1.into(), v2: v2.into() } } }

pub const XP_ST_RADIUS_NE_1: u32 = 1;
pub const XP_ST_ASSERT_NEG_DISTANCE: u32 = 2;
pub const XP_ST_ITER_CAP: u32 = 4;

#[repr(C)] pub struct XpScene { _private: [u8; 0] }

extern "C" {
    fn xp_scene_create(device: c_int, out: *mut *mut XpScene) -> c_int;
    fn xp_scene_destroy(s: *mut XpScene);
    fn xp_scene_set_triangles(s: *mut XpScene, aos: *const Triangle, n: u64) -> c_int;
    fn xp_scene_set_positions(s: *mut XpScene, soup: *const Vec3f, n_vec3: u64) -> c_int;
    fn xp_scene_build(s: *mut XpScene) -> c_int;
    fn xp_detect_batch(s: *mut XpScene, input: *const Response, n: u64, horizon_mode: u32, out: *mut Collision,
                       hit: *mut u8, tri_index: *mut u32, status: *mut u32) -> c_int;
    fn xp_detect_batch_submit(s: *mut XpScene, input: *const Response, n: u64, horizon_mode: u32, out: *mut Collision,
                              hit: *mut u8, tri_index: *mut u32, status: *mut u32, ticket: *mut c_int) -> c_int;
    fn xp_detect_batch_wait(s: *mut XpScene, ticket: c_int) -> c_int;
    // compact results (8 B per sphere over the host link); Collision::from_hit rebuilds the full record
    fn xp_detect_batch_compact(s: *mut XpScene, input: *const Response, n: u64, horizon_mode: u32, out: *mut Hit,
                               status: *mut u32) -> c_int;
    fn xp_scene_set_heightmap(s: *mut XpScene, heights: *const f32, nx: u32, nz: u32, x0: f32, z0: f32, spacing: f32) -> c_int;
    fn xp_collision_response_batch(s: *mut XpScene, input: *const Response, n: u64, max_iter: u32, horizon_mode: u32,
                                   out: *mut Response, iterations: *mut u32, status: *mut u32,
                                   last_slide_normal: *mut Vec3f) -> c_int;
    fn xp_sphere_triangle_detect_collision(device: c_int, response: *const Response, triangle: *const Triangle,
                                           out: *mut Collision) -> c_int;
    fn xp_sphere_triangle_calculate_response(device: c_int, response: *const Response, collision: *const Collision,
                                             out: *mut Response) -> c_int;
}

/// The triangle soup uploaded once plus the device LBVH (what `&[Triangle]` becomes on a GPU).
pub struct Scene { raw: *mut XpScene }
impl Scene {
    pub fn new(device: i32, triangles: &[Triangle]) -> Scene {
        let mut raw = std::ptr::null_mut();
        unsafe {
            assert_eq!(xp_scene_create(device, &mut raw), 0, "xp_scene_create");
            assert_eq!(xp_scene_set_triangles(raw, triangles.as_ptr(), triangles.len() as u64), 0);
            assert_eq!(xp_scene_build(raw), 0, "xp_scene_build");
        }
        Scene { raw }
    }
    /// Stream of batches: queue one batch (copies and sweep run asynchronously, up to 2 in flight) and get a
    /// ticket.  Safety: `input` and the four output slices must stay alive and untouched until `wait(ticket)`
    /// returns, and should be page-locked for the copies to be truly asynchronous.
    pub unsafe fn submit(&self, input: &[Response], out: &mut [Collision], hit: &mut [u8], tri_index: &mut [u32],
                         status: &mut [u32]) -> i32 {
        let n = input.len();
        assert!(out.len() >= n && hit.len() >= n && tri_index.len() >= n && status.len() >= n);
        let mut ticket: c_int = -1;
        let rc = xp_detect_batch_submit(self.raw, input.as_ptr(), n as u64, 0, out.as_mut_ptr(), hit.as_mut_ptr(),
                                        tri_index.as_mut_ptr(), status.as_mut_ptr(), &mut ticket);
        assert_eq!(rc, 0, "xp_detect_batch_submit");
        ticket
    }
    pub fn wait(&self, ticket: i32) {
        let rc = unsafe { xp_detect_batch_wait(self.raw, ticket) };
        assert_eq!(rc, 0, "xp_detect_batch_wait");
    }
    pub fn from_positions(device: i32, soup: &[Vec3f]) -> Scene {
        let mut raw = std::ptr::null_mut();
        unsafe {
            assert_eq!(xp_scene_create(device, &mut raw), 0, "xp_scene_create");
            // the reference indexes vs[2] of a short chunk and panics (lib.rs:19-24)
            assert_eq!(xp_scene_set_positions(raw, soup.as_ptr(), soup.len() as u64), 0, "index out of bounds");
            assert_eq!(xp_scene_build(raw), 0, "xp_scene_build");
        }
        Scene { raw }
    }
    /// collision_response (lib.rs:29-62) for every sphere; returns (responses, iterations, status).
    pub fn collision_response_batch(&self, input: &[Response], max_iter: u32) -> (Vec<Response>, Vec<u32>, Vec<u32>) {
        let n = input.len();
        let mut out = input.to_vec();
        let (mut it, mut st) = (vec![0u32; n], vec![0u32; n]);
        let rc = unsafe {
            xp_collision_response_batch(self.raw, input.as_ptr(), n as u64, max_iter, 0, out.as_mut_ptr(),
                                        it.as_mut_ptr(), st.as_mut_ptr(), std::ptr::null_mut())
        };
        assert_eq!(rc, 0, "xp_collision_response_batch");
        (out, it, st)
    }
    /// Closest collision per sphere (collision_detect.rs:155 + the min_by of lib.rs:33-42).
    pub fn detect_batch(&self, input: &[Response]) -> Vec<Option<Collision>> {
        let n = input.len();
        let zero = Vec3f { x: 0.0, y: 0.0, z: 0.0 };
        let mut col = vec![Collision { time_to: 0.0, distance_to: 0.0, position: zero, intersection: zero }; n];
        let mut hit = vec![0u8; n];
        let rc = unsafe {
            xp_detect_batch(self.raw, input.as_ptr(), n as u64, 0, col.as_mut_ptr(), hit.as_mut_ptr(),
                            std::ptr::null_mut(), std::ptr::null_mut())
        };
        assert_eq!(rc, 0, "xp_detect_batch");
        col.into_iter().zip(hit).map(|(c, h)| if h != 0 { Some(c) } else { None }).collect()
    }
}
impl Drop for Scene { fn drop(&mut self) { unsafe { xp_scene_destroy(self.raw) } } }

fn panic_on(status: u32) {
    assert!(status & XP_ST_RADIUS_NE_1 == 0, "assertion failed: `(left == right)` sphere.r == 1.0");
    assert!(status & XP_ST_ASSERT_NEG_DISTANCE == 0, "assertion failed: original_destination_to_plane_distance >= 0.0");
}

/// xp_physics::collision_response (lib.rs:29)
pub fn collision_response(response: Response, triangles: &[Triangle]) -> Response {
    let scene = Scene::new(0, triangles);
    let (out, _, st) = scene.collision_response_batch(&[response], 0);
    panic_on(st[0]);
    out[0]
}
/// xp_physics::collision_response_non_trianulated (lib.rs:17)
pub fn collision_response_non_trianulated(response: Response, triangles: &[Vec3]) -> Response {
    let soup: Vec<Vec3f> = triangles.iter().map(|v| (*v).into()).collect();
    let scene = Scene::from_positions(0, &soup);
    let (out, _, st) = scene.collision_response_batch(&[response], 0);
    panic_on(st[0]);
    out[0]
}
/// xp_physics::sphere_triangle_detect_collision (collision_detect.rs:155)
pub fn sphere_triangle_detect_collision(response: &Response, triangle: &Triangle) -> Option<Collision> {
    let zero = Vec3f { x: 0.0, y: 0.0, z: 0.0 };
    let mut c = Collision { time_to: 0.0, distance_to: 0.0, position: zero, intersection: zero };
    let rc = unsafe { xp_sphere_triangle_detect_collision(0, response, triangle, &mut c) };
    assert!(rc != -1, "assertion failed: `(left == right)` sphere.r == 1.0");
    assert!(rc >= 0, "xp_sphere_triangle_detect_collision failed: {}", rc);
    if rc == 1 { Some(c) } else { None }
}
/// xp_physics::sphere_triangle_calculate_response (collision_response.rs:6)
pub fn sphere_triangle_calculate_response(response: &Response, collision: &Collision) -> Response {
    let mut out = *response;
    let rc = unsafe { xp_sphere_triangle_calculate_response(0, response, collision, &mut out) };
    assert!(rc != 1, "assertion failed: original_destination_to_plane_distance >= 0.0");
    assert!(rc == 0, "xp_sphere_triangle_calculate_response failed: {}", rc);
    out
}

#[cfg(test)]
mod tests {
    use super::*;
    use nalgebra_glm::vec3;
    // the reference's own tests (collision_detect.rs:222-260), unchanged
    #[test] fn test_detect_where_collision_inside_triangle() {
        let t = Triangle::new(vec3(-2.0, 2.0, -2.0), vec3(-2.0, 2.0, 2.0), vec3(2.0, 2.0, 0.0));
        let r = Response { sphere: Sphere::new(vec3(0.0, 4.0, 0.0), 1.0), movement: vec3(0.0, -2.0, 0.0).into() };
        assert_eq!(sphere_triangle_detect_collision(&r, &t).unwrap().time_to, 0.5);
    }
    #[test] fn test_detect_where_collision_against_vertex() {
        let t = Triangle::new(vec3(0.0, 0.0, 0.0), vec3(-2.0, -1.0, 0.0), vec3(2.0, -1.0, 0.0));
        let r = Response { sphere: Sphere::new(vec3(0.0, 4.0, 0.0), 1.0), movement: vec3(0.0, -8.0, 0.0).into() };
        assert_eq!(sphere_triangle_detect_collision(&r, &t).unwrap().time_to, 0.375);
    }
    #[test] fn test_detect_where_collision_against_edge() {
        let t = Triangle::new(vec3(0.0, -2.0, 0.0), vec3(-2.0, -1.0, 0.0), vec3(2.0, -1.0, 0.0));
        let r = Response { sphere: Sphere::new(vec3(0.0, 4.0, 0.0), 1.0), movement: vec3(0.0, -8.0, 0.0).into() };
        assert_eq!(sphere_triangle_detect_collision(&r, &t).unwrap().time_to, 0.5);
    }
}
