//! Rust shim over the C ABI of include/bam3d_gpu.h: packs bam3d primitives and glam `Mat4`s into the SoA host arrays
//! the ABI takes and exposes batched `GJK` / `SweepAndPrune3` calls with the crate's own types.
//! SOURCE ONLY: this repository's build image has no cargo / rustc, so this file is not compiled or tested here; the
//! compiled host layer with the same shape is include/bam3d.hpp (C++), exercised by tests/cpp/kat_gpu.cpp.
use bam3d::{Aabb3, CollisionStrategy, Contact, Primitive3, Ray, Sphere};
use glam::{Mat4, Vec3};
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct GjkSettings { pub distance_tolerance: f32, pub continuous_tolerance: f32, pub epa_tolerance: f32, pub max_iterations: u32, pub toi_iteration_cap: u32 }
#[repr(C)] #[derive(Clone, Copy, Default)]
pub struct RawContact { pub normal: [f32; 3], pub penetration_depth: f32, pub contact_point: [f32; 3], pub time_of_impact: f32 }

#[allow(non_camel_case_types)] type ctx = c_void;
extern "C" {
    fn bam3d_ctx_create(device: i32, out: *mut *mut ctx) -> i32;
    fn bam3d_ctx_destroy(c: *mut ctx);
    fn bam3d_last_error(c: *const ctx) -> *const c_char;
    fn bam3d_default_settings(out: *mut GjkSettings);
    fn bam3d_meshes_upload(c: *mut ctx, verts: *const f32, offsets: *const u32, n: u32, out: *mut *mut c_void) -> i32;
    fn bam3d_shapes_upload(c: *mut ctx, kind: *const u8, params: *const f32, xform: *const f32, mesh_id: *const u32, n: u32,
                           meshes: *const c_void, out: *mut *mut c_void) -> i32;
    fn bam3d_shapes_set_transforms(c: *mut ctx, s: *mut c_void, xform: *const f32) -> i32;
    fn bam3d_shapes_free(c: *mut ctx, s: *mut c_void);
    fn bam3d_gjk_intersect(c: *mut ctx, s: *const c_void, pairs: *const u32, n: u64, cfg: *const GjkSettings, status: *mut u8) -> i32;
    fn bam3d_gjk_contact(c: *mut ctx, s: *const c_void, pairs: *const u32, n: u64, cfg: *const GjkSettings, strategy: i32,
                         out: *mut RawContact, status: *mut u8) -> i32;
    fn bam3d_gjk_distance(c: *mut ctx, s: *const c_void, pairs: *const u32, n: u64, cfg: *const GjkSettings, ref_value: *mut f32,
                          distance: *mut f32, status: *mut u8) -> i32;
    fn bam3d_gjk_time_of_impact(c: *mut ctx, s0: *const c_void, s1: *const c_void, pairs: *const u32, n: u64, cfg: *const GjkSettings,
                                out: *mut RawContact, status: *mut u8) -> i32;
    fn bam3d_sap_find_pairs(c: *mut ctx, aabbs: *const f32, n: u64, axis: *mut i32, perm: *mut u32, pairs: *mut u32, cap: u64, count: *mut u64) -> i32;
    fn bam3d_sap_find_pairs_shard(c: *mut ctx, aabbs: *const f32, n: u64, axis: *mut i32, perm: *mut u32, pairs: *mut u32, cap: u64,
                                  count: *mut u64, shard: u32, n_shards: u32) -> i32;
    fn bam3d_ray_cast(c: *mut ctx, s: *const c_void, rays: *const f32, n_rays: u64, casts: *const u32, n: u64, query: i32,
                      points: *mut f32, status: *mut u8) -> i32;
    fn bam3d_shapes_world_spheres(c: *mut ctx, s: *const c_void, out: *mut f32) -> i32;
    fn bam3d_sap_find_pairs_spheres(c: *mut ctx, spheres: *const f32, n: u64, axis: *mut i32, perm: *mut u32, pairs: *mut u32, cap: u64,
                                    count: *mut u64) -> i32;
    fn bam3d_host_alloc(c: *mut ctx, bytes: u64, out: *mut *mut c_void) -> i32;
    fn bam3d_host_free(c: *mut ctx, p: *mut c_void);
    fn bam3d_frames_create(c: *mut ctx, kind: *const u8, params: *const f32, xform: *const f32, mesh_id: *const u32, n_shapes: u32,
                           meshes: *const c_void, max_pairs: u64, depth: u32, with_end: i32, out: *mut *mut c_void) -> i32;
    fn bam3d_frames_submit(c: *mut ctx, f: *mut c_void, query: i32, xform: *const f32, xform_end: *const f32, pairs: *const u32, n: u64,
                           cfg: *const GjkSettings, strategy: i32, out: *mut RawContact, status: *mut u8, frame: *mut u64) -> i32;
    fn bam3d_frames_wait(c: *mut ctx, f: *mut c_void, frame: u64) -> i32;
    fn bam3d_frames_free(c: *mut ctx, f: *mut c_void);
}

/// Batched `ContinuousTransformed<Ray> for Primitive3` (ray.rs:68-78): `casts[i] = (ray index, shape index)`.
/// `None` where the reference returns `None`; a ConvexPolyhedron (status 6) is not computed on the device and should be
/// routed to the crate's own `polyhedron.rs` implementation.
pub fn ray_intersection_batch(ctx: &Context, shapes: &ShapeSet, rays: &[Ray], casts: &[(u32, u32)]) -> Result<Vec<Option<Vec3>>, i32> {
    let flat: Vec<f32> = rays.iter().flat_map(|r| [r.origin.x(), r.origin.y(), r.origin.z(), r.direction.x(), r.direction.y(), r.direction.z()]).collect();
    let idx: Vec<u32> = casts.iter().flat_map(|&(r, s)| [r, s]).collect();
    let mut pts = vec![0f32; 3 * casts.len()];
    let mut st = vec![0u8; casts.len()];
    let rc = unsafe { bam3d_ray_cast(ctx.0, shapes.h, flat.as_ptr(), rays.len() as u64, idx.as_ptr(), casts.len() as u64, 1, pts.as_mut_ptr(), st.as_mut_ptr()) };
    if rc != 0 { return Err(rc); }
    Ok(st.iter().enumerate().map(|(i, &s)| if s == 0 { Some(Vec3::new(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2])) } else { None }).collect())
}

pub struct Context(*mut ctx);
impl Context {
    pub fn new(device: i32) -> Result<Self, i32> { let mut h = std::ptr::null_mut(); let rc = unsafe { bam3d_ctx_create(device, &mut h) }; if rc == 0 { Ok(Context(h)) } else { Err(rc) } }
}
impl Drop for Context { fn drop(&mut self) { unsafe { bam3d_ctx_destroy(self.0) } } }

/// Primitive3 -> (kind, params) of the ABI (include/bam3d_gpu.h BAM3D_KIND_*).
fn pack(p: &Primitive3) -> (u8, [f32; 4]) {
    match p {
        Primitive3::Particle(_) => (0, [0.; 4]),
        Primitive3::Quad(q) => (1, [q.half_dim().x(), q.half_dim().y(), 0., 0.]),
        Primitive3::Sphere(s) => (2, [s.radius, 0., 0., 0.]),
        Primitive3::Cuboid(c) => (3, [c.half_dim().x(), c.half_dim().y(), c.half_dim().z(), 0.]),
        Primitive3::Cube(c) => (3, [c.half_dim(), c.half_dim(), c.half_dim(), 0.]),
        Primitive3::Cylinder(c) => (4, [c.height() / 2., c.radius(), 0., 0.]),
        Primitive3::Capsule(c) => (5, [c.height() / 2., c.radius(), 0., 0.]),
        Primitive3::ConvexPolyhedron(_) => (6, [0.; 4]), // vertices go through bam3d_meshes_upload
    }
}

pub struct ShapeSet<'a> { ctx: &'a Context, h: *mut c_void }
impl<'a> ShapeSet<'a> {
    pub fn new(ctx: &'a Context, shapes: &[(Primitive3, Mat4)]) -> Result<Self, i32> {
        let n = shapes.len();
        let (mut kind, mut params, mut xf) = (Vec::with_capacity(n), Vec::with_capacity(4 * n), Vec::with_capacity(16 * n));
        for (p, t) in shapes { let (k, q) = pack(p); kind.push(k); params.extend_from_slice(&q); xf.extend_from_slice(&t.to_cols_array()); }
        let mut h = std::ptr::null_mut();
        let rc = unsafe { bam3d_shapes_upload(ctx.0, kind.as_ptr(), params.as_ptr(), xf.as_ptr(), std::ptr::null(), n as u32, std::ptr::null(), &mut h) };
        if rc == 0 { Ok(ShapeSet { ctx, h }) } else { Err(rc) }
    }
}
impl<'a> Drop for ShapeSet<'a> { fn drop(&mut self) { unsafe { bam3d_shapes_free(self.ctx.0, self.h) } } }

/// Batched counterpart of `GJK::intersection` (gjk/mod.rs:317-341): one `Option<Contact>` per pair.
pub fn intersection_batch(ctx: &Context, strategy: &CollisionStrategy, shapes: &ShapeSet, pairs: &[(u32, u32)]) -> Result<Vec<Option<Contact>>, i32> {
    let mut cfg = GjkSettings { distance_tolerance: 0., continuous_tolerance: 0., epa_tolerance: 0., max_iterations: 0, toi_iteration_cap: 0 };
    unsafe { bam3d_default_settings(&mut cfg) };
    let (mut raw, mut st) = (vec![RawContact::default(); pairs.len()], vec![0u8; pairs.len()]);
    let s = match strategy { CollisionStrategy::FullResolution => 0, CollisionStrategy::CollisionOnly => 1 };
    let rc = unsafe { bam3d_gjk_contact(ctx.0, shapes.h, pairs.as_ptr() as *const u32, pairs.len() as u64, &cfg, s, raw.as_mut_ptr(), st.as_mut_ptr()) };
    if rc != 0 { return Err(rc); }
    Ok(raw.iter().zip(&st).map(|(c, &s)| if s == 0 {
        let mut k = Contact::new_with_point(strategy.clone(), Vec3::new(c.normal[0], c.normal[1], c.normal[2]), c.penetration_depth,
                                            Vec3::new(c.contact_point[0], c.contact_point[1], c.contact_point[2]));
        k.time_of_impact = c.time_of_impact; Some(k) } else { None }).collect())
}

/// `SweepAndPrune3<volume::Sphere>::find_collider_pairs` on the device (`bam3d::Sphere` = the bounding sphere of volume/sphere.rs):
/// re-sorts `shapes` like the reference and returns indices into it; touching spheres are reported (`Sphere::intersects` is `<=`).
pub fn sap_find_collider_pairs_spheres(ctx: &Context, sweep_axis: &mut usize, shapes: &mut [Sphere]) -> Result<Vec<(usize, usize)>, i32> {
    let n = shapes.len();
    let flat: Vec<f32> = shapes.iter().flat_map(|s| [s.center.x(), s.center.y(), s.center.z(), s.radius]).collect();
    let (mut perm, mut cap, mut count) = (vec![0u32; n], 8 * n as u64 + 1024, 0u64);
    loop {
        let mut pairs = vec![0u32; 2 * cap as usize];
        let mut axis = *sweep_axis as i32;
        let rc = unsafe { bam3d_sap_find_pairs_spheres(ctx.0, flat.as_ptr(), n as u64, &mut axis, perm.as_mut_ptr(), pairs.as_mut_ptr(), cap, &mut count) };
        if rc == -4 && count > cap { cap = count; continue; } // BAM3D_ERR_CAPACITY: call again with the reported size
        if rc != 0 { return Err(rc); }
        *sweep_axis = axis as usize;
        let sorted: Vec<Sphere> = perm.iter().map(|&k| shapes[k as usize]).collect();
        shapes.copy_from_slice(&sorted);
        return Ok((0..count as usize).map(|k| (pairs[2 * k] as usize, pairs[2 * k + 1] as usize)).collect());
    }
}

/// `ComputeBound<volume::Sphere>` + `Bound::transform_volume` for every shape of a set.
pub fn world_spheres(ctx: &Context, shapes: &ShapeSet, n: usize) -> Result<Vec<Sphere>, i32> {
    let mut out = vec![0f32; 4 * n];
    let rc = unsafe { bam3d_shapes_world_spheres(ctx.0, shapes.h, out.as_mut_ptr()) };
    if rc != 0 { return Err(rc); }
    Ok(out.chunks(4).map(|c| Sphere { center: Vec3::new(c[0], c[1], c[2]), radius: c[3] }).collect())
}

/// Page-locked host buffer (`bam3d_host_alloc`): what a frame's transforms / pairs / results should live in so that the
/// PCIe copies are asynchronous and run at full rate.
pub struct Pinned<'a, T: Copy> { ctx: &'a Context, p: *mut T, len: usize }
impl<'a, T: Copy> Pinned<'a, T> {
    pub fn new(ctx: &'a Context, len: usize) -> Result<Self, i32> {
        let mut p = std::ptr::null_mut();
        let rc = unsafe { bam3d_host_alloc(ctx.0, (len.max(1) * std::mem::size_of::<T>()) as u64, &mut p) };
        if rc == 0 { Ok(Pinned { ctx, p: p as *mut T, len }) } else { Err(rc) }
    }
    pub fn as_slice(&self) -> &[T] { unsafe { std::slice::from_raw_parts(self.p, self.len) } }
    pub fn as_mut_slice(&mut self) -> &mut [T] { unsafe { std::slice::from_raw_parts_mut(self.p, self.len) } }
}
impl<'a, T: Copy> Drop for Pinned<'a, T> { fn drop(&mut self) { unsafe { bam3d_host_free(self.ctx.0, self.p as *mut c_void) } } }

/// The per-frame sequence of a simulation loop — new `Mat4` for every shape, `GJK::intersection` over the frame's pairs,
/// contacts back in host memory — with `depth` frames in flight (`bam3d_frames_*`): frame k+1's upload and frame k-1's
/// download run beside frame k's kernels. `submit` borrows its buffers until `wait(frame)` returns, which is why they are
/// `Pinned` values owned by the caller for the queue's lifetime.
pub struct FrameQueue<'a> { ctx: &'a Context, h: *mut c_void, cfg: GjkSettings }
impl<'a> FrameQueue<'a> {
    pub fn new(ctx: &'a Context, shapes: &[(Primitive3, Mat4)], max_pairs: usize, depth: u32) -> Result<Self, i32> {
        let n = shapes.len();
        let (mut kind, mut params, mut xf) = (Vec::with_capacity(n), Vec::with_capacity(4 * n), Vec::with_capacity(16 * n));
        for (p, t) in shapes { let (k, q) = pack(p); kind.push(k); params.extend_from_slice(&q); xf.extend_from_slice(&t.to_cols_array()); }
        let mut cfg = GjkSettings { distance_tolerance: 0., continuous_tolerance: 0., epa_tolerance: 0., max_iterations: 0, toi_iteration_cap: 0 };
        unsafe { bam3d_default_settings(&mut cfg) };
        let mut h = std::ptr::null_mut();
        let rc = unsafe { bam3d_frames_create(ctx.0, kind.as_ptr(), params.as_ptr(), xf.as_ptr(), std::ptr::null(), n as u32, std::ptr::null(),
                                              max_pairs as u64, depth, 0, &mut h) };
        if rc == 0 { Ok(FrameQueue { ctx, h, cfg }) } else { Err(rc) }
    }
    /// Enqueue `GJK::intersection(strategy, ..)` for every pair under this frame's transforms; returns the frame number.
    pub fn submit_intersection(&mut self, strategy: &CollisionStrategy, xform: &Pinned<[f32; 16]>, pairs: &Pinned<(u32, u32)>,
                               out: &mut Pinned<RawContact>, status: &mut Pinned<u8>) -> Result<u64, i32> {
        let s = match strategy { CollisionStrategy::FullResolution => 0, CollisionStrategy::CollisionOnly => 1 };
        let mut frame = 0u64;
        let rc = unsafe { bam3d_frames_submit(self.ctx.0, self.h, 1, xform.p as *const f32, std::ptr::null(), pairs.p as *const u32, pairs.len as u64,
                                              &self.cfg, s, out.p, status.p, &mut frame) };
        if rc == 0 { Ok(frame) } else { Err(rc) }
    }
    /// Block until that frame's contacts and status are in the buffers given to its `submit_intersection`.
    pub fn wait(&mut self, frame: u64) -> Result<(), i32> { let rc = unsafe { bam3d_frames_wait(self.ctx.0, self.h, frame) }; if rc == 0 { Ok(()) } else { Err(rc) } }
}
impl<'a> Drop for FrameQueue<'a> { fn drop(&mut self) { unsafe { bam3d_frames_free(self.ctx.0, self.h) } } }

/// `SweepAndPrune3::find_collider_pairs` on the device: re-sorts `shapes` like the reference and returns indices into it.
pub fn sap_find_collider_pairs(ctx: &Context, sweep_axis: &mut usize, shapes: &mut [Aabb3]) -> Result<Vec<(usize, usize)>, i32> {
    let n = shapes.len();
    let flat: Vec<f32> = shapes.iter().flat_map(|b| vec![b.min.x(), b.min.y(), b.min.z(), b.max.x(), b.max.y(), b.max.z()]).collect();
    let (mut perm, mut cap, mut count, mut axis) = (vec![0u32; n], (8 * n + 1024) as u64, 0u64, *sweep_axis as i32);
    let mut pairs;
    loop {
        pairs = vec![0u32; 2 * cap as usize];
        let rc = unsafe { bam3d_sap_find_pairs(ctx.0, flat.as_ptr(), n as u64, &mut axis, perm.as_mut_ptr(), pairs.as_mut_ptr(), cap, &mut count) };
        if rc == -4 && count > cap { cap = count; continue; }
        if rc != 0 { return Err(rc); }
        break;
    }
    let sorted: Vec<Aabb3> = perm.iter().map(|&k| shapes[k as usize]).collect();
    shapes.copy_from_slice(&sorted);
    *sweep_axis = axis as usize;
    Ok((0..count as usize).map(|k| (pairs[2 * k] as usize, pairs[2 * k + 1] as usize)).collect())
}
