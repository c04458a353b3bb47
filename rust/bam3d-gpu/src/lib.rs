//! Rust host layer over the C ABI of include/bam3d_gpu.h (north_star subsystem 1): packs bam3d primitives and glam `Mat4`s
//! into the SoA host arrays the ABI takes, and mirrors the crate's API surface for the accelerated path —
//!   * `GJK` (`gjk.rs`): `new`, `new_with_settings`, `intersect`, `intersection`, `distance`,
//!     `intersection_time_of_impact`, the three `*_complex` forms (gjk/mod.rs:32-524), each also as a `_batch` call;
//!   * `SweepAndPrune3`, `DynamicBoundingVolumeTree` queries, `DbvtBroadPhase`, the broad -> narrow pipeline (`broad.rs`);
//!   * `Discrete` / `Continuous` / `DiscreteTransformed` / `ContinuousTransformed` impls for device-resident shapes
//!     (`traits.rs`; the crate's traits.rs:7-147, ray.rs:58-78);
//!   * the frame queue (`frames.rs`) and the NCCL contact gather (`multi.rs`).
//! `ffi.rs` is generated from the header (tools/gen_rust_ffi.py): every entry point of the ABI is declared.
//!
//! SOURCE ONLY in this repository: the build image has no cargo / rustc, so nothing here is compiled or tested; the
//! compiled host layer with the same shape is include/bam3d.hpp (C++17, tests/cpp/kat_gpu.cpp), and the Python mirror
//! (bam3d_b200/api.py, broad.py) is what the parity tests drive.
pub mod ffi;
pub mod gjk;
pub mod broad;
pub mod frames;
pub mod multi;
pub mod traits;

use bam3d::Primitive3;
use glam::Mat4;
use std::ffi::CStr;
use std::os::raw::c_void;

/// A failed ABI call: the code (`ffi::ERR_*`) and `bam3d_last_error`'s text.
#[derive(Debug, Clone)]
pub struct Error { pub code: i32, pub message: String }
pub type Result<T> = std::result::Result<T, Error>;

/// One per host thread: owns the CUDA stream and the device scratch (`bam3d_ctx`). There is no CPU fallback: `new` fails
/// when no CUDA device is present.
pub struct Context(pub(crate) *mut ffi::Ctx);
impl Context {
    pub fn new(device: i32) -> Result<Self> {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { ffi::bam3d_ctx_create(device, &mut h) };
        if rc == ffi::OK { Ok(Context(h)) } else { Err(Error { code: rc, message: "bam3d_ctx_create: no usable CUDA device".into() }) }
    }
    pub(crate) fn check(&self, rc: i32) -> Result<()> {
        if rc == ffi::OK { return Ok(()); }
        let msg = unsafe { CStr::from_ptr(ffi::bam3d_last_error(self.0)) }.to_string_lossy().into_owned();
        Err(Error { code: rc, message: msg })
    }
    /// Wait for everything enqueued by `_dev` calls; also where an out-of-range pair index of such a call is reported.
    pub fn synchronize(&self) -> Result<()> { self.check(unsafe { ffi::bam3d_ctx_synchronize(self.0) }) }
    pub fn set_option(&self, name: &str, value: i32) -> Result<()> {
        let c = std::ffi::CString::new(name).unwrap();
        self.check(unsafe { ffi::bam3d_ctx_set_option(self.0, c.as_ptr(), value) })
    }
    pub fn launch_count(&self) -> u64 { unsafe { ffi::bam3d_ctx_launch_count(self.0) } }
}
impl Drop for Context { fn drop(&mut self) { unsafe { ffi::bam3d_ctx_destroy(self.0) } } }

/// Primitive3 -> (kind, params) of the ABI (`BAM3D_KIND_*`); the vertices of a ConvexPolyhedron go through `MeshLibrary`.
pub(crate) fn pack(p: &Primitive3) -> (u8, [f32; 4]) {
    match p {
        Primitive3::Particle(_) => (ffi::KIND_PARTICLE, [0.; 4]),
        Primitive3::Quad(q) => (ffi::KIND_QUAD, [q.half_dim().x(), q.half_dim().y(), 0., 0.]),
        Primitive3::Sphere(s) => (ffi::KIND_SPHERE, [s.radius, 0., 0., 0.]),
        Primitive3::Cuboid(c) => (ffi::KIND_CUBOID, [c.half_dim().x(), c.half_dim().y(), c.half_dim().z(), 0.]),
        Primitive3::Cube(c) => (ffi::KIND_CUBOID, [c.half_dim(), c.half_dim(), c.half_dim(), 0.]),
        Primitive3::Cylinder(c) => (ffi::KIND_CYLINDER, [c.height() / 2., c.radius(), 0., 0.]),
        Primitive3::Capsule(c) => (ffi::KIND_CAPSULE, [c.height() / 2., c.radius(), 0., 0.]),
        Primitive3::ConvexPolyhedron(_) => (ffi::KIND_POLYHEDRON, [0.; 4]),
    }
}

/// The SoA host arrays of a slice of `(Primitive3, Mat4)`: kinds, params, column-major transforms, mesh ids, and the
/// vertex / face library of the ConvexPolyhedron values among them (one mesh per polyhedron, in order).
pub(crate) struct Packed { pub kind: Vec<u8>, pub params: Vec<f32>, pub xform: Vec<f32>, pub mesh_id: Vec<u32>, pub verts: Vec<f32>, pub voff: Vec<u32>,
                           pub faces: Vec<u32>, pub foff: Vec<u32> }
pub(crate) fn pack_all(shapes: &[(Primitive3, Mat4)]) -> Packed {
    let n = shapes.len();
    let mut p = Packed { kind: Vec::with_capacity(n), params: Vec::with_capacity(4 * n), xform: Vec::with_capacity(16 * n), mesh_id: vec![0; n],
                         verts: Vec::new(), voff: vec![0], faces: Vec::new(), foff: vec![0] };
    for (i, (prim, t)) in shapes.iter().enumerate() {
        let (k, q) = pack(prim);
        p.kind.push(k); p.params.extend_from_slice(&q); p.xform.extend_from_slice(&t.to_cols_array());
        if let Primitive3::ConvexPolyhedron(poly) = prim {
            p.mesh_id[i] = (p.voff.len() - 1) as u32;
            for v in poly.vertices() { p.verts.extend_from_slice(&[v.x(), v.y(), v.z()]); }
            p.voff.push((p.verts.len() / 3) as u32);
            for f in poly.faces() { p.faces.extend_from_slice(&[f.0 as u32, f.1 as u32, f.2 as u32]); }   // empty in VertexOnly mode
            p.foff.push((p.faces.len() / 3) as u32);
        }
    }
    p
}

/// Vertex (and, for `new_with_faces` polyhedra, half-edge) library resident on the device (`bam3d_meshes`).
pub struct MeshLibrary<'a> { pub(crate) ctx: &'a Context, pub(crate) h: *mut ffi::Meshes }
impl<'a> MeshLibrary<'a> {
    pub(crate) fn from_packed(ctx: &'a Context, p: &Packed) -> Result<Option<Self>> {
        let n = (p.voff.len() - 1) as u32;
        if n == 0 { return Ok(None); }
        let mut h = std::ptr::null_mut();
        let rc = if p.faces.is_empty() { unsafe { ffi::bam3d_meshes_upload(ctx.0, p.verts.as_ptr(), p.voff.as_ptr(), n, &mut h) } }
                 else { unsafe { ffi::bam3d_meshes_upload_faces(ctx.0, p.verts.as_ptr(), p.voff.as_ptr(), n, p.faces.as_ptr(), p.foff.as_ptr(), &mut h) } };
        ctx.check(rc)?;
        Ok(Some(MeshLibrary { ctx, h }))
    }
}
impl<'a> Drop for MeshLibrary<'a> { fn drop(&mut self) { unsafe { ffi::bam3d_meshes_free(self.ctx.0, self.h) } } }

/// Primitives + model-to-world `Mat4`, packed into device buffers (`bam3d_shapes`); `Mat4::inverse()` is evaluated once per
/// shape there instead of once per `support_point` call.
pub struct ShapeSet<'a> { pub(crate) ctx: &'a Context, pub(crate) h: *mut ffi::Shapes, pub(crate) n: usize, _meshes: Option<MeshLibrary<'a>> }
impl<'a> ShapeSet<'a> {
    pub fn new(ctx: &'a Context, shapes: &[(Primitive3, Mat4)]) -> Result<Self> {
        let p = pack_all(shapes);
        let meshes = MeshLibrary::from_packed(ctx, &p)?;
        let mut h = std::ptr::null_mut();
        let mh = meshes.as_ref().map_or(std::ptr::null(), |m| m.h as *const ffi::Meshes);
        ctx.check(unsafe { ffi::bam3d_shapes_upload(ctx.0, p.kind.as_ptr(), p.params.as_ptr(), p.xform.as_ptr(), p.mesh_id.as_ptr(), shapes.len() as u32, mh, &mut h) })?;
        Ok(ShapeSet { ctx, h, n: shapes.len(), _meshes: meshes })
    }
    pub fn len(&self) -> usize { self.n }
    /// Next frame's transforms for the same primitives.
    pub fn set_transforms(&mut self, transforms: &[Mat4]) -> Result<()> {
        assert_eq!(transforms.len(), self.n);
        let flat: Vec<f32> = transforms.iter().flat_map(|t| t.to_cols_array().to_vec()).collect();
        self.ctx.check(unsafe { ffi::bam3d_shapes_set_transforms(self.ctx.0, self.h, flat.as_ptr()) })
    }
    /// The same in the 48-byte layout (the four columns' xyz — a model transform's fourth row is (0, 0, 0, 1)): a quarter
    /// less to upload per frame, bit-identical records on the device.
    pub fn set_transforms_affine(&mut self, transforms: &[Mat4]) -> Result<()> {
        assert_eq!(transforms.len(), self.n);
        let flat: Vec<f32> = transforms.iter().flat_map(|t| { let c = t.to_cols_array(); [c[0], c[1], c[2], c[4], c[5], c[6], c[8], c[9], c[10], c[12], c[13], c[14]] }).collect();
        self.ctx.check(unsafe { ffi::bam3d_shapes_set_transforms_affine(self.ctx.0, self.h, flat.as_ptr()) })
    }
}
impl<'a> Drop for ShapeSet<'a> { fn drop(&mut self) { unsafe { ffi::bam3d_shapes_free(self.ctx.0, self.h) } } }

/// Page-locked host buffer (`bam3d_host_alloc`): what a frame's transforms / pairs / results should live in so that the
/// PCIe copies are asynchronous and run at full rate.
pub struct Pinned<'a, T: Copy> { ctx: &'a Context, pub(crate) p: *mut T, pub(crate) len: usize }
impl<'a, T: Copy> Pinned<'a, T> {
    pub fn new(ctx: &'a Context, len: usize) -> Result<Self> {
        let mut p: *mut c_void = std::ptr::null_mut();
        ctx.check(unsafe { ffi::bam3d_host_alloc(ctx.0, (len.max(1) * std::mem::size_of::<T>()) as u64, &mut p) })?;
        Ok(Pinned { ctx, p: p as *mut T, len })
    }
    pub fn as_slice(&self) -> &[T] { unsafe { std::slice::from_raw_parts(self.p, self.len) } }
    pub fn as_mut_slice(&mut self) -> &mut [T] { unsafe { std::slice::from_raw_parts_mut(self.p, self.len) } }
}
impl<'a, T: Copy> Drop for Pinned<'a, T> { fn drop(&mut self) { unsafe { ffi::bam3d_host_free(self.ctx.0, self.p as *mut c_void) } } }
