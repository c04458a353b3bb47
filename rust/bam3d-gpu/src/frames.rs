//! The per-frame sequence of a simulation loop — new `Mat4` for every shape, the narrow phase over the frame's pairs, results
//! back in host memory — with `depth` frames in flight (`bam3d_frames_*`): uploads, kernels (two compute lanes) and
//! downloads of neighbouring frames overlap. Buffers handed to `submit_*` are borrowed until `wait(frame)` returns, which is
//! why they are `Pinned` values owned by the caller for the queue's lifetime.
use crate::{ffi, gjk::GJK, multi::Comm, pack_all, Context, MeshLibrary, Pinned, Result};
use bam3d::{CollisionStrategy, Primitive3};
use glam::Mat4;

pub struct FrameQueue<'a> { ctx: &'a Context, h: *mut ffi::Frames, cfg: ffi::GjkSettings, _meshes: Option<MeshLibrary<'a>> }
fn code(s: &CollisionStrategy) -> i32 { match s { CollisionStrategy::FullResolution => ffi::FULL_RESOLUTION, CollisionStrategy::CollisionOnly => ffi::COLLISION_ONLY } }

impl<'a> FrameQueue<'a> {
    /// `with_end`: also allocate the end-of-step shape sets `intersection_time_of_impact` needs.
    pub fn new(gjk: &GJK<'a>, shapes: &[(Primitive3, Mat4)], max_pairs: usize, depth: u32, with_end: bool) -> Result<Self> {
        let ctx = gjk.context();
        let p = pack_all(shapes);
        let meshes = MeshLibrary::from_packed(ctx, &p)?;
        let mh = meshes.as_ref().map_or(std::ptr::null(), |m| m.h as *const ffi::Meshes);
        let mut h = std::ptr::null_mut();
        ctx.check(unsafe { ffi::bam3d_frames_create(ctx.0, p.kind.as_ptr(), p.params.as_ptr(), p.xform.as_ptr(), p.mesh_id.as_ptr(), shapes.len() as u32, mh,
                                                    max_pairs as u64, depth, with_end as i32, &mut h) })?;
        Ok(FrameQueue { ctx, h, cfg: *gjk.settings(), _meshes: meshes })
    }
    /// Later submits pass 12 floats per shape (the four columns' xyz; `ShapeSet::set_transforms_affine` describes the layout)
    /// through the `*_affine` methods: a quarter less to upload per frame.
    pub fn use_affine_layout(&mut self) -> Result<()> {
        self.ctx.check(unsafe { ffi::bam3d_frames_set_transform_layout(self.ctx.0, self.h, ffi::XFORM_AFFINE3X4) })
    }
    /// `submit_intersection` in the 48-byte layout (after `use_affine_layout`).
    pub fn submit_intersection_affine(&mut self, strategy: &CollisionStrategy, xform: &Pinned<[f32; 12]>, pairs: &Pinned<(u32, u32)>,
                                      out: &mut Pinned<ffi::RawContact>, status: &mut Pinned<u8>) -> Result<u64> {
        let mut f = 0u64;
        self.ctx.check(unsafe { ffi::bam3d_frames_submit(self.ctx.0, self.h, ffi::FRAME_CONTACT, xform.p as *const f32, std::ptr::null(), pairs.p as *const u32,
                                                         pairs.len as u64, &self.cfg, code(strategy), out.p, status.p, &mut f) })?;
        Ok(f)
    }
    /// `GJK::intersect` for every pair under this frame's transforms; returns the frame number.
    pub fn submit_intersect(&mut self, xform: &Pinned<[f32; 16]>, pairs: &Pinned<(u32, u32)>, status: &mut Pinned<u8>) -> Result<u64> {
        let mut f = 0u64;
        self.ctx.check(unsafe { ffi::bam3d_frames_submit(self.ctx.0, self.h, ffi::FRAME_INTERSECT, xform.p as *const f32, std::ptr::null(), pairs.p as *const u32,
                                                         pairs.len as u64, &self.cfg, 0, std::ptr::null_mut(), status.p, &mut f) })?;
        Ok(f)
    }
    /// `GJK::intersection(strategy, ..)` for every pair.
    pub fn submit_intersection(&mut self, strategy: &CollisionStrategy, xform: &Pinned<[f32; 16]>, pairs: &Pinned<(u32, u32)>, out: &mut Pinned<ffi::RawContact>,
                               status: &mut Pinned<u8>) -> Result<u64> {
        let mut f = 0u64;
        self.ctx.check(unsafe { ffi::bam3d_frames_submit(self.ctx.0, self.h, ffi::FRAME_CONTACT, xform.p as *const f32, std::ptr::null(), pairs.p as *const u32,
                                                         pairs.len as u64, &self.cfg, code(strategy), out.p, status.p, &mut f) })?;
        Ok(f)
    }
    /// `GJK::intersection_time_of_impact` between `xform` (start) and `xform_end`.
    pub fn submit_time_of_impact(&mut self, xform: &Pinned<[f32; 16]>, xform_end: &Pinned<[f32; 16]>, pairs: &Pinned<(u32, u32)>, out: &mut Pinned<ffi::RawContact>,
                                 status: &mut Pinned<u8>) -> Result<u64> {
        let mut f = 0u64;
        self.ctx.check(unsafe { ffi::bam3d_frames_submit(self.ctx.0, self.h, ffi::FRAME_TIME_OF_IMPACT, xform.p as *const f32, xform_end.p as *const f32,
                                                         pairs.p as *const u32, pairs.len as u64, &self.cfg, 0, out.p, status.p, &mut f) })?;
        Ok(f)
    }
    /// Multi-GPU: like `submit_intersection`, but `wait` leaves the hits of ALL ranks in `gathered` (rank order) instead of this
    /// rank's dense contacts; `pair_index_base` = global index of this rank's first pair. Every rank must submit and wait in the same order.
    pub fn submit_intersection_gather(&mut self, comm: &Comm, strategy: &CollisionStrategy, xform: &Pinned<[f32; 16]>, pairs: &Pinned<(u32, u32)>,
                                      pair_index_base: u64, gathered: &mut Pinned<ffi::RawContact40>, status: &mut Pinned<u8>) -> Result<u64> {
        let mut f = 0u64;
        self.ctx.check(unsafe { ffi::bam3d_frames_submit_gather(self.ctx.0, self.h, comm.h, ffi::FRAME_CONTACT, xform.p as *const f32, std::ptr::null(),
                                                                pairs.p as *const u32, pairs.len as u64, &self.cfg, code(strategy), pair_index_base, gathered.p,
                                                                gathered.len as u64, status.p, &mut f) })?;
        Ok(f)
    }
    /// Block until that frame's results are in the buffers given to its submit.
    pub fn wait(&mut self, frame: u64) -> Result<()> { self.ctx.check(unsafe { ffi::bam3d_frames_wait(self.ctx.0, self.h, frame) }) }
    /// After `wait` on a gather-mode frame: the per-rank record counts and their sum.
    pub fn gathered(&self, frame: u64, n_ranks: usize) -> Result<(Vec<u64>, u64)> {
        let (mut counts, mut total) = (vec![0u64; n_ranks], 0u64);
        self.ctx.check(unsafe { ffi::bam3d_frames_gathered(self.ctx.0, self.h, frame, counts.as_mut_ptr(), &mut total) })?;
        Ok((counts, total))
    }
}
impl<'a> Drop for FrameQueue<'a> { fn drop(&mut self) { unsafe { ffi::bam3d_frames_free(self.ctx.0, self.h) } } }
