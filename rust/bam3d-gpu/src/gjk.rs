//! `GJK` on the device: the crate's `impl GJK` (algorithm/minkowski/gjk/mod.rs:32-524) with the same method names and
//! `Option` behaviour. The scalar methods are batches of one (they exist so that code written against the crate compiles
//! unchanged); the `_batch` methods are what makes the path data parallel.
use crate::{ffi, pack_all, Context, MeshLibrary, Result, ShapeSet};
use bam3d::{CollisionStrategy, Contact, Primitive3};
use glam::{Mat4, Vec3};
use std::ops::Range;

pub struct GJK<'a> { ctx: &'a Context, cfg: ffi::GjkSettings }

fn strategy_code(s: &CollisionStrategy) -> i32 { match s { CollisionStrategy::FullResolution => ffi::FULL_RESOLUTION, CollisionStrategy::CollisionOnly => ffi::COLLISION_ONLY } }
fn contact_of(strategy: &CollisionStrategy, c: &ffi::RawContact) -> Contact {
    let mut k = Contact::new_with_point(strategy.clone(), Vec3::new(c.normal[0], c.normal[1], c.normal[2]), c.penetration_depth,
                                        Vec3::new(c.contact_point[0], c.contact_point[1], c.contact_point[2]));
    k.time_of_impact = c.time_of_impact;
    k
}
fn flat_pairs(pairs: &[(u32, u32)]) -> Vec<u32> { pairs.iter().flat_map(|&(l, r)| vec![l, r]).collect() }

/// Compound shapes resident on the device (`bam3d_compounds`): the crate's `&[(Primitive, Mat4)]` slices, one per compound.
pub struct CompoundSet<'a> { ctx: &'a Context, h: *mut ffi::Compounds, n: usize, _meshes: Option<MeshLibrary<'a>> }
impl<'a> CompoundSet<'a> {
    pub fn new(ctx: &'a Context, compounds: &[&[(Primitive3, Mat4)]]) -> Result<Self> {
        let all: Vec<(Primitive3, Mat4)> = compounds.iter().flat_map(|c| c.iter().cloned()).collect();
        let mut offsets = vec![0u32];
        for c in compounds { offsets.push(offsets.last().unwrap() + c.len() as u32); }
        let p = pack_all(&all);
        let meshes = MeshLibrary::from_packed(ctx, &p)?;
        let mh = meshes.as_ref().map_or(std::ptr::null(), |m| m.h as *const ffi::Meshes);
        let mut h = std::ptr::null_mut();
        ctx.check(unsafe { ffi::bam3d_compounds_upload(ctx.0, p.kind.as_ptr(), p.params.as_ptr(), p.xform.as_ptr(), p.mesh_id.as_ptr(), all.len() as u32,
                                                       offsets.as_ptr(), compounds.len() as u32, mh, &mut h) })?;
        Ok(CompoundSet { ctx, h, n: compounds.len(), _meshes: meshes })
    }
}
impl<'a> Drop for CompoundSet<'a> { fn drop(&mut self) { unsafe { ffi::bam3d_compounds_free(self.ctx.0, self.h) } } }

impl<'a> GJK<'a> {
    /// `GJK::new()` (gjk/mod.rs:34-42)
    pub fn new(ctx: &'a Context) -> Self {
        let mut cfg = ffi::GjkSettings { distance_tolerance: 0., continuous_tolerance: 0., epa_tolerance: 0., max_iterations: 0, toi_iteration_cap: 0 };
        unsafe { ffi::bam3d_default_settings(&mut cfg) };
        GJK { ctx, cfg }
    }
    /// `GJK::new_with_settings` (gjk/mod.rs:45-58)
    pub fn new_with_settings(ctx: &'a Context, distance_tolerance: f32, continuous_tolerance: f32, epa_tolerance: f32, max_iterations: u32) -> Self {
        GJK { ctx, cfg: ffi::GjkSettings { distance_tolerance, continuous_tolerance, epa_tolerance, max_iterations, toi_iteration_cap: 100 } }
    }

    // ---- batched forms -------------------------------------------------------------------------------------------
    /// `GJK::intersect` per pair (gjk/mod.rs:74-113): `true` where the reference returns `Some(simplex)`.
    pub fn intersect_batch(&self, shapes: &ShapeSet, pairs: &[(u32, u32)]) -> Result<Vec<bool>> {
        let (flat, mut st) = (flat_pairs(pairs), vec![0u8; pairs.len()]);
        self.ctx.check(unsafe { ffi::bam3d_gjk_intersect(self.ctx.0, shapes.h, flat.as_ptr(), pairs.len() as u64, &self.cfg, st.as_mut_ptr()) })?;
        Ok(st.iter().map(|&s| s == ffi::STATUS_OK).collect())
    }
    /// `GJK::intersection` per pair (gjk/mod.rs:317-341).
    pub fn intersection_batch(&self, strategy: &CollisionStrategy, shapes: &ShapeSet, pairs: &[(u32, u32)]) -> Result<Vec<Option<Contact>>> {
        let flat = flat_pairs(pairs);
        let (mut raw, mut st) = (vec![ffi::RawContact::default(); pairs.len()], vec![0u8; pairs.len()]);
        self.ctx.check(unsafe { ffi::bam3d_gjk_contact(self.ctx.0, shapes.h, flat.as_ptr(), pairs.len() as u64, &self.cfg, strategy_code(strategy), raw.as_mut_ptr(), st.as_mut_ptr()) })?;
        Ok(raw.iter().zip(&st).map(|(c, &s)| if s == ffi::STATUS_OK { Some(contact_of(strategy, c)) } else { None }).collect())
    }
    /// `GJK::distance` per pair (gjk/mod.rs:232-280): the value the reference returns (its convergence residual), `None` when the
    /// shapes collide or the iteration budget runs out. Panics where the reference panics (status `REF_PANIC`).
    pub fn distance_batch(&self, shapes: &ShapeSet, pairs: &[(u32, u32)]) -> Result<Vec<Option<f32>>> {
        let flat = flat_pairs(pairs);
        let (mut rv, mut geo, mut st) = (vec![0f32; pairs.len()], vec![0f32; pairs.len()], vec![0u8; pairs.len()]);
        self.ctx.check(unsafe { ffi::bam3d_gjk_distance(self.ctx.0, shapes.h, flat.as_ptr(), pairs.len() as u64, &self.cfg, rv.as_mut_ptr(), geo.as_mut_ptr(), st.as_mut_ptr()) })?;
        Ok(rv.iter().zip(&st).map(|(&v, &s)| { assert!(s != ffi::STATUS_REF_PANIC, "GJK::distance: the reference panics on this pair (simplex3d.rs:134)"); if s == ffi::STATUS_OK { Some(v) } else { None } }).collect())
    }
    /// `GJK::intersection_time_of_impact` per pair (gjk/mod.rs:130-217); `start` / `end` hold the same primitives at the two ends
    /// of the step (`Range<&Mat4>`).
    pub fn intersection_time_of_impact_batch(&self, start: &ShapeSet, end: &ShapeSet, pairs: &[(u32, u32)]) -> Result<Vec<Option<Contact>>> {
        let flat = flat_pairs(pairs);
        let (mut raw, mut st) = (vec![ffi::RawContact::default(); pairs.len()], vec![0u8; pairs.len()]);
        self.ctx.check(unsafe { ffi::bam3d_gjk_time_of_impact(self.ctx.0, start.h, end.h, flat.as_ptr(), pairs.len() as u64, &self.cfg, raw.as_mut_ptr(), st.as_mut_ptr()) })?;
        Ok(raw.iter().zip(&st).map(|(c, &s)| if s == ffi::STATUS_OK { Some(contact_of(&CollisionStrategy::FullResolution, c)) } else { None }).collect())
    }
    /// `GJK::intersection_complex` per compound pair (gjk/mod.rs:362-407); `transforms[c]` = model-to-world of compound c.
    pub fn intersection_complex_batch(&self, strategy: &CollisionStrategy, compounds: &CompoundSet, transforms: &[Mat4], pairs: &[(u32, u32)]) -> Result<Vec<Option<Contact>>> {
        assert_eq!(transforms.len(), compounds.n);
        let (flat, xf): (Vec<u32>, Vec<f32>) = (flat_pairs(pairs), transforms.iter().flat_map(|t| t.to_cols_array().to_vec()).collect());
        let (mut raw, mut st) = (vec![ffi::RawContact::default(); pairs.len()], vec![0u8; pairs.len()]);
        self.ctx.check(unsafe { ffi::bam3d_gjk_contact_complex(self.ctx.0, compounds.h, xf.as_ptr(), flat.as_ptr(), pairs.len() as u64, &self.cfg, strategy_code(strategy), raw.as_mut_ptr(), st.as_mut_ptr()) })?;
        Ok(raw.iter().zip(&st).map(|(c, &s)| { assert!(s != ffi::STATUS_REF_PANIC, "intersection_complex: partial_cmp(..).unwrap() on a NaN depth"); if s == ffi::STATUS_OK { Some(contact_of(strategy, c)) } else { None } }).collect())
    }
    /// `GJK::distance_complex` per compound pair (gjk/mod.rs:422-456).
    pub fn distance_complex_batch(&self, compounds: &CompoundSet, transforms: &[Mat4], pairs: &[(u32, u32)]) -> Result<Vec<Option<f32>>> {
        let (flat, xf): (Vec<u32>, Vec<f32>) = (flat_pairs(pairs), transforms.iter().flat_map(|t| t.to_cols_array().to_vec()).collect());
        let (mut v, mut st) = (vec![0f32; pairs.len()], vec![0u8; pairs.len()]);
        self.ctx.check(unsafe { ffi::bam3d_gjk_distance_complex(self.ctx.0, compounds.h, xf.as_ptr(), flat.as_ptr(), pairs.len() as u64, &self.cfg, v.as_mut_ptr(), st.as_mut_ptr()) })?;
        Ok(v.iter().zip(&st).map(|(&d, &s)| if s == ffi::STATUS_OK { Some(d) } else { None }).collect())
    }
    /// `GJK::intersection_complex_time_of_impact` per compound pair (gjk/mod.rs:478-523).
    pub fn intersection_complex_time_of_impact_batch(&self, strategy: &CollisionStrategy, compounds: &CompoundSet, start: &[Mat4], end: &[Mat4],
                                                     pairs: &[(u32, u32)]) -> Result<Vec<Option<Contact>>> {
        let flat = flat_pairs(pairs);
        let x0: Vec<f32> = start.iter().flat_map(|t| t.to_cols_array().to_vec()).collect();
        let x1: Vec<f32> = end.iter().flat_map(|t| t.to_cols_array().to_vec()).collect();
        let (mut raw, mut st) = (vec![ffi::RawContact::default(); pairs.len()], vec![0u8; pairs.len()]);
        self.ctx.check(unsafe { ffi::bam3d_gjk_time_of_impact_complex(self.ctx.0, compounds.h, x0.as_ptr(), x1.as_ptr(), flat.as_ptr(), pairs.len() as u64, &self.cfg,
                                                                      strategy_code(strategy), raw.as_mut_ptr(), st.as_mut_ptr()) })?;
        Ok(raw.iter().zip(&st).map(|(c, &s)| if s == ffi::STATUS_OK { Some(contact_of(strategy, c)) } else { None }).collect())
    }

    // ---- the crate's scalar methods, as batches of one -------------------------------------------------------------
    fn two(&self, left: &Primitive3, lt: &Mat4, right: &Primitive3, rt: &Mat4) -> ShapeSet<'a> {
        ShapeSet::new(self.ctx, &[(left.clone(), *lt), (right.clone(), *rt)]).expect("bam3d_shapes_upload")
    }
    /// gjk/mod.rs:74-113 (the simplex itself stays on the device: `Some(())` stands for `Some(simplex)`)
    pub fn intersect(&self, left: &Primitive3, left_transform: &Mat4, right: &Primitive3, right_transform: &Mat4) -> Option<()> {
        let s = self.two(left, left_transform, right, right_transform);
        if self.intersect_batch(&s, &[(0, 1)]).ok()?[0] { Some(()) } else { None }
    }
    /// gjk/mod.rs:317-341
    pub fn intersection(&self, strategy: &CollisionStrategy, left: &Primitive3, left_transform: &Mat4, right: &Primitive3, right_transform: &Mat4) -> Option<Contact> {
        let s = self.two(left, left_transform, right, right_transform);
        self.intersection_batch(strategy, &s, &[(0, 1)]).ok()?.pop().unwrap()
    }
    /// gjk/mod.rs:232-280
    pub fn distance(&self, left: &Primitive3, left_transform: &Mat4, right: &Primitive3, right_transform: &Mat4) -> Option<f32> {
        let s = self.two(left, left_transform, right, right_transform);
        self.distance_batch(&s, &[(0, 1)]).ok()?[0]
    }
    /// gjk/mod.rs:130-217
    pub fn intersection_time_of_impact(&self, left: &Primitive3, left_transform: Range<&Mat4>, right: &Primitive3, right_transform: Range<&Mat4>) -> Option<Contact> {
        let s0 = self.two(left, left_transform.start, right, right_transform.start);
        let s1 = self.two(left, left_transform.end, right, right_transform.end);
        self.intersection_time_of_impact_batch(&s0, &s1, &[(0, 1)]).ok()?.pop().unwrap()
    }
    /// gjk/mod.rs:362-407
    pub fn intersection_complex(&self, strategy: &CollisionStrategy, left: &[(Primitive3, Mat4)], left_transform: &Mat4, right: &[(Primitive3, Mat4)],
                                right_transform: &Mat4) -> Option<Contact> {
        let c = CompoundSet::new(self.ctx, &[left, right]).ok()?;
        self.intersection_complex_batch(strategy, &c, &[*left_transform, *right_transform], &[(0, 1)]).ok()?.pop().unwrap()
    }
    /// gjk/mod.rs:422-456
    pub fn distance_complex(&self, left: &[(Primitive3, Mat4)], left_transform: &Mat4, right: &[(Primitive3, Mat4)], right_transform: &Mat4) -> Option<f32> {
        let c = CompoundSet::new(self.ctx, &[left, right]).ok()?;
        self.distance_complex_batch(&c, &[*left_transform, *right_transform], &[(0, 1)]).ok()?[0]
    }
    /// gjk/mod.rs:478-523
    pub fn intersection_complex_time_of_impact(&self, strategy: &CollisionStrategy, left: &[(Primitive3, Mat4)], left_transform: Range<&Mat4>,
                                               right: &[(Primitive3, Mat4)], right_transform: Range<&Mat4>) -> Option<Contact> {
        let c = CompoundSet::new(self.ctx, &[left, right]).ok()?;
        self.intersection_complex_time_of_impact_batch(strategy, &c, &[*left_transform.start, *right_transform.start],
                                                       &[*left_transform.end, *right_transform.end], &[(0, 1)]).ok()?.pop().unwrap()
    }
    pub(crate) fn settings(&self) -> &ffi::GjkSettings { &self.cfg }
    pub(crate) fn context(&self) -> &'a Context { self.ctx }
}
