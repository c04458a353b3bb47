//! Broad phase on the device with the crate's names: `SweepAndPrune3` (algorithm/broad_phase/sweep_prune.rs:27-128),
//! the query side of `DynamicBoundingVolumeTree` (dbvt/mod.rs:481-515 with DiscreteVisitor / ContinuousVisitor / FrustumVisitor,
//! dbvt/util.rs query_ray / query_ray_closest) on a static device BVH, `DbvtBroadPhase` (algorithm/broad_phase/dbvt.rs:14-63),
//! and the broad -> narrow pipeline (BASELINE configs[3]).
use crate::{ffi, gjk::GJK, Context, Result, ShapeSet};
use bam3d::{Aabb3, CollisionStrategy, Contact, Frustum, Ray, Relation, Sphere};
use glam::Vec3;

fn flat_boxes(b: &[Aabb3]) -> Vec<f32> { b.iter().flat_map(|b| vec![b.min.x(), b.min.y(), b.min.z(), b.max.x(), b.max.y(), b.max.z()]).collect() }

/// Calls `f(capacity)` again with the size the ABI reports when the first guess is too small (`BAM3D_ERR_CAPACITY`).
fn with_capacity<T>(ctx: &Context, guess: u64, mut f: impl FnMut(u64, &mut u64) -> (i32, T)) -> Result<(T, u64)> {
    let mut cap = guess.max(1);
    loop {
        let mut count = 0u64;
        let (rc, out) = f(cap, &mut count);
        if rc == ffi::ERR_CAPACITY && count > cap { cap = count; continue; }
        ctx.check(rc)?;
        return Ok((out, count));
    }
}

/// `SweepAndPrune3` (sweep_prune.rs:27-30): keeps `sweep_axis` across calls like the reference.
pub struct SweepAndPrune3<'a> { ctx: &'a Context, sweep_axis: usize }
impl<'a> SweepAndPrune3<'a> {
    pub fn new(ctx: &'a Context) -> Self { SweepAndPrune3 { ctx, sweep_axis: 0 } }                                     // :37-39
    pub fn with_sweep_axis(ctx: &'a Context, sweep_axis: usize) -> Self { SweepAndPrune3 { ctx, sweep_axis } }         // :42-47
    pub fn get_sweep_axis(&self) -> usize { self.sweep_axis }                                                          // :50-52
    /// `find_collider_pairs` (:69-128): re-sorts `shapes` in place like the reference and returns indices into the sorted slice,
    /// in the reference's emission order; the next sweep axis is Variance3's choice.
    pub fn find_collider_pairs(&mut self, shapes: &mut [Aabb3]) -> Result<Vec<(usize, usize)>> { self.find_shard(shapes, 0, 1) }
    /// Multi-GPU form: only the pairs whose higher sorted index lies in shard `shard` of `n_shards`; the shards' lists
    /// concatenate to the full list (the sort is replicated on every GPU, the sweep is sharded).
    pub fn find_shard(&mut self, shapes: &mut [Aabb3], shard: u32, n_shards: u32) -> Result<Vec<(usize, usize)>> {
        let n = shapes.len();
        let flat = flat_boxes(shapes);
        let mut perm = vec![0u32; n];
        let mut axis = self.sweep_axis as i32;
        let start_axis = axis;
        let ((pairs, ax), count) = with_capacity(self.ctx, 8 * n as u64 + 1024, |cap, count| {
            let mut pairs = vec![0u32; 2 * cap as usize];
            axis = start_axis;
            let rc = unsafe { ffi::bam3d_sap_find_pairs_shard(self.ctx.0, flat.as_ptr(), n as u64, &mut axis, perm.as_mut_ptr(), pairs.as_mut_ptr(), cap, count, shard, n_shards) };
            (rc, (pairs, axis))
        })?;
        self.sweep_axis = ax as usize;
        let sorted: Vec<Aabb3> = perm.iter().map(|&k| shapes[k as usize].clone()).collect();
        shapes.clone_from_slice(&sorted);
        Ok((0..count as usize).map(|k| (pairs[2 * k] as usize, pairs[2 * k + 1] as usize)).collect())
    }
    /// `SweepAndPrune3<volume::Sphere>` (Bound for Sphere, volume/sphere.rs:17-48): touching spheres are reported.
    pub fn find_collider_pairs_spheres(&mut self, shapes: &mut [Sphere]) -> Result<Vec<(usize, usize)>> {
        let n = shapes.len();
        let flat: Vec<f32> = shapes.iter().flat_map(|s| vec![s.center.x(), s.center.y(), s.center.z(), s.radius]).collect();
        let mut perm = vec![0u32; n];
        let mut axis = self.sweep_axis as i32;
        let start_axis = axis;
        let ((pairs, ax), count) = with_capacity(self.ctx, 8 * n as u64 + 1024, |cap, count| {
            let mut pairs = vec![0u32; 2 * cap as usize];
            axis = start_axis;
            let rc = unsafe { ffi::bam3d_sap_find_pairs_spheres(self.ctx.0, flat.as_ptr(), n as u64, &mut axis, perm.as_mut_ptr(), pairs.as_mut_ptr(), cap, count) };
            (rc, (pairs, axis))
        })?;
        self.sweep_axis = ax as usize;
        let sorted: Vec<Sphere> = perm.iter().map(|&k| shapes[k as usize].clone()).collect();
        shapes.clone_from_slice(&sorted);
        Ok((0..count as usize).map(|k| (pairs[2 * k] as usize, pairs[2 * k + 1] as usize)).collect())
    }
}

/// `ComputeBound<Aabb3>` + `Aabb::transform` for every shape of a set (primitive3.rs:83-96, aabb3.rs:89-96).
pub fn world_aabbs(shapes: &ShapeSet) -> Result<Vec<Aabb3>> {
    let mut out = vec![0f32; 6 * shapes.len()];
    shapes.ctx.check(unsafe { ffi::bam3d_shapes_world_aabbs(shapes.ctx.0, shapes.h, out.as_mut_ptr()) })?;
    Ok(out.chunks(6).map(|c| Aabb3::new(Vec3::new(c[0], c[1], c[2]), Vec3::new(c[3], c[4], c[5]))).collect())
}
/// `ComputeBound<volume::Sphere>` + `Bound::transform_volume` for every shape of a set (primitive3.rs:98-114).
pub fn world_spheres(shapes: &ShapeSet) -> Result<Vec<Sphere>> {
    let mut out = vec![0f32; 4 * shapes.len()];
    shapes.ctx.check(unsafe { ffi::bam3d_shapes_world_spheres(shapes.ctx.0, shapes.h, out.as_mut_ptr()) })?;
    Ok(out.chunks(4).map(|c| Sphere { center: Vec3::new(c[0], c[1], c[2]), radius: c[3] }).collect())
}

/// One candidate of the broad -> narrow pipeline: the two SHAPE indices (not sorted positions) and the narrow-phase outcome.
pub struct Candidate { pub left: usize, pub right: usize, pub contact: Option<Contact> }
/// World AABBs -> `SweepAndPrune3::find_collider_pairs` -> `GJK::intersection` on every candidate, resident on the device
/// (`bam3d_pipeline_contacts`). The reference leaves this glue to its caller; candidates come in the reference's emission order.
pub fn collide_all(gjk: &GJK, strategy: &CollisionStrategy, shapes: &ShapeSet, sweep_axis: &mut usize) -> Result<Vec<Candidate>> {
    let ctx = gjk.context();
    let s = match strategy { CollisionStrategy::FullResolution => ffi::FULL_RESOLUTION, CollisionStrategy::CollisionOnly => ffi::COLLISION_ONLY };
    let start_axis = *sweep_axis as i32;
    let mut axis = start_axis;
    let ((pairs, raw, st, ax), count) = with_capacity(ctx, 8 * shapes.len() as u64 + 1024, |cap, count| {
        let (mut pairs, mut raw, mut st) = (vec![0u32; 2 * cap as usize], vec![ffi::RawContact::default(); cap as usize], vec![0u8; cap as usize]);
        axis = start_axis;
        let rc = unsafe { ffi::bam3d_pipeline_contacts(ctx.0, shapes.h, &mut axis, gjk.settings(), s, pairs.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut(),
                                                       raw.as_mut_ptr(), st.as_mut_ptr(), cap, count) };
        (rc, (pairs, raw, st, axis))
    })?;
    *sweep_axis = ax as usize;
    Ok((0..count as usize).map(|k| Candidate { left: pairs[2 * k] as usize, right: pairs[2 * k + 1] as usize,
        contact: if st[k] == ffi::STATUS_OK {
            let c = &raw[k];
            Some(Contact::new_with_point(strategy.clone(), Vec3::new(c.normal[0], c.normal[1], c.normal[2]), c.penetration_depth,
                                         Vec3::new(c.contact_point[0], c.contact_point[1], c.contact_point[2])))
        } else { None } }).collect())
}

/// The query side of `DynamicBoundingVolumeTree` over a static set of values (value index = position in `bounds`). Result
/// sets equal the reference's (they do not depend on its randomised tree shape); their order is the device tree's.
pub struct DynamicBoundingVolumeTree<'a> { ctx: &'a Context, h: *mut ffi::Bvh, n: usize }
impl<'a> DynamicBoundingVolumeTree<'a> {
    pub fn new(ctx: &'a Context, bounds: &[Aabb3]) -> Result<Self> {
        let flat = flat_boxes(bounds);
        let mut h = std::ptr::null_mut();
        ctx.check(unsafe { ffi::bam3d_bvh_build(ctx.0, flat.as_ptr(), bounds.len() as u64, &mut h) })?;
        Ok(DynamicBoundingVolumeTree { ctx, h, n: bounds.len() })
    }
    /// The same tree over values bounded by `volume::Sphere` (`DynamicBoundingVolumeTree` is generic over the bound): `query_ray`
    /// `query_ray_closest` and `query_frustum` then apply `Continuous<Ray>` / `PlaneBound for Sphere`, `query_discrete_sphere` and
    /// `DbvtBroadPhase` `Discrete<Sphere>`; `query_discrete` (an Aabb3 query) and `refit` report an error on such a tree.
    pub fn new_with_spheres(ctx: &'a Context, bounds: &[Sphere]) -> Result<Self> {
        let flat: Vec<f32> = bounds.iter().flat_map(|s| vec![s.center.x(), s.center.y(), s.center.z(), s.radius]).collect();
        let mut h = std::ptr::null_mut();
        ctx.check(unsafe { ffi::bam3d_bvh_build_spheres(ctx.0, flat.as_ptr(), bounds.len() as u64, &mut h) })?;
        Ok(DynamicBoundingVolumeTree { ctx, h, n: bounds.len() })
    }
    /// `query(&mut DiscreteVisitor::new(&sphere))` on a tree over sphere bounds (sphere.rs:82-91; touching spheres intersect).
    pub fn query_discrete_sphere(&self, queries: &[Sphere]) -> Result<Vec<Vec<usize>>> {
        let flat: Vec<f32> = queries.iter().flat_map(|s| vec![s.center.x(), s.center.y(), s.center.z(), s.radius]).collect();
        let mut off = vec![0u64; queries.len() + 1];
        let (vals, _) = with_capacity(self.ctx, 16 * queries.len() as u64 + 1024, |cap, total| {
            let mut vals = vec![0u32; cap as usize];
            let rc = unsafe { ffi::bam3d_bvh_query_sphere(self.ctx.0, self.h, flat.as_ptr(), queries.len() as u64, off.as_mut_ptr(), vals.as_mut_ptr(), cap, total) };
            (rc, vals)
        })?;
        Ok((0..queries.len()).map(|q| vals[off[q] as usize..off[q + 1] as usize].iter().map(|&v| v as usize).collect()).collect())
    }
    /// `update_node` + `do_refit` for every value at once (dbvt/mod.rs:561-589): new bounds, same topology.
    pub fn refit(&mut self, bounds: &[Aabb3]) -> Result<()> {
        assert_eq!(bounds.len(), self.n);
        let flat = flat_boxes(bounds);
        self.ctx.check(unsafe { ffi::bam3d_bvh_refit(self.ctx.0, self.h, flat.as_ptr()) })
    }
    /// `query(&mut DiscreteVisitor::new(&bound))` (dbvt/visitor.rs:57-90) for every query box: value indices per query.
    pub fn query_discrete(&self, queries: &[Aabb3]) -> Result<Vec<Vec<usize>>> {
        let flat = flat_boxes(queries);
        let mut off = vec![0u64; queries.len() + 1];
        let (vals, _) = with_capacity(self.ctx, 16 * queries.len() as u64 + 1024, |cap, total| {
            let mut vals = vec![0u32; cap as usize];
            let rc = unsafe { ffi::bam3d_bvh_query_aabb(self.ctx.0, self.h, flat.as_ptr(), queries.len() as u64, off.as_mut_ptr(), vals.as_mut_ptr(), cap, total) };
            (rc, vals)
        })?;
        Ok((0..queries.len()).map(|q| vals[off[q] as usize..off[q + 1] as usize].iter().map(|&v| v as usize).collect()).collect())
    }
    /// `query_ray` (dbvt/util.rs:114-129): (value index, entry point) of every value the ray hits, per ray.
    pub fn query_ray(&self, rays: &[Ray]) -> Result<Vec<Vec<(usize, Vec3)>>> {
        let flat: Vec<f32> = rays.iter().flat_map(|r| vec![r.origin.x(), r.origin.y(), r.origin.z(), r.direction.x(), r.direction.y(), r.direction.z()]).collect();
        let mut off = vec![0u64; rays.len() + 1];
        let ((vals, pts), _) = with_capacity(self.ctx, 16 * rays.len() as u64 + 1024, |cap, total| {
            let (mut vals, mut pts) = (vec![0u32; cap as usize], vec![0f32; 3 * cap as usize]);
            let rc = unsafe { ffi::bam3d_bvh_query_ray(self.ctx.0, self.h, flat.as_ptr(), rays.len() as u64, off.as_mut_ptr(), vals.as_mut_ptr(), pts.as_mut_ptr(), cap, total) };
            (rc, (vals, pts))
        })?;
        Ok((0..rays.len()).map(|q| (off[q] as usize..off[q + 1] as usize).map(|i| (vals[i] as usize, Vec3::new(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]))).collect()).collect())
    }
    /// `query_ray_closest` (dbvt/util.rs:77-103): the hit of smallest t over ALL values the ray hits (the reference's own answer
    /// can be larger when a branch containing the ray's origin is pruned by its exit distance: see tests/test_broad_gpu.py).
    pub fn query_ray_closest(&self, rays: &[Ray]) -> Result<Vec<Option<(usize, Vec3)>>> {
        let flat: Vec<f32> = rays.iter().flat_map(|r| vec![r.origin.x(), r.origin.y(), r.origin.z(), r.direction.x(), r.direction.y(), r.direction.z()]).collect();
        let (mut val, mut pt) = (vec![0u32; rays.len()], vec![0f32; 4 * rays.len()]);
        self.ctx.check(unsafe { ffi::bam3d_bvh_query_ray_closest(self.ctx.0, self.h, flat.as_ptr(), rays.len() as u64, val.as_mut_ptr(), pt.as_mut_ptr()) })?;
        Ok((0..rays.len()).map(|q| if val[q] == u32::MAX { None } else { Some((val[q] as usize, Vec3::new(pt[4 * q], pt[4 * q + 1], pt[4 * q + 2]))) }).collect())
    }
    /// `query(&mut FrustumVisitor::new(&frustum))` (dbvt/visitor.rs:99-134): (value index, Relation) per frustum; Out is never reported.
    pub fn query_frustum(&self, frusta: &[Frustum]) -> Result<Vec<Vec<(usize, Relation)>>> {
        let flat: Vec<f32> = frusta.iter().flat_map(|f| [&f.left, &f.right, &f.bottom, &f.top, &f.near, &f.far].iter()
            .flat_map(|p| vec![p.n.x(), p.n.y(), p.n.z(), p.d]).collect::<Vec<f32>>()).collect();
        let mut off = vec![0u64; frusta.len() + 1];
        let ((vals, rel), _) = with_capacity(self.ctx, 1024, |cap, total| {
            let (mut vals, mut rel) = (vec![0u32; cap as usize], vec![0u8; cap as usize]);
            let rc = unsafe { ffi::bam3d_bvh_query_frustum(self.ctx.0, self.h, flat.as_ptr(), frusta.len() as u64, off.as_mut_ptr(), vals.as_mut_ptr(), rel.as_mut_ptr(), cap, total) };
            (rc, (vals, rel))
        })?;
        Ok((0..frusta.len()).map(|q| (off[q] as usize..off[q + 1] as usize).map(|i| (vals[i] as usize, if rel[i] == 0 { Relation::In } else { Relation::Cross })).collect()).collect())
    }
}
impl<'a> Drop for DynamicBoundingVolumeTree<'a> { fn drop(&mut self) { unsafe { ffi::bam3d_bvh_free(self.ctx.0, self.h) } } }

/// `DbvtBroadPhase` (algorithm/broad_phase/dbvt.rs:14-63): every dirty value queries the tree; sorted unique (low, high) pairs.
pub struct DbvtBroadPhase;
impl DbvtBroadPhase {
    pub fn new() -> Self { DbvtBroadPhase }
    pub fn find_collider_pairs(&self, tree: &DynamicBoundingVolumeTree, dirty: &[bool]) -> Result<Vec<(usize, usize)>> {
        assert_eq!(dirty.len(), tree.n);
        let d: Vec<u8> = dirty.iter().map(|&b| b as u8).collect();
        let (pairs, count) = with_capacity(tree.ctx, 8 * tree.n as u64 + 1024, |cap, count| {
            let mut pairs = vec![0u32; 2 * cap as usize];
            let rc = unsafe { ffi::bam3d_bvh_find_collider_pairs(tree.ctx.0, tree.h, d.as_ptr(), pairs.as_mut_ptr(), cap, count) };
            (rc, pairs)
        })?;
        Ok((0..count as usize).map(|k| (pairs[2 * k] as usize, pairs[2 * k + 1] as usize)).collect())
    }
}
