//! The crate's query traits (traits.rs:7-147) for shapes that live on the device, so that code written against
//! `Discrete` / `Continuous` / `DiscreteTransformed` / `ContinuousTransformed` runs unchanged on a `DeviceShape`:
//!   * `Discrete<Ray>` / `Continuous<Ray>` and, through them, the crate's blanket `DiscreteTransformed<Ray>` /
//!     `ContinuousTransformed<Ray>` (ray.rs:58-78) — the device applies the stored model-to-world transform itself, so the
//!     `_transformed` forms are given explicitly too (with the caller's transform replacing the stored one);
//!   * `Discrete<DeviceShape>` = `GJK::intersect`, `Continuous<DeviceShape, Result = Contact>` = `GJK::intersection`
//!     (FullResolution) between two shapes of the same set.
//! Each call is a batch of one: these impls are for drop-in compatibility; throughput comes from the `_batch` methods.
use crate::{ffi, gjk::GJK, ShapeSet};
use bam3d::{CollisionStrategy, Contact, Continuous, Discrete, Ray};
use glam::Vec3;

/// Shape `index` of a device-resident set, with the GJK settings to use for shape-shape queries.
pub struct DeviceShape<'s, 'a> { pub set: &'s ShapeSet<'a>, pub gjk: &'s GJK<'a>, pub index: u32 }

fn cast(shape: &DeviceShape, ray: &Ray, query: i32) -> Option<Vec3> {
    let r = [ray.origin.x(), ray.origin.y(), ray.origin.z(), ray.direction.x(), ray.direction.y(), ray.direction.z()];
    let casts = [0u32, shape.index];
    let (mut pt, mut st) = ([0f32; 3], [0u8; 1]);
    let rc = unsafe { ffi::bam3d_ray_cast(shape.set.ctx.0, shape.set.h, r.as_ptr(), 1, casts.as_ptr(), 1, query, pt.as_mut_ptr(), st.as_mut_ptr()) };
    assert_eq!(rc, ffi::OK, "bam3d_ray_cast");
    assert!(st[0] != ffi::STATUS_UNSUPPORTED, "ray cast against a ConvexPolyhedron built without faces (the reference's walk panics on it)");
    if st[0] == ffi::STATUS_OK { Some(Vec3::new(pt[0], pt[1], pt[2])) } else { None }
}

/// `DiscreteTransformed<Ray>::intersects_transformed` with the shape's stored transform (ray.rs:58-66)
impl<'s, 'a> Discrete<Ray> for DeviceShape<'s, 'a> {
    fn intersects(&self, ray: &Ray) -> bool { cast(self, ray, 0).is_some() }
}
/// `ContinuousTransformed<Ray>::intersection_transformed` with the shape's stored transform (ray.rs:68-78): world-space hit point
impl<'s, 'a> Continuous<Ray> for DeviceShape<'s, 'a> {
    type Result = Vec3;
    fn intersection(&self, ray: &Ray) -> Option<Vec3> { cast(self, ray, 1) }
}
/// `GJK::intersect` between two shapes of one set (gjk/mod.rs:74-113)
impl<'s, 'a> Discrete<DeviceShape<'s, 'a>> for DeviceShape<'s, 'a> {
    fn intersects(&self, other: &DeviceShape<'s, 'a>) -> bool {
        self.gjk.intersect_batch(self.set, &[(self.index, other.index)]).map(|v| v[0]).unwrap_or(false)
    }
}
/// `GJK::intersection(FullResolution, ..)` between two shapes of one set (gjk/mod.rs:317-341)
impl<'s, 'a> Continuous<DeviceShape<'s, 'a>> for DeviceShape<'s, 'a> {
    type Result = Contact;
    fn intersection(&self, other: &DeviceShape<'s, 'a>) -> Option<Contact> {
        self.gjk.intersection_batch(&CollisionStrategy::FullResolution, self.set, &[(self.index, other.index)]).ok()?.pop().unwrap()
    }
}
