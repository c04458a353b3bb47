// bam3d.hpp — header-only C++17 host layer over the C ABI (bam3d_gpu.h), mirroring the public interface of the
// reference crate for the accelerated path: same type and method names, argument meaning and None/Some behaviour.
//   primitives            reference src/primitive/*.rs          ->  bam3d::Primitive3 (Sphere, Cuboid, Cube, Cylinder, ...)
//   Contact / strategy    reference src/contact.rs:12-69        ->  bam3d::Contact, bam3d::CollisionStrategy
//   GJK                   reference src/algorithm/minkowski/gjk/mod.rs:32-524
//   SweepAndPrune3        reference src/algorithm/broad_phase/sweep_prune.rs:27-128
//   DBVT queries          reference src/dbvt/mod.rs:451-515, dbvt/visitor.rs, dbvt/util.rs
// The reference is Rust; no Rust toolchain exists in this build image, so this C++ layer (the language the image does
// have) is the compiled host side, and rust/ holds the source of the equivalent Rust shim.
// There is no CPU fallback: every call runs on the GPU or throws bam3d::Error.
#pragma once
#include <array>
#include <cstdint>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "bam3d_gpu.h"

namespace bam3d {

struct Error : std::runtime_error {
    int32_t code;
    Error(int32_t c, const std::string& m) : std::runtime_error(m), code(c) {}
};

using Vec3 = std::array<float, 3>;
using Mat4 = std::array<float, 16>;  // glam Mat4::to_cols_array (column-major)

inline Mat4 mat4_from_translation(float x, float y, float z) { return {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, x, y, z, 1}; }

enum class CollisionStrategy : int32_t { FullResolution = BAM3D_FULL_RESOLUTION, CollisionOnly = BAM3D_COLLISION_ONLY };

struct Contact {  // contact.rs:22-38
    CollisionStrategy strategy;
    Vec3 normal;
    float penetration_depth;
    Vec3 contact_point;
    float time_of_impact;
};

// enum Primitive3 (primitive/primitive3.rs:16-33) as a tagged value; constructors follow the reference's `new`.
struct Primitive3 {
    uint8_t kind = BAM3D_KIND_PARTICLE;
    std::array<float, 4> params{0, 0, 0, 0};
    std::vector<float> vertices;  // ConvexPolyhedron only (xyz triples)
    static Primitive3 Particle() { return {}; }
    static Primitive3 Sphere(float radius) { Primitive3 p; p.kind = BAM3D_KIND_SPHERE; p.params = {radius, 0, 0, 0}; return p; }
    static Primitive3 Cuboid(float dx, float dy, float dz) { Primitive3 p; p.kind = BAM3D_KIND_CUBOID; p.params = {dx / 2.f, dy / 2.f, dz / 2.f, 0}; return p; }
    static Primitive3 Cube(float d) { return Cuboid(d, d, d); }
    static Primitive3 Quad(float dx, float dy) { Primitive3 p; p.kind = BAM3D_KIND_QUAD; p.params = {dx / 2.f, dy / 2.f, 0, 0}; return p; }
    static Primitive3 Cylinder(float half_height, float radius) { Primitive3 p; p.kind = BAM3D_KIND_CYLINDER; p.params = {half_height, radius, 0, 0}; return p; }
    static Primitive3 Capsule(float half_height, float radius) { Primitive3 p; p.kind = BAM3D_KIND_CAPSULE; p.params = {half_height, radius, 0, 0}; return p; }
    static Primitive3 ConvexPolyhedron(std::vector<float> xyz) { Primitive3 p; p.kind = BAM3D_KIND_POLYHEDRON; p.vertices = std::move(xyz); return p; }
};

class Context {
public:
    explicit Context(int device = 0) {
        int32_t rc = bam3d_ctx_create(device, &h_);
        if (rc != BAM3D_OK) throw Error(rc, "bam3d_ctx_create failed: no usable CUDA device (there is no CPU fallback)");
    }
    ~Context() { bam3d_ctx_destroy(h_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    bam3d_ctx* handle() const { return h_; }
    void check(int32_t rc) const { if (rc != BAM3D_OK) throw Error(rc, bam3d_last_error(h_)); }
private:
    bam3d_ctx* h_ = nullptr;
};

// Primitives + transforms packed into device buffers.
class ShapeSet {
public:
    ShapeSet(const Context& ctx, const std::vector<Primitive3>& prims, const std::vector<Mat4>& xforms) : ctx_(ctx) {
        const uint32_t n = (uint32_t)prims.size();
        std::vector<uint8_t> kind(n);
        std::vector<float> params(4 * (size_t)n), xf(16 * (size_t)n), verts;
        std::vector<uint32_t> mesh_id(n, 0), offsets{0};
        for (uint32_t i = 0; i < n; ++i) {
            kind[i] = prims[i].kind;
            for (int k = 0; k < 4; ++k) params[4 * i + k] = prims[i].params[k];
            for (int k = 0; k < 16; ++k) xf[16 * i + k] = xforms[i][k];
            if (prims[i].kind == BAM3D_KIND_POLYHEDRON) {
                mesh_id[i] = (uint32_t)offsets.size() - 1;
                verts.insert(verts.end(), prims[i].vertices.begin(), prims[i].vertices.end());
                offsets.push_back((uint32_t)(verts.size() / 3));
            }
        }
        if (offsets.size() > 1) ctx_.check(bam3d_meshes_upload(ctx_.handle(), verts.data(), offsets.data(), (uint32_t)offsets.size() - 1, &meshes_));
        ctx_.check(bam3d_shapes_upload(ctx_.handle(), kind.data(), params.data(), xf.data(), mesh_id.data(), n, meshes_, &h_));
    }
    ~ShapeSet() { bam3d_shapes_free(ctx_.handle(), h_); bam3d_meshes_free(ctx_.handle(), meshes_); }
    ShapeSet(const ShapeSet&) = delete;
    ShapeSet& operator=(const ShapeSet&) = delete;
    bam3d_shapes* handle() const { return h_; }
    void set_transforms(const std::vector<Mat4>& xforms) { ctx_.check(bam3d_shapes_set_transforms(ctx_.handle(), h_, xforms.data()->data())); }
    /// The same from the 48-byte layout (the four columns' xyz; see affine3x4): a quarter less to upload per frame.
    void set_transforms(const std::vector<std::array<float, 12>>& xforms) { ctx_.check(bam3d_shapes_set_transforms_affine(ctx_.handle(), h_, xforms.data()->data())); }
private:
    const Context& ctx_;
    bam3d_shapes* h_ = nullptr;
    bam3d_meshes* meshes_ = nullptr;
};

using Pair = std::pair<uint32_t, uint32_t>;

// Gilbert-Johnson-Keerthi narrow phase (gjk/mod.rs:24-58).
class GJK {
public:
    explicit GJK(const Context& ctx) : ctx_(ctx) { bam3d_default_settings(&s_); }                     // GJK::new()
    GJK(const Context& ctx, float distance_tolerance, float continuous_tolerance, float epa_tolerance,  // new_with_settings
        uint32_t max_iterations) : ctx_(ctx) {
        bam3d_default_settings(&s_);
        s_.distance_tolerance = distance_tolerance; s_.continuous_tolerance = continuous_tolerance;
        s_.epa_tolerance = epa_tolerance; s_.max_iterations = max_iterations;
    }

    // ---- batched additions ----
    std::vector<uint8_t> intersect_batch(const ShapeSet& shapes, const std::vector<Pair>& pairs) const {
        std::vector<uint8_t> st(pairs.size());
        ctx_.check(bam3d_gjk_intersect(ctx_.handle(), shapes.handle(), flat(pairs), pairs.size(), &s_, st.data()));
        return st;
    }
    std::vector<std::optional<Contact>> intersection_batch(CollisionStrategy strategy, const ShapeSet& shapes, const std::vector<Pair>& pairs) const {
        std::vector<uint8_t> st(pairs.size());
        std::vector<bam3d_contact> c(pairs.size());
        ctx_.check(bam3d_gjk_contact(ctx_.handle(), shapes.handle(), flat(pairs), pairs.size(), &s_, (int32_t)strategy, c.data(), st.data()));
        return to_optional(strategy, c, st);
    }
    // Some(x): x.first = what the reference returns (the residual dp - d0), x.second = the geometric distance
    std::vector<std::optional<std::pair<float, float>>> distance_batch(const ShapeSet& shapes, const std::vector<Pair>& pairs) const {
        std::vector<uint8_t> st(pairs.size());
        std::vector<float> ref(pairs.size()), dist(pairs.size());
        ctx_.check(bam3d_gjk_distance(ctx_.handle(), shapes.handle(), flat(pairs), pairs.size(), &s_, ref.data(), dist.data(), st.data()));
        std::vector<std::optional<std::pair<float, float>>> out(pairs.size());
        for (size_t i = 0; i < pairs.size(); ++i) if (st[i] == BAM3D_STATUS_OK) out[i] = std::make_pair(ref[i], dist[i]);
        return out;
    }
    std::vector<std::optional<Contact>> intersection_time_of_impact_batch(const ShapeSet& start, const ShapeSet& end, const std::vector<Pair>& pairs) const {
        std::vector<uint8_t> st(pairs.size());
        std::vector<bam3d_contact> c(pairs.size());
        ctx_.check(bam3d_gjk_time_of_impact(ctx_.handle(), start.handle(), end.handle(), flat(pairs), pairs.size(), &s_, c.data(), st.data()));
        return to_optional(CollisionStrategy::FullResolution, c, st);
    }

    // ---- the reference's scalar methods (batch of one) ----
    bool intersect(const Primitive3& left, const Mat4& lt, const Primitive3& right, const Mat4& rt) const {  // gjk/mod.rs:74
        ShapeSet s(ctx_, {left, right}, {lt, rt});
        return intersect_batch(s, {{0, 1}})[0] == BAM3D_STATUS_OK;
    }
    std::optional<Contact> intersection(CollisionStrategy strategy, const Primitive3& left, const Mat4& lt, const Primitive3& right, const Mat4& rt) const {  // :317
        ShapeSet s(ctx_, {left, right}, {lt, rt});
        return intersection_batch(strategy, s, {{0, 1}})[0];
    }
    std::optional<float> distance(const Primitive3& left, const Mat4& lt, const Primitive3& right, const Mat4& rt) const {  // :232
        ShapeSet s(ctx_, {left, right}, {lt, rt});
        auto d = distance_batch(s, {{0, 1}})[0];
        return d ? std::optional<float>(d->first) : std::nullopt;
    }
    std::optional<Contact> intersection_time_of_impact(const Primitive3& left, const Mat4& lt_start, const Mat4& lt_end, const Primitive3& right,
                                                       const Mat4& rt_start, const Mat4& rt_end) const {  // :130
        ShapeSet s0(ctx_, {left, right}, {lt_start, rt_start}), s1(ctx_, {left, right}, {lt_end, rt_end});
        return intersection_time_of_impact_batch(s0, s1, {{0, 1}})[0];
    }

    const bam3d_gjk_settings& settings() const { return s_; }

private:
    static const uint32_t* flat(const std::vector<Pair>& p) { return reinterpret_cast<const uint32_t*>(p.data()); }
    static std::vector<std::optional<Contact>> to_optional(CollisionStrategy strategy, const std::vector<bam3d_contact>& c, const std::vector<uint8_t>& st) {
        std::vector<std::optional<Contact>> out(c.size());
        for (size_t i = 0; i < c.size(); ++i)
            if (st[i] == BAM3D_STATUS_OK)
                out[i] = Contact{strategy, {c[i].normal[0], c[i].normal[1], c[i].normal[2]}, c[i].penetration_depth,
                                 {c[i].contact_point[0], c[i].contact_point[1], c[i].contact_point[2]}, c[i].time_of_impact};
        return out;
    }
    const Context& ctx_;
    bam3d_gjk_settings s_;
};

// The per-frame sequence with several frames in flight (bam3d_frames_*): new Mat4 for every shape, the narrow phase over
// the frame's pairs, results in host memory — uploads and downloads run beside the kernels of the neighbouring frames.
// The buffers given to submit_* must stay alive and untouched until wait(frame) returns.
class FrameQueue {
public:
    FrameQueue(const Context& ctx, const std::vector<Primitive3>& prims, const std::vector<Mat4>& xforms, uint64_t max_pairs,
               uint32_t depth = 2, bool with_end = false) : ctx_(ctx) {
        const uint32_t n = (uint32_t)prims.size();
        std::vector<uint8_t> kind(n);
        std::vector<float> params(4 * (size_t)n), verts;
        std::vector<uint32_t> mesh_id(n, 0), offsets{0};
        for (uint32_t i = 0; i < n; ++i) {
            kind[i] = prims[i].kind;
            for (int k = 0; k < 4; ++k) params[4 * i + k] = prims[i].params[k];
            if (prims[i].kind == BAM3D_KIND_POLYHEDRON) {
                mesh_id[i] = (uint32_t)offsets.size() - 1;
                verts.insert(verts.end(), prims[i].vertices.begin(), prims[i].vertices.end());
                offsets.push_back((uint32_t)(verts.size() / 3));
            }
        }
        if (offsets.size() > 1) ctx_.check(bam3d_meshes_upload(ctx_.handle(), verts.data(), offsets.data(), (uint32_t)offsets.size() - 1, &meshes_));
        ctx_.check(bam3d_frames_create(ctx_.handle(), kind.data(), params.data(), xforms.data()->data(), mesh_id.data(), n, meshes_, max_pairs,
                                       depth, with_end ? 1 : 0, &h_));
    }
    ~FrameQueue() { bam3d_frames_free(ctx_.handle(), h_); bam3d_meshes_free(ctx_.handle(), meshes_); }
    FrameQueue(const FrameQueue&) = delete;
    FrameQueue& operator=(const FrameQueue&) = delete;
    uint64_t submit_intersect(const GJK& gjk, const std::vector<Mat4>& xforms, const std::vector<Pair>& pairs, uint8_t* out_status) {
        return submit(BAM3D_FRAME_INTERSECT, gjk, xforms.data()->data(), nullptr, pairs, CollisionStrategy::FullResolution, nullptr, out_status);
    }
    uint64_t submit_intersection(const GJK& gjk, CollisionStrategy strategy, const std::vector<Mat4>& xforms, const std::vector<Pair>& pairs,
                                 bam3d_contact* out_contacts, uint8_t* out_status) {
        return submit(BAM3D_FRAME_CONTACT, gjk, xforms.data()->data(), nullptr, pairs, strategy, out_contacts, out_status);
    }
    uint64_t submit_intersection_time_of_impact(const GJK& gjk, const std::vector<Mat4>& start, const std::vector<Mat4>& end,
                                                const std::vector<Pair>& pairs, bam3d_contact* out_contacts, uint8_t* out_status) {
        return submit(BAM3D_FRAME_TIME_OF_IMPACT, gjk, start.data()->data(), end.data()->data(), pairs, CollisionStrategy::FullResolution, out_contacts, out_status);
    }
    void wait(uint64_t frame) { ctx_.check(bam3d_frames_wait(ctx_.handle(), h_, frame)); }
private:
    uint64_t submit(int32_t query, const GJK& gjk, const float* x0, const float* x1, const std::vector<Pair>& pairs, CollisionStrategy strategy,
                    bam3d_contact* out_contacts, uint8_t* out_status) {
        uint64_t frame = 0;
        ctx_.check(bam3d_frames_submit(ctx_.handle(), h_, query, x0, x1, reinterpret_cast<const uint32_t*>(pairs.data()), pairs.size(),
                                       &gjk.settings(), (int32_t)strategy, out_contacts, out_status, &frame));
        return frame;
    }
    const Context& ctx_;
    bam3d_frames* h_ = nullptr;
    bam3d_meshes* meshes_ = nullptr;
};

struct Aabb3 { Vec3 min, max; };  // volume/aabb/aabb3.rs:15-21 (6 contiguous f32)

// SweepAndPrune3 (sweep_prune.rs:27-128): keeps sweep_axis across calls; re-sorts the caller's slice like the reference.
class SweepAndPrune3 {
public:
    explicit SweepAndPrune3(const Context& ctx, int sweep_axis = 0) : ctx_(ctx), axis_(sweep_axis) {}  // new / with_sweep_axis
    int get_sweep_axis() const { return axis_; }
    std::vector<std::pair<size_t, size_t>> find_collider_pairs(std::vector<Aabb3>& shapes) {
        const uint64_t n = shapes.size();
        std::vector<uint32_t> perm(n), pairs;
        uint64_t cap = 8 * n + 1024, count = 0;
        for (;;) {
            pairs.resize(2 * cap);
            int32_t axis = axis_;
            int32_t rc = bam3d_sap_find_pairs(ctx_.handle(), reinterpret_cast<const float*>(shapes.data()), n, &axis, perm.data(), pairs.data(), cap, &count);
            if (rc == BAM3D_ERR_CAPACITY && count > cap) { cap = count; continue; }
            ctx_.check(rc);
            axis_ = axis;
            break;
        }
        std::vector<Aabb3> sorted(n);
        for (uint64_t k = 0; k < n; ++k) sorted[k] = shapes[perm[k]];
        shapes.swap(sorted);  // "The shapes list might have been resorted" (sweep_prune.rs:65-68)
        std::vector<std::pair<size_t, size_t>> out(count);
        for (uint64_t k = 0; k < count; ++k) out[k] = {pairs[2 * k], pairs[2 * k + 1]};
        return out;
    }
private:
    const Context& ctx_;
    int axis_;
};

struct Sphere { Vec3 center; float radius; };  // volume/sphere.rs:8-15 (the bounding sphere; 4 contiguous f32)

// ComputeBound<volume::Sphere> + Bound::transform_volume for every shape of a set (primitive3.rs:98-114, volume/sphere.rs:34-39)
inline std::vector<Sphere> world_spheres(const Context& ctx, const ShapeSet& shapes) {
    std::vector<Sphere> out(bam3d_shapes_count(shapes.handle()));
    ctx.check(bam3d_shapes_world_spheres(ctx.handle(), shapes.handle(), reinterpret_cast<float*>(out.data())));
    return out;
}

// SweepAndPrune3<volume::Sphere>: same interface as SweepAndPrune3; touching spheres are reported (Sphere::intersects is <=).
class SweepAndPrune3Sphere {
public:
    explicit SweepAndPrune3Sphere(const Context& ctx, int sweep_axis = 0) : ctx_(ctx), axis_(sweep_axis) {}
    int get_sweep_axis() const { return axis_; }
    std::vector<std::pair<size_t, size_t>> find_collider_pairs(std::vector<Sphere>& shapes) {
        const uint64_t n = shapes.size();
        std::vector<uint32_t> perm(n), pairs;
        uint64_t cap = 8 * n + 1024, count = 0;
        for (;;) {
            pairs.resize(2 * cap);
            int32_t axis = axis_;
            int32_t rc = bam3d_sap_find_pairs_spheres(ctx_.handle(), reinterpret_cast<const float*>(shapes.data()), n, &axis, perm.data(), pairs.data(), cap, &count);
            if (rc == BAM3D_ERR_CAPACITY && count > cap) { cap = count; continue; }
            ctx_.check(rc);
            axis_ = axis;
            break;
        }
        std::vector<Sphere> sorted(n);
        for (uint64_t k = 0; k < n; ++k) sorted[k] = shapes[perm[k]];
        shapes.swap(sorted);
        std::vector<std::pair<size_t, size_t>> out(count);
        for (uint64_t k = 0; k < count; ++k) out[k] = {pairs[2 * k], pairs[2 * k + 1]};
        return out;
    }
private:
    const Context& ctx_;
    int axis_;
};

// ---- DBVT queries (dbvt/mod.rs:451-515, visitor.rs, util.rs) ---------------------------------------------------------------------
struct Ray { Vec3 origin, direction; };            // ray.rs (6 contiguous f32)
struct Plane { Vec3 n; float d; };                 // plane.rs:17-22: n . x = d (4 contiguous f32)
enum class Relation : uint8_t { In = 0, Cross = 1, Out = 2 };  // bound.rs:9-17
struct Frustum {                                   // frustum.rs:17-24 (24 contiguous f32 in this order)
    Plane left, right, bottom, top, near, far;
    /// Frustum::from_mat4 (frustum.rs:46-75, including its near / far = left / right rows); nullopt where the reference returns None.
    static std::optional<Frustum> from_mat4(const Mat4& m) {
        Frustum f;
        if (bam3d_frustum_from_mat4(m.data(), reinterpret_cast<float*>(&f)) != BAM3D_OK) return std::nullopt;
        return f;
    }
};
static_assert(sizeof(Frustum) == 96 && sizeof(Ray) == 24, "query layouts are the C ABI's");

/// The query side of DynamicBoundingVolumeTree over a static set of values (value index = position in `bounds`), bounds being
/// Aabb3 or volume::Sphere. Result sets equal the reference's (they do not depend on its randomised tree shape).
class DynamicBoundingVolumeTree {
public:
    DynamicBoundingVolumeTree(const Context& ctx, const std::vector<Aabb3>& bounds) : ctx_(ctx), n_(bounds.size()) {
        ctx_.check(bam3d_bvh_build(ctx_.handle(), reinterpret_cast<const float*>(bounds.data()), n_, &h_));
    }
    DynamicBoundingVolumeTree(const Context& ctx, const std::vector<Sphere>& bounds) : ctx_(ctx), n_(bounds.size()) {
        ctx_.check(bam3d_bvh_build_spheres(ctx_.handle(), reinterpret_cast<const float*>(bounds.data()), n_, &h_));
    }
    ~DynamicBoundingVolumeTree() { bam3d_bvh_free(ctx_.handle(), h_); }
    DynamicBoundingVolumeTree(const DynamicBoundingVolumeTree&) = delete;
    DynamicBoundingVolumeTree& operator=(const DynamicBoundingVolumeTree&) = delete;
    /// update_node + do_refit for every value at once (dbvt/mod.rs:561-589): new bounds, same topology (box trees).
    void refit(const std::vector<Aabb3>& bounds) { ctx_.check(bam3d_bvh_refit(ctx_.handle(), h_, reinterpret_cast<const float*>(bounds.data()))); }
    /// query(&mut DiscreteVisitor::new(&bound)) (visitor.rs:57-90): the values whose bound intersects `bound`.
    std::vector<uint32_t> query_discrete(const Aabb3& bound) const {
        return var([&](uint64_t* off, uint32_t* v, uint64_t cap, uint64_t* total) {
            return bam3d_bvh_query_aabb(ctx_.handle(), h_, reinterpret_cast<const float*>(&bound), 1, off, v, cap, total); });
    }
    std::vector<uint32_t> query_discrete(const Sphere& bound) const {
        return var([&](uint64_t* off, uint32_t* v, uint64_t cap, uint64_t* total) {
            return bam3d_bvh_query_sphere(ctx_.handle(), h_, reinterpret_cast<const float*>(&bound), 1, off, v, cap, total); });
    }
    /// query_ray (util.rs:114-129): (value, hit point) of every value the ray meets.
    std::vector<std::pair<uint32_t, Vec3>> query_ray(const Ray& ray) const {
        std::vector<Vec3> pts;
        auto vals = var([&](uint64_t* off, uint32_t* v, uint64_t cap, uint64_t* total) {
            pts.resize(cap);
            return bam3d_bvh_query_ray(ctx_.handle(), h_, reinterpret_cast<const float*>(&ray), 1, off, v, reinterpret_cast<float*>(pts.data()), cap, total); });
        std::vector<std::pair<uint32_t, Vec3>> out(vals.size());
        for (size_t k = 0; k < vals.size(); ++k) out[k] = {vals[k], pts[k]};
        return out;
    }
    /// query_ray_closest (util.rs:77-103): the hit of smallest t over all values the ray meets (box trees).
    std::optional<std::pair<uint32_t, Vec3>> query_ray_closest(const Ray& ray) const {
        uint32_t v = 0; float pt[4];
        ctx_.check(bam3d_bvh_query_ray_closest(ctx_.handle(), h_, reinterpret_cast<const float*>(&ray), 1, &v, pt));
        if (v == 0xFFFFFFFFu) return std::nullopt;
        return std::make_pair(v, Vec3{pt[0], pt[1], pt[2]});
    }
    /// query(&mut FrustumVisitor::new(&frustum)) (visitor.rs:99-134): (value, Relation) — Out is never reported.
    std::vector<std::pair<uint32_t, Relation>> query_frustum(const Frustum& f) const {
        std::vector<uint8_t> rel;
        auto vals = var([&](uint64_t* off, uint32_t* v, uint64_t cap, uint64_t* total) {
            rel.resize(cap);
            return bam3d_bvh_query_frustum(ctx_.handle(), h_, reinterpret_cast<const float*>(&f), 1, off, v, rel.data(), cap, total); });
        std::vector<std::pair<uint32_t, Relation>> out(vals.size());
        for (size_t k = 0; k < vals.size(); ++k) out[k] = {vals[k], static_cast<Relation>(rel[k])};
        return out;
    }
    /// DbvtBroadPhase::find_collider_pairs (algorithm/broad_phase/dbvt.rs:26-63): sorted unique (low, high) value pairs.
    std::vector<std::pair<uint32_t, uint32_t>> find_collider_pairs(const std::vector<uint8_t>& dirty) const {
        uint64_t cap = 8 * n_ + 1024, count = 0;
        std::vector<uint32_t> pairs;
        for (;;) {
            pairs.resize(2 * cap);
            int32_t rc = bam3d_bvh_find_collider_pairs(ctx_.handle(), h_, dirty.data(), pairs.data(), cap, &count);
            if (rc == BAM3D_ERR_CAPACITY && count > cap) { cap = count; continue; }
            ctx_.check(rc);
            break;
        }
        std::vector<std::pair<uint32_t, uint32_t>> out(count);
        for (uint64_t k = 0; k < count; ++k) out[k] = {pairs[2 * k], pairs[2 * k + 1]};
        return out;
    }
private:
    template <class F> std::vector<uint32_t> var(F call) const {  // one query, retried with the reported size
        uint64_t cap = 1024, total = 0, off[2] = {0, 0};
        std::vector<uint32_t> vals;
        for (;;) {
            vals.resize(cap);
            int32_t rc = call(off, vals.data(), cap, &total);
            if (rc == BAM3D_ERR_CAPACITY && total > cap) { cap = total; continue; }
            ctx_.check(rc);
            break;
        }
        vals.resize(total);
        return vals;
    }
    const Context& ctx_;
    uint64_t n_;
    bam3d_bvh* h_ = nullptr;
};

/// ComputeBound<Aabb3> + Aabb::transform of every shape of the set (primitive3.rs:77-96, aabb3.rs:89-96).
inline std::vector<Aabb3> world_aabbs(const Context& ctx, const ShapeSet& shapes) {
    std::vector<Aabb3> out(bam3d_shapes_count(shapes.handle()));
    ctx.check(bam3d_shapes_world_aabbs(ctx.handle(), shapes.handle(), reinterpret_cast<float*>(out.data())));
    return out;
}

// Compound shapes: the reference's `&[(Primitive, Mat4)]` slices (gjk/mod.rs:362-371), one per compound, resident on the device,
// and GJK::intersection_complex / distance_complex / intersection_complex_time_of_impact (gjk/mod.rs:362-523) over pairs of them.
class CompoundSet {
public:
    using Part = std::pair<Primitive3, Mat4>;  // (primitive, local-to-model transform)
    CompoundSet(const Context& ctx, const std::vector<std::vector<Part>>& compounds) : ctx_(ctx), n_((uint32_t)compounds.size()) {
        std::vector<uint8_t> kind;
        std::vector<float> params, local, verts;
        std::vector<uint32_t> mesh_id, voff{0}, offsets{0};
        for (const auto& c : compounds) {
            for (const auto& part : c) {
                kind.push_back(part.first.kind);
                params.insert(params.end(), part.first.params.begin(), part.first.params.end());
                local.insert(local.end(), part.second.begin(), part.second.end());
                mesh_id.push_back(0);
                if (part.first.kind == BAM3D_KIND_POLYHEDRON) {
                    mesh_id.back() = (uint32_t)voff.size() - 1;
                    verts.insert(verts.end(), part.first.vertices.begin(), part.first.vertices.end());
                    voff.push_back((uint32_t)(verts.size() / 3));
                }
            }
            offsets.push_back((uint32_t)kind.size());
        }
        if (voff.size() > 1) ctx_.check(bam3d_meshes_upload(ctx_.handle(), verts.data(), voff.data(), (uint32_t)voff.size() - 1, &meshes_));
        ctx_.check(bam3d_compounds_upload(ctx_.handle(), kind.data(), params.data(), local.data(), mesh_id.data(), (uint32_t)kind.size(), offsets.data(), n_, meshes_, &h_));
    }
    ~CompoundSet() { bam3d_compounds_free(ctx_.handle(), h_); bam3d_meshes_free(ctx_.handle(), meshes_); }
    CompoundSet(const CompoundSet&) = delete;
    CompoundSet& operator=(const CompoundSet&) = delete;
    uint32_t size() const { return n_; }
    // transforms[c] = model-to-world Mat4 of compound c
    std::vector<std::optional<Contact>> intersection_complex_batch(const GJK& gjk, CollisionStrategy strategy, const std::vector<Mat4>& transforms,
                                                                   const std::vector<Pair>& pairs) const {
        std::vector<uint8_t> st(pairs.size());
        std::vector<bam3d_contact> c(pairs.size());
        ctx_.check(bam3d_gjk_contact_complex(ctx_.handle(), h_, transforms.data()->data(), reinterpret_cast<const uint32_t*>(pairs.data()), pairs.size(), &gjk.settings(),
                                             (int32_t)strategy, c.data(), st.data()));
        return opt(strategy, c, st);
    }
    std::vector<std::optional<float>> distance_complex_batch(const GJK& gjk, const std::vector<Mat4>& transforms, const std::vector<Pair>& pairs) const {
        std::vector<uint8_t> st(pairs.size());
        std::vector<float> v(pairs.size());
        ctx_.check(bam3d_gjk_distance_complex(ctx_.handle(), h_, transforms.data()->data(), reinterpret_cast<const uint32_t*>(pairs.data()), pairs.size(), &gjk.settings(),
                                              v.data(), st.data()));
        std::vector<std::optional<float>> out(pairs.size());
        for (size_t i = 0; i < pairs.size(); ++i) if (st[i] == BAM3D_STATUS_OK) out[i] = v[i];
        return out;
    }
    std::vector<std::optional<Contact>> intersection_complex_time_of_impact_batch(const GJK& gjk, CollisionStrategy strategy, const std::vector<Mat4>& start,
                                                                                  const std::vector<Mat4>& end, const std::vector<Pair>& pairs) const {
        std::vector<uint8_t> st(pairs.size());
        std::vector<bam3d_contact> c(pairs.size());
        ctx_.check(bam3d_gjk_time_of_impact_complex(ctx_.handle(), h_, start.data()->data(), end.data()->data(), reinterpret_cast<const uint32_t*>(pairs.data()),
                                                    pairs.size(), &gjk.settings(), (int32_t)strategy, c.data(), st.data()));
        return opt(strategy, c, st);
    }
private:
    static std::vector<std::optional<Contact>> opt(CollisionStrategy strategy, const std::vector<bam3d_contact>& c, const std::vector<uint8_t>& st) {
        std::vector<std::optional<Contact>> out(c.size());
        for (size_t i = 0; i < c.size(); ++i)
            if (st[i] == BAM3D_STATUS_OK)
                out[i] = Contact{strategy, {c[i].normal[0], c[i].normal[1], c[i].normal[2]}, c[i].penetration_depth,
                                 {c[i].contact_point[0], c[i].contact_point[1], c[i].contact_point[2]}, c[i].time_of_impact};
        return out;
    }
    const Context& ctx_;
    bam3d_compounds* h_ = nullptr;
    bam3d_meshes* meshes_ = nullptr;
    uint32_t n_;
};

// The scalar forms of the reference (gjk/mod.rs:362, :422, :478), as batches of one compound pair.
inline std::optional<Contact> intersection_complex(const Context& ctx, const GJK& gjk, CollisionStrategy strategy, const std::vector<CompoundSet::Part>& left,
                                                   const Mat4& left_transform, const std::vector<CompoundSet::Part>& right, const Mat4& right_transform) {
    CompoundSet c(ctx, {left, right});
    return c.intersection_complex_batch(gjk, strategy, {left_transform, right_transform}, {{0, 1}})[0];
}
inline std::optional<float> distance_complex(const Context& ctx, const GJK& gjk, const std::vector<CompoundSet::Part>& left, const Mat4& left_transform,
                                             const std::vector<CompoundSet::Part>& right, const Mat4& right_transform) {
    CompoundSet c(ctx, {left, right});
    return c.distance_complex_batch(gjk, {left_transform, right_transform}, {{0, 1}})[0];
}

// Broad phase -> narrow phase in one resident call (BASELINE configs[3]): world AABBs -> SweepAndPrune3 -> perm remap -> GJK::intersection.
struct Candidate { uint32_t left, right; std::optional<Contact> contact; };  // SHAPE indices, not sorted positions
inline std::vector<Candidate> collide_all(const Context& ctx, const GJK& gjk, CollisionStrategy strategy, const ShapeSet& shapes, int& sweep_axis, uint64_t capacity_guess = 0) {
    uint64_t cap = capacity_guess ? capacity_guess : 4096, count = 0;
    std::vector<uint32_t> pairs;
    std::vector<bam3d_contact> c;
    std::vector<uint8_t> st;
    for (;;) {
        pairs.resize(2 * cap); c.resize(cap); st.resize(cap);
        int32_t axis = sweep_axis;
        int32_t rc = bam3d_pipeline_contacts(ctx.handle(), shapes.handle(), &axis, &gjk.settings(), (int32_t)strategy, pairs.data(), nullptr, nullptr, c.data(), st.data(), cap, &count);
        if (rc == BAM3D_ERR_CAPACITY && count > cap) { cap = count; continue; }
        ctx.check(rc);
        sweep_axis = axis;
        break;
    }
    std::vector<Candidate> out(count);
    for (uint64_t k = 0; k < count; ++k) {
        out[k].left = pairs[2 * k]; out[k].right = pairs[2 * k + 1];
        if (st[k] == BAM3D_STATUS_OK)
            out[k].contact = Contact{strategy, {c[k].normal[0], c[k].normal[1], c[k].normal[2]}, c[k].penetration_depth,
                                     {c[k].contact_point[0], c[k].contact_point[1], c[k].contact_point[2]}, c[k].time_of_impact};
    }
    return out;
}

// Multi-GPU (one process per GPU): the NCCL gather of the ranks' contacts (bam3d_comm_*). The 256-byte id is made by rank 0
// (unique_id) and handed to the other ranks by the application's own channel.
class Comm {
public:
    static std::array<uint8_t, BAM3D_UNIQUE_ID_BYTES> unique_id() {
        std::array<uint8_t, BAM3D_UNIQUE_ID_BYTES> id{};
        int32_t rc = bam3d_comm_unique_id(id.data());
        if (rc != BAM3D_OK) throw Error(rc, std::string("bam3d_comm_unique_id: ") + bam3d_comm_unavailable_reason());
        return id;
    }
    Comm(int device, const std::array<uint8_t, BAM3D_UNIQUE_ID_BYTES>& id, int rank, int n_ranks) : n_ranks_(n_ranks) {
        int32_t rc = bam3d_comm_create(device, id.data(), rank, n_ranks, &h_);
        if (rc != BAM3D_OK) throw Error(rc, std::string("bam3d_comm_create: ") + bam3d_comm_unavailable_reason());
    }
    ~Comm() { bam3d_comm_destroy(h_); }
    Comm(const Comm&) = delete;
    Comm& operator=(const Comm&) = delete;
    bam3d_comm* handle() const { return h_; }
    // this rank's dense results (for its slice of the pair list, global indices from pair_index_base) -> the hits of ALL ranks, rank order
    std::vector<bam3d_contact40> gather_contacts(const Context& ctx, const std::vector<bam3d_contact>& contacts, const std::vector<uint8_t>& status,
                                                 uint64_t pair_index_base, uint64_t capacity, std::vector<uint64_t>* counts = nullptr) const {
        std::vector<bam3d_contact40> out(capacity ? capacity : 1);
        std::vector<uint64_t> cnt((size_t)n_ranks_);
        uint64_t total = 0;
        ctx.check(bam3d_gather_contacts(ctx.handle(), h_, contacts.data(), status.data(), status.size(), pair_index_base, out.data(), capacity, cnt.data(), &total));
        out.resize(total);
        if (counts) *counts = cnt;
        return out;
    }
private:
    bam3d_comm* h_ = nullptr;
    int n_ranks_;
};

}  // namespace bam3d
