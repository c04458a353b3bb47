// ORACLE — TEST INFRASTRUCTURE ONLY (see glam_mini.hpp header). CPU restatement of the bam3d broad phase:
// Aabb3, Ray, Plane, Frustum, Relation, SweepAndPrune3 + Variance3, and the DBVT query path with its
// visitors. Paths relative to /root/reference/src.
//
// Pinning: Aabb3 / ray slab / plane relation / frustum are pinned by the reference's integration tests
// (tests/aabb.rs, tests/dbvt.rs, tests/frustum.rs, tests/plane.rs), ported in oracle/kat_main.cpp.
// PARITY UNPINNED (the reference has no test at all): SweepAndPrune / Variance3, DbvtBroadPhase,
// DiscreteVisitor, ContinuousVisitor / query_ray / query_ray_closest.
//
// DBVT: insert (dbvt/mod.rs:622-731) and the bound refit are restated; tree rotations are NOT (they are gated
// by rand::thread_rng(), dbvt/mod.rs:901, so the reference's tree shape is non-deterministic). Query result
// SETS do not depend on the tree shape (every branch bound contains its descendants), so parity on DBVT
// queries is defined on results sorted by value index.
#pragma once
#include <algorithm>
#include <cstdint>
#include <utility>
#include <vector>
#include "glam_mini.hpp"

namespace orc {

// ---------------------------------------------------------------------------------------------------------
// ray.rs:8-31
// ---------------------------------------------------------------------------------------------------------
struct Ray {
    Vec3 origin, direction;
    Ray() {}
    Ray(Vec3 o, Vec3 d) : origin(o), direction(d) {}
    Ray transform(const Mat4& t) const { return Ray(t.transform_point3(origin), t.transform_vector3(direction)); }
};

// ---------------------------------------------------------------------------------------------------------
// plane.rs
// ---------------------------------------------------------------------------------------------------------
struct Plane {
    Vec3 n;
    float d = 0.f;
    Plane() {}
    Plane(Vec3 n_, float d_) : n(n_), d(d_) {}
    static Plane from_vector4_alt(Vec4 v) { return Plane(Vec3(v.x, v.y, v.z), -v.w); }  // plane.rs:61-66
    static Plane from_point_normal(Vec3 p, Vec3 n) { return Plane(n, dot(p, n)); }        // plane.rs:90-92
    // plane.rs:95-103
    bool normalize(Plane& out) const {
        if (all_eq(n, Vec3::zero())) return false;
        float denom = 1.0f / std::sqrt(dot(n, n));
        out = Plane(n * denom, d * denom);
        return true;
    }
    // plane.rs:69-86 (note the `n.x + d = 0` convention here)
    static bool from_points(Vec3 p1, Vec3 p2, Vec3 p3, Plane& out) {
        Vec3 v0 = p2 - p1, v1 = p3 - p1;
        Vec3 normal = cross(v0, v1);
        if (all_eq(normal, Vec3::zero())) return false;
        Vec3 n = orc::normalize(normal);
        out = Plane(n, -dot(p1, n));
        return true;
    }
    // plane.rs:117-129 Continuous<Ray> for Plane
    bool intersection(const Ray& r, Vec3& out) const {
        float t = -(d + dot(r.origin, n)) / dot(r.direction, n);
        if (t < 0.0f) return false;
        out = r.origin + r.direction * t;
        return true;
    }
    // plane.rs:132-138 Discrete<Ray> for Plane
    bool intersects(const Ray& r) const {
        float t = -(d + dot(r.origin, n)) / dot(r.direction, n);
        return t >= 0.0f;
    }
};

// bound.rs:9-18 — ordering In < Cross < Out
enum Relation : uint8_t { In = 0, Cross = 1, Out = 2 };

// bound.rs:48-58 PlaneBound for Vec3
inline Relation relate_plane(Vec3 p, const Plane& plane) {
    float dist = dot(p, plane.n);
    if (dist > plane.d) return In;
    else if (dist < plane.d) return Out;
    else return Cross;
}

// ---------------------------------------------------------------------------------------------------------
// volume/aabb/aabb3.rs, volume/aabb/mod.rs
// ---------------------------------------------------------------------------------------------------------
struct Aabb3 {
    Vec3 min, max;
    Aabb3() {}
    // aabb3.rs:26-31
    Aabb3(Vec3 p1, Vec3 p2)
        : min(rmin(p1.x, p2.x), rmin(p1.y, p2.y), rmin(p1.z, p2.z)),
          max(rmax(p1.x, p2.x), rmax(p1.y, p2.y), rmax(p1.z, p2.z)) {}
    static Aabb3 zero() { return Aabb3(Vec3::zero(), Vec3::zero()); }
    // aabb3.rs:35-46
    void to_corners(Vec3 c[8]) const {
        c[0] = min;
        c[1] = Vec3(max.x, min.y, min.z); c[2] = Vec3(min.x, max.y, min.z); c[3] = Vec3(max.x, max.y, min.z);
        c[4] = Vec3(min.x, min.y, max.z); c[5] = Vec3(max.x, min.y, max.z); c[6] = Vec3(min.x, max.y, max.z);
        c[7] = max;
    }
    Vec3 dim() const { return max - min; }                               // aabb/mod.rs:48-50
    float volume() const { Vec3 d = dim(); return d.x * d.y * d.z; }     // aabb/mod.rs:54-57
    Vec3 center() const { return min + dim() / 2.0f; }                   // aabb/mod.rs:61-63
    // aabb/mod.rs:66-68
    Aabb3 grow(Vec3 p) const {
        return Aabb3(Vec3(rmin(min.x, p.x), rmin(min.y, p.y), rmin(min.z, p.z)),
                     Vec3(rmax(max.x, p.x), rmax(max.y, p.y), rmax(max.z, p.z)));
    }
    Aabb3 add_v(Vec3 v) const { return Aabb3(min + v, max + v); }        // aabb/mod.rs:72-74
    Aabb3 mul_s(float s) const { return Aabb3(min * s, max * s); }       // aabb/mod.rs:81-83
    Aabb3 mul_v(Vec3 v) const { return Aabb3(min * v, max * v); }        // aabb/mod.rs:86-90
    // aabb3.rs:72-86
    Aabb3 add_margin(Vec3 m) const {
        return Aabb3(Vec3(min.x - m.x, min.y - m.y, min.z - m.z), Vec3(max.x + m.x, max.y + m.y, max.z + m.z));
    }
    // aabb3.rs:89-96
    Aabb3 transform(const Mat4& t) const {
        Vec3 c[8]; to_corners(c);
        Vec3 first = t.transform_point3(c[0]);
        Aabb3 u(first, first);
        for (int i = 1; i < 8; ++i) u = u.grow(t.transform_point3(c[i]));
        return u;
    }
    // aabb3.rs:105-111 Contains<Vec3>
    bool contains(Vec3 p) const {
        return min.x <= p.x && p.x < max.x && min.y <= p.y && p.y < max.y && min.z <= p.z && p.z < max.z;
    }
    // aabb3.rs:113-123 Contains<Aabb3>
    bool contains(const Aabb3& o) const {
        return o.min.x >= min.x && o.min.y >= min.y && o.min.z >= min.z && o.max.x <= max.x && o.max.y <= max.y &&
               o.max.z <= max.z;
    }
    Aabb3 union_(const Aabb3& o) const { return grow(o.min).grow(o.max); }  // aabb3.rs:146-152
    Vec3 min_extent() const { return min; }  // Bound for Aabb3, aabb3.rs:50-57
    Vec3 max_extent() const { return max; }
    // aabb3.rs:258-264
    float surface_area() const {
        Vec3 d = dim();
        return 2.0f * ((d.x * d.y) + (d.x * d.z) + (d.y * d.z));
    }
    // aabb3.rs:237-244 Discrete<Aabb3> for Aabb3 — strict on every axis
    bool intersects(const Aabb3& b) const {
        return max.x > b.min.x && min.x < b.max.x && max.y > b.min.y && min.y < b.max.y && max.z > b.min.z &&
               min.z < b.max.z;
    }
    // aabb3.rs:246-257 PlaneBound for Aabb3
    Relation relate_plane(const Plane& plane) const {
        Vec3 c[8]; to_corners(c);
        Relation first = orc::relate_plane(c[0], plane);
        for (int i = 1; i < 8; ++i)
            if (orc::relate_plane(c[i], plane) != first) return Cross;
        return first;
    }
};

// aabb3.rs:165-195 Continuous<Aabb3> for Ray
inline bool ray_aabb_intersection(const Ray& ray, const Aabb3& aabb, Vec3& out) {
    Vec3 inv_dir = Vec3(1.0f, 1.0f, 1.0f) / ray.direction;
    float t1 = (aabb.min.x - ray.origin.x) * inv_dir.x;
    float t2 = (aabb.max.x - ray.origin.x) * inv_dir.x;
    float tmin = rmin(t1, t2);
    float tmax = rmax(t1, t2);
    t1 = (aabb.min.y - ray.origin.y) * inv_dir.y;
    t2 = (aabb.max.y - ray.origin.y) * inv_dir.y;
    tmin = rmax(tmin, rmin(t1, t2));
    tmax = rmin(tmax, rmax(t1, t2));
    t1 = (aabb.min.z - ray.origin.z) * inv_dir.z;
    t2 = (aabb.max.z - ray.origin.z) * inv_dir.z;
    tmin = rmax(tmin, rmin(t1, t2));
    tmax = rmin(tmax, rmax(t1, t2));
    if ((tmin < 0.0f && tmax < 0.0f) || tmax < tmin) return false;
    float t = (tmin >= 0.0f) ? tmin : tmax;
    out = ray.origin + ray.direction * t;
    return true;
}
// aabb3.rs:205-229 Discrete<Aabb3> for Ray
inline bool ray_aabb_intersects(const Ray& ray, const Aabb3& aabb) {
    Vec3 inv_dir = Vec3(1.0f, 1.0f, 1.0f) / ray.direction;
    float t1 = (aabb.min.x - ray.origin.x) * inv_dir.x;
    float t2 = (aabb.max.x - ray.origin.x) * inv_dir.x;
    float tmin = rmin(t1, t2);
    float tmax = rmax(t1, t2);
    t1 = (aabb.min.y - ray.origin.y) * inv_dir.y;
    t2 = (aabb.max.y - ray.origin.y) * inv_dir.y;
    tmin = rmax(tmin, rmin(t1, t2));
    tmax = rmin(tmax, rmax(t1, t2));
    t1 = (aabb.min.z - ray.origin.z) * inv_dir.z;
    t2 = (aabb.max.z - ray.origin.z) * inv_dir.z;
    tmin = rmax(tmin, rmin(t1, t2));
    tmax = rmin(tmax, rmax(t1, t2));
    return tmax >= tmin && (tmin >= 0.0f || tmax >= 0.0f);
}

// ---------------------------------------------------------------------------------------------------------
// volume/sphere.rs — the bounding sphere (named BSphere here: the primitive Sphere lives in bam3d_oracle.hpp)
// ---------------------------------------------------------------------------------------------------------
struct BSphere {
    Vec3 center;
    float radius = 0.f;
    BSphere() {}
    BSphere(Vec3 c, float r) : center(c), radius(r) {}
    // Bound for Sphere, sphere.rs:17-48
    Vec3 min_extent() const { return center + Vec3(-radius, -radius, -radius); }
    Vec3 max_extent() const { return center + Vec3(radius, radius, radius); }
    BSphere with_margin(Vec3 add) const { return BSphere(center, radius + rmax(rmax(add.x, add.y), add.z)); }
    BSphere transform_volume(const Mat4& t) const { return BSphere(t.transform_point3(center), radius); }
    static BSphere empty() { return BSphere(Vec3::zero(), 0.f); }
    // Continuous<Ray>, sphere.rs:50-67
    bool intersection(const Ray& r, Vec3& out) const {
        Vec3 l = center - r.origin;
        float tca = dot(l, r.direction);
        if (tca < 0.f) return false;
        float d2 = dot(l, l) - tca * tca;
        if (d2 > radius * radius) return false;
        float thc = std::sqrt(radius * radius - d2);
        out = r.origin + r.direction * (tca - thc);
        return true;
    }
    // Discrete<Ray>, sphere.rs:69-80
    bool intersects(const Ray& r) const {
        Vec3 l = center - r.origin;
        float tca = dot(l, r.direction);
        if (tca < 0.f) return false;
        float d2 = dot(l, l) - tca * tca;
        return d2 <= radius * radius;
    }
    // Discrete<Sphere>, sphere.rs:82-91 — touching spheres intersect
    bool intersects(const BSphere& s2) const {
        Vec3 distance = center - s2.center;
        float radiuses = radius + s2.radius;
        return dot(distance, distance) <= radiuses * radiuses;
    }
    // PlaneBound, sphere.rs:93-104
    Relation relate_plane(const Plane& plane) const {
        float dist = dot(center, plane.n) - plane.d;
        if (dist > radius) return In;
        else if (dist < -radius) return Out;
        else return Cross;
    }
    // Contains<Aabb3>, sphere.rs:106-119
    bool contains(const Aabb3& aabb) const {
        float radius_sq = radius * radius;
        Vec3 c[8]; aabb.to_corners(c);
        for (int i = 0; i < 8; ++i) {
            Vec3 distance = c[i] - center;
            if (dot(distance, distance) > radius_sq) return false;
        }
        return true;
    }
    // Contains<Vec3>, sphere.rs:121-127
    bool contains(Vec3 p) const {
        Vec3 distance = p - center;
        return dot(distance, distance) <= radius * radius;
    }
    // Contains<Line>, sphere.rs:129-134
    bool contains_line(Vec3 origin, Vec3 dest) const { return contains(origin) && contains(dest); }
    // Contains<Sphere>, sphere.rs:136-143
    bool contains(const BSphere& other) const {
        Vec3 center_dist = other.center - center;
        float distance_sq = dot(center_dist, center_dist);
        return std::sqrt(distance_sq) + other.radius <= radius;
    }
    // Union, sphere.rs:145-163
    BSphere union_(const BSphere& other) const {
        if (contains(other)) return *this;
        if (other.contains(*this)) return other;
        Vec3 center_diff = other.center - center;
        float center_diff_s = std::sqrt(dot(center_diff, center_diff));
        float r = (radius + other.radius + center_diff_s) / 2.f;
        return BSphere(center + center_diff * (r - radius) / center_diff_s, r);
    }
    // Union<Aabb3>, sphere.rs:165-190
    BSphere union_(const Aabb3& aabb) const;
    // SurfaceArea, sphere.rs:192-196
    float surface_area() const { return 4.f * 3.14159265358979323846f * radius * radius; }
};
// Contains<Sphere> for Aabb3, aabb3.rs:125-136 (border hits count on both extents)
inline bool aabb_contains_sphere(const Aabb3& a, const BSphere& s) {
    return (s.center.x - s.radius) >= a.min.x && (s.center.y - s.radius) >= a.min.y && (s.center.z - s.radius) >= a.min.z &&
           (s.center.x + s.radius) <= a.max.x && (s.center.y + s.radius) <= a.max.y && (s.center.z + s.radius) <= a.max.z;
}
// Union<Sphere> for Aabb3, aabb3.rs:153-163
inline Aabb3 aabb_union_sphere(const Aabb3& a, const BSphere& s) {
    return a.grow(Vec3(s.center.x - s.radius, s.center.y - s.radius, s.center.z - s.radius)).grow(s.center + Vec3(s.radius, s.radius, s.radius));
}
inline BSphere BSphere::union_(const Aabb3& aabb) const {
    if (contains(aabb)) return *this;
    Vec3 dist = aabb.max - aabb.center();
    float aabb_radius = std::sqrt(dot(dist, dist));
    if (aabb_contains_sphere(aabb, *this)) return BSphere(aabb.center(), aabb_radius);
    Vec3 center_diff = aabb.center() - center;
    float center_diff_s = std::sqrt(dot(center_diff, center_diff));
    float r = (radius + aabb_radius + center_diff_s) / 2.f;
    return BSphere(center + center_diff * (r - radius) / center_diff_s, r);
}

// ---------------------------------------------------------------------------------------------------------
// frustum.rs
// ---------------------------------------------------------------------------------------------------------
struct Frustum {
    Plane left, right, bottom, top, near_, far_;
    // frustum.rs:46-75 — NB near/far are built from the SAME rows as left/right (w+x, w-x); reproduced as-is.
    static bool from_mat4(const Mat4& matrix, Frustum& f) {
        Mat4 mat = matrix.transpose();
        if (!Plane::from_vector4_alt(mat.w_axis + mat.x_axis).normalize(f.left)) return false;
        if (!Plane::from_vector4_alt(mat.w_axis - mat.x_axis).normalize(f.right)) return false;
        if (!Plane::from_vector4_alt(mat.w_axis + mat.y_axis).normalize(f.bottom)) return false;
        if (!Plane::from_vector4_alt(mat.w_axis - mat.y_axis).normalize(f.top)) return false;
        if (!Plane::from_vector4_alt(mat.w_axis + mat.x_axis).normalize(f.near_)) return false;
        if (!Plane::from_vector4_alt(mat.w_axis - mat.x_axis).normalize(f.far_)) return false;
        return true;
    }
    // frustum.rs:78-95: fold max over [left, right, top, bottom, near, far]
    template <class B>  // any PlaneBound: Aabb3 or BSphere
    Relation contains(const B& b) const {
        const Plane* ps[6] = {&left, &right, &top, &bottom, &near_, &far_};
        Relation cur = In;
        for (int i = 0; i < 6; ++i) {
            Relation r = b.relate_plane(*ps[i]);
            if (r > cur) cur = r;
        }
        return cur;
    }
};

// ---------------------------------------------------------------------------------------------------------
// algorithm/broad_phase/sweep_prune.rs
// ---------------------------------------------------------------------------------------------------------
struct Variance3 {  // sweep_prune.rs:166-216
    Vec3 csum, csumsq;
    void clear() { csum = Vec3::zero(); csumsq = Vec3::zero(); }
    template <class B>
    void add_bound(const B& b) {  // :192-199
        Vec3 sum = b.min_extent() + b.max_extent();
        Vec3 c = sum / 2.0f;
        csum = csum + c;
        csumsq = csumsq + c * c;
    }
    // :202-215 — NB "variance" is csumsq * (csum*csum/n): a product, not a difference; reproduced as-is.
    int compute_axis(float n, float* var_out = nullptr) const {
        Vec3 square_n = csum * csum / n;
        Vec3 variance = csumsq * square_n;
        int sweep_axis = 0;
        float sweep_variance = variance[0];
        for (int i = 1; i < 3; ++i) {
            float v = variance[i];
            if (v > sweep_variance) { sweep_axis = i; sweep_variance = v; }
        }
        if (var_out) *var_out = sweep_variance;
        return sweep_axis;
    }
};

template <class Bnd>
struct SweepAndPruneT {  // sweep_prune.rs:27-128, generic over the bound type like the reference (A::Bound: Bound + Discrete<A::Bound>)
    typedef Bnd Aabb3;  // (the body below was written for Aabb3; it only uses min_extent / max_extent / intersects)
    int sweep_axis = 0;
    Variance3 variance;
    // find_collider_pairs, :69-128. `bounds` is re-sorted in place (stable), `perm[k]` = original index of the
    // shape now at sorted position k. Pair indices refer to the SORTED order, emitted (j asc, i asc).
    std::vector<std::pair<uint32_t, uint32_t>> find_collider_pairs(std::vector<Aabb3>& bounds,
                                                                   std::vector<uint32_t>* perm = nullptr) {
        std::vector<std::pair<uint32_t, uint32_t>> pairs;
        size_t n = bounds.size();
        std::vector<uint32_t> order(n);
        for (size_t i = 0; i < n; ++i) order[i] = (uint32_t)i;
        if (n <= 1) { if (perm) *perm = order; return pairs; }
        const int ax = sweep_axis;
        // slice::sort_by is a stable merge sort; comparator :80-90 (NaN compares Equal)
        auto cmp3 = [&](const Aabb3& a, const Aabb3& b) -> int {
            float am = a.min_extent()[ax], bm = b.min_extent()[ax];
            if (am < bm) return -1;
            if (am > bm) return 1;
            if (am == bm) {
                float ax_ = a.max_extent()[ax], bx_ = b.max_extent()[ax];
                if (ax_ < bx_) return -1;
                if (ax_ > bx_) return 1;
                return 0;  // Equal, or NaN -> unwrap_or(Equal)
            }
            return 0;  // partial_cmp == None -> Equal
        };
        std::stable_sort(order.begin(), order.end(),
                         [&](uint32_t a, uint32_t b) { return cmp3(bounds[a], bounds[b]) < 0; });
        std::vector<Aabb3> sorted(n);
        for (size_t i = 0; i < n; ++i) sorted[i] = bounds[order[i]];
        bounds.swap(sorted);
        if (perm) *perm = order;

        variance.clear();
        variance.add_bound(bounds[0]);
        std::vector<uint32_t> active;
        active.push_back(0);
        for (size_t shape_index = 1; shape_index < n; ++shape_index) {
            const Aabb3& shape = bounds[shape_index];
            // active.retain(...) :102-105
            size_t w = 0;
            for (size_t k = 0; k < active.size(); ++k)
                if (bounds[active[k]].max_extent()[ax] >= shape.min_extent()[ax]) active[w++] = active[k];
            active.resize(w);
            for (size_t k = 0; k < active.size(); ++k)
                if (bounds[active[k]].intersects(shape)) pairs.emplace_back(active[k], (uint32_t)shape_index);
            active.push_back((uint32_t)shape_index);
            variance.add_bound(shape);
        }
        sweep_axis = variance.compute_axis((float)n);
        return pairs;
    }
};
typedef SweepAndPruneT<Aabb3> SweepAndPrune3;          // SweepAndPrune3<Aabb3>
typedef SweepAndPruneT<BSphere> SweepAndPrune3Sphere;  // SweepAndPrune3<volume::Sphere>

// ---------------------------------------------------------------------------------------------------------
// dbvt/mod.rs (insert + refit of bounds + query), dbvt/visitor.rs, dbvt/util.rs
// ---------------------------------------------------------------------------------------------------------
struct TreeValue {  // TreeValueWrapped, dbvt/wrapped.rs:10-51
    Aabb3 bound;
    Aabb3 fat;  // get_bound_with_margin()
};

struct Dbvt {
    enum NodeKind : uint8_t { Nil = 0, Leaf = 1, Branch = 2 };
    struct Node {
        NodeKind kind = Nil;
        uint32_t parent = 0, left = 0, right = 0, value = 0;
        Aabb3 bound;
    };
    std::vector<Node> nodes;       // nodes[0] is Nil (dbvt/mod.rs:257)
    std::vector<TreeValue> values;
    uint32_t root_index = 0;
    bool stack_overflow = false;   // the reference would panic (index out of bounds on the 256-entry stack)

    Dbvt() { nodes.emplace_back(); }

    // dbvt/mod.rs:622-731 (without mark_for_refit bookkeeping; call refit() afterwards)
    uint32_t insert(const TreeValue& value) {
        uint32_t value_index = (uint32_t)values.size();
        values.push_back(value);
        Node new_leaf; new_leaf.kind = Leaf; new_leaf.value = value_index; new_leaf.bound = value.fat;
        if (root_index == 0) {
            root_index = (uint32_t)nodes.size();
            nodes.push_back(new_leaf);
            return root_index;
        }
        uint32_t node_index = root_index;
        uint32_t new_branch_index = (uint32_t)nodes.size();
        uint32_t new_leaf_index = new_branch_index + 1;
        nodes.emplace_back(); nodes.emplace_back();
        new_leaf.parent = new_branch_index;
        for (;;) {
            Node& cur = nodes[node_index];
            if (cur.kind == Leaf) {
                Node br; br.kind = Branch; br.left = node_index; br.right = new_leaf_index; br.parent = cur.parent;
                br.bound = cur.bound.union_(new_leaf.bound);
                uint32_t leaf_index = node_index;
                cur.parent = new_branch_index;
                if (br.parent != 0) {
                    Node& par = nodes[br.parent];
                    if (par.left == leaf_index) par.left = new_branch_index; else par.right = new_branch_index;
                }
                nodes[new_branch_index] = br;
                nodes[new_leaf_index] = new_leaf;
                if (leaf_index == root_index) root_index = new_branch_index;
                break;
            } else if (cur.kind == Branch) {
                float left_area = nodes[cur.left].bound.union_(new_leaf.bound).surface_area();
                float right_area = nodes[cur.right].bound.union_(new_leaf.bound).surface_area();
                node_index = (left_area < right_area) ? cur.left : cur.right;
            } else {
                break;
            }
        }
        return new_leaf_index;
    }

    // Bottom-up recomputation of every branch bound = union(left, right) (what do_refit converges to,
    // dbvt/mod.rs:1083-1137, minus rotations).
    void refit() { if (root_index) refit_rec(root_index); }
    void refit_rec(uint32_t root) {
        // iterative post-order to survive deep trees
        std::vector<std::pair<uint32_t, bool>> st;
        st.emplace_back(root, false);
        while (!st.empty()) {
            auto [i, done] = st.back(); st.pop_back();
            Node& nd = nodes[i];
            if (nd.kind != Branch) continue;
            if (!done) { st.emplace_back(i, true); st.emplace_back(nd.left, false); st.emplace_back(nd.right, false); }
            else nd.bound = nodes[nd.left].bound.union_(nodes[nd.right].bound);
        }
    }

    // dbvt/mod.rs:481-515 query_for_indices: DFS, fixed 256-entry stack; leaves test the value's REAL bound,
    // branches the node bound. `accept(bound, is_leaf, result&) -> bool`.
    template <class Visitor, class Result>
    void query_for_indices(Visitor& visitor, std::vector<std::pair<uint32_t, Result>>& out) {
        uint32_t stack[256];
        stack[0] = root_index;
        int sp = 1;
        while (sp > 0) {
            --sp;
            const Node& node = nodes[stack[sp]];
            if (node.kind == Leaf) {
                Result r;
                if (visitor.accept(values[node.value].bound, true, r)) out.emplace_back(node.value, r);
            } else if (node.kind == Branch) {
                Result r;
                if (visitor.accept(node.bound, false, r)) {
                    if (sp + 2 > 256) { stack_overflow = true; return; }
                    stack[sp] = node.left; stack[sp + 1] = node.right; sp += 2;
                }
            }
        }
    }
};

struct DiscreteVisitor {  // visitor.rs:57-90
    Aabb3 bound;
    bool accept(const Aabb3& b, bool, uint8_t& r) { r = 0; return b.intersects(bound); }
};
struct ContinuousRayVisitor {  // visitor.rs:16-47 with B = Ray -> aabb3.rs:197-203
    Ray ray;
    bool accept(const Aabb3& b, bool, Vec3& r) { return ray_aabb_intersection(ray, b, r); }
};
struct FrustumVisitor {  // visitor.rs:99-134
    Frustum frustum;
    bool accept(const Aabb3& b, bool, uint8_t& r) {
        Relation rel = frustum.contains(b);
        if (rel == Out) return false;
        r = (uint8_t)rel;
        return true;
    }
};
struct RayClosestVisitor {  // util.rs:12-62
    Ray ray;
    float min = std::numeric_limits<float>::infinity();
    bool accept(const Aabb3& b, bool is_leaf, Vec3& r) {
        Vec3 point;
        if (!ray_aabb_intersection(ray, b, point)) return false;
        Vec3 offset = point - ray.origin;
        float t = dot(offset, ray.direction);
        if (t < min) {
            if (is_leaf) min = t;
            r = point;
            return true;
        }
        return false;
    }
};

// util.rs:77-103 query_ray_closest
inline bool query_ray_closest(Dbvt& tree, const Ray& ray, uint32_t& value_index, Vec3& point, float& t_out) {
    RayClosestVisitor vis; vis.ray = ray;
    std::vector<std::pair<uint32_t, Vec3>> hits;
    tree.query_for_indices<RayClosestVisitor, Vec3>(vis, hits);
    bool saved = false;
    float tmin = std::numeric_limits<float>::infinity();
    for (auto& h : hits) {
        Vec3 offset = h.second - ray.origin;
        float t = dot(offset, ray.direction);
        if (t < tmin) { tmin = t; value_index = h.first; point = h.second; saved = true; }
    }
    t_out = tmin;
    return saved;
}

// algorithm/broad_phase/dbvt.rs:26-63 DbvtBroadPhase::find_collider_pairs
inline std::vector<std::pair<uint32_t, uint32_t>> dbvt_find_collider_pairs(Dbvt& tree, const uint8_t* dirty) {
    std::vector<std::pair<uint32_t, uint32_t>> potentials;
    for (uint32_t vi = 0; vi < tree.values.size(); ++vi) {
        if (!dirty[vi]) continue;
        DiscreteVisitor vis; vis.bound = tree.values[vi].bound;
        std::vector<std::pair<uint32_t, uint8_t>> hits;
        tree.query_for_indices<DiscreteVisitor, uint8_t>(vis, hits);
        for (auto& h : hits) {
            if (h.first == vi) continue;
            std::pair<uint32_t, uint32_t> pr = (vi < h.first) ? std::make_pair(vi, h.first) : std::make_pair(h.first, vi);
            auto pos = std::lower_bound(potentials.begin(), potentials.end(), pr);
            if (pos == potentials.end() || *pos != pr) potentials.insert(pos, pr);
        }
    }
    return potentials;
}

}  // namespace orc
