// GJK simplex state and the 3-D simplex processor, held entirely in registers.
// Reference: SimplexProcessor3 (gjk/simplex/simplex3d.rs:16-201) over `Simplex = SmallVec<[SupportPoint; 5]>`
// (gjk/simplex/mod.rs:10). Points are stored oldest-first exactly like the reference (newest = last), and every
// `remove(i)` / `swap(0, 1)` is spelled out with compile-time indices so the arrays never leave registers.
#pragma once
#include "vecmath.cuh"

namespace b3 {

// Per-item outcome codes (bam3d_gpu.h: BAM3D_STATUS_*)
enum Status : uint8_t { ST_OK = 0, ST_MISS = 1, ST_MAX_ITER = 2, ST_REF_PANIC = 3, ST_NAN = 4, ST_CAPACITY = 5 };

// WITH_A: also carry sup_a (the support point on the left shape), which only EPA's contact point reads.
template <bool WITH_A>
struct Simplex {
    V3 v[4];
    V3 a[WITH_A ? 4 : 1];
    int n;

    B3_DEV void clear() { n = 0; }
    B3_DEV void push(V3 pv, V3 pa) {
        if (n == 0) { v[0] = pv; if (WITH_A) a[0] = pa; }
        else if (n == 1) { v[1] = pv; if (WITH_A) a[1] = pa; }
        else if (n == 2) { v[2] = pv; if (WITH_A) a[2] = pa; }
        else { v[3] = pv; if (WITH_A) a[3] = pa; }
        ++n;
    }
    B3_DEV void mov(int dst, int src) { v[dst] = v[src]; if (WITH_A) a[dst] = a[src]; }
    B3_DEV void swap01() {
        V3 t = v[0]; v[0] = v[1]; v[1] = t;
        if (WITH_A) { V3 u = a[0]; a[0] = a[1]; a[1] = u; }
    }
};

B3_DEV V3 cross_aba(V3 a, V3 b) { return cross(cross(a, b), a); }  // simplex3d.rs:157-159

// check_side (simplex3d.rs:163-201) on a 3-point simplex [c, b, a] (indices 0, 1, 2).
template <bool WITH_A>
B3_DEV void check_side(V3 abc, V3 ab, V3 ac, V3 ao, Simplex<WITH_A>& s, V3& v, bool above, bool ignore_ab) {
    V3 ab_perp = cross(ab, abc);
    if (!ignore_ab && dot(ab_perp, ao) > 0.f) {
        s.mov(0, 1); s.mov(1, 2); s.n = 2;  // simplex.remove(0)
        v = cross_aba(ab, ao);
        return;
    }
    V3 ac_perp = cross(abc, ac);
    if (dot(ac_perp, ao) > 0.f) {
        s.mov(1, 2); s.n = 2;               // simplex.remove(1)
        v = cross_aba(ac, ao);
        return;
    }
    if (above || dot(abc, ao) > 0.f) {
        v = abc;
    } else {
        s.swap01();
        v = -abc;
    }
}

// reduce_to_closest_feature (simplex3d.rs:16-90). Returns true when the tetrahedron encloses the origin.
template <bool WITH_A>
B3_DEV bool reduce_to_closest_feature(Simplex<WITH_A>& s, V3& v) {
    if (s.n == 4) {
        V3 a = s.v[3], b = s.v[2], c = s.v[1], d = s.v[0];
        V3 ao = -a, ab = b - a, ac = c - a;
        V3 abc = cross(ab, ac);
        if (dot(abc, ao) > 0.f) {
            s.mov(0, 1); s.mov(1, 2); s.mov(2, 3); s.n = 3;  // remove(0)
            check_side(abc, ab, ac, ao, s, v, true, false);
        } else {
            V3 ad = d - a;
            V3 acd = cross(ac, ad);
            if (dot(acd, ao) > 0.f) {
                s.mov(2, 3); s.n = 3;                        // remove(2)
                check_side(acd, ac, ad, ao, s, v, true, true);
            } else {
                V3 adb = cross(ad, ab);
                if (dot(adb, ao) > 0.f) {
                    s.mov(1, 2); s.mov(2, 3); s.n = 3;       // remove(1)
                    s.swap01();
                    v = adb;
                } else {
                    return true;
                }
            }
        }
    } else if (s.n == 3) {
        V3 a = s.v[2], b = s.v[1], c = s.v[0];
        V3 ao = -a, ab = b - a, ac = c - a;
        check_side(cross(ab, ac), ab, ac, ao, s, v, false, false);
    } else if (s.n == 2) {
        V3 a = s.v[1], b = s.v[0];
        V3 ao = -a, ab = b - a;
        v = cross_aba(ab, ao);
        if (is_zero(v)) v.x = 0.1f;
    }
    return false;
}

// barycentric_vector (primitive/util.rs:44-59)
B3_DEV void barycentric_vector(V3 p, V3 a, V3 b, V3 c, float& u, float& v, float& w) {
    V3 v0 = b - a, v1 = c - a, v2 = p - a;
    float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    float inv_denom = 1.0f / (d00 * d11 - d01 * d01);
    v = (d11 * d20 - d01 * d21) * inv_denom;
    w = (d00 * d21 - d01 * d20) * inv_denom;
    u = 1.0f - v - w;
}

// get_closest_point_on_edge (primitive/util.rs:81-93) with point = origin
B3_DEV V3 closest_point_on_edge_to_origin(V3 start, V3 end) {
    V3 line = end - start;
    V3 line_dir = normalize(line);
    V3 v = v3(0.f, 0.f, 0.f) - start;
    float d = dot(v, line_dir);
    if (d < 0.f) return start;
    else if ((d * d) > dot(line, line)) return end;
    else return start + line_dir * d;
}

B3_DEV bool in_range01(float v) { return v >= 0.f && v <= 1.f; }  // simplex3d.rs:152-154

// get_closest_point_on_face (simplex3d.rs:122-149) = Plane::from_points (plane.rs:69-86) + ray/plane
// intersection (plane.rs:117-129) + the barycentric assert!. Where the reference would panic
// (`unwrap` on a degenerate face, or the assert), st is set to ST_REF_PANIC instead.
B3_DEV V3 closest_point_on_face(V3 a, V3 b, V3 c, V3 normal, uint8_t& st) {
    V3 origin = v3(0.f, 0.f, 0.f);
    V3 dir = -normal;
    V3 pn = cross(b - a, c - a);
    if (is_zero(pn)) { st = ST_REF_PANIC; return origin; }
    V3 n = normalize(pn);
    float d = -dot(a, n);
    float t = -(d + dot(origin, n)) / dot(dir, n);
    if (t < 0.0f) return origin;
    V3 point = origin + dir * t;
    float u, v, w;
    barycentric_vector(point, a, b, c, u, v, w);
    if (!(in_range01(u) && in_range01(v) && in_range01(w))) st = ST_REF_PANIC;
    return point;
}

// get_closest_point_to_origin (simplex3d.rs:95-113)
template <bool WITH_A>
B3_DEV V3 closest_point_to_origin(Simplex<WITH_A>& s, uint8_t& st) {
    V3 d = v3(0.f, 0.f, 0.f);
    if (reduce_to_closest_feature(s, d)) return d;
    if (s.n == 1) return s.v[0];
    else if (s.n == 2) return closest_point_on_edge_to_origin(s.v[1], s.v[0]);
    else return closest_point_on_face(s.v[2], s.v[1], s.v[0], d, st);
}

}  // namespace b3
