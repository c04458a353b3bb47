// ORACLE — TEST INFRASTRUCTURE ONLY (see glam_mini.hpp header). Restatement of the reference's ray casts against
// primitives, f32, same operation order:
//   Discrete/Continuous<Ray> for Vec3 (ray.rs:35-57), Sphere (primitive/sphere.rs:45-76), Cuboid (cuboid.rs:85-103,
//   through Aabb3: volume/aabb/aabb3.rs:165-235), Quad (quad.rs:80-96), Cylinder (cylinder.rs:82-204), Capsule
//   (capsule.rs:69-178), Particle (particle.rs:53-69), cylinder_ray_quadratic_solve (primitive/util.rs:95-109);
//   the blanket Discrete/ContinuousTransformed<Ray> (ray.rs:58-78); Primitive3's dispatch (primitive3.rs:133-166);
//   the swept particle (Particle, Range<Vec3>) (particle.rs:39-128).
// ConvexPolyhedron (polyhedron.rs:373-522): the face walk over the half-edge structure, for polyhedra built with
// new_with_faces; a VertexOnly polyhedron has no faces, and the walk's `Some(0)` start then indexes an empty list (the
// reference panics): reported as unsupported.
#pragma once
#include "bam3d_oracle.hpp"
#include "broad_oracle.hpp"

namespace orc {

// ray.rs:46-57 Discrete<Ray> for Vec3
inline bool point_ray_intersects(Vec3 p, const Ray& ray) {
    Vec3 l = p - ray.origin;
    float tca = dot(l, ray.direction);
    return tca > 0.f && (std::fabs((tca * tca) - dot(l, l)) < F32_EPSILON);
}

// primitive/util.rs:95-109
inline bool cylinder_ray_quadratic_solve(const Ray& r, float radius, float& t1, float& t2) {
    float a = r.direction.x * r.direction.x + r.direction.z * r.direction.z;
    float b = 2.f * (r.direction.x * r.origin.x + r.direction.z * r.origin.z);
    float c = r.origin.x * r.origin.x + r.origin.z * r.origin.z - radius * radius;
    float dr = b * b - 4.f * a * c;
    if (dr < 0.f) return false;
    float drsqrt = std::sqrt(dr);
    t1 = (-b + drsqrt) / (2.f * a);
    t2 = (-b - drsqrt) / (2.f * a);
    return true;
}

// primitive/util.rs:63-79 barycentric_point (same expressions as barycentric_vector)
inline void barycentric_point(Vec3 p, Vec3 a, Vec3 b, Vec3 c, float& u, float& v, float& w) {
    Vec3 v0 = b - a, v1 = c - a, v2 = p - a;
    float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    float inv_denom = 1.f / (d00 * d11 - d01 * d01);
    v = (d11 * d20 - d01 * d21) * inv_denom;
    w = (d00 * d21 - d01 * d20) * inv_denom;
    u = 1.f - v - w;
}

// find_intersecting_face (polyhedron.rs:401-425) with intersect_ray_face (:497-521) and next_face_classify (:433-494).
// Returns the face index or -1; (u, v, w) are the hit's barycentric coordinates.
inline int poly_find_intersecting_face(const Shape& s, const Ray& ray, float& u, float& v, float& w) {
    const HalfEdgeMesh& m = *s.he;
    const size_t nfaces = m.faces.size();
    auto pos = [&](uint32_t i) { return Vec3(s.verts[3 * i], s.verts[3 * i + 1], s.verts[3 * i + 2]); };
    std::vector<bool> checked(nfaces, false);
    long face = 0;
    while (face >= 0) {
        const size_t fi = (size_t)face;
        checked[fi] = true;
        const HalfEdgeMesh::Face& f = m.faces[fi];
        bool have = false;
        if (dot(f.n, ray.direction) < 0.f) {           // intersect_ray_face
            const float t = -(f.d + dot(ray.origin, f.n)) / dot(ray.direction, f.n);  // plane.rs:117-129
            if (!(t < 0.f)) {
                const Vec3 p = ray.origin + ray.direction * t;
                barycentric_point(p, pos(f.v[0]), pos(f.v[1]), pos(f.v[2]), u, v, w);
                have = true;
                auto in_range = [](float x) { return x >= 0.f && x <= 1.f; };
                if (in_range(u) && in_range(v) && in_range(w)) return (int)fi;
            }
        }
        // next_face_classify
        long next = -1;
        if (nfaces < 10) {
            next = (fi == nfaces - 1) ? -1 : (long)fi + 1;
        } else if (!have) {
            size_t nx = fi + 1;
            while (nx < nfaces && checked[nx]) ++nx;
            next = (nx == nfaces) ? -1 : (long)nx;
        } else {
            const uint32_t target = (u < 0.f) ? f.v[2] : ((v < 0.f) ? f.v[0] : f.v[1]);
            const HalfEdgeMesh::Edge& fe = m.edges[f.edge];
            uint32_t e3[3];
            if (fe.target_vertex == target) { e3[0] = f.edge; e3[1] = fe.next_edge; e3[2] = fe.previous_edge; }
            else if (m.edges[fe.previous_edge].target_vertex == target) { e3[0] = fe.previous_edge; e3[1] = f.edge; e3[2] = fe.next_edge; }
            else { e3[0] = fe.next_edge; e3[1] = fe.previous_edge; e3[2] = f.edge; }
            for (int k = 0; k < 3 && next < 0; ++k) {
                const uint32_t lf = m.edges[m.edges[e3[k]].twin_edge].left_face;
                if (!checked[lf]) next = (long)lf;
            }
            if (next < 0)
                for (size_t i = 0; i < nfaces; ++i)
                    if (!checked[i]) { next = (long)i; break; }
        }
        face = next;
    }
    return -1;
}

// Discrete<Ray> in the primitive's object space
inline bool ray_intersects_local(const Shape& s, const Ray& r) {
    switch (s.kind) {
        case PARTICLE: return point_ray_intersects(Vec3::zero(), r);  // particle.rs:53-58
        case SPHERE: {  // sphere.rs:45-56
            Vec3 l(-r.origin.x, -r.origin.y, -r.origin.z);
            float tca = dot(l, r.direction);
            if (tca < 0.f) return false;
            float d2 = dot(l, l) - tca * tca;
            return d2 <= s.p0 * s.p0;
        }
        case CUBOID: return ray_aabb_intersects(r, Aabb3(-Vec3(s.p0, s.p1, s.p2), Vec3(s.p0, s.p1, s.p2)));  // cuboid.rs:85-92
        case QUAD: return ray_aabb_intersects(r, Aabb3(Vec3(-s.p0, -s.p1, 0.f), Vec3(s.p0, s.p1, 0.f)));     // quad.rs:80-86
        case CYLINDER: {  // cylinder.rs:82-132; p0 = half_height, p1 = radius
            const float hh = s.p0, radius = s.p1;
            if (std::fabs(r.direction.x - 0.f) < F32_EPSILON && std::fabs(r.direction.z - 0.f) < F32_EPSILON) {
                if (std::fabs(r.direction.y - 0.f) < F32_EPSILON) return false;
                return (r.origin.y >= -hh && r.direction.y <= 0.f) || (r.origin.y <= hh && r.direction.y >= 0.f);
            }
            float t1, t2;
            if (!cylinder_ray_quadratic_solve(r, radius, t1, t2)) return false;
            if (t1 < 0.f && t2 < 0.f) return false;
            float t = (t1 < 0.f) ? t2 : ((t2 < 0.f) ? t1 : rmin(t1, t2));
            Vec3 pc = r.origin + r.direction * t;
            if (pc.y <= hh && pc.y >= -hh) return true;
            Vec3 n = -Vec3(0.f, 1.f, 0.f);
            float tp = -(hh + dot(r.origin, n)) / dot(r.direction, n);
            if (tp >= 0.f) {
                Vec3 p = r.origin + r.direction * tp;
                if (p.x * p.x + p.z * p.z < radius * radius) return true;
            }
            n = Vec3(0.f, 1.f, 0.f);
            float tb = -(-hh + dot(r.origin, n)) / dot(r.direction, n);
            if (tb >= 0.f) {
                Vec3 p = r.origin + r.direction * tb;
                if (p.x * p.x + p.z * p.z < radius * radius) return true;
            }
            return false;
        }
        case CAPSULE: {  // capsule.rs:69-113
            const float hh = s.p0, radius = s.p1;
            float t1, t2;
            if (!cylinder_ray_quadratic_solve(r, radius, t1, t2)) return false;
            if (t1 < 0.f && t2 < 0.f) return false;
            float t = (t1 < 0.f) ? t2 : ((t2 < 0.f) ? t1 : rmin(t1, t2));
            Vec3 pc = r.origin + r.direction * t;
            if (pc.y <= hh && pc.y >= -hh) return true;
            Vec3 l(-r.origin.x, hh - r.origin.y, -r.origin.z);
            float tca = dot(l, r.direction);
            if (tca > 0.f) {
                float d2 = dot(l, l) - tca * tca;
                if (d2 <= radius * radius) return true;
            }
            l = Vec3(-r.origin.x, -hh - r.origin.y, -r.origin.z);
            tca = dot(l, r.direction);
            if (tca > 0.f) {
                float d2 = dot(l, l) - tca * tca;
                if (d2 <= radius * radius) return true;
            }
            return false;
        }
        case POLYHEDRON: {  // polyhedron.rs:373-378
            float u, v, w;
            return s.he && !s.he->faces.empty() && poly_find_intersecting_face(s, r, u, v, w) >= 0;
        }
        default: return false;
    }
}

// Continuous<Ray> in the primitive's object space
inline bool ray_intersection_local(const Shape& s, const Ray& r, Vec3& out) {
    switch (s.kind) {
        case PARTICLE:  // particle.rs:61-68 -> ray.rs:35-44
            if (point_ray_intersects(Vec3::zero(), r)) { out = Vec3::zero(); return true; }
            return false;
        case SPHERE: {  // sphere.rs:58-76
            Vec3 l(-r.origin.x, -r.origin.y, -r.origin.z);
            float tca = dot(l, r.direction);
            if (tca < 0.f) return false;
            float d2 = dot(l, l) - tca * tca;
            if (d2 > s.p0 * s.p0) return false;
            float thc = std::sqrt(s.p0 * s.p0 - d2);
            out = r.origin + r.direction * (tca - thc);
            return true;
        }
        case CUBOID: return ray_aabb_intersection(r, Aabb3(-Vec3(s.p0, s.p1, s.p2), Vec3(s.p0, s.p1, s.p2)), out);  // cuboid.rs:94-103
        case QUAD: return ray_aabb_intersection(r, Aabb3(Vec3(-s.p0, -s.p1, 0.f), Vec3(s.p0, s.p1, 0.f)), out);     // quad.rs:88-96
        case CYLINDER: {  // cylinder.rs:134-204
            const float hh = s.p0, radius = s.p1;
            if (std::fabs(r.direction.x - 0.f) < F32_EPSILON && std::fabs(r.direction.z - 0.f) < F32_EPSILON) {
                if (std::fabs(r.direction.y - 0.f) < F32_EPSILON) return false;
                if (r.origin.y >= hh && r.direction.y < 0.f) { out = Vec3(0.f, hh, 0.f); return true; }
                if (r.origin.y >= -hh && r.direction.y < 0.f) { out = Vec3(0.f, -hh, 0.f); return true; }
                if (r.origin.y <= -hh && r.direction.y > 0.f) { out = Vec3(0.f, -hh, 0.f); return true; }
                if (r.origin.y <= hh && r.direction.y > 0.f) { out = Vec3(0.f, hh, 0.f); return true; }
                return false;
            }
            float t1, t2;
            if (!cylinder_ray_quadratic_solve(r, radius, t1, t2)) return false;
            if (t1 < 0.f && t2 < 0.f) return false;
            float t = (t1 < 0.f) ? t2 : ((t2 < 0.f) ? t1 : rmin(t1, t2));
            Vec3 n = -Vec3(0.f, 1.f, 0.f);
            float tp = -(hh + dot(r.origin, n)) / dot(r.direction, n);
            if (tp >= 0.f && tp < t) {
                Vec3 p = r.origin + r.direction * tp;
                if (p.x * p.x + p.z * p.z < radius * radius) t = tp;
            }
            n = Vec3(0.f, 1.f, 0.f);
            float tb = -(-hh + dot(r.origin, n)) / dot(r.direction, n);
            if (tb >= 0.f && tb < t) {
                Vec3 p = r.origin + r.direction * tb;
                if (p.x * p.x + p.z * p.z < radius * radius) t = tb;
            }
            Vec3 pc = r.origin + r.direction * t;
            if (pc.y > hh || pc.y < -hh) return false;
            out = pc;
            return true;
        }
        case CAPSULE: {  // capsule.rs:115-178
            const float hh = s.p0, radius = s.p1;
            float t1, t2;
            if (!cylinder_ray_quadratic_solve(r, radius, t1, t2)) return false;
            if (t1 < 0.f && t2 < 0.f) return false;
            float t = (t1 < 0.f) ? t2 : ((t2 < 0.f) ? t1 : rmin(t1, t2));
            Vec3 l(-r.origin.x, hh - r.origin.y, -r.origin.z);
            float tca = dot(l, r.direction);
            if (tca > 0.f) {
                float d2 = dot(l, l) - tca * tca;
                if (d2 <= radius * radius) {
                    float thc = std::sqrt(radius * radius - d2);
                    float t0 = tca - thc;
                    if (t0 >= 0.f && (t != t || t0 < t)) t = t0;
                }
            }
            l = Vec3(-r.origin.x, -hh - r.origin.y, -r.origin.z);
            tca = dot(l, r.direction);
            if (tca > 0.f) {
                float d2 = dot(l, l) - tca * tca;
                if (d2 <= radius * radius) {
                    float thc = std::sqrt(radius * radius - d2);
                    float t0 = tca - thc;
                    if (t0 >= 0.f && (t != t || t0 < t)) t = t0;
                }
            }
            if (t != t) return false;
            Vec3 pc = r.origin + r.direction * t;
            float full = hh + radius;
            if (pc.y > full || pc.y < -full) return false;
            out = pc;
            return true;
        }
        case POLYHEDRON: {  // polyhedron.rs:380-394
            if (!s.he || s.he->faces.empty()) return false;
            float u, v, w;
            const int fi = poly_find_intersecting_face(s, r, u, v, w);
            if (fi < 0) return false;
            const HalfEdgeMesh::Face& f = s.he->faces[(size_t)fi];
            auto pos = [&](uint32_t i) { return Vec3(s.verts[3 * i], s.verts[3 * i + 1], s.verts[3 * i + 2]); };
            out = (pos(f.v[0]) * u) + (pos(f.v[1]) * v) + (pos(f.v[2]) * w);
            return true;
        }
        default: return false;
    }
}

// ray.rs:58-66 DiscreteTransformed<Ray>, ray.rs:68-78 ContinuousTransformed<Ray>
inline bool ray_intersects_transformed(const Shape& s, const Ray& ray, const Mat4& transform) {
    return ray_intersects_local(s, ray.transform(transform.inverse()));
}
inline bool ray_intersection_transformed(const Shape& s, const Ray& ray, const Mat4& transform, Vec3& out) {
    Vec3 p;
    if (!ray_intersection_local(s, ray.transform(transform.inverse()), p)) return false;
    out = transform.transform_point3(p);
    return true;
}

// particle.rs:71-87 / :108-128: the particle moving from `start` to `end` (both Discrete and Continuous go through
// intersection_transformed and keep the hit only if it lies within the travelled distance)
inline bool particle_sweep_transformed(const Shape& s, Vec3 start, Vec3 end, const Mat4& transform, Vec3& out) {
    Vec3 direction = end - start;
    Ray ray(start, normalize(direction));
    Vec3 p;
    if (!ray_intersection_transformed(s, ray, transform, p)) return false;
    if (!(dot(p - start, p - start) <= dot(direction, direction))) return false;
    out = p;
    return true;
}

}  // namespace orc
