// Ray casts against primitives, one thread per (ray, shape) cast. Reference:
//   Discrete/ContinuousTransformed<Ray> (ray.rs:58-78): the ray goes to object space with transform.inverse(), the
//   primitive's Discrete/Continuous<Ray> runs there, the hit point comes back with transform.transform_point3;
//   Vec3 (ray.rs:35-57), Sphere (primitive/sphere.rs:45-76), Cuboid (cuboid.rs:85-103) and Quad (quad.rs:80-96) through
//   Aabb3 (volume/aabb/aabb3.rs:165-235), Cylinder (cylinder.rs:82-204), Capsule (capsule.rs:69-178), Particle
//   (particle.rs:53-69), cylinder_ray_quadratic_solve (primitive/util.rs:95-109), Primitive3's dispatch
//   (primitive3.rs:133-166) and the swept particle (Particle, Range<Vec3>) (particle.rs:39-128).
// Every expression is evaluated in the reference's (glam 0.8 scalar) order, unfused: results are bit-identical.
// ConvexPolyhedron (polyhedron.rs:373-522): the face walk over the half-edge structure, for meshes uploaded with faces
// (up to POLY_RAY_MAX_FACES of them); a VertexOnly polyhedron has no faces to walk (the reference indexes an empty list and
// panics) and is reported as unsupported.
#pragma once
#include "simplex.cuh"
#include "support.cuh"

namespace b3 {

constexpr uint8_t ST_UNSUPPORTED = 6;

// The translation column (w_axis.xyz) of glam's Mat4::inverse for the affine transform in `s` (bottom row 0 0 0 1),
// with the same cofactor expressions and evaluation order pack_shapes_kernel uses for the linear part.
B3_DEV V3 inverse_translation(const Shape& s) {
    const float m00 = s.X.x, m01 = s.X.y, m02 = s.X.z, m03 = 0.f;
    const float m10 = s.Y.x, m11 = s.Y.y, m12 = s.Y.z, m13 = 0.f;
    const float m20 = s.Z.x, m21 = s.Z.y, m22 = s.Z.z, m23 = 0.f;
    const float m30 = s.W.x, m31 = s.W.y, m32 = s.W.z, m33 = 1.f;
    const float coef00 = m22 * m33 - m32 * m23, coef04 = m21 * m33 - m31 * m23, coef08 = m21 * m32 - m31 * m22;
    const float coef10 = m11 * m32 - m31 * m12;
    const float coef12 = m20 * m33 - m30 * m23, coef16 = m20 * m32 - m30 * m22, coef18 = m10 * m32 - m30 * m12;
    const float coef20 = m20 * m31 - m30 * m21, coef22 = m10 * m31 - m30 * m11;
    const float inv0x = m11 * coef00 - m12 * coef04 + m13 * coef08;
    const float inv1x = m10 * coef00 - m12 * coef12 + m13 * coef16;
    const float inv2x = m10 * coef04 - m11 * coef12 + m13 * coef20;
    const float inv3x = m10 * coef08 - m11 * coef16 + m12 * coef20;
    const float inv3y = m00 * coef08 - m01 * coef16 + m02 * coef20;
    const float inv3z = m00 * coef10 - m01 * coef18 + m02 * coef22;
    const float i0x = inv0x * 1.0f, i1x = inv1x * -1.0f, i2x = inv2x * 1.0f, i3x = inv3x * -1.0f;
    const float i3y = inv3y * 1.0f, i3z = inv3z * -1.0f;
    const float dot1 = ((m00 * i0x + m01 * i1x) + m02 * i2x) + m03 * i3x;
    const float rcp = 1.0f / dot1;
    return v3(i3x * rcp, i3y * rcp, i3z * rcp);
}

struct RayD { V3 o, d; };

// Continuous<Aabb3> for Ray (aabb3.rs:165-195); `discrete` selects the Discrete form's acceptance rule (:205-229).
// The box is Aabb3::new(p1, p2) (aabb3.rs:26-31: component-wise min / max).
B3_DEV bool ray_box(const RayD& r, V3 p1, V3 p2, bool discrete, V3& out) {
    const V3 mn = v3(rmin(p1.x, p2.x), rmin(p1.y, p2.y), rmin(p1.z, p2.z));
    const V3 mx = v3(rmax(p1.x, p2.x), rmax(p1.y, p2.y), rmax(p1.z, p2.z));
    const V3 inv = v3(1.0f / r.d.x, 1.0f / r.d.y, 1.0f / r.d.z);
    float t1 = (mn.x - r.o.x) * inv.x, t2 = (mx.x - r.o.x) * inv.x;
    float tmin = rmin(t1, t2), tmax = rmax(t1, t2);
    t1 = (mn.y - r.o.y) * inv.y; t2 = (mx.y - r.o.y) * inv.y;
    tmin = rmax(tmin, rmin(t1, t2)); tmax = rmin(tmax, rmax(t1, t2));
    t1 = (mn.z - r.o.z) * inv.z; t2 = (mx.z - r.o.z) * inv.z;
    tmin = rmax(tmin, rmin(t1, t2)); tmax = rmin(tmax, rmax(t1, t2));
    if (discrete) return tmax >= tmin && (tmin >= 0.0f || tmax >= 0.0f);
    if ((tmin < 0.0f && tmax < 0.0f) || tmax < tmin) return false;
    const float t = (tmin >= 0.0f) ? tmin : tmax;
    out = r.o + r.d * t;
    return true;
}

// ray.rs:46-57
B3_DEV bool ray_point(const RayD& r, V3 p) {
    const V3 l = p - r.o;
    const float tca = dot(l, r.d);
    return tca > 0.f && (fabsf((tca * tca) - dot(l, l)) < F32_EPSILON);
}

// primitive/util.rs:95-109
B3_DEV bool cylinder_quadratic(const RayD& r, float radius, float& t1, float& t2) {
    const float a = r.d.x * r.d.x + r.d.z * r.d.z;
    const float b = 2.f * (r.d.x * r.o.x + r.d.z * r.o.z);
    const float c = r.o.x * r.o.x + r.o.z * r.o.z - radius * radius;
    const float dr = b * b - 4.f * a * c;
    if (dr < 0.f) return false;
    const float drsqrt = sqrtf(dr);
    t1 = (-b + drsqrt) / (2.f * a);
    t2 = (-b - drsqrt) / (2.f * a);
    return true;
}

B3_DEV float pick_t(float t1, float t2) { return (t1 < 0.f) ? t2 : ((t2 < 0.f) ? t1 : rmin(t1, t2)); }

// top / bottom cap sphere of a capsule centred at (0, cy, 0): capsule.rs:91-110 (discrete), :134-160 (continuous)
B3_DEV bool capsule_cap_discrete(const RayD& r, float cy, float radius) {
    const V3 l = v3(-r.o.x, cy - r.o.y, -r.o.z);
    const float tca = dot(l, r.d);
    if (tca > 0.f) {
        const float d2 = dot(l, l) - tca * tca;
        if (d2 <= radius * radius) return true;
    }
    return false;
}
B3_DEV void capsule_cap_continuous(const RayD& r, float cy, float radius, float& t) {
    const V3 l = v3(-r.o.x, cy - r.o.y, -r.o.z);
    const float tca = dot(l, r.d);
    if (tca > 0.f) {
        const float d2 = dot(l, l) - tca * tca;
        if (d2 <= radius * radius) {
            const float thc = sqrtf(radius * radius - d2);
            const float t0 = tca - thc;
            if (t0 >= 0.f && (t != t || t0 < t)) t = t0;
        }
    }
}
// cylinder cap plane with normal (0, sy, 0) (exactly +-Vec3::unit_y()) at height h0: cylinder.rs:114-129, :169-187
B3_DEV float cylinder_cap_t(const RayD& r, float h0, float sy) {
    const V3 n = v3(sy * 0.f, sy, sy * 0.f);
    return -(h0 + dot(r.o, n)) / dot(r.d, n);
}

constexpr int POLY_RAY_MAX_FACES = 4096;  // `checked` BitSet of the walk: 128 words of local memory per thread

// barycentric_point (primitive/util.rs:63-79): the same expressions as barycentric_vector (simplex.cuh)
// find_intersecting_face (polyhedron.rs:401-425) + intersect_ray_face (:497-521) + next_face_classify (:433-494).
// Returns the face index or -1; (u, v, w) = barycentric coordinates of the hit.
B3_DEV int poly_find_intersecting_face(const Shape& s, const RayD& ray, float& u, float& v, float& w) {
    const MeshHeader* h = mesh_header(s.verts);
    const uint32_t nfaces = h->n_faces;
    const uint4* __restrict__ faces = h->faces;
    const uint4* __restrict__ edges = h->edges;
    uint32_t checked[POLY_RAY_MAX_FACES / 32];
    for (uint32_t i = 0; i < (nfaces + 31) / 32; ++i) checked[i] = 0u;
    int face = 0;
    while (face >= 0) {
        const uint32_t fi = (uint32_t)face;
        checked[fi >> 5] |= 1u << (fi & 31);
        const uint4 fv = __ldg(faces + 2 * fi);
        const float4 pl = __ldg(reinterpret_cast<const float4*>(faces + 2 * fi + 1));
        const V3 n = v3(pl.x, pl.y, pl.z);
        bool have = false;
        if (dot(n, ray.d) < 0.f) {
            const float t = -(pl.w + dot(ray.o, n)) / dot(ray.d, n);  // Continuous<Ray> for Plane (plane.rs:117-129)
            if (!(t < 0.f)) {
                const V3 p = ray.o + ray.d * t;
                const float4 a = __ldg(s.verts + fv.x), b = __ldg(s.verts + fv.y), c = __ldg(s.verts + fv.z);
                barycentric_vector(p, v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), u, v, w);
                have = true;
                if (u >= 0.f && u <= 1.f && v >= 0.f && v <= 1.f && w >= 0.f && w <= 1.f) return face;
            }
        }
        int next = -1;
        if (nfaces < 10) {
            next = (fi == nfaces - 1) ? -1 : (int)fi + 1;
        } else if (!have) {
            uint32_t nx = fi + 1;
            while (nx < nfaces && ((checked[nx >> 5] >> (nx & 31)) & 1u)) ++nx;
            next = (nx == nfaces) ? -1 : (int)nx;
        } else {
            const uint32_t target = (u < 0.f) ? fv.z : ((v < 0.f) ? fv.x : fv.y);
            const uint4 fe = __ldg(edges + 2 * fv.w);  // (target_vertex, twin, next, previous)
            uint32_t e3[3];
            if (fe.x == target) { e3[0] = fv.w; e3[1] = fe.z; e3[2] = fe.w; }
            else if (__ldg(edges + 2 * fe.w).x == target) { e3[0] = fe.w; e3[1] = fv.w; e3[2] = fe.z; }
            else { e3[0] = fe.z; e3[1] = fe.w; e3[2] = fv.w; }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (next < 0) {
                    const uint32_t twin = __ldg(edges + 2 * e3[k]).y;
                    const uint32_t lf = __ldg(edges + 2 * twin + 1).x;
                    if (!((checked[lf >> 5] >> (lf & 31)) & 1u)) next = (int)lf;
                }
            }
            if (next < 0) {
                for (uint32_t i = 0; i < (nfaces + 31) / 32 && next < 0; ++i) {
                    uint32_t freebits = ~checked[i];
                    if (i == nfaces / 32 && (nfaces & 31)) freebits &= (1u << (nfaces & 31)) - 1u;
                    if (i * 32 >= nfaces) freebits = 0u;
                    if (freebits) next = (int)(i * 32 + __ffs(freebits) - 1);
                }
            }
        }
        face = next;
    }
    return -1;
}

// Discrete<Ray> in object space
B3_DEV bool ray_intersects_local(const Shape& s, const RayD& r) {
    V3 unused;
    switch (s.kind) {
        case K_PARTICLE: return ray_point(r, v3(0.f, 0.f, 0.f));
        case K_SPHERE: {
            const V3 l = v3(-r.o.x, -r.o.y, -r.o.z);
            const float tca = dot(l, r.d);
            if (tca < 0.f) return false;
            const float d2 = dot(l, l) - tca * tca;
            return d2 <= s.p0 * s.p0;
        }
        case K_CUBOID: return ray_box(r, v3(-s.p0, -s.p1, -s.p2), v3(s.p0, s.p1, s.p2), true, unused);
        case K_QUAD: return ray_box(r, v3(-s.p0, -s.p1, 0.f), v3(s.p0, s.p1, 0.f), true, unused);
        case K_CYLINDER: {
            const float hh = s.p0, radius = s.p1;
            if (fabsf(r.d.x - 0.f) < F32_EPSILON && fabsf(r.d.z - 0.f) < F32_EPSILON) {
                if (fabsf(r.d.y - 0.f) < F32_EPSILON) return false;
                return (r.o.y >= -hh && r.d.y <= 0.f) || (r.o.y <= hh && r.d.y >= 0.f);
            }
            float t1, t2;
            if (!cylinder_quadratic(r, radius, t1, t2)) return false;
            if (t1 < 0.f && t2 < 0.f) return false;
            const float t = pick_t(t1, t2);
            const V3 pc = r.o + r.d * t;
            if (pc.y <= hh && pc.y >= -hh) return true;
            const float tp = cylinder_cap_t(r, hh, -1.f);
            if (tp >= 0.f) {
                const V3 p = r.o + r.d * tp;
                if (p.x * p.x + p.z * p.z < radius * radius) return true;
            }
            const float tb = cylinder_cap_t(r, -hh, 1.f);
            if (tb >= 0.f) {
                const V3 p = r.o + r.d * tb;
                if (p.x * p.x + p.z * p.z < radius * radius) return true;
            }
            return false;
        }
        case K_CAPSULE: {
            const float hh = s.p0, radius = s.p1;
            float t1, t2;
            if (!cylinder_quadratic(r, radius, t1, t2)) return false;
            if (t1 < 0.f && t2 < 0.f) return false;
            const float t = pick_t(t1, t2);
            const V3 pc = r.o + r.d * t;
            if (pc.y <= hh && pc.y >= -hh) return true;
            return capsule_cap_discrete(r, hh, radius) || capsule_cap_discrete(r, -hh, radius);
        }
        case K_POLYHEDRON: {  // polyhedron.rs:373-378
            float u, v, w;
            return poly_find_intersecting_face(s, r, u, v, w) >= 0;
        }
        default: return false;
    }
}

// Continuous<Ray> in object space
B3_DEV bool ray_intersection_local(const Shape& s, const RayD& r, V3& out) {
    switch (s.kind) {
        case K_PARTICLE:
            if (ray_point(r, v3(0.f, 0.f, 0.f))) { out = v3(0.f, 0.f, 0.f); return true; }
            return false;
        case K_SPHERE: {
            const V3 l = v3(-r.o.x, -r.o.y, -r.o.z);
            const float tca = dot(l, r.d);
            if (tca < 0.f) return false;
            const float d2 = dot(l, l) - tca * tca;
            if (d2 > s.p0 * s.p0) return false;
            const float thc = sqrtf(s.p0 * s.p0 - d2);
            out = r.o + r.d * (tca - thc);
            return true;
        }
        case K_CUBOID: return ray_box(r, v3(-s.p0, -s.p1, -s.p2), v3(s.p0, s.p1, s.p2), false, out);
        case K_QUAD: return ray_box(r, v3(-s.p0, -s.p1, 0.f), v3(s.p0, s.p1, 0.f), false, out);
        case K_CYLINDER: {
            const float hh = s.p0, radius = s.p1;
            if (fabsf(r.d.x - 0.f) < F32_EPSILON && fabsf(r.d.z - 0.f) < F32_EPSILON) {
                if (fabsf(r.d.y - 0.f) < F32_EPSILON) return false;
                if (r.o.y >= hh && r.d.y < 0.f) { out = v3(0.f, hh, 0.f); return true; }
                if (r.o.y >= -hh && r.d.y < 0.f) { out = v3(0.f, -hh, 0.f); return true; }
                if (r.o.y <= -hh && r.d.y > 0.f) { out = v3(0.f, -hh, 0.f); return true; }
                if (r.o.y <= hh && r.d.y > 0.f) { out = v3(0.f, hh, 0.f); return true; }
                return false;
            }
            float t1, t2;
            if (!cylinder_quadratic(r, radius, t1, t2)) return false;
            if (t1 < 0.f && t2 < 0.f) return false;
            float t = pick_t(t1, t2);
            const float tp = cylinder_cap_t(r, hh, -1.f);
            if (tp >= 0.f && tp < t) {
                const V3 p = r.o + r.d * tp;
                if (p.x * p.x + p.z * p.z < radius * radius) t = tp;
            }
            const float tb = cylinder_cap_t(r, -hh, 1.f);
            if (tb >= 0.f && tb < t) {
                const V3 p = r.o + r.d * tb;
                if (p.x * p.x + p.z * p.z < radius * radius) t = tb;
            }
            const V3 pc = r.o + r.d * t;
            if (pc.y > hh || pc.y < -hh) return false;
            out = pc;
            return true;
        }
        case K_CAPSULE: {
            const float hh = s.p0, radius = s.p1;
            float t1, t2;
            if (!cylinder_quadratic(r, radius, t1, t2)) return false;
            if (t1 < 0.f && t2 < 0.f) return false;
            float t = pick_t(t1, t2);
            capsule_cap_continuous(r, hh, radius, t);
            capsule_cap_continuous(r, -hh, radius, t);
            if (t != t) return false;
            const V3 pc = r.o + r.d * t;
            const float full = hh + radius;
            if (pc.y > full || pc.y < -full) return false;
            out = pc;
            return true;
        }
        case K_POLYHEDRON: {  // polyhedron.rs:380-394
            float u, v, w;
            const int fi = poly_find_intersecting_face(s, r, u, v, w);
            if (fi < 0) return false;
            const uint4 fv = __ldg(mesh_header(s.verts)->faces + 2 * fi);
            const float4 a = __ldg(s.verts + fv.x), b = __ldg(s.verts + fv.y), c = __ldg(s.verts + fv.z);
            out = (v3(a.x, a.y, a.z) * u) + (v3(b.x, b.y, b.z) * v) + (v3(c.x, c.y, c.z) * w);
            return true;
        }
        default: return false;
    }
}

// query: 0 = intersects_transformed, 1 = intersection_transformed, 2 = swept particle (a = start, b = end)
B3_DEV uint8_t ray_cast_one(const Shape& s, V3 a, V3 b, int query, V3& out) {
    out = v3(0.f, 0.f, 0.f);
    if (s.kind == K_POLYHEDRON) {
        const uint32_t nfaces = (s.nverts & MESH_HALF_EDGE) ? mesh_header(s.verts)->n_faces : 0u;
        if (nfaces == 0 || nfaces > (uint32_t)POLY_RAY_MAX_FACES) return ST_UNSUPPORTED;
    }
    RayD world;
    V3 travel = v3(0.f, 0.f, 0.f);
    if (query == 2) { travel = b - a; world.o = a; world.d = normalize(travel); }
    else { world.o = a; world.d = b; }
    // Ray::transform(transform.inverse()) (ray.rs:26-31): transform_point3 on the origin, transform_vector3 on the direction
    const V3 iw = inverse_translation(s);
    RayD local;
    V3 t = s.iX * world.o.x;
    t = s.iY * world.o.y + t;
    t = s.iZ * world.o.z + t;
    local.o = iw + t;
    local.d = inv_transform_vector3(s, world.d);
    if (query == 0) return ray_intersects_local(s, local) ? ST_OK : ST_MISS;
    V3 p;
    if (!ray_intersection_local(s, local, p)) return ST_MISS;
    p = transform_point3(s, p);
    if (query == 2) {
        const V3 dp = p - a;
        if (!(dot(dp, dp) <= dot(travel, travel))) return ST_MISS;
    }
    out = p;
    return ST_OK;
}

}  // namespace b3
