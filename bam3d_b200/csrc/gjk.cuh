// GJK drivers on the device: intersect, distance, time of impact. Reference: `impl GJK` in
// algorithm/minkowski/gjk/mod.rs (intersect :74-113, intersection_time_of_impact :130-217, distance :232-280).
#pragma once
#include "epa.cuh"

namespace b3 {

struct GjkSettings {  // GJK::new_with_settings (gjk/mod.rs:45-58); max_iterations is shared by GJK and EPA
    float distance_tolerance, continuous_tolerance, epa_tolerance;
    uint32_t max_iterations;
};

// transform.transform_point3(Vec3::zero()) evaluated literally (gjk/mod.rs:85-86)
B3_DEV V3 shape_origin(const Shape& s) { return transform_point3(s, v3(0.f, 0.f, 0.f)); }

// Out-of-line copy of the simplex processor for the warp-per-pair kernels (instruction-fetch bound: C3 22.6 -> 18.1 ms with
// this and the un-unrolled edge loop of epa_warp.cuh; outlining EPA3::process as a whole instead made it 31 ms).
template <bool WITH_A>
__device__ __noinline__ bool reduce_to_closest_feature_outlined(Simplex<WITH_A>& s, V3& v) { return reduce_to_closest_feature(s, v); }

// GJK::intersect (gjk/mod.rs:74-113). On a hit the enclosing tetrahedron is left in `s`.
template <bool WARP, bool WITH_A>
B3_DEV uint8_t gjk_intersect(const Shape& L, const Shape& R, uint32_t max_iterations, Simplex<WITH_A>& s) {
    V3 d = shape_origin(R) - shape_origin(L);
    if (is_zero(d)) d = v3(1.f, 1.f, 1.f);
    V3 pv, pa;
    minkowski_support<WARP>(L, R, d, pv, pa);
    if (dot(pv, d) <= 0.f) return ST_MISS;
    s.clear();
    s.push(pv, pa);
    d = -d;
    for (uint32_t it = 0; it < max_iterations; ++it) {
        minkowski_support<WARP>(L, R, d, pv, pa);
        if (dot(pv, d) <= 0.f) return ST_MISS;
        s.push(pv, pa);
        if (WARP ? reduce_to_closest_feature_outlined(s, d) : reduce_to_closest_feature(s, d)) return ST_OK;
    }
    return ST_MAX_ITER;  // the reference returns None here too (gjk/mod.rs:112)
}

// GJK::distance (gjk/mod.rs:232-280). ST_OK => Some: ref_value = dp - d0 (what the reference returns: the
// convergence residual), geometric = sqrt(d.d) (the return the reference has commented out at :274).
// ST_MISS => None because the shapes collide; ST_MAX_ITER => None after max_iterations.
template <bool WARP>
B3_DEV uint8_t gjk_distance(const Shape& L, const Shape& R, float distance_tolerance, uint32_t max_iterations,
                            float& ref_value, float& geometric) {
    V3 d = shape_origin(R) - shape_origin(L);
    if (fabsf(d.x - 0.f) < F32_EPSILON && fabsf(d.y - 0.f) < F32_EPSILON && fabsf(d.z - 0.f) < F32_EPSILON) d = v3(1.f, 1.f, 1.f);
    Simplex<false> s;
    s.clear();
    V3 pv, pa;
    minkowski_support<WARP>(L, R, d, pv, pa);
    s.push(pv, pa);
    minkowski_support<WARP>(L, R, -d, pv, pa);
    s.push(pv, pa);
    for (uint32_t it = 0; it < max_iterations; ++it) {
        uint8_t st = ST_OK;
        V3 c = closest_point_to_origin(s, st);
        if (st != ST_OK) return st;
        if (fabsf(c.x - 0.f) < F32_EPSILON && fabsf(c.y - 0.f) < F32_EPSILON && fabsf(c.z - 0.f) < F32_EPSILON) return ST_MISS;
        V3 nd = -c;
        minkowski_support<WARP>(L, R, nd, pv, pa);
        float dp = dot(pv, nd);
        float d0 = dot(s.v[0], nd);
        if (dp - d0 < distance_tolerance) {
            ref_value = dp - d0;
            geometric = sqrtf(dot(nd, nd));
            return ST_OK;
        }
        if (s.n >= 4) return ST_REF_PANIC;  // unreachable: a non-enclosing reduction leaves <= 3 points
        s.push(pv, pa);
    }
    return ST_MAX_ITER;
}

// Quat::from_rotation_axes (glam 0.8) — only used to rebuild the right shape's transform at the time of impact.
B3_DEV void quat_from_axes(V3 xa, V3 ya, V3 za, float& qx, float& qy, float& qz, float& qw) {
    float m00 = xa.x, m01 = xa.y, m02 = xa.z, m10 = ya.x, m11 = ya.y, m12 = ya.z, m20 = za.x, m21 = za.y, m22 = za.z;
    if (m22 <= 0.0f) {
        float dif10 = m11 - m00, omm22 = 1.0f - m22;
        if (dif10 <= 0.0f) {
            float f = omm22 - dif10, inv = 0.5f / sqrtf(f);
            qx = f * inv; qy = (m01 + m10) * inv; qz = (m02 + m20) * inv; qw = (m12 - m21) * inv;
        } else {
            float f = omm22 + dif10, inv = 0.5f / sqrtf(f);
            qx = (m01 + m10) * inv; qy = f * inv; qz = (m12 + m21) * inv; qw = (m20 - m02) * inv;
        }
    } else {
        float sum10 = m11 + m00, opm22 = 1.0f + m22;
        if (sum10 <= 0.0f) {
            float f = opm22 - sum10, inv = 0.5f / sqrtf(f);
            qx = (m02 + m20) * inv; qy = (m12 + m21) * inv; qz = f * inv; qw = (m01 - m10) * inv;
        } else {
            float f = opm22 + sum10, inv = 0.5f / sqrtf(f);
            qx = (m12 - m21) * inv; qy = (m20 - m02) * inv; qz = (m01 - m10) * inv; qw = f * inv;
        }
    }
}

// right_start.translation_interpolate(right_end, lambda).transform_point3(p)  (traits.rs:180-190,
// gjk/mod.rs:203-210): decompose the END transform into scale / rotation, lerp the translations, recompose
// (glam to_scale_rotation_translation / from_scale_rotation_translation), then transform p. Affine input:
// the 4th row is (0,0,0,1), so |axis| over 4 components equals |axis.xyz| and det(Mat4) = det(3x3).
B3_DEV V3 toi_contact_point(const Shape& R0, const Shape& R1, float lambda, V3 p) {
    // determinant by cofactor expansion along the first column, with the affine bottom row folded in
    V3 X = R1.X, Y = R1.Y, Z = R1.Z;
    float a2323 = Z.z * 1.0f - 0.0f * R1.W.z;
    float a1323 = Z.y * 1.0f - 0.0f * R1.W.y;
    float a1223 = Z.y * R1.W.z - Z.z * R1.W.y;
    float a0323 = Z.x * 1.0f - 0.0f * R1.W.x;
    float a0223 = Z.x * R1.W.z - Z.z * R1.W.x;
    float a0123 = Z.x * R1.W.y - Z.y * R1.W.x;
    float det = X.x * (Y.y * a2323 - Y.z * a1323 + 0.0f * a1223) - X.y * (Y.x * a2323 - Y.z * a0323 + 0.0f * a0223) +
                X.z * (Y.x * a1323 - Y.y * a0323 + 0.0f * a0123) - 0.0f * (Y.x * a1223 - Y.y * a0223 + Y.z * a0123);
    auto len4 = [](V3 v) { return sqrtf(((v.x * v.x + v.y * v.y) + v.z * v.z) + 0.0f * 0.0f); };
    float sx = len4(X) * rsignum(det), sy = len4(Y), sz = len4(Z);
    float isx = 1.0f / sx, isy = 1.0f / sy, isz = 1.0f / sz;
    float x, y, z, w;
    quat_from_axes(X * isx, Y * isy, Z * isz, x, y, z, w);
    V3 t = R0.W + ((R1.W - R0.W) * lambda);  // Vec3::lerp
    // quat_to_axes + scale
    float x2 = x + x, y2 = y + y, z2 = z + z;
    float xx = x * x2, xy = x * y2, xz = x * z2, yy = y * y2, yz = y * z2, zz = z * z2, wx = w * x2, wy = w * y2, wz = w * z2;
    Shape T;
    T.X = v3(1.0f - (yy + zz), xy + wz, xz - wy) * sx;
    T.Y = v3(xy - wz, 1.0f - (xx + zz), yz + wx) * sy;
    T.Z = v3(xz + wy, yz - wx, 1.0f - (xx + yy)) * sz;
    T.W = t;
    return transform_point3(T, p);
}

// GJK::intersection_time_of_impact (gjk/mod.rs:130-217). L0/R0 = start transforms, L1/R1 = end transforms.
// The reference's `while` has no iteration cap (:166); `toi_cap` bounds it here and reports ST_MAX_ITER.
template <bool WARP>
B3_DEV uint8_t gjk_time_of_impact(const Shape& L0, const Shape& L1, const Shape& R0, const Shape& R1,
                                  float continuous_tolerance, uint32_t toi_cap, ContactOut& out, float& toi) {
    V3 left_lin_vel = shape_origin(L1) - shape_origin(L0);
    V3 right_lin_vel = shape_origin(R1) - shape_origin(R0);
    V3 ray = right_lin_vel - left_lin_vel;
    float lambda = 0.f;
    V3 normal = v3(0.f, 0.f, 0.f);
    V3 ray_origin = v3(0.f, 0.f, 0.f);
    Simplex<false> s;
    s.clear();
    V3 pv, pa;
    minkowski_support<WARP>(L0, R0, -ray, pv, pa);
    V3 v = pv;
    uint32_t iters = 0;
    while (dot(v, v) > continuous_tolerance) {
        if (iters++ >= toi_cap) return ST_MAX_ITER;
        minkowski_support<WARP>(L0, R0, -v, pv, pa);
        float vp = dot(v, pv);
        float vr = dot(v, ray);
        if (vp > lambda * vr) {
            if (vr > 0.f) {
                lambda = vp / vr;
                if (lambda > 1.f) return ST_MISS;
                ray_origin = ray * lambda;
                s.clear();
                normal = -v;
            } else {
                return ST_MISS;
            }
        }
        if (s.n >= 4) return ST_REF_PANIC;  // unreachable, see gjk_distance
        s.push(pv - ray_origin, pa);
        uint8_t st = ST_OK;
        v = closest_point_to_origin(s, st);
        if (st != ST_OK) return st;
    }
    if (dot(v, v) <= continuous_tolerance) {
        out.normal = -normalize(normal);
        out.depth = sqrtf(dot(v, v));
        out.point = toi_contact_point(R0, R1, lambda, ray_origin);
        toi = lambda;
        return ST_OK;
    }
    return ST_NAN;  // v.v is NaN: the loop test and the <= test both fail -> None
}

}  // namespace b3
