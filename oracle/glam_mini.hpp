// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into, imported by or executed from the product
// path (bam3d_b200/, include/). Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use anything under oracle/.
//
// glam_mini.hpp — strict-f32 restatement of the subset of `glam = "0.8"` (feature `scalar-math`,
// reference Cargo.toml:10) that the bam3d hot path calls. glam is an un-vendored crates.io dependency
// (its source is NOT under /root/reference and Cargo.lock is git-ignored, reference .gitignore:2), so the
// formulas below restate glam 0.8's published scalar implementation; they are pinned by the reference's
// own known-answer tests that pass through them (cylinder.rs:257-264, capsule.rs:231-238,
// sphere.rs:86-123, gjk/mod.rs:576-644, epa3d.rs:216-247) — see oracle/kat_main.cpp.
// UNPINNED (no reference KAT distinguishes the alternatives): Vec3::normalize as v*(1/len) vs v/len,
// to_scale_rotation_translation / quaternion extraction beyond identity rotation.
//
// Build with: -O2 -ffp-contract=off -fno-fast-math (no FMA contraction; IEEE sqrt/div).
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace orc {

constexpr float F32_EPSILON = 1.1920929e-7f;  // Rust std::f32::EPSILON

// Rust f32::min / f32::max: if one operand is NaN the other is returned.
inline float rmin(float a, float b) { return (a != a) ? b : ((b != b) ? a : ((b < a) ? b : a)); }
inline float rmax(float a, float b) { return (a != a) ? b : ((b != b) ? a : ((b > a) ? b : a)); }
// Rust f32::signum: NaN -> NaN, else copysign(1, x) (so +0 -> 1, -0 -> -1).
inline float rsignum(float x) { return (x != x) ? x : std::copysign(1.0f, x); }

struct Vec3 {
    float x, y, z;
    Vec3() : x(0.f), y(0.f), z(0.f) {}
    Vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    static Vec3 zero() { return Vec3(0.f, 0.f, 0.f); }
    static Vec3 splat(float v) { return Vec3(v, v, v); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};

inline Vec3 operator+(Vec3 a, Vec3 b) { return Vec3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 operator-(Vec3 a, Vec3 b) { return Vec3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 operator*(Vec3 a, Vec3 b) { return Vec3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline Vec3 operator/(Vec3 a, Vec3 b) { return Vec3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline Vec3 operator*(Vec3 a, float s) { return Vec3(a.x * s, a.y * s, a.z * s); }
inline Vec3 operator/(Vec3 a, float s) { return Vec3(a.x / s, a.y / s, a.z / s); }
inline Vec3 operator-(Vec3 a) { return Vec3(-a.x, -a.y, -a.z); }
// Vec3::cmpeq(..).all() and PartialEq: component-wise ==  (so -0 == +0, NaN != NaN)
inline bool all_eq(Vec3 a, Vec3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }
// glam 0.8 scalar Vec3::dot: (x*x' + y*y') + z*z'
inline float dot(Vec3 a, Vec3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// glam 0.8 scalar Vec3::cross
inline Vec3 cross(Vec3 a, Vec3 b) {
    return Vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
inline float length(Vec3 a) { return std::sqrt(dot(a, a)); }
// glam 0.8: normalize = self * self.length_reciprocal(), length_reciprocal = 1.0 / length  (UNPINNED, see header)
inline Vec3 normalize(Vec3 a) { return a * (1.0f / length(a)); }
// glam Vec3::lerp: self + ((other - self) * s)
inline Vec3 lerp(Vec3 a, Vec3 b, float s) { return a + ((b - a) * s); }

struct Vec4 {
    float x, y, z, w;
    Vec4() : x(0.f), y(0.f), z(0.f), w(0.f) {}
    Vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    Vec3 truncate() const { return Vec3(x, y, z); }
};
inline Vec4 operator+(Vec4 a, Vec4 b) { return Vec4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
inline Vec4 operator-(Vec4 a, Vec4 b) { return Vec4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
inline Vec4 operator*(Vec4 a, Vec4 b) { return Vec4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
inline Vec4 operator*(Vec4 a, float s) { return Vec4(a.x * s, a.y * s, a.z * s, a.w * s); }

struct Quat {
    float x, y, z, w;
    Quat() : x(0.f), y(0.f), z(0.f), w(1.f) {}
    Quat(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    // glam: from_rotation_{x,y,z}(angle): (s, c) = (angle * 0.5).sin_cos()
    static Quat from_rotation_x(float a) { float h = a * 0.5f; return Quat(std::sin(h), 0.f, 0.f, std::cos(h)); }
    static Quat from_rotation_y(float a) { float h = a * 0.5f; return Quat(0.f, std::sin(h), 0.f, std::cos(h)); }
    static Quat from_rotation_z(float a) { float h = a * 0.5f; return Quat(0.f, 0.f, std::sin(h), std::cos(h)); }
};

// Rust f32::to_radians: self * (PI / 180)
inline float to_radians(float deg) { return deg * (3.14159265358979323846f / 180.0f); }

struct Mat4 {
    Vec4 x_axis, y_axis, z_axis, w_axis;  // columns

    static Mat4 identity() {
        Mat4 m;
        m.x_axis = Vec4(1, 0, 0, 0); m.y_axis = Vec4(0, 1, 0, 0);
        m.z_axis = Vec4(0, 0, 1, 0); m.w_axis = Vec4(0, 0, 0, 1);
        return m;
    }
    static Mat4 from_cols_array(const float* a) {
        Mat4 m;
        m.x_axis = Vec4(a[0], a[1], a[2], a[3]);   m.y_axis = Vec4(a[4], a[5], a[6], a[7]);
        m.z_axis = Vec4(a[8], a[9], a[10], a[11]); m.w_axis = Vec4(a[12], a[13], a[14], a[15]);
        return m;
    }
    void to_cols_array(float* a) const {
        a[0] = x_axis.x; a[1] = x_axis.y; a[2] = x_axis.z; a[3] = x_axis.w;
        a[4] = y_axis.x; a[5] = y_axis.y; a[6] = y_axis.z; a[7] = y_axis.w;
        a[8] = z_axis.x; a[9] = z_axis.y; a[10] = z_axis.z; a[11] = z_axis.w;
        a[12] = w_axis.x; a[13] = w_axis.y; a[14] = w_axis.z; a[15] = w_axis.w;
    }

    // glam 0.8 quat_to_axes + from_scale_rotation_translation
    static Mat4 from_scale_rotation_translation(Vec3 scale, Quat q, Vec3 t) {
        float x = q.x, y = q.y, z = q.z, w = q.w;
        float x2 = x + x, y2 = y + y, z2 = z + z;
        float xx = x * x2, xy = x * y2, xz = x * z2;
        float yy = y * y2, yz = y * z2, zz = z * z2;
        float wx = w * x2, wy = w * y2, wz = w * z2;
        Vec4 xa(1.0f - (yy + zz), xy + wz, xz - wy, 0.0f);
        Vec4 ya(xy - wz, 1.0f - (xx + zz), yz + wx, 0.0f);
        Vec4 za(xz + wy, yz - wx, 1.0f - (xx + yy), 0.0f);
        Mat4 m;
        m.x_axis = xa * scale.x; m.y_axis = ya * scale.y; m.z_axis = za * scale.z;
        m.w_axis = Vec4(t.x, t.y, t.z, 1.0f);
        return m;
    }

    // glam 0.8: r = X*px; r = Y*py + r; r = Z*pz + r; r = W + r   (no perspective divide)
    Vec3 transform_point3(Vec3 p) const {
        Vec3 r = x_axis.truncate() * p.x;
        r = y_axis.truncate() * p.y + r;
        r = z_axis.truncate() * p.z + r;
        r = w_axis.truncate() + r;
        return r;
    }
    Vec3 transform_vector3(Vec3 v) const {
        Vec3 r = x_axis.truncate() * v.x;
        r = y_axis.truncate() * v.y + r;
        r = z_axis.truncate() * v.z + r;
        return r;
    }
    Vec4 mul_vec4(Vec4 v) const {
        Vec4 r = x_axis * v.x;
        r = y_axis * v.y + r;
        r = z_axis * v.z + r;
        r = w_axis * v.w + r;
        return r;
    }
    Mat4 mul_mat4(const Mat4& o) const {
        Mat4 m;
        m.x_axis = mul_vec4(o.x_axis); m.y_axis = mul_vec4(o.y_axis);
        m.z_axis = mul_vec4(o.z_axis); m.w_axis = mul_vec4(o.w_axis);
        return m;
    }
    Mat4 transpose() const {
        Mat4 m;
        m.x_axis = Vec4(x_axis.x, y_axis.x, z_axis.x, w_axis.x);
        m.y_axis = Vec4(x_axis.y, y_axis.y, z_axis.y, w_axis.y);
        m.z_axis = Vec4(x_axis.z, y_axis.z, z_axis.z, w_axis.z);
        m.w_axis = Vec4(x_axis.w, y_axis.w, z_axis.w, w_axis.w);
        return m;
    }

    // glam 0.8 scalar Mat4::inverse: GLM-style cofactor expansion.
    Mat4 inverse() const {
        float m00 = x_axis.x, m01 = x_axis.y, m02 = x_axis.z, m03 = x_axis.w;
        float m10 = y_axis.x, m11 = y_axis.y, m12 = y_axis.z, m13 = y_axis.w;
        float m20 = z_axis.x, m21 = z_axis.y, m22 = z_axis.z, m23 = z_axis.w;
        float m30 = w_axis.x, m31 = w_axis.y, m32 = w_axis.z, m33 = w_axis.w;

        float coef00 = m22 * m33 - m32 * m23;
        float coef02 = m12 * m33 - m32 * m13;
        float coef03 = m12 * m23 - m22 * m13;

        float coef04 = m21 * m33 - m31 * m23;
        float coef06 = m11 * m33 - m31 * m13;
        float coef07 = m11 * m23 - m21 * m13;

        float coef08 = m21 * m32 - m31 * m22;
        float coef10 = m11 * m32 - m31 * m12;
        float coef11 = m11 * m22 - m21 * m12;

        float coef12 = m20 * m33 - m30 * m23;
        float coef14 = m10 * m33 - m30 * m13;
        float coef15 = m10 * m23 - m20 * m13;

        float coef16 = m20 * m32 - m30 * m22;
        float coef18 = m10 * m32 - m30 * m12;
        float coef19 = m10 * m22 - m20 * m12;

        float coef20 = m20 * m31 - m30 * m21;
        float coef22 = m10 * m31 - m30 * m11;
        float coef23 = m10 * m21 - m20 * m11;

        Vec4 fac0(coef00, coef00, coef02, coef03);
        Vec4 fac1(coef04, coef04, coef06, coef07);
        Vec4 fac2(coef08, coef08, coef10, coef11);
        Vec4 fac3(coef12, coef12, coef14, coef15);
        Vec4 fac4(coef16, coef16, coef18, coef19);
        Vec4 fac5(coef20, coef20, coef22, coef23);

        Vec4 vec0(m10, m00, m00, m00);
        Vec4 vec1(m11, m01, m01, m01);
        Vec4 vec2(m12, m02, m02, m02);
        Vec4 vec3(m13, m03, m03, m03);

        Vec4 inv0 = vec1 * fac0 - vec2 * fac1 + vec3 * fac2;
        Vec4 inv1 = vec0 * fac0 - vec2 * fac3 + vec3 * fac4;
        Vec4 inv2 = vec0 * fac1 - vec1 * fac3 + vec3 * fac5;
        Vec4 inv3 = vec0 * fac2 - vec1 * fac4 + vec2 * fac5;

        Vec4 sign_a(1.0f, -1.0f, 1.0f, -1.0f);
        Vec4 sign_b(-1.0f, 1.0f, -1.0f, 1.0f);

        Mat4 inv;
        inv.x_axis = inv0 * sign_a; inv.y_axis = inv1 * sign_b;
        inv.z_axis = inv2 * sign_a; inv.w_axis = inv3 * sign_b;

        Vec4 col0(inv.x_axis.x, inv.y_axis.x, inv.z_axis.x, inv.w_axis.x);
        Vec4 dot0 = x_axis * col0;
        float dot1 = ((dot0.x + dot0.y) + dot0.z) + dot0.w;
        float rcp_det = 1.0f / dot1;
        inv.x_axis = inv.x_axis * rcp_det; inv.y_axis = inv.y_axis * rcp_det;
        inv.z_axis = inv.z_axis * rcp_det; inv.w_axis = inv.w_axis * rcp_det;
        return inv;
    }

    // glam 0.8 Mat4::determinant (cofactor expansion along the first column)
    float determinant() const {
        float m00 = x_axis.x, m01 = x_axis.y, m02 = x_axis.z, m03 = x_axis.w;
        float m10 = y_axis.x, m11 = y_axis.y, m12 = y_axis.z, m13 = y_axis.w;
        float m20 = z_axis.x, m21 = z_axis.y, m22 = z_axis.z, m23 = z_axis.w;
        float m30 = w_axis.x, m31 = w_axis.y, m32 = w_axis.z, m33 = w_axis.w;
        float a2323 = m22 * m33 - m23 * m32;
        float a1323 = m21 * m33 - m23 * m31;
        float a1223 = m21 * m32 - m22 * m31;
        float a0323 = m20 * m33 - m23 * m30;
        float a0223 = m20 * m32 - m22 * m30;
        float a0123 = m20 * m31 - m21 * m30;
        return m00 * (m11 * a2323 - m12 * a1323 + m13 * a1223)
             - m01 * (m10 * a2323 - m12 * a0323 + m13 * a0223)
             + m02 * (m10 * a1323 - m11 * a0323 + m13 * a0123)
             - m03 * (m10 * a1223 - m11 * a0223 + m12 * a0123);
    }

    // glam 0.8 Quat::from_rotation_axes (DirectXMath XMQuaternionRotationMatrix branches). UNPINNED.
    static Quat quat_from_axes(Vec3 xa, Vec3 ya, Vec3 za) {
        float m00 = xa.x, m01 = xa.y, m02 = xa.z;
        float m10 = ya.x, m11 = ya.y, m12 = ya.z;
        float m20 = za.x, m21 = za.y, m22 = za.z;
        if (m22 <= 0.0f) {
            float dif10 = m11 - m00;
            float omm22 = 1.0f - m22;
            if (dif10 <= 0.0f) {
                float four_xsq = omm22 - dif10;
                float inv4x = 0.5f / std::sqrt(four_xsq);
                return Quat(four_xsq * inv4x, (m01 + m10) * inv4x, (m02 + m20) * inv4x, (m12 - m21) * inv4x);
            } else {
                float four_ysq = omm22 + dif10;
                float inv4y = 0.5f / std::sqrt(four_ysq);
                return Quat((m01 + m10) * inv4y, four_ysq * inv4y, (m12 + m21) * inv4y, (m20 - m02) * inv4y);
            }
        } else {
            float sum10 = m11 + m00;
            float opm22 = 1.0f + m22;
            if (sum10 <= 0.0f) {
                float four_zsq = opm22 - sum10;
                float inv4z = 0.5f / std::sqrt(four_zsq);
                return Quat((m02 + m20) * inv4z, (m12 + m21) * inv4z, four_zsq * inv4z, (m01 - m10) * inv4z);
            } else {
                float four_wsq = opm22 + sum10;
                float inv4w = 0.5f / std::sqrt(four_wsq);
                return Quat((m12 - m21) * inv4w, (m20 - m02) * inv4w, (m01 - m10) * inv4w, four_wsq * inv4w);
            }
        }
    }

    // glam 0.8 Mat4::to_scale_rotation_translation. UNPINNED beyond identity rotation.
    void to_scale_rotation_translation(Vec3& scale, Quat& rot, Vec3& trans) const {
        float det = determinant();
        auto len4 = [](Vec4 v) { return std::sqrt(((v.x * v.x + v.y * v.y) + v.z * v.z) + v.w * v.w); };
        scale = Vec3(len4(x_axis) * rsignum(det), len4(y_axis), len4(z_axis));
        Vec3 inv_scale(1.0f / scale.x, 1.0f / scale.y, 1.0f / scale.z);
        rot = quat_from_axes(x_axis.truncate() * inv_scale.x, y_axis.truncate() * inv_scale.y,
                             z_axis.truncate() * inv_scale.z);
        trans = w_axis.truncate();
    }

    // glam 0.8 Mat4::perspective_rh_gl (used only by the reference's frustum/dbvt tests)
    static Mat4 perspective_rh_gl(float fov_y, float aspect, float z_near, float z_far) {
        float inv_length = 1.0f / (z_near - z_far);
        float f = 1.0f / std::tan(0.5f * fov_y);
        float a = f / aspect;
        float b = (z_near + z_far) * inv_length;
        float c = (2.0f * z_near * z_far) * inv_length;
        Mat4 m;
        m.x_axis = Vec4(a, 0, 0, 0); m.y_axis = Vec4(0, f, 0, 0);
        m.z_axis = Vec4(0, 0, b, -1.0f); m.w_axis = Vec4(0, 0, c, 0);
        return m;
    }
};

}  // namespace orc
