// Device-side f32 vector math for the batched collision kernels (sm_100a).
//
// Every decision in GJK/EPA (hit/miss, which simplex feature, which polytope face) is a comparison of f32 values,
// so to return the same answers as bam3d (which computes in glam 0.8 `scalar-math`: plain scalar f32, no FMA) the
// kernels evaluate every expression in glam's operation order and the translation unit is compiled with
// `-fmad=false` (no FMA contraction), IEEE division and IEEE sqrt (`-prec-div=true -prec-sqrt=true`, nvcc
// defaults). Do not reassociate anything in this file.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace b3 {

#define B3_DEV __device__ __forceinline__

constexpr float F32_EPSILON = 1.1920929e-7f;  // Rust std::f32::EPSILON (cylinder.rs:45, gjk/mod.rs:247)

struct V3 { float x, y, z; };

B3_DEV V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
B3_DEV V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
B3_DEV V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
B3_DEV V3 operator*(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
B3_DEV V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
// glam Vec3::dot: (x*x' + y*y') + z*z'
B3_DEV float dot(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
// glam Vec3::cross
B3_DEV V3 cross(V3 a, V3 b) { return v3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
// glam Vec3::normalize = self * (1.0 / length)
B3_DEV V3 normalize(V3 a) { return a * (1.0f / sqrtf(dot(a, a))); }
B3_DEV bool is_zero(V3 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f; }  // cmpeq(zero).all(): -0 == +0, NaN != 0
// Rust f32::min/max (NaN-ignoring), same selection rule as the CPU oracle
B3_DEV float rmin(float a, float b) { return (a != a) ? b : ((b != b) ? a : ((b < a) ? b : a)); }
B3_DEV float rmax(float a, float b) { return (a != a) ? b : ((b != b) ? a : ((b > a) ? b : a)); }
// Rust f32::signum: NaN -> NaN, else copysign(1, x)
B3_DEV float rsignum(float x) { return (x != x) ? x : copysignf(1.0f, x); }

// Device-resident shape record: 6 x float4 = 96 B, written once per shape by pack_shapes_kernel.
//   q0..q2 : model-to-world affine, 12 floats = columns X.xyz, Y.xyz, Z.xyz, W.xyz of the Mat4
//   q3..q5 : 9 floats of the linear part of Mat4::inverse() (columns X.xyz, Y.xyz, Z.xyz), then p0, p1, p2
//            (shape parameters), all packed contiguously: inv[9], p[3]
// kind / mesh id live in a separate u32 array (read by the binning pass too).
struct ShapeRec { float4 q[6]; };

struct Shape {
    // columns of the model-to-world transform (affine part)
    V3 X, Y, Z, W;
    // columns of the linear part of the inverse transform
    V3 iX, iY, iZ;
    float p0, p1, p2;
    uint32_t kind;
    const float4* verts;  // polyhedron vertices (xyz, w unused) or nullptr
    uint32_t nverts;
};

B3_DEV void load_shape(Shape& s, const ShapeRec* __restrict__ recs, uint32_t idx) {
    const float4* p = reinterpret_cast<const float4*>(recs + idx);
    float4 a = __ldg(p + 0), b = __ldg(p + 1), c = __ldg(p + 2), d = __ldg(p + 3), e = __ldg(p + 4), f = __ldg(p + 5);
    s.X = v3(a.x, a.y, a.z); s.Y = v3(a.w, b.x, b.y); s.Z = v3(b.z, b.w, c.x); s.W = v3(c.y, c.z, c.w);
    s.iX = v3(d.x, d.y, d.z); s.iY = v3(d.w, e.x, e.y); s.iZ = v3(e.z, e.w, f.x);
    s.p0 = f.y; s.p1 = f.z; s.p2 = f.w;
}

// glam Mat4::transform_point3: r = X*px; r = Y*py + r; r = Z*pz + r; r = W + r
B3_DEV V3 transform_point3(const Shape& s, V3 p) {
    V3 r = s.X * p.x;
    r = s.Y * p.y + r;
    r = s.Z * p.z + r;
    r = s.W + r;
    return r;
}
// glam Mat4::transform_vector3 applied with the inverse matrix (primitive/util.rs:11 et al.)
B3_DEV V3 inv_transform_vector3(const Shape& s, V3 v) {
    V3 r = s.iX * v.x;
    r = s.iY * v.y + r;
    r = s.iZ * v.z + r;
    return r;
}

}  // namespace b3
