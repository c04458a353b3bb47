// Minkowski support functions for the bam3d primitives, one device function per primitive.
// Reference semantics: Primitive::support_point (traits.rs:78-100) as implemented in primitive/*.rs — the
// world direction goes to local space with transform.inverse().transform_vector3, the local farthest point is
// found, and transform.transform_point3 brings it back. The inverse is a pure function of the transform, so
// it is computed once per shape (pack_shapes_kernel) instead of once per call; results are bit-identical.
#pragma once
#include "vecmath.cuh"

namespace b3 {

// Kind numbering follows `enum Primitive3` (primitive3.rs:16-33); Cube is a Cuboid (cuboid.rs:133-137).
enum Kind : uint32_t { K_PARTICLE = 0, K_QUAD = 1, K_SPHERE = 2, K_CUBOID = 3, K_CYLINDER = 4, K_CAPSULE = 5, K_POLYHEDRON = 6, K_COUNT = 7 };

// get_max_point over the 8 cached cuboid corners (cuboid.rs:46-64, primitive/util.rs:7-23): fold from
// (zero, -inf) with strict `>`, so the FIRST maximum in the reference's corner order wins. The eight dot
// products (cx*dx + cy*dy) + cz*dz share the three products a=hx*dx, b=hy*dy, c=hz*dz: negating a factor
// negates the rounded product, and (-a)+(-b) == -(a+b), (-a)+b == -(a-b) in IEEE arithmetic (up to the sign
// of an exact zero, which no comparison sees), so 2 adds + 8 adds reproduce all eight dots bit-for-bit.
B3_DEV V3 cuboid_local_support(float hx, float hy, float hz, V3 d) {
    float a = hx * d.x, b = hy * d.y, c = hz * d.z;
    float s1 = a + b, s2 = a - b;
    float best = -INFINITY;
    int idx = 8;
    float t;
    t = s1 + c;  if (t > best) { best = t; idx = 0; }   // ( hx,  hy,  hz)
    t = -s2 + c; if (t > best) { best = t; idx = 1; }   // (-hx,  hy,  hz)
    t = -s1 + c; if (t > best) { best = t; idx = 2; }   // (-hx, -hy,  hz)
    t = s2 + c;  if (t > best) { best = t; idx = 3; }   // ( hx, -hy,  hz)
    t = s1 - c;  if (t > best) { best = t; idx = 4; }   // ( hx,  hy, -hz)
    t = -s2 - c; if (t > best) { best = t; idx = 5; }   // (-hx,  hy, -hz)
    t = -s1 - c; if (t > best) { best = t; idx = 6; }   // (-hx, -hy, -hz)
    t = s2 - c;  if (t > best) { best = t; idx = 7; }   // ( hx, -hy, -hz)
    if (idx == 8) return v3(0.f, 0.f, 0.f);             // every dot was NaN: the fold keeps Vec3::zero()
    // sign masks over idx: x negative for {1,2,5,6}, y negative for {2,3,6,7}, z negative for {4..7}
    float px = ((0x66 >> idx) & 1) ? -hx : hx;
    float py = ((0xCC >> idx) & 1) ? -hy : hy;
    float pz = ((0xF0 >> idx) & 1) ? -hz : hz;
    return v3(px, py, pz);
}

// Quad corners (quad.rs:46-53): (hx,hy,0), (-hx,hy,0), (-hx,-hy,0), (hx,-hy,0); same fold.
B3_DEV V3 quad_local_support(float hx, float hy, V3 d) {
    float a = hx * d.x, b = hy * d.y, c = 0.f * d.z;
    float s1 = a + b, s2 = a - b;
    float best = -INFINITY;
    int idx = 4;
    float t;
    t = s1 + c;  if (t > best) { best = t; idx = 0; }
    t = -s2 + c; if (t > best) { best = t; idx = 1; }
    t = -s1 + c; if (t > best) { best = t; idx = 2; }
    t = s2 + c;  if (t > best) { best = t; idx = 3; }
    if (idx == 4) return v3(0.f, 0.f, 0.f);
    float px = ((0x6 >> idx) & 1) ? -hx : hx;
    float py = ((0xC >> idx) & 1) ? -hy : hy;
    return v3(px, py, 0.f);
}

// Brute-force vertex scan (polyhedron.rs:135-150, VertexOnly mode): strict `>`, first maximum wins.
// WARP = true: the 32 lanes of the warp stride the vertex list, each keeping its own first maximum; the
// shuffle reduction prefers the larger dot and, on equal dots, the lower vertex index — the same vertex the
// sequential scan returns. All lanes must call this convergently with identical arguments.
template <bool WARP>
B3_DEV V3 polyhedron_local_support(const float4* __restrict__ verts, uint32_t n, V3 d) {
    float best = -INFINITY;
    uint32_t idx = 0xFFFFFFFFu;
    if (WARP) {
        const uint32_t lane = threadIdx.x & 31u;
#ifndef B3_SUPPORT_BATCH
#define B3_SUPPORT_BATCH 4
#endif
        // B3_SUPPORT_BATCH vertices per lane per step, loads first: one L1/L2 round trip per 32 x BATCH vertices
        for (uint32_t i0 = lane; i0 < n; i0 += 32 * B3_SUPPORT_BATCH) {
            float4 p[B3_SUPPORT_BATCH];
#pragma unroll
            for (int j = 0; j < B3_SUPPORT_BATCH; ++j) p[j] = (i0 + 32 * j < n) ? __ldg(verts + i0 + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < B3_SUPPORT_BATCH; ++j) {
                const uint32_t i = i0 + 32 * j;
                const float t = dot(v3(p[j].x, p[j].y, p[j].z), d);
                if (i < n && t > best) { best = t; idx = i; }
            }
        }
#ifdef B3_SUPPORT_SHUFFLE_REDUCE
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            float ob = __shfl_xor_sync(0xFFFFFFFFu, best, off);
            uint32_t oi = __shfl_xor_sync(0xFFFFFFFFu, idx, off);
            // a lane that saw no valid dot has idx = 0xFFFFFFFF and best = -inf and never wins over a valid one
            if (ob > best || (ob == best && oi < idx)) { best = ob; idx = oi; }
        }
#else
        // Two warp-wide integer reductions (redux.sync) instead of five shuffle rounds: the largest dot through an
        // order-preserving map of f32 onto u32 (-0 folded onto +0 first, as == would; a lane's `best` is never NaN),
        // then the lowest vertex index among the lanes that hold it. A lane that saw no valid dot keeps
        // (-inf, 0xFFFFFFFF) and can only win when every lane did.
        const uint32_t bits = __float_as_uint(best + 0.0f);
        const uint32_t key = bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u);
        const uint32_t kmax = __reduce_max_sync(0xFFFFFFFFu, key);
        idx = __reduce_min_sync(0xFFFFFFFFu, key == kmax ? idx : 0xFFFFFFFFu);
#endif
    } else {
        for (uint32_t i = 0; i < n; ++i) {
            float4 p = __ldg(verts + i);
            float t = dot(v3(p.x, p.y, p.z), d);
            if (t > best) { best = t; idx = i; }
        }
    }
    if (idx == 0xFFFFFFFFu) return v3(0.f, 0.f, 0.f);
    float4 p = __ldg(verts + idx);
    return v3(p.x, p.y, p.z);
}

// ConvexPolyhedron::new_with_faces (HalfEdge mode). The two float4 slots in front of a mesh's vertices hold this header
// (written by the host layer, capi.cu); n_faces = 0 for a VertexOnly mesh. Shape::nverts carries MESH_HALF_EDGE then.
constexpr uint32_t MESH_HALF_EDGE = 0x80000000u;
constexpr int MESH_HEADER_SLOTS = 2;
struct MeshHeader {
    const uint32_t* vert_edge;  // per vertex: an outgoing half-edge
    const uint4* edges;         // 2 x uint4 per half-edge: (target_vertex, twin_edge, next_edge, previous_edge), (left_face, 0, 0, 0)
    const uint4* faces;         // 2 x 16 B per face: (v0, v1, v2, edge), then the plane (n.xyz, d) as float4
    uint32_t n_faces, n_edges;
};
static_assert(sizeof(MeshHeader) == MESH_HEADER_SLOTS * 16, "mesh header must fill its float4 slots");
B3_DEV const MeshHeader* mesh_header(const float4* verts) { return reinterpret_cast<const MeshHeader*>(verts - MESH_HEADER_SLOTS); }

// hill_climb_support_point (polyhedron.rs:153-183): from vertex 0, walk the ring of neighbours of the best vertex, move
// to the best of them, stop when a pass improves the dot product by less than f32::EPSILON. Serial by nature (in the warp
// kernels every lane runs it on identical data). The caps only bound malformed input the reference would hang on (a ring
// that does not close, NaN directions).
// (not inlined: the VertexOnly path, which C3 measures, should not carry this loop nest at every support call site)
static __device__ __noinline__ V3 polyhedron_hill_climb(const float4* __restrict__ verts, uint32_t nverts, V3 d) {
    const MeshHeader* h = mesh_header(verts);
    const uint32_t* __restrict__ ve = h->vert_edge;
    const uint4* __restrict__ ed = h->edges;
    uint32_t best = 0;
    float4 p = __ldg(verts);
    float best_dot = dot(v3(p.x, p.y, p.z), d);
    for (uint32_t pass = 0; pass <= nverts; ++pass) {
        const float previous = best_dot;
        uint32_t e = __ldg(ve + best);
        const uint32_t start = e;
        for (uint32_t step = 0; step <= h->n_edges; ++step) {
            const uint4 E = __ldg(ed + 2 * e);
            const float4 q = __ldg(verts + E.x);
            const float t = dot(v3(q.x, q.y, q.z), d);
            if (t > best_dot) { best = E.x; best_dot = t; }
            e = __ldg(ed + 2 * E.y).z;
            if (e == start) break;
        }
        if (fabsf(best_dot - previous) < F32_EPSILON) break;
    }
    p = __ldg(verts + best);
    return v3(p.x, p.y, p.z);
}

// World-space support point of shape `s` along world direction `dir`.
template <bool WARP>
B3_DEV V3 support_point(const Shape& s, V3 dir) {
    if (s.kind == K_PARTICLE) {
        // primitive3.rs:119: transform.transform_point3(Vec3::zero())
        return transform_point3(s, v3(0.f, 0.f, 0.f));
    }
    V3 d = inv_transform_vector3(s, dir);
    V3 p;
    switch (s.kind) {
        case K_SPHERE:  // sphere.rs:20-25
            p = d * (s.p0 / sqrtf(dot(d, d)));
            break;
        case K_CUBOID:
            p = cuboid_local_support(s.p0, s.p1, s.p2, d);
            break;
        case K_QUAD:
            p = quad_local_support(s.p0, s.p1, d);
            break;
        case K_CYLINDER: {  // cylinder.rs:37-62; p0 = half_height, p1 = radius
            bool negative = signbit(d.y);
            V3 r = v3(d.x, 0.f, d.z);
            if (fabsf(dot(r, r) - 0.f) < F32_EPSILON) {
                r = v3(0.f, 0.f, 0.f);
            } else {
                r = normalize(r);
                if (fabsf(r.y - 0.f) < F32_EPSILON && fabsf(r.x - 0.f) < F32_EPSILON && fabsf(r.z - 0.f) < F32_EPSILON) {
                    r = v3(0.f, 0.f, 0.f);
                } else {
                    r = r * s.p1;
                }
            }
            r.y = negative ? -s.p0 : s.p0;
            p = r;
            break;
        }
        case K_CAPSULE: {  // capsule.rs:40-49; p0 = half_height, p1 = radius
            V3 r = v3(0.f, rsignum(d.y) * s.p0, 0.f);
            p = r + d * (s.p1 / sqrtf(dot(d, d)));
            break;
        }
        default:  // K_POLYHEDRON, polyhedron.rs:342-355
#ifndef B3_NO_HALF_EDGE
            if (s.nverts & MESH_HALF_EDGE) p = polyhedron_hill_climb(s.verts, s.nverts & ~MESH_HALF_EDGE, d);
            else
#endif
                p = polyhedron_local_support<WARP>(s.verts, s.nverts, d);
            break;
    }
    return transform_point3(s, p);
}

// SupportPoint::from_minkowski (minkowski/mod.rs:33-51): v = l - r, sup_a = l. sup_b is never read by the
// reference and is not kept.
// One out-of-line copy for the warp-per-pair kernels: with the support code inlined at every call site (two shapes x three
// sites in the contact kernel, every primitive kind in each) those kernels stalled 43 % of the time on instruction fetch
// (ncu: stall_no_inst) — 16 warps per SM at unrelated places of a code image far larger than the instruction cache.
template <bool WARP>
__device__ __noinline__ V3 support_point_outlined(const Shape& s, V3 dir) { return support_point<WARP>(s, dir); }

// The polyhedron case of the warp kernels gets its own out-of-line function with BY-VALUE arguments (six registers): a
// `const Shape&` parameter of a real call lives in local memory, and the callee's loads of the transform from it were a
// quarter of the contact kernel's stall samples. The two small transforms stay inline at the call site.
static __device__ __noinline__ V3 polyhedron_scan_outlined(const float4* __restrict__ verts, uint32_t nverts, V3 d) {
    if (nverts & MESH_HALF_EDGE) return polyhedron_hill_climb(verts, nverts & ~MESH_HALF_EDGE, d);
    return polyhedron_local_support<true>(verts, nverts, d);
}
B3_DEV V3 warp_support_point(const Shape& s, V3 dir) {
    if (s.kind == K_POLYHEDRON) return transform_point3(s, polyhedron_scan_outlined(s.verts, s.nverts, inv_transform_vector3(s, dir)));
    return support_point_outlined<true>(s, dir);
}

template <bool WARP>
B3_DEV void minkowski_support(const Shape& L, const Shape& R, V3 dir, V3& v, V3& sup_a) {
#ifndef B3_INLINE_WARP_SUPPORT
    if (WARP) {
        V3 l = warp_support_point(L, dir);
        V3 r = warp_support_point(R, -dir);
        v = l - r;
        sup_a = l;
        return;
    }
#endif
    V3 l = support_point<WARP>(L, dir);
    V3 r = support_point<WARP>(R, -dir);
    v = l - r;
    sup_a = l;
}

}  // namespace b3
