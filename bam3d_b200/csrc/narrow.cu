// Narrow-phase kernels for sm_100a: shape packing, pair binning, and batched GJK intersect / contact (GJK+EPA) /
// distance / time of impact. Two execution shapes per query:
//   * thread-per-pair  — analytic supports (sphere, cuboid, cylinder, capsule, quad, particle); simplex in
//     registers; EPA polytope in a global-memory slab, one contiguous record per resident thread (persistent kernels in which
//     every lane owns one pair and idle lanes refill: contact, time of impact, distance), with a warp-per-pair and a
//     block-per-pair overflow tier behind it for polytopes beyond 256 / 1024 faces;
//   * warp-per-pair    — any pair that involves a ConvexPolyhedron: the 32 lanes stride the vertex list and
//     arg-max by shuffle (support.cuh); everything else runs redundantly on identical data in all lanes, with the
//     EPA polytope in a per-warp shared-memory slice.
// Pairs are first binned by (kindL, kindR) so that a warp of the thread-per-pair kernel runs one support code
// path; polyhedron pairs form the last bin and are routed to the warp kernel.
//
// Plain FP32 SIMT: there is no dense contraction anywhere in this path, so tensor cores / TMEM are not used.
// Compiled with -fmad=false (see vecmath.cuh).
#include <cstdio>
#include <cstdlib>
#include "narrow.h"
#include "gjk.cuh"
#include "epa_big.cuh"
#include "epa_huge.cuh"
#ifdef B3_EXPERIMENTS
#include "epa_pipe.cuh"
#endif
#include "epa_warp.cuh"
#include "raycast.cuh"

namespace b3 {

constexpr int TPB = 128;            // threads per block, every kernel in this file
constexpr int WARPS_PER_BLOCK = TPB / 32;
constexpr int NUM_SMS = 148;        // B200
constexpr uint32_t KEY_WARP = 49;   // bin of the pairs that involve a polyhedron (thread bins are 0..48)
constexpr uint32_t ERR_NON_AFFINE = 1u, ERR_BAD_KIND = 2u, ERR_BAD_MESH = 4u;
constexpr uint32_t OVF_MAX = 65536;  // capacity of the EPA overflow list; pairs beyond it keep ST_CAPACITY
// Overflow list (head of the overflow scratch): u32 words [0] = entries appended, [1] = consumer cursor,
// [2] = "every producer kernel has finished", [3] = producer blocks that have exited; entries (pair indices) from
// byte 256, preset to 0xFFFFFFFF so that a consumer that reserved a position can wait for its value to land.
constexpr int OVF_COUNT = 0, OVF_CURSOR = 1, OVF_DONE = 2, OVF_EXITED = 3;
// third tier (epa_huge.cuh): [4] = pairs the overflow tier could not finish either, [5] = consumer cursor; entries behind the overflow list
constexpr int HUGE_COUNT = 4, HUGE_CURSOR = 5;
constexpr uint32_t HUGE_MAX = 4096;
constexpr uint32_t HUGE_SORT_MAX = 1024;  // third-tier lists up to this long are taken longest-first

// ------------------------------------------------------------------------------------------------------------
// pack_shapes: Mat4 (16 f32, column-major) + params -> 96-byte ShapeRec with the inverse's linear part.
// The inverse is glam 0.8's scalar Mat4::inverse (GLM cofactor expansion) evaluated on all 16 inputs in glam's
// order; only its upper-left 3x3 is kept because the reference only ever applies the inverse through
// transform_vector3 (primitive/util.rs:11, sphere.rs:22, cylinder.rs:39, capsule.rs:43, polyhedron.rs:346-350).
// ------------------------------------------------------------------------------------------------------------
// AFFINE48: the transforms arrive as 12 floats per shape — the xyz of the four columns, i.e. the Mat4 without its fourth row,
// which a model transform has as (0, 0, 0, 1) (anything else is rejected below) — a quarter less to move over PCIe per frame.
// The same expressions are then evaluated on the same 16 values: bit-identical records.
template <bool AFFINE48>
__global__ void __launch_bounds__(TPB) pack_shapes_kernel(const uint8_t* __restrict__ kind, const float4* __restrict__ params,
                                                          const float4* __restrict__ xform, const uint32_t* __restrict__ mesh_id,
                                                          uint32_t n, uint32_t n_meshes, ShapeRec* __restrict__ recs,
                                                          uint32_t* __restrict__ meta, uint32_t* __restrict__ err) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float m00, m01, m02, m03, m10, m11, m12, m13, m20, m21, m22, m23, m30, m31, m32, m33;
    if (AFFINE48) {
        const float4 a = xform[3 * (size_t)i + 0], b = xform[3 * (size_t)i + 1], c = xform[3 * (size_t)i + 2];
        m00 = a.x; m01 = a.y; m02 = a.z; m03 = 0.f;
        m10 = a.w; m11 = b.x; m12 = b.y; m13 = 0.f;
        m20 = b.z; m21 = b.w; m22 = c.x; m23 = 0.f;
        m30 = c.y; m31 = c.z; m32 = c.w; m33 = 1.f;
    } else {
        const float4 cx = xform[4 * (size_t)i + 0], cy = xform[4 * (size_t)i + 1], cz = xform[4 * (size_t)i + 2], cw = xform[4 * (size_t)i + 3];
        m00 = cx.x; m01 = cx.y; m02 = cx.z; m03 = cx.w;
        m10 = cy.x; m11 = cy.y; m12 = cy.z; m13 = cy.w;
        m20 = cz.x; m21 = cz.y; m22 = cz.z; m23 = cz.w;
        m30 = cw.x; m31 = cw.y; m32 = cw.z; m33 = cw.w;
    }
    uint32_t e = 0;
    if (!(m03 == 0.f && m13 == 0.f && m23 == 0.f && m33 == 1.f)) e |= ERR_NON_AFFINE;

    float coef00 = m22 * m33 - m32 * m23, coef02 = m12 * m33 - m32 * m13;
    float coef04 = m21 * m33 - m31 * m23, coef06 = m11 * m33 - m31 * m13;
    float coef08 = m21 * m32 - m31 * m22, coef10 = m11 * m32 - m31 * m12;
    float coef12 = m20 * m33 - m30 * m23, coef14 = m10 * m33 - m30 * m13;
    float coef16 = m20 * m32 - m30 * m22, coef18 = m10 * m32 - m30 * m12;
    float coef20 = m20 * m31 - m30 * m21, coef22 = m10 * m31 - m30 * m11;
    // (coef03/07/11/15/19/23 only feed the w row of the inverse, which is never applied)
    // inv_k = vec * fac - vec * fac + vec * fac, per component (x, y, z of each column)
    // fac0 = (c00,c00,c02,c03) fac1 = (c04,c04,c06,c07) fac2 = (c08,c08,c10,c11)
    // fac3 = (c12,c12,c14,c15) fac4 = (c16,c16,c18,c19) fac5 = (c20,c20,c22,c23)
    // vec0 = (m10,m00,m00,m00) vec1 = (m11,m01,m01,m01) vec2 = (m12,m02,m02,m02) vec3 = (m13,m03,m03,m03)
    float inv0x = m11 * coef00 - m12 * coef04 + m13 * coef08;
    float inv0y = m01 * coef00 - m02 * coef04 + m03 * coef08;
    float inv0z = m01 * coef02 - m02 * coef06 + m03 * coef10;
    float inv1x = m10 * coef00 - m12 * coef12 + m13 * coef16;
    float inv1y = m00 * coef00 - m02 * coef12 + m03 * coef16;
    float inv1z = m00 * coef02 - m02 * coef14 + m03 * coef18;
    float inv2x = m10 * coef04 - m11 * coef12 + m13 * coef20;
    float inv2y = m00 * coef04 - m01 * coef12 + m03 * coef20;
    float inv2z = m00 * coef06 - m01 * coef14 + m03 * coef22;
    float inv3x = m10 * coef08 - m11 * coef16 + m12 * coef20;
    // signs: columns 0 and 2 use (+,-,+,-), columns 1 and 3 use (-,+,-,+)
    float i0x = inv0x * 1.0f, i0y = inv0y * -1.0f, i0z = inv0z * 1.0f;
    float i1x = inv1x * -1.0f, i1y = inv1y * 1.0f, i1z = inv1z * -1.0f;
    float i2x = inv2x * 1.0f, i2y = inv2y * -1.0f, i2z = inv2z * 1.0f;
    float i3x = inv3x * -1.0f;
    // det = dot(x_axis, (inv0.x, inv1.x, inv2.x, inv3.x)), summed left to right
    float dot1 = ((m00 * i0x + m01 * i1x) + m02 * i2x) + m03 * i3x;
    float rcp = 1.0f / dot1;
    i0x *= rcp; i0y *= rcp; i0z *= rcp;
    i1x *= rcp; i1y *= rcp; i1z *= rcp;
    i2x *= rcp; i2y *= rcp; i2z *= rcp;

    float4 p = params[i];
    uint32_t k = kind[i];
    if (k >= K_COUNT) e |= ERR_BAD_KIND;
    uint32_t mid = 0;
    if (k == K_POLYHEDRON) {
        mid = mesh_id ? mesh_id[i] : 0xFFFFFFFFu;
        if (mid >= n_meshes || mid >= (1u << 24)) { e |= ERR_BAD_MESH; mid = 0; }
    }
    float4* q = reinterpret_cast<float4*>(recs + i);
    q[0] = make_float4(m00, m01, m02, m10);
    q[1] = make_float4(m11, m12, m20, m21);
    q[2] = make_float4(m22, m30, m31, m32);
    q[3] = make_float4(i0x, i0y, i0z, i1x);
    q[4] = make_float4(i1y, i1z, i2x, i2y);
    q[5] = make_float4(i2z, p.x, p.y, p.z);
    meta[i] = k | (mid << 8);
    if (e) atomicOr(err, e);
}

// Face planes of the half-edge meshes: Plane::from_points (plane.rs:69-86) with n = normalize(cross(p2 - p1, p3 - p1)),
// d = -p1.dot(n). On entry the second uint4 of a face holds the slot of its mesh's vertex 0 in `verts`.
__global__ void __launch_bounds__(TPB) mesh_face_planes_kernel(const float4* __restrict__ verts, uint4* __restrict__ faces, uint64_t nfaces,
                                                               uint32_t* __restrict__ err) {
    uint64_t f = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nfaces) return;
    const uint4 fv = faces[2 * f];
    const uint32_t base = faces[2 * f + 1].x;
    const float4 a = verts[base + fv.x], b = verts[base + fv.y], c = verts[base + fv.z];
    const V3 p1 = v3(a.x, a.y, a.z);
    const V3 normal = cross(v3(b.x, b.y, b.z) - p1, v3(c.x, c.y, c.z) - p1);
    if (is_zero(normal)) { atomicOr(err, 1u); return; }
    const V3 n = normalize(normal);
    const float d = -dot(p1, n);
    reinterpret_cast<float4*>(faces)[2 * f + 1] = make_float4(n.x, n.y, n.z, d);
}

// ------------------------------------------------------------------------------------------------------------
// pair binning: counting sort of pair indices on key = kindL * 7 + kindR (polyhedron pairs -> KEY_WARP)
// ------------------------------------------------------------------------------------------------------------
B3_DEV uint32_t pair_key(uint32_t ml, uint32_t mr) {
    uint32_t kl = ml & 0xFF, kr = mr & 0xFF;
    return (kl == K_POLYHEDRON || kr == K_POLYHEDRON) ? KEY_WARP : kl * 7 + kr;
}

__global__ void __launch_bounds__(TPB) bin_hist_kernel(const uint32_t* __restrict__ meta, uint32_t n_shapes, const uint2* __restrict__ pairs, uint32_t n,
                                                       uint8_t* __restrict__ keys, uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[64];
    if (threadIdx.x < 64) sh[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint2 pr = __ldg(pairs + i);
        // (an out-of-range index is reported by the kernel that loads the shapes; here it only must not be dereferenced)
        uint32_t key = pair_key(__ldg(meta + min(pr.x, n_shapes - 1)), __ldg(meta + min(pr.y, n_shapes - 1)));
        keys[i] = (uint8_t)key;
        atomicAdd(&sh[key], 1u);
    }
    __syncthreads();
    if (threadIdx.x < 64 && sh[threadIdx.x]) atomicAdd(&hist[threadIdx.x], sh[threadIdx.x]);
}

__global__ void bin_scan_kernel(const uint32_t* __restrict__ hist, uint32_t* __restrict__ start, uint32_t* __restrict__ cursor,
                                uint32_t* __restrict__ ranges, uint32_t n) {
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int k = 0; k < 64; ++k) { start[k] = acc; cursor[k] = acc; acc += hist[k]; }
        ranges[0] = 0; ranges[1] = start[KEY_WARP]; ranges[2] = start[KEY_WARP]; ranges[3] = n;
    }
}

__global__ void __launch_bounds__(TPB) bin_scatter_kernel(const uint8_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ cursor,
                                                          uint32_t* __restrict__ order) {
    // one tile of TPB pairs per block iteration: count per bin in shared memory, reserve a global range per bin,
    // then every thread writes its pair index at (range base + its rank inside the tile's bin)
    __shared__ uint32_t cnt[64];
    __shared__ uint32_t base[64];
    for (uint32_t tile = blockIdx.x * TPB; tile < n; tile += gridDim.x * TPB) {
        if (threadIdx.x < 64) cnt[threadIdx.x] = 0;
        __syncthreads();
        uint32_t i = tile + threadIdx.x;
        uint32_t key = 0, rank = 0;
        bool valid = i < n;
        if (valid) {
            key = keys[i];
            // warp-aggregated shared atomic: one atomic per distinct key per warp
            uint32_t peers = __match_any_sync(__activemask(), key);
            uint32_t lane = threadIdx.x & 31u;
            uint32_t leader = __ffs(peers) - 1;
            uint32_t wbase = 0;
            if (lane == leader) wbase = atomicAdd(&cnt[key], __popc(peers));
            wbase = __shfl_sync(peers, wbase, leader);
            rank = wbase + __popc(peers & ((1u << lane) - 1u));
        }
        __syncthreads();
        if (threadIdx.x < 64 && cnt[threadIdx.x]) base[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], cnt[threadIdx.x]);
        __syncthreads();
        if (valid) order[base[key] + rank] = i;
        __syncthreads();
    }
}

B3_DEV uint32_t ld_volatile_u32(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

__device__ __forceinline__ void overflow_append(unsigned char* ovf_scratch, uint32_t p) {
    __threadfence();  // the caller's ST_CAPACITY status / zero contact must land before the consumer can overwrite them
    const uint32_t slot = atomicAdd(reinterpret_cast<uint32_t*>(ovf_scratch) + OVF_COUNT, 1u);
    if (slot < OVF_MAX) { reinterpret_cast<uint32_t*>(ovf_scratch + 256)[slot] = p; __threadfence(); }
}
// Last thing a producer block does: the block that exits last raises the "producers finished" flag.
__device__ __forceinline__ void overflow_producer_exit(unsigned char* ovf_scratch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t* h = reinterpret_cast<uint32_t*>(ovf_scratch);
        __threadfence();
        if (atomicAdd(h + OVF_EXITED, 1u) == gridDim.x - 1) atomicExch(h + OVF_DONE, 1u);
    }
}

// ------------------------------------------------------------------------------------------------------------
// pair loading
// ------------------------------------------------------------------------------------------------------------
B3_DEV uint32_t checked_index(const ShapeSetDev& S, uint32_t idx) {
    if (idx >= S.n) { atomicOr(S.err, 1u); idx = S.n - 1; }  // (the entry points refuse pairs over an empty shape set)
    return idx;
}
B3_DEV void load_pair_shape(Shape& s, const ShapeSetDev& S, uint32_t idx) {
    idx = checked_index(S, idx);
    load_shape(s, S.recs, idx);
    uint32_t m = __ldg(S.meta + idx);
    s.kind = m & 0xFF;
    s.verts = nullptr;
    s.nverts = 0;
    if (s.kind == K_POLYHEDRON) {
        uint32_t mid = m >> 8;
        // device offsets count the header slots in front of every mesh's vertices (support.cuh: MeshHeader)
        uint32_t b = __ldg(S.mesh_offsets + mid), e = __ldg(S.mesh_offsets + mid + 1);
        s.verts = S.mesh_verts + b + MESH_HEADER_SLOTS;
        s.nverts = e - b - MESH_HEADER_SLOTS;
#ifndef B3_NO_HALF_EDGE  // (development: measures what the HalfEdge-mode check costs the VertexOnly path)
        if (mesh_header(s.verts)->n_faces) s.nverts |= MESH_HALF_EDGE;
#endif
    }
}

// Work distribution. Thread kernels: item t of [begin, end) per thread, grid-stride. Warp kernels: one item per
// warp, grid-stride over warps. `order` (nullable) maps a sorted slot to the pair index.
template <bool WARP>
B3_DEV void work_range(const uint32_t* ranges, uint32_t n, uint32_t& first, uint32_t& end, uint32_t& stride) {
    uint32_t begin = 0;
    end = n;
    if (ranges) { begin = ranges[WARP ? 2 : 0]; end = ranges[WARP ? 3 : 1]; }
    if (WARP) {
        first = begin + (blockIdx.x * blockDim.x + threadIdx.x) / 32;
        stride = gridDim.x * blockDim.x / 32;
    } else {
        first = begin + blockIdx.x * blockDim.x + threadIdx.x;
        stride = gridDim.x * blockDim.x;
    }
}

#ifndef B3_WARP_SHAPES_SMEM
#define B3_WARP_SHAPES_SMEM 1  // warp-per-pair kernels: one copy of a pair's shape records per warp in shared memory (C3: 15.6 -> 13.4 ms)
#endif
// Warp-per-pair kernels: every lane would hold the same shape records (28 registers each) for the whole pair, and a call of an
// out-of-line support function would copy them to local memory first. Lane 0 stages one copy per warp in shared memory instead;
// the GJK / EPA code reads it with broadcast loads. `sh` = this warp's slots, idx[k] = shape index to load from sets[k].
template <int N>
B3_DEV void stage_warp_shapes(Shape* sh, const ShapeSetDev* const (&sets)[N], const uint32_t (&idx)[N]) {
    __syncwarp();  // the previous pair's readers are done
    if ((threadIdx.x & 31u) == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            Shape t;
            load_pair_shape(t, *sets[k], idx[k]);
            sh[k] = t;
        }
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------------------
// GJK::intersect
// ------------------------------------------------------------------------------------------------------------
template <bool WARP>
__global__ void __launch_bounds__(TPB) gjk_intersect_kernel(ShapeSetDev S, const uint2* __restrict__ pairs, uint32_t n,
                                                            const uint32_t* __restrict__ order, const uint32_t* __restrict__ ranges,
                                                            SettingsDev cfg, uint8_t* __restrict__ status) {
    uint32_t first, end, stride;
    work_range<WARP>(ranges, n, first, end, stride);
    for (uint32_t slot = first; slot < end; slot += stride) {
        uint32_t p = order ? __ldg(order + slot) : slot;
        uint2 pr = __ldg(pairs + p);
        Simplex<false> s;
        uint8_t st;
        if constexpr (WARP && B3_WARP_SHAPES_SMEM) {
            __shared__ Shape sh_all[WARPS_PER_BLOCK * 2];
            Shape* sh = sh_all + 2 * (threadIdx.x / 32);
            const ShapeSetDev* const sets[2] = {&S, &S};
            const uint32_t idx[2] = {pr.x, pr.y};
            stage_warp_shapes<2>(sh, sets, idx);
            st = gjk_intersect<WARP, false>(sh[0], sh[1], cfg.max_iterations, s);
        } else {
            Shape L, R;
            load_pair_shape(L, S, pr.x);
            load_pair_shape(R, S, pr.y);
            st = gjk_intersect<WARP, false>(L, R, cfg.max_iterations, s);
        }
        if (!WARP || (threadIdx.x & 31) == 0) status[p] = st;
    }
}

#ifdef B3_EXPERIMENTS  // measured slower than the single kernel: kept as the record of the experiment
// Thread-per-pair intersect in two launches. Most misses are decided by the very first support point
// (gjk/mod.rs:92: `a.v . d <= 0`), so a single kernel spends most of its issue slots with a third of the lanes alive
// (ncu: 11 of 32). Launch 1 streams every pair through that first test with all lanes busy — misses get their status,
// survivors are appended (warp-aggregated) to a list of pair indices; launch 2 runs the full GJK::intersect on the
// survivors only, so its warps start full. Survivors repeat the first support point (about a tenth of their work).
// Measured on C1 (2^20 cuboid pairs, 33 % hits): 0.176 ms against 0.167 ms for the single kernel — the second stream
// over the shape records costs more than the fuller warps save — so this path is opt-in ("intersect_split").
__global__ void __launch_bounds__(TPB) gjk_first_test_kernel(ShapeSetDev S, const uint2* __restrict__ pairs, uint32_t n,
                                                             const uint32_t* __restrict__ order, const uint32_t* __restrict__ ranges,
                                                             uint32_t* __restrict__ survivors, uint32_t* __restrict__ n_survivors,
                                                             uint8_t* __restrict__ status) {
    uint32_t first, end, stride;
    work_range<false>(ranges, n, first, end, stride);
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t slot0 = first - lane; slot0 < end; slot0 += stride) {
        const uint32_t slot = slot0 + lane;
        const bool valid = slot < end;
        uint32_t p = 0;
        bool alive = false;
        if (valid) {
            p = order ? __ldg(order + slot) : slot;
            const uint2 pr = __ldg(pairs + p);
            Shape L, R;
            load_pair_shape(L, S, pr.x);
            load_pair_shape(R, S, pr.y);
            V3 d = shape_origin(R) - shape_origin(L);
            if (is_zero(d)) d = v3(1.f, 1.f, 1.f);
            V3 pv, pa;
            minkowski_support<false>(L, R, d, pv, pa);
            alive = !(dot(pv, d) <= 0.f);
            if (!alive) status[p] = ST_MISS;
        }
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, alive);
        if (mask) {
            uint32_t base = 0;
            const uint32_t leader = __ffs(mask) - 1;
            if (lane == leader) base = atomicAdd(n_survivors, (uint32_t)__popc(mask));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (alive) survivors[base + __popc(mask & ((1u << lane) - 1u))] = p;
        }
    }
}

__global__ void __launch_bounds__(TPB) gjk_intersect_survivors_kernel(ShapeSetDev S, const uint2* __restrict__ pairs,
                                                                      const uint32_t* __restrict__ survivors,
                                                                      const uint32_t* __restrict__ n_survivors, SettingsDev cfg,
                                                                      uint8_t* __restrict__ status) {
    const uint32_t count = *n_survivors;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x) {
        const uint32_t p = __ldg(survivors + i);
        const uint2 pr = __ldg(pairs + p);
        Shape L, R;
        load_pair_shape(L, S, pr.x);
        load_pair_shape(R, S, pr.y);
        Simplex<false> s;
        status[p] = gjk_intersect<false, false>(L, R, cfg.max_iterations, s);
    }
}

#endif  // B3_EXPERIMENTS

// ------------------------------------------------------------------------------------------------------------
// GJK::intersection = intersect + EPA3::process (FullResolution) or Contact::new(CollisionOnly)
// ------------------------------------------------------------------------------------------------------------
#ifndef B3_WARP_MIN_BLOCKS
#define B3_WARP_MIN_BLOCKS (B3_WARP_SHAPES_SMEM ? 6 : 1)  // 6 x (4 polytope slices + shapes) = 205 KB of shared memory: the most that fits
#endif
// dynamic shared memory of the warp-per-pair contact kernel: the polytope slices, then (B3_WARP_SHAPES_SMEM) two Shape per warp
__host__ __device__ constexpr size_t warp_shapes_offset() { return (WARPS_PER_BLOCK * sizeof(WarpPolytopeSmem) + 15) & ~(size_t)15; }
constexpr size_t WARP_CONTACT_SMEM = B3_WARP_SHAPES_SMEM ? warp_shapes_offset() + WARPS_PER_BLOCK * 2 * sizeof(Shape) : WARPS_PER_BLOCK * sizeof(WarpPolytopeSmem);
template <bool WARP>
__global__ void __launch_bounds__(TPB, WARP ? B3_WARP_MIN_BLOCKS : 1) gjk_contact_kernel(ShapeSetDev S, const uint2* __restrict__ pairs, uint32_t n,
                                                          const uint32_t* __restrict__ order, const uint32_t* __restrict__ ranges,
                                                          SettingsDev cfg, int strategy, unsigned char* __restrict__ ovf_scratch,
                                                          float4* __restrict__ contacts, uint8_t* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];  // WARP: WARPS_PER_BLOCK polytope slices; else unused
    // thread-per-pair polytope: per-thread local memory (L1/L2-backed), indexed identically by all lanes
    float lvv[WARP ? 1 : EPA_VCAP * 3];
    float lva[WARP ? 1 : EPA_VCAP * 3];
    float4 lfn[WARP ? 1 : EPA_FCAP];
    uint32_t lfi[WARP ? 1 : EPA_FCAP];
    uint16_t led[WARP ? 1 : EPA_ECAP];
    PolytopeStore P;
    if (WARP) {
        WarpPolytopeSmem& w = reinterpret_cast<WarpPolytopeSmem*>(dyn_smem)[threadIdx.x / 32];
        P.vv = w.vv; P.va = w.va; P.fn = w.fn; P.fi = w.fi; P.ed = w.ed;
    } else {
        P.vv = lvv; P.va = lva; P.fn = lfn; P.fi = lfi; P.ed = led;
    }
    P.fcap = EPA_FCAP; P.ecap = EPA_ECAP; P.estride = 1;
    uint32_t first, end, stride;
    work_range<WARP>(ranges, n, first, end, stride);
    for (uint32_t slot = first; slot < end; slot += stride) {
        uint32_t p = order ? __ldg(order + slot) : slot;
        uint2 pr = __ldg(pairs + p);
        auto run_pair = [&](const Shape& L, const Shape& R) {
            Simplex<true> s;
            uint8_t st = gjk_intersect<WARP, true>(L, R, cfg.max_iterations, s);
            ContactOut c;
            c.normal = v3(0.f, 0.f, 0.f); c.depth = 0.f; c.point = v3(0.f, 0.f, 0.f);
            if (st == ST_OK && strategy == 0) {
                if (WARP) st = epa_process_warp(L, R, s, cfg.epa_tolerance, cfg.max_iterations, reinterpret_cast<WarpPolytopeSmem*>(dyn_smem)[threadIdx.x / 32], c);
                else st = epa_process<false>(L, R, s, cfg.epa_tolerance, cfg.max_iterations, P, c);
                if (st != ST_OK) { c.normal = v3(0.f, 0.f, 0.f); c.depth = 0.f; c.point = v3(0.f, 0.f, 0.f); }
            }
            if (!WARP || (threadIdx.x & 31) == 0) {
                contacts[2 * (size_t)p] = make_float4(c.normal.x, c.normal.y, c.normal.z, c.depth);
                contacts[2 * (size_t)p + 1] = make_float4(c.point.x, c.point.y, c.point.z, 0.f);
                status[p] = st;
                if (st == ST_CAPACITY) {  // hand the pair to the overflow tier
                    overflow_append(ovf_scratch, p);
                }
            }
        };
        if constexpr (WARP && B3_WARP_SHAPES_SMEM) {
            // one copy of the pair's two shape records per warp, behind the polytope slices (stage_warp_shapes)
            Shape* sh = reinterpret_cast<Shape*>(dyn_smem + warp_shapes_offset()) + 2 * (threadIdx.x / 32);
            const ShapeSetDev* const sets[2] = {&S, &S};
            const uint32_t idx[2] = {pr.x, pr.y};
            stage_warp_shapes<2>(sh, sets, idx);
            run_pair(sh[0], sh[1]);
        } else {
            Shape L, R;
            load_pair_shape(L, S, pr.x);
            load_pair_shape(R, S, pr.y);
            run_pair(L, R);
        }
        if (WARP) __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------------------
// Thread-per-pair contact path, split in two kernels so that lanes stay busy:
//   gjk_phase_kernel   GJK::intersect per pair; misses are final (zero contact + status); every hit that needs
//                      EPA is appended (warp-aggregated) to a device queue with its 4-point simplex (v, sup_a);
//   epa_refill_kernel  persistent warps; each LANE owns one queued pair at a time and all 32 lanes advance by one
//                      EPA iteration per trip; a lane whose pair finishes refills from the queue at once. Without
//                      this, a warp ran as long as its slowest pair (iteration counts span 2..100) and ncu showed
//                      4.4 of 32 lanes active on average.
// ------------------------------------------------------------------------------------------------------------
constexpr int EPAQ_REC_F4 = 7;             // queue record: 4 x (v, sup_a) = 24 floats, pair index, 3 pad words
constexpr size_t EPAQ_HEADER_BYTES = 1024;  // u32 words: [1] = consumer cursor, [64 + b] = records queued for bin b
constexpr int EPA_ECAP_SMEM = 24;            // horizon capacity of the thread-per-pair EPA tier (shared memory)
constexpr int EPAQ_BINS = 49;               // thread-path bins (kindL * 7 + kindR)

// The queue is segmented by pair bin: bin b's records live at [bin_start[b], bin_start[b] + count[b]) — the bin's
// own slot range of the sorted pair order, so a segment can never overflow — and the EPA kernel consumes the
// segments in `BinOrder` (heaviest expected EPA work first). Lanes of an EPA warp therefore hold pairs of the same
// shape-kind combination (same support code, similar polytope sizes: their local-memory loads stay
// sector-coalesced) except where one segment runs out. Without binning everything is bin 0 at base 0.
// bin[i] = the i-th bin to consume; tier[i] = the shared-memory EPA tier its pairs start in (epa_tiers.cuh).
// refill[i] = idle lanes a warp waits for before it fetches from segment i (1 = every finished lane refills at once;
// near 32 = "cohort" mode: the warp's lanes start their pairs together and stay at the same EPA iteration).
struct BinOrder { uint8_t bin[64]; uint8_t tier[64]; uint8_t refill[64]; };
// queue header words
constexpr int QH_CURSOR = 1;        // epa_refill_kernel's cursor
constexpr int QH_DONE = 4;          // pairs the EPA kernel has finished
constexpr int QH_STAGE_CURSOR = 32; // + static stage (epa_tiers_kernel)
constexpr int QH_PROMO_COUNT = 8;   // + tier (1..3)
constexpr int QH_PROMO_CURSOR = 12; // + tier (1..3)
constexpr int QH_BIN_COUNT = 64;    // + bin

__global__ void __launch_bounds__(TPB) gjk_phase_kernel(ShapeSetDev S, const uint2* __restrict__ pairs, uint32_t n,
                                                        const uint32_t* __restrict__ order, const uint32_t* __restrict__ ranges,
                                                        const uint8_t* __restrict__ keys, uint32_t default_key,
                                                        const uint32_t* __restrict__ bin_start, SettingsDev cfg, int strategy,
                                                        unsigned char* __restrict__ queue,
                                                        float4* __restrict__ contacts, uint8_t* __restrict__ status) {
    uint32_t first, end, stride;
    work_range<false>(ranges, n, first, end, stride);
    uint32_t* q_count = reinterpret_cast<uint32_t*>(queue) + QH_BIN_COUNT;
    float4* q_recs = reinterpret_cast<float4*>(queue + EPAQ_HEADER_BYTES);
    const uint32_t lane = threadIdx.x & 31u;
    // keep the warp convergent for the ballot below: iterate until every lane is past the end
    for (uint32_t slot0 = first - lane; slot0 < end; slot0 += stride) {
        const uint32_t slot = slot0 + lane;
        const bool valid = slot < end;
        uint32_t p = 0;
        uint8_t st = ST_MISS;
        Simplex<true> s;
        if (valid) {
            p = order ? __ldg(order + slot) : slot;
            uint2 pr = __ldg(pairs + p);
            Shape L, R;
            load_pair_shape(L, S, pr.x);
            load_pair_shape(R, S, pr.y);
            st = gjk_intersect<false, true>(L, R, cfg.max_iterations, s);
        }
        const bool to_epa = valid && st == ST_OK && strategy == 0;
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, to_epa);
        if (mask) {
            if (to_epa) {
                // one atomic per distinct bin per warp (a warp's slots are consecutive in bin order: usually one bin)
                const uint32_t key = keys ? (uint32_t)keys[p] : default_key;
                const uint32_t peers = __match_any_sync(mask, key);
                const uint32_t leader = __ffs(peers) - 1;
                uint32_t base = 0;
                if (lane == leader) base = atomicAdd(q_count + key, (uint32_t)__popc(peers));
                base = __shfl_sync(peers, base, leader) + (bin_start ? __ldg(bin_start + key) : 0u);
                float4* r = q_recs + (size_t)(base + __popc(peers & ((1u << lane) - 1u))) * EPAQ_REC_F4;
                r[0] = make_float4(s.v[0].x, s.v[0].y, s.v[0].z, s.a[0].x);
                r[1] = make_float4(s.a[0].y, s.a[0].z, s.v[1].x, s.v[1].y);
                r[2] = make_float4(s.v[1].z, s.a[1].x, s.a[1].y, s.a[1].z);
                r[3] = make_float4(s.v[2].x, s.v[2].y, s.v[2].z, s.a[2].x);
                r[4] = make_float4(s.a[2].y, s.a[2].z, s.v[3].x, s.v[3].y);
                r[5] = make_float4(s.v[3].z, s.a[3].x, s.a[3].y, s.a[3].z);
                r[6] = make_float4(__uint_as_float(p), 0.f, 0.f, 0.f);
            }
        }
        if (valid && !to_epa) {
            contacts[2 * (size_t)p] = make_float4(0.f, 0.f, 0.f, 0.f);
            contacts[2 * (size_t)p + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
            status[p] = st;
        }
    }
}

// GLOBAL_POLY: the polytopes live in a global-memory slab, one contiguous EPA_SLAB_BYTES record per thread, instead of
// per-thread local memory (which the hardware interleaves by lane at 4-byte granularity). Contiguous records mean every
// 32-byte sector a lane fetches holds only its own data, whatever the other lanes of the warp are doing.
constexpr size_t EPA_SLAB_VV = 1280, EPA_SLAB_FN = (size_t)EPA_FCAP * 16, EPA_SLAB_FI = (size_t)EPA_FCAP * 4;
constexpr size_t EPA_SLAB_BYTES = 2 * EPA_SLAB_VV + EPA_SLAB_FN + EPA_SLAB_FI;  // 7680 = 60 x 128
static_assert(EPA_SLAB_VV >= (size_t)EPA_VCAP * 12 && EPA_SLAB_BYTES % 128 == 0, "slab layout");

#ifndef B3_EPA_BLOCKS
#define B3_EPA_BLOCKS 4  // resident blocks per SM the EPA kernel is compiled for (register cap) and its slab is sized for
#endif
#ifndef B3_EXPERIMENTS
struct PipeStore { float4 *vv4, *va4, *ring; uint32_t* proc; uint16_t* mov; float* stage; int stride; };  // (epa_pipe.cuh is an experiment)
constexpr int EPA_RING = 0, EPA_PCAP = 0;
constexpr size_t EPA_PIPE_SLAB_BYTES = 0;
#define B3_PIPE_A 0
#endif
// dynamic shared memory of the PIPE form: face ring | horizon edges | moves | removed-face slots (epa_pipe.cuh)
constexpr size_t EPA_PIPE_RING_BYTES = B3_PIPE_A ? (size_t)EPA_RING * 8 * TPB * 16 : 0;
constexpr size_t EPA_PIPE_SMEM = EPA_PIPE_RING_BYTES + (size_t)EPA_ECAP_SMEM * TPB * 2 + (size_t)EPA_PCAP * TPB * 2 + (size_t)EPA_PCAP * TPB * 4 + (size_t)24 * TPB * 4;
template <bool GLOBAL_POLY, bool PIPE>
__global__ void __launch_bounds__(TPB, B3_EPA_BLOCKS) epa_refill_kernel(ShapeSetDev S, const uint2* __restrict__ pairs, SettingsDev cfg,
                                                         const uint32_t* __restrict__ bin_start, BinOrder bo,
                                                         unsigned char* __restrict__ queue, unsigned char* __restrict__ ovf_scratch,
                                                         unsigned char* __restrict__ slab,
                                                         float4* __restrict__ contacts, uint8_t* __restrict__ status) {
    // segment table in consumption order: s_pref[i] = records before segment i, s_base[i] = its first record
    __shared__ uint32_t s_pref[EPAQ_BINS + 1];
    __shared__ uint32_t s_base[EPAQ_BINS];
    __shared__ uint8_t s_refill[EPAQ_BINS + 1];
    if (threadIdx.x == 0) {
        const uint32_t* q_count = reinterpret_cast<const uint32_t*>(queue) + QH_BIN_COUNT;
        uint32_t acc = 0;
        for (int i = 0; i < EPAQ_BINS; ++i) {
            const uint32_t b = bo.bin[i];
            s_pref[i] = acc;
            s_base[i] = bin_start ? bin_start[b] : 0u;
            s_refill[i] = bo.refill[i];
            acc += q_count[b];
        }
        s_pref[EPAQ_BINS] = acc;
        s_refill[EPAQ_BINS] = 1;
    }
    __syncthreads();
    float lvv[GLOBAL_POLY ? 1 : EPA_VCAP * 3];
    float lva[GLOBAL_POLY ? 1 : EPA_VCAP * 3];
    float4 lfn[GLOBAL_POLY ? 1 : EPA_FCAP];
    uint32_t lfi[GLOBAL_POLY ? 1 : EPA_FCAP];
    PolytopeStore P;
    PipeStore Q;
    if (GLOBAL_POLY) {
        unsigned char* mine = slab + ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * (PIPE ? EPA_PIPE_SLAB_BYTES : EPA_SLAB_BYTES);
        P.fn = reinterpret_cast<float4*>(mine);
        P.fi = reinterpret_cast<uint32_t*>(mine + EPA_SLAB_FN);
        P.vv = reinterpret_cast<float*>(mine + EPA_SLAB_FN + EPA_SLAB_FI);
        P.va = reinterpret_cast<float*>(mine + EPA_SLAB_FN + EPA_SLAB_FI + EPA_SLAB_VV);
        Q.vv4 = reinterpret_cast<float4*>(mine + EPA_SLAB_FN + EPA_SLAB_FI);
        Q.va4 = Q.vv4 + EPA_VCAP;
    } else {
        P.vv = lvv; P.va = lva; P.fn = lfn; P.fi = lfi;
    }
    P.fcap = EPA_FCAP;
#ifdef B3_EPA_LOCAL_EDGES
    uint16_t led[EPA_ECAP];
    P.ed = led; P.ecap = EPA_ECAP; P.estride = 1;
#else
    // Horizon edge list: one shared-memory column per thread. Pass B indexes it data-dependently (every lane at its
    // own position), which shared memory serves without the 32-lines-per-request cost of divergent local-memory
    // addresses. The horizon of a well-formed polytope is short (<= 12 edges for 99.7 % of mixed pairs, <= 16 for
    // all but ~1e-4); a pair that needs more than EPA_ECAP_SMEM goes to the overflow tier like one that outgrows
    // the face list.
    extern __shared__ __align__(16) unsigned char epa_dyn[];
    __shared__ uint16_t s_ed[PIPE ? 1 : EPA_ECAP_SMEM * TPB];
    if (PIPE) {
        Q.ring = reinterpret_cast<float4*>(epa_dyn) + threadIdx.x;
        uint32_t* w32 = reinterpret_cast<uint32_t*>(epa_dyn + EPA_PIPE_RING_BYTES);
        Q.proc = w32 + threadIdx.x;
        Q.stage = reinterpret_cast<float*>(w32 + EPA_PCAP * TPB) + threadIdx.x;
        uint16_t* w16 = reinterpret_cast<uint16_t*>(w32 + EPA_PCAP * TPB + 24 * TPB);
        P.ed = w16 + threadIdx.x;
        Q.mov = w16 + EPA_ECAP_SMEM * TPB + threadIdx.x;
        Q.stride = TPB;
    } else {
        P.ed = s_ed + threadIdx.x;
        Q.ring = nullptr; Q.proc = nullptr; Q.mov = nullptr; Q.stage = nullptr; Q.stride = TPB;
    }
    P.ecap = EPA_ECAP_SMEM; P.estride = TPB;
#endif
    const uint32_t count = s_pref[EPAQ_BINS];
#ifdef B3_TIER_TRACE
    if (threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); atomicMin(reinterpret_cast<unsigned long long*>(queue) + 8, t_); }
#endif
    uint32_t* q_cursor = reinterpret_cast<uint32_t*>(queue) + 1;
    const float4* q_recs = reinterpret_cast<const float4*>(queue + EPAQ_HEADER_BYTES);
    const uint32_t lane = threadIdx.x & 31u;
    bool active = false, exhausted = false;
    int seg = 0;  // the cursor only grows, so each lane walks the segment table forward
    Shape L, R;
    int nv = 0, nf = 0;
    uint32_t it = 0, p = 0;
    V3 vmax = v3(0.f, 0.f, 0.f);
    int best = 0;
    // Idle lanes a warp waits for before it fetches (warp-uniform; follows the segment of the warp's last fetch). The
    // segments of ball-like pairs (Sphere / Capsule on both sides) run as cohorts: their EPA iteration counts are nearly
    // constant (5th..95th percentile 64..68 for Sphere-Sphere at the default tolerance), so lanes that start together
    // keep the same face count and pass A's loop runs full — refilled one by one they sit at random phases of a
    // 4 -> 136-face growth and the warp pays the largest polytope for every lane.
    int refill_min = 1;
    for (;;) {
        // refill idle lanes from the queue (one atomic per warp per fetch)
        const uint32_t need = __ballot_sync(0xFFFFFFFFu, !active && !exhausted);
        if (need && (__popc(need) >= refill_min || !__any_sync(0xFFFFFFFFu, active))) {
            uint32_t base = 0;
            const uint32_t leader = __ffs(need) - 1;
            if (lane == leader) base = atomicAdd(q_cursor, (uint32_t)__popc(need));
            base = __shfl_sync(0xFFFFFFFFu, base, leader);
            if (!active && !exhausted) {
                const uint32_t idx = base + __popc(need & ((1u << lane) - 1u));
                if (idx < count) {
                    while (idx >= s_pref[seg + 1]) ++seg;
                    const float4* r = q_recs + (size_t)(s_base[seg] + (idx - s_pref[seg])) * EPAQ_REC_F4;
                    float4 a = __ldg(r), b = __ldg(r + 1), c = __ldg(r + 2), d = __ldg(r + 3), e = __ldg(r + 4), f = __ldg(r + 5), g = __ldg(r + 6);
                    Simplex<true> s;
                    s.n = 4;
                    s.v[0] = v3(a.x, a.y, a.z); s.a[0] = v3(a.w, b.x, b.y);
                    s.v[1] = v3(b.z, b.w, c.x); s.a[1] = v3(c.y, c.z, c.w);
                    s.v[2] = v3(d.x, d.y, d.z); s.a[2] = v3(d.w, e.x, e.y);
                    s.v[3] = v3(e.z, e.w, f.x); s.a[3] = v3(f.y, f.z, f.w);
                    p = __float_as_uint(g.x);
                    uint2 pr = __ldg(pairs + p);
                    load_pair_shape(L, S, pr.x);
                    load_pair_shape(R, S, pr.y);
#ifdef B3_EXPERIMENTS
                    if constexpr (PIPE) epa_init_pipe(s, P, Q, nv, nf, it, vmax, best);
                    else
#endif
                        epa_init<false>(s, P, nv, nf, it, vmax, best);
                    active = true;
                } else {
                    seg = EPAQ_BINS;
                    exhausted = true;
#ifdef B3_TIER_TRACE
                    { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); atomicMin(reinterpret_cast<unsigned long long*>(queue) + 9, t_); atomicMax(reinterpret_cast<unsigned long long*>(queue) + 10, t_); }
#endif
                }
            }
            // the threshold of the segment the first fetching lane landed in
            refill_min = s_refill[__shfl_sync(0xFFFFFFFFu, seg, leader)];
        }
        if (!__any_sync(0xFFFFFFFFu, active)) break;
        if (active) {
            ContactOut c;
            uint8_t st;
#ifdef B3_EXPERIMENTS
            if constexpr (PIPE) st = epa_iterate_pipe(L, R, cfg.epa_tolerance, cfg.max_iterations, P, Q, nv, nf, it, vmax, best, c);
            else
#endif
                st = epa_iterate<false>(L, R, cfg.epa_tolerance, cfg.max_iterations, P, nv, nf, it, vmax, best, c);
            if (st != EPA_CONTINUE) {
                if (st != ST_OK) { c.normal = v3(0.f, 0.f, 0.f); c.depth = 0.f; c.point = v3(0.f, 0.f, 0.f); }
                contacts[2 * (size_t)p] = make_float4(c.normal.x, c.normal.y, c.normal.z, c.depth);
                contacts[2 * (size_t)p + 1] = make_float4(c.point.x, c.point.y, c.point.z, 0.f);
                status[p] = st;
                if (st == ST_CAPACITY) {  // hand the pair to the overflow tier
                    overflow_append(ovf_scratch, p);
                }
                active = false;
            }
        }
    }
#ifdef B3_TIER_TRACE
    if ((threadIdx.x & 31) == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); atomicMax(reinterpret_cast<unsigned long long*>(queue) + 11, t_); atomicMin(reinterpret_cast<unsigned long long*>(queue) + 13, t_); }
#endif
    overflow_producer_exit(ovf_scratch);
}

#ifdef B3_EXPERIMENTS
}  // namespace b3
#include "epa_tiers.cuh"  // needs load_pair_shape
namespace b3 {

// Shared-memory EPA in size tiers (epa_tiers.cuh). One warp per block, TIER_WARPS_PER_SM blocks per SM, persistent:
// every warp walks the static stages (the queue segments that start in tier 2, then tier 1, then tier 0), then
// serves the promotion lists until every queued pair has a final status.
constexpr size_t TIER_VA_BYTES = (size_t)NUM_SMS * TIER_WARPS_PER_SM * EPA_VCAP * 32 * sizeof(float4);
B3_DEV size_t align16(size_t x) { return (x + 15) & ~(size_t)15; }

__global__ void __launch_bounds__(32, TIER_WARPS_PER_SM) epa_tiers_kernel(ShapeSetDev S, const uint2* __restrict__ pairs, SettingsDev cfg,
                                                                         const uint32_t* __restrict__ bin_start, BinOrder bo,
                                                                         unsigned char* __restrict__ queue, uint32_t n_cap,
                                                                         unsigned char* __restrict__ ovf_scratch, float4* __restrict__ contacts,
                                                                         uint8_t* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char tier_pool[];
    __shared__ uint32_t s_pref[EPAQ_BINS + 1];
    __shared__ uint32_t s_base[EPAQ_BINS];
    uint32_t* qh = reinterpret_cast<uint32_t*>(queue);
    if (threadIdx.x == 0) {
        uint32_t acc = 0;
        for (int i = 0; i < EPAQ_BINS; ++i) {
            const uint32_t b = bo.bin[i];
            s_pref[i] = acc;
            s_base[i] = bin_start ? bin_start[b] : 0u;
            acc += qh[QH_BIN_COUNT + b];
        }
        s_pref[EPAQ_BINS] = acc;
    }
    __syncwarp();
    const uint32_t total = s_pref[EPAQ_BINS];
    if (total == 0) { overflow_producer_exit(ovf_scratch); return; }

    TierShared sh;
    sh.S = &S;
    sh.pairs = pairs;
    sh.cfg = cfg;
    sh.q_recs = reinterpret_cast<const float4*>(queue + EPAQ_HEADER_BYTES);
    unsigned char* lists = queue + EPAQ_HEADER_BYTES + (size_t)n_cap * EPAQ_REC_F4 * 16;
    sh.promo_count = qh + QH_PROMO_COUNT;
    sh.promo_list[0] = nullptr;
    for (int t = 1; t <= 3; ++t) sh.promo_list[t] = reinterpret_cast<uint32_t*>(lists) + (size_t)(t - 1) * n_cap;
    sh.done = qh + QH_DONE;
    sh.va = reinterpret_cast<float4*>(lists + align16((size_t)n_cap * 12)) + (size_t)blockIdx.x * EPA_VCAP * 32;
    sh.ovf_scratch = ovf_scratch;
    sh.contacts = contacts;
    sh.status = status;
    sh.pool = tier_pool;

    TierSource src;
    src.promo = false;
    src.count = nullptr; src.list = nullptr;
    src.s_pref = s_pref; src.s_base = s_base;
#ifdef B3_TIER_TRACE
    // development: [8..] u64 header words = earliest start, latest end of each static stage, latest exit (globaltimer ns)
    unsigned long long* tr = reinterpret_cast<unsigned long long*>(queue) + 8;
#define B3_TRACE(i, op) do { if (threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); op(tr + (i), t_); } } while (0)
    B3_TRACE(0, atomicMin);
#else
#define B3_TRACE(i, op) do { } while (0)
#endif
    // static stages: maximal runs of consecutive queue segments that start in the same tier (at most 8 cursors)
    int stage = 0;
    for (int i = 0; i < EPAQ_BINS && stage < 8;) {
        int j = i + 1;
        while (j < EPAQ_BINS && bo.tier[j] == bo.tier[i]) ++j;
        src.lo = s_pref[i]; src.hi = s_pref[j]; src.cursor = qh + QH_STAGE_CURSOR + stage;
        if (src.hi > src.lo) {
            if (bo.tier[i] == 2) tier_run<4>(src, sh);
            else if (bo.tier[i] == 1) tier_run<2>(src, sh);
            else tier_run<1>(src, sh);
        }
        ++stage;
        i = j;
    }
    B3_TRACE(3, atomicMax);

    src.promo = true;
    src.lo = 0; src.hi = 0;
    for (;;) {
        bool any = false;
        src.cursor = qh + QH_PROMO_CURSOR + 1; src.count = qh + QH_PROMO_COUNT + 1; src.list = sh.promo_list[1];
        any |= tier_run<2>(src, sh);
        src.cursor = qh + QH_PROMO_CURSOR + 2; src.count = qh + QH_PROMO_COUNT + 2; src.list = sh.promo_list[2];
        any |= tier_run<4>(src, sh);
        src.cursor = qh + QH_PROMO_CURSOR + 3; src.count = qh + QH_PROMO_COUNT + 3; src.list = sh.promo_list[3];
        any |= tier_run<8>(src, sh);
        uint32_t d = 0;
        if (threadIdx.x == 0) d = ld_volatile_u32(sh.done);
        d = __shfl_sync(FULL, d, 0);
        if (d >= total) break;
        if (!any) __nanosleep(1000);
    }
    B3_TRACE(4, atomicMax);
    overflow_producer_exit(ovf_scratch);
}

#else
constexpr size_t TIER_VA_BYTES = 0;
#endif  // B3_EXPERIMENTS

// Overflow tier of the contact kernels (epa_big.cuh): the fast-tier kernels append every pair they leave at
// ST_CAPACITY to a device list; one warp per listed pair redoes GJK + EPA against a per-warp global-memory slab.
// Rare (~6e-5 of mixed pairs), but those pairs are the long tail of the reference's cost.
constexpr int OVF_WARPS = NUM_SMS;        // one 32-thread block per SM
constexpr size_t OVF_HEADER_BYTES = 256 + (size_t)OVF_MAX * 4 + (size_t)HUGE_MAX * 8;  // header, overflow list, third-tier list + its keys

__global__ void __launch_bounds__(32) gjk_contact_overflow_kernel(ShapeSetDev S, const uint2* __restrict__ pairs, SettingsDev cfg,
                                                                  unsigned char* __restrict__ scratch, int producers_running,
                                                                  float4* __restrict__ contacts, uint8_t* __restrict__ status) {
    // producers_running = 0: every producer kernel finished before this launch (stream order), the list is final.
    // producers_running = 1: launched on a side stream next to the thread-per-pair EPA kernel; entries are taken as
    // they appear, and the warp leaves once the producers have raised OVF_DONE and the list is drained. (The pairs
    // that end here are the longest jobs of the whole call, so they should start as early as possible.)
    uint32_t* hdr = reinterpret_cast<uint32_t*>(scratch);
    const uint32_t* list = reinterpret_cast<const uint32_t*>(scratch + 256);
    const uint32_t warp = blockIdx.x, lane = threadIdx.x;
    if (!producers_running && warp >= min(hdr[OVF_COUNT], OVF_MAX)) return;
    extern __shared__ __align__(16) unsigned char ovf_smem[];
    BigStore B = carve_big_store(scratch + OVF_HEADER_BYTES + (size_t)warp * EPA_BIG_SLAB_BYTES, ovf_smem);
    for (int k = lane; k < EPA_KEYS; k += 32) { B.head[k] = EPA_NIL; B.tail[k] = EPA_NIL; }
    __syncwarp();
    for (;;) {
        uint32_t i = 0xFFFFFFFFu;
        if (lane == 0) {
            for (;;) {
                // read the flag BEFORE the count: once the flag is up the count is final
                const uint32_t done = producers_running ? ld_volatile_u32(hdr + OVF_DONE) : 1u;
                const uint32_t cnt = min(ld_volatile_u32(hdr + OVF_COUNT), OVF_MAX);
                const uint32_t cur = ld_volatile_u32(hdr + OVF_CURSOR);
                if (cur < cnt) {
                    if (atomicCAS(hdr + OVF_CURSOR, cur, cur + 1) == cur) { i = cur; break; }
                } else if (done) {
                    break;
                } else {
                    __nanosleep(2000);
                }
            }
        }
        i = __shfl_sync(0xFFFFFFFFu, i, 0);
        if (i == 0xFFFFFFFFu) break;
        uint32_t p = 0;
        if (lane == 0) { while ((p = ld_volatile_u32(list + i)) == 0xFFFFFFFFu) __nanosleep(200); }
        p = __shfl_sync(0xFFFFFFFFu, p, 0);
        uint2 pr = __ldg(pairs + p);
        Shape L, R;
        load_pair_shape(L, S, pr.x);
        load_pair_shape(R, S, pr.y);
        Simplex<true> s;
        uint8_t st = gjk_intersect<true, true>(L, R, cfg.max_iterations, s);
        ContactOut c;
        c.normal = v3(0.f, 0.f, 0.f); c.depth = 0.f; c.point = v3(0.f, 0.f, 0.f);
        uint32_t grown_at = 0;  // vertices the polytope had when it outgrew this tier
        if (st == ST_OK) {
            st = epa_process_big(L, R, s, cfg.epa_tolerance, cfg.max_iterations, B, c);
            if (st == ST_CAPACITY) grown_at = (uint32_t)c.depth;
            if (st != ST_OK) { c.normal = v3(0.f, 0.f, 0.f); c.depth = 0.f; c.point = v3(0.f, 0.f, 0.f); }
        }
        if (lane == 0) {
            contacts[2 * (size_t)p] = make_float4(c.normal.x, c.normal.y, c.normal.z, c.depth);
            contacts[2 * (size_t)p + 1] = make_float4(c.point.x, c.point.y, c.point.z, 0.f);
            status[p] = st;
            if (st == ST_CAPACITY) {  // more than EPA_BIG_FCAP faces: the block-per-pair tier takes it (launched behind this kernel)
                const uint32_t slot = atomicAdd(hdr + HUGE_COUNT, 1u);
                if (slot < HUGE_MAX) {
                    uint32_t* hl = reinterpret_cast<uint32_t*>(scratch + 256 + (size_t)OVF_MAX * 4);
                    hl[slot] = p;
                    hl[HUGE_MAX + slot] = grown_at;
                }
            }
        }
        __syncwarp();
    }
}

// Third tier (epa_huge.cuh): one block per listed pair, launched behind the overflow kernel (the list is final by then).
__global__ void __launch_bounds__(HUGE_THREADS) gjk_contact_huge_kernel(ShapeSetDev S, const uint2* __restrict__ pairs, SettingsDev cfg,
                                                                        unsigned char* __restrict__ scratch, unsigned char* __restrict__ slab,
                                                                        float4* __restrict__ contacts, uint8_t* __restrict__ status) {
    __shared__ HugeShared sh;
    __shared__ uint32_t s_item, s_pair;
    __shared__ uint8_t s_st;
    extern __shared__ __align__(16) unsigned char huge_smem[];  // per-edge tables + the short lists of one add() (epa_huge.cuh)
    uint32_t* hdr = reinterpret_cast<uint32_t*>(scratch);
    const uint32_t* list = reinterpret_cast<const uint32_t*>(scratch + 256 + (size_t)OVF_MAX * 4);
    const uint32_t count = min(hdr[HUGE_COUNT], HUGE_MAX);
    if (count == 0) return;
    const HugeStore H = carve_huge_store(slab + (size_t)blockIdx.x * HUGE_SLAB_BYTES);
    huge_zero_tables(huge_smem, sh);
    for (;;) {
        if (threadIdx.x == 0) { s_item = atomicAdd(hdr + HUGE_CURSOR, 1u); s_pair = 0xFFFFFFFFu; }
        __syncthreads();
        const uint32_t i = s_item;
        if (i >= count) break;
        // Longest job first: a polytope that outgrew the overflow tier early keeps growing for more trips (final faces fall
        // steeply with the trip it passed 1024 at: 121 962 faces for trip 48, 1 034 for trip 99), and the largest pair alone is
        // as long as ten ordinary ones. The i-th pair in the order (grown_at, pair index): every block ranks the list itself.
        if (count <= HUGE_SORT_MAX) {
            for (uint32_t t = threadIdx.x; t < count; t += HUGE_THREADS) {
                const unsigned long long kt = ((unsigned long long)list[HUGE_MAX + t] << 32) | list[t];
                uint32_t rank = 0;
                for (uint32_t u = 0; u < count; ++u) rank += (((unsigned long long)list[HUGE_MAX + u] << 32) | list[u]) < kt ? 1u : 0u;
                if (rank == i) s_pair = list[t];
            }
            __syncthreads();
        }
        const uint32_t p = s_pair != 0xFFFFFFFFu ? s_pair : list[i];
        if (threadIdx.x < 32) {  // warp 0: the pair's records into shared memory, GJK, the terminal simplex into the vertex lists
            const uint2 pr = __ldg(pairs + p);
            if (threadIdx.x == 0) { load_pair_shape(sh.L, S, pr.x); load_pair_shape(sh.R, S, pr.y); }
            __syncwarp();
            Simplex<true> s;
            s.n = 0;
            const uint8_t st0 = gjk_intersect<true, true>(sh.L, sh.R, cfg.max_iterations, s);
            if (threadIdx.x == 0) {
                s_st = st0;
                if (st0 == ST_OK) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) { st3(sh.vv, k, s.v[k]); st3(sh.va, k, s.a[k]); }
                }
            }
        }
        __syncthreads();
        uint8_t st = s_st;
        ContactOut c;
        c.normal = v3(0.f, 0.f, 0.f); c.depth = 0.f; c.point = v3(0.f, 0.f, 0.f);
        if (st == ST_OK) {
            st = epa_process_huge(cfg.epa_tolerance, cfg.max_iterations, H, sh, huge_smem, c);
            if (st != ST_OK) { c.normal = v3(0.f, 0.f, 0.f); c.depth = 0.f; c.point = v3(0.f, 0.f, 0.f); }
        }
        if (threadIdx.x == 0) {
            contacts[2 * (size_t)p] = make_float4(c.normal.x, c.normal.y, c.normal.z, c.depth);
            contacts[2 * (size_t)p + 1] = make_float4(c.point.x, c.point.y, c.point.z, 0.f);
            status[p] = st;
        }
        __syncthreads();
    }
}

#ifdef B3_EXPERIMENTS
// Thread-per-pair intersect with lane refill ("intersect_refill" = k): persistent grid, every lane owns one pair at a time, all
// lanes advance by one iteration of the loop of gjk/mod.rs:96-111 per pass; lanes whose pair is decided wait until k lanes of the
// warp are idle (or none is busy) and then fetch together, so the fetch path (two record loads + the first support test, which
// alone decides about half of the pairs) runs with at least k lanes instead of once per finished pair. Same device functions in
// the same order per pair as gjk_intersect<false, false>: bit-identical results.
__global__ void __launch_bounds__(TPB) gjk_intersect_refill_kernel(ShapeSetDev S, const uint2* __restrict__ pairs, uint32_t n,
                                                                   const uint32_t* __restrict__ order, const uint32_t* __restrict__ ranges,
                                                                   SettingsDev cfg, uint32_t* __restrict__ cursor, int refill_min,
                                                                   uint8_t* __restrict__ status) {
    uint32_t begin = 0, end = n;
    if (ranges) { begin = ranges[0]; end = ranges[1]; }
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t FULL = 0xFFFFFFFFu;
    bool busy = false, drained = false;
    uint32_t p = 0, it = 0;
    Shape L, R;
    V3 d = v3(0.f, 0.f, 0.f);
    Simplex<false> s;
    s.clear();
    for (;;) {
        const uint32_t need = __ballot_sync(FULL, !busy && !drained);
        if (need && (__popc(need) >= refill_min || !__any_sync(FULL, busy))) {
            const int leader = __ffs(need) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(need));
            base = __shfl_sync(FULL, base, leader);
            if (!busy && !drained) {
                const uint32_t off = base + (uint32_t)__popc(need & ((1u << lane) - 1u));
                if (off >= end - begin) {
                    drained = true;
                } else {
                    p = order ? __ldg(order + begin + off) : begin + off;
                    const uint2 pr = __ldg(pairs + p);
                    load_pair_shape(L, S, pr.x);
                    load_pair_shape(R, S, pr.y);
                    d = shape_origin(R) - shape_origin(L);
                    if (is_zero(d)) d = v3(1.f, 1.f, 1.f);
                    V3 pv, pa;
                    minkowski_support<false>(L, R, d, pv, pa);
                    if (dot(pv, d) <= 0.f) {
                        status[p] = ST_MISS;
                    } else {
                        s.clear();
                        s.push(pv, pa);
                        d = -d;
                        it = 0;
                        busy = true;
                    }
                }
            }
        }
        if (!__any_sync(FULL, busy)) {
            if (!__any_sync(FULL, !drained)) break;
            continue;  // every fetched pair was decided by its first support point: fetch again
        }
        if (busy) {
            uint8_t st = 0xFF;  // 0xFF: the pair goes on
            if (it >= cfg.max_iterations) {
                st = ST_MAX_ITER;  // the reference returns None here too (gjk/mod.rs:112)
            } else {
                V3 pv, pa;
                minkowski_support<false>(L, R, d, pv, pa);
                if (dot(pv, d) <= 0.f) {
                    st = ST_MISS;
                } else {
                    s.push(pv, pa);
                    if (reduce_to_closest_feature(s, d)) st = ST_OK;
                    ++it;
                }
            }
            if (st != 0xFF) { status[p] = st; busy = false; }
        }
    }
}

#endif  // B3_EXPERIMENTS

// ------------------------------------------------------------------------------------------------------------
// GJK::distance
// ------------------------------------------------------------------------------------------------------------
template <bool WARP>
__global__ void __launch_bounds__(TPB) gjk_distance_kernel(ShapeSetDev S, const uint2* __restrict__ pairs, uint32_t n,
                                                           const uint32_t* __restrict__ order, const uint32_t* __restrict__ ranges,
                                                           SettingsDev cfg, float* __restrict__ out_ref, float* __restrict__ out_geo,
                                                           uint8_t* __restrict__ status) {
    uint32_t first, end, stride;
    work_range<WARP>(ranges, n, first, end, stride);
    for (uint32_t slot = first; slot < end; slot += stride) {
        uint32_t p = order ? __ldg(order + slot) : slot;
        uint2 pr = __ldg(pairs + p);
        float rv = 0.f, geo = 0.f;
        uint8_t st;
        if constexpr (WARP && B3_WARP_SHAPES_SMEM) {
            __shared__ Shape sh_all[WARPS_PER_BLOCK * 2];
            Shape* sh = sh_all + 2 * (threadIdx.x / 32);
            const ShapeSetDev* const sets[2] = {&S, &S};
            const uint32_t idx[2] = {pr.x, pr.y};
            stage_warp_shapes<2>(sh, sets, idx);
            st = gjk_distance<WARP>(sh[0], sh[1], cfg.distance_tolerance, cfg.max_iterations, rv, geo);
        } else {
            Shape L, R;
            load_pair_shape(L, S, pr.x);
            load_pair_shape(R, S, pr.y);
            st = gjk_distance<WARP>(L, R, cfg.distance_tolerance, cfg.max_iterations, rv, geo);
        }
        if (st != ST_OK) { rv = 0.f; geo = 0.f; }
        if (!WARP || (threadIdx.x & 31) == 0) { out_ref[p] = rv; out_geo[p] = geo; status[p] = st; }
    }
}

// Thread-per-pair distance with lane refill ("distance_refill"), the scheme of gjk_toi_refill_kernel: on a mixed scene a
// fifth of the pairs run to the iteration cap (curved shapes do not reach the 1e-6 absolute tolerance in f32) and hold 80 % of
// all trips of gjk/mod.rs:246-279's loop while the median pair takes 4, so a warp of gjk_distance_kernel<false> lives as
// long as its slowest pair. Here the grid is persistent, every LANE owns one pair, all lanes advance one trip per pass, and
// lanes whose pair is decided fetch together from one cursor once `refill_min` of a warp are idle. Per pair the arithmetic
// is gjk_distance<false>'s, operation for operation (same device functions, no contraction): bit-identical results.
__global__ void __launch_bounds__(TPB) gjk_distance_refill_kernel(ShapeSetDev S, const uint2* __restrict__ pairs, uint32_t n,
                                                                  const uint32_t* __restrict__ order, const uint32_t* __restrict__ ranges,
                                                                  SettingsDev cfg, uint32_t* __restrict__ cursor, int refill_min,
                                                                  float* __restrict__ out_ref, float* __restrict__ out_geo,
                                                                  uint8_t* __restrict__ status) {
    uint32_t begin = 0, end = n;
    if (ranges) { begin = ranges[0]; end = ranges[1]; }
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t FULL = 0xFFFFFFFFu;
    bool busy = false, drained = false;
    uint32_t p = 0, it = 0;
    Shape L, R;
    Simplex<false> s;
    s.clear();
    for (;;) {
        const uint32_t need = __ballot_sync(FULL, !busy && !drained);
        if (need && (__popc(need) >= refill_min || !__any_sync(FULL, busy))) {
            const int leader = __ffs(need) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(need));
            base = __shfl_sync(FULL, base, leader);
            if (!busy && !drained) {
                const uint32_t off = base + (uint32_t)__popc(need & ((1u << lane) - 1u));
                if (off >= end - begin) {
                    drained = true;
                } else {
                    p = order ? __ldg(order + begin + off) : begin + off;
                    const uint2 pr = __ldg(pairs + p);
                    load_pair_shape(L, S, pr.x);
                    load_pair_shape(R, S, pr.y);
                    V3 d = shape_origin(R) - shape_origin(L);
                    if (fabsf(d.x - 0.f) < F32_EPSILON && fabsf(d.y - 0.f) < F32_EPSILON && fabsf(d.z - 0.f) < F32_EPSILON) d = v3(1.f, 1.f, 1.f);
                    s.clear();
                    V3 pv, pa;
                    minkowski_support<false>(L, R, d, pv, pa);
                    s.push(pv, pa);
                    minkowski_support<false>(L, R, -d, pv, pa);
                    s.push(pv, pa);
                    it = 0;
                    busy = true;
                }
            }
        }
        if (!__any_sync(FULL, busy)) break;
        if (busy) {
            uint8_t st = 0xFF;  // 0xFF: the pair goes on
            float rv = 0.f, geo = 0.f;
            if (it >= cfg.max_iterations) {
                st = ST_MAX_ITER;
            } else {
                uint8_t cst = ST_OK;
                const V3 c = closest_point_to_origin(s, cst);
                if (cst != ST_OK) {
                    st = cst;
                } else if (fabsf(c.x - 0.f) < F32_EPSILON && fabsf(c.y - 0.f) < F32_EPSILON && fabsf(c.z - 0.f) < F32_EPSILON) {
                    st = ST_MISS;
                } else {
                    const V3 nd = -c;
                    V3 pv, pa;
                    minkowski_support<false>(L, R, nd, pv, pa);
                    const float dp = dot(pv, nd);
                    const float d0 = dot(s.v[0], nd);
                    if (dp - d0 < cfg.distance_tolerance) {
                        rv = dp - d0;
                        geo = sqrtf(dot(nd, nd));
                        st = ST_OK;
                    } else if (s.n >= 4) {
                        st = ST_REF_PANIC;  // unreachable, see gjk_distance
                    } else {
                        s.push(pv, pa);
                        ++it;
                    }
                }
            }
            if (st != 0xFF) {
                out_ref[p] = rv; out_geo[p] = geo; status[p] = st;
                busy = false;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// GJK::intersection_time_of_impact
// ------------------------------------------------------------------------------------------------------------
template <bool WARP>
__global__ void __launch_bounds__(TPB) gjk_toi_kernel(ShapeSetDev S0, ShapeSetDev S1, const uint2* __restrict__ pairs, uint32_t n,
                                                      const uint32_t* __restrict__ order, const uint32_t* __restrict__ ranges,
                                                      SettingsDev cfg, float4* __restrict__ contacts, uint8_t* __restrict__ status) {
    uint32_t first, end, stride;
    work_range<WARP>(ranges, n, first, end, stride);
    for (uint32_t slot = first; slot < end; slot += stride) {
        uint32_t p = order ? __ldg(order + slot) : slot;
        uint2 pr = __ldg(pairs + p);
        ContactOut c;
        c.normal = v3(0.f, 0.f, 0.f); c.depth = 0.f; c.point = v3(0.f, 0.f, 0.f);
        float toi = 0.f;
        uint8_t st;
        if constexpr (WARP && B3_WARP_SHAPES_SMEM) {
            __shared__ Shape sh_all[WARPS_PER_BLOCK * 4];
            Shape* sh = sh_all + 4 * (threadIdx.x / 32);
            const ShapeSetDev* const sets[4] = {&S0, &S0, &S1, &S1};
            const uint32_t idx[4] = {pr.x, pr.y, pr.x, pr.y};
            stage_warp_shapes<4>(sh, sets, idx);
            st = gjk_time_of_impact<WARP>(sh[0], sh[2], sh[1], sh[3], cfg.continuous_tolerance, cfg.toi_iteration_cap, c, toi);
        } else {
            Shape L0, R0, L1, R1;
            load_pair_shape(L0, S0, pr.x);
            load_pair_shape(R0, S0, pr.y);
            // only the end transforms' affine columns are read (translation deltas; the right end transform's
            // rotation/scale for the contact point)
            load_shape(L1, S1.recs, checked_index(S0, pr.x));
            load_shape(R1, S1.recs, checked_index(S0, pr.y));
            st = gjk_time_of_impact<WARP>(L0, L1, R0, R1, cfg.continuous_tolerance, cfg.toi_iteration_cap, c, toi);
        }
        if (st != ST_OK) { c.normal = v3(0.f, 0.f, 0.f); c.depth = 0.f; c.point = v3(0.f, 0.f, 0.f); toi = 0.f; }
        if (!WARP || (threadIdx.x & 31) == 0) {
            contacts[2 * (size_t)p] = make_float4(c.normal.x, c.normal.y, c.normal.z, c.depth);
            contacts[2 * (size_t)p + 1] = make_float4(c.point.x, c.point.y, c.point.z, toi);
            status[p] = st;
        }
    }
}

// Thread-per-pair time of impact with lane refill ("toi_refill"). In gjk_toi_kernel<false> a warp lives as long as its
// longest sweep: 90 % of the pairs of a mixed scene leave after at most two trips of the `while` loop (gjk/mod.rs:166) while
// 4.6 % take ten or more (up to 82) and hold 82 % of all trips, so ncu sees 2.6 of 32 lanes active
// (profiles/r1c_toi_ncu_full.txt). Here the grid is persistent, every LANE owns one pair at a time, all lanes advance by one
// trip per pass, and a lane whose pair is finished takes the next slot of the analytic range from one cursor (one atomic per
// warp and pass). The arithmetic of a pair is gjk_time_of_impact<false>'s, operation for operation (same device functions,
// no contraction), so the results are bit-identical; only the interleaving across lanes changes. The end transforms are
// read at the start (translation deltas) and the right one again for the contact point, instead of living in registers.
__global__ void __launch_bounds__(TPB) gjk_toi_refill_kernel(ShapeSetDev S0, ShapeSetDev S1, const uint2* __restrict__ pairs, uint32_t n,
                                                             const uint32_t* __restrict__ order, const uint32_t* __restrict__ ranges,
                                                             SettingsDev cfg, uint32_t* __restrict__ cursor, int refill_min,
                                                             float4* __restrict__ contacts, uint8_t* __restrict__ status) {
    uint32_t begin = 0, end = n;
    if (ranges) { begin = ranges[0]; end = ranges[1]; }
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t FULL = 0xFFFFFFFFu;
    bool busy = false, drained = false;
    uint32_t p = 0, right_idx = 0, iters = 0;
    Shape L0, R0;
    V3 ray = v3(0.f, 0.f, 0.f), v = ray, normal = ray, ray_origin = ray;
    float lambda = 0.f;
    Simplex<false> s;
    s.clear();
    for (;;) {
        // lanes without a pair fetch together once `refill_min` of them are waiting (or nothing else is running): the fetch path
        // (four record loads) then runs with that many lanes instead of once per finished pair
        const uint32_t need = __ballot_sync(FULL, !busy && !drained);
        if (need && (__popc(need) >= refill_min || !__any_sync(FULL, busy))) {
            const int leader = __ffs(need) - 1;
            uint32_t base = 0;
            if ((int)lane == leader) base = atomicAdd(cursor, (uint32_t)__popc(need));
            base = __shfl_sync(FULL, base, leader);
            if (!busy && !drained) {
                const uint32_t off = base + (uint32_t)__popc(need & ((1u << lane) - 1u));
                if (off >= end - begin) {
                    drained = true;
                } else {
                    p = order ? __ldg(order + begin + off) : begin + off;
                    const uint2 pr = __ldg(pairs + p);
                    right_idx = checked_index(S0, pr.y);
                    load_pair_shape(L0, S0, pr.x);
                    load_pair_shape(R0, S0, pr.y);
                    Shape L1, R1;
                    load_shape(L1, S1.recs, checked_index(S0, pr.x));
                    load_shape(R1, S1.recs, right_idx);
                    const V3 left_lin_vel = shape_origin(L1) - shape_origin(L0);
                    const V3 right_lin_vel = shape_origin(R1) - shape_origin(R0);
                    ray = right_lin_vel - left_lin_vel;
                    lambda = 0.f;
                    normal = v3(0.f, 0.f, 0.f);
                    ray_origin = v3(0.f, 0.f, 0.f);
                    s.clear();
                    V3 pv, pa;
                    minkowski_support<false>(L0, R0, -ray, pv, pa);
                    v = pv;
                    iters = 0;
                    busy = true;
                }
            }
        }
        if (!__any_sync(FULL, busy)) break;
        if (busy) {
            uint8_t st = 0xFF;  // 0xFF: the pair goes on
            if (dot(v, v) > cfg.continuous_tolerance) {
                if (iters++ >= cfg.toi_iteration_cap) {
                    st = ST_MAX_ITER;
                } else {
                    V3 pv, pa;
                    minkowski_support<false>(L0, R0, -v, pv, pa);
                    const float vp = dot(v, pv);
                    const float vr = dot(v, ray);
                    if (vp > lambda * vr) {
                        if (vr > 0.f) {
                            lambda = vp / vr;
                            if (lambda > 1.f) {
                                st = ST_MISS;
                            } else {
                                ray_origin = ray * lambda;
                                s.clear();
                                normal = -v;
                            }
                        } else {
                            st = ST_MISS;
                        }
                    }
                    if (st == 0xFF) {
                        if (s.n >= 4) {
                            st = ST_REF_PANIC;  // unreachable, see gjk_distance
                        } else {
                            s.push(pv - ray_origin, pa);
                            uint8_t cst = ST_OK;
                            v = closest_point_to_origin(s, cst);
                            if (cst != ST_OK) st = cst;
                        }
                    }
                }
            } else {
                st = (dot(v, v) <= cfg.continuous_tolerance) ? ST_OK : ST_NAN;  // NaN: the loop test and the <= test both fail -> None
            }
            if (st != 0xFF) {
                float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
                if (st == ST_OK) {
                    Shape R1;
                    load_shape(R1, S1.recs, right_idx);
                    const V3 nn = -normalize(normal);
                    const float depth = sqrtf(dot(v, v));
                    const V3 pt = toi_contact_point(R0, R1, lambda, ray_origin);
                    c0 = make_float4(nn.x, nn.y, nn.z, depth);
                    c1 = make_float4(pt.x, pt.y, pt.z, lambda);
                }
                contacts[2 * (size_t)p] = c0;
                contacts[2 * (size_t)p + 1] = c1;
                status[p] = st;
                busy = false;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// compound shapes: GJK::intersection_complex / distance_complex / intersection_complex_time_of_impact (gjk/mod.rs:362-523)
// ------------------------------------------------------------------------------------------------------------
// world transform of every primitive = compound transform * local transform, glam 0.8 Mat4::mul_mat4 (each result column is
// self.mul_vec4(other column): x_axis * v.x, then y_axis * v.y + r, z_axis * v.z + r, w_axis * v.w + r — four components each)
__global__ void __launch_bounds__(TPB) compound_compose_kernel(const float4* __restrict__ comp_xform, const float4* __restrict__ local,
                                                               const uint32_t* __restrict__ owner, uint32_t n_prims, float4* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_prims) return;
    const float4* a = comp_xform + 4 * (size_t)owner[i];
    const float4 ax = a[0], ay = a[1], az = a[2], aw = a[3];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 v = local[4 * (size_t)i + j];
        float4 r = make_float4(ax.x * v.x, ax.y * v.x, ax.z * v.x, ax.w * v.x);
        r = make_float4(ay.x * v.y + r.x, ay.y * v.y + r.y, ay.z * v.y + r.z, ay.w * v.y + r.w);
        r = make_float4(az.x * v.z + r.x, az.y * v.z + r.y, az.z * v.z + r.z, az.w * v.z + r.w);
        r = make_float4(aw.x * v.w + r.x, aw.y * v.w + r.y, aw.z * v.w + r.z, aw.w * v.w + r.w);
        out[4 * (size_t)i + j] = r;
    }
}

// every primitive x primitive combination of every compound pair, in the reference's loop order (left outer, right inner);
// seg[k] = first flat pair of compound pair k (seg[n] = total)
__global__ void __launch_bounds__(TPB) compound_expand_kernel(const uint2* __restrict__ comp_pairs, const unsigned long long* __restrict__ seg,
                                                              uint32_t n_pairs, const uint32_t* __restrict__ offsets, uint2* __restrict__ flat) {
    const unsigned long long total = seg[n_pairs];
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t lo = 0, hi = n_pairs;  // the last k with seg[k] <= t
        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (seg[mid] <= t) lo = mid; else hi = mid; }
        const uint2 cp = comp_pairs[lo];
        const uint32_t nr = offsets[cp.y + 1] - offsets[cp.y];
        const unsigned long long w = t - seg[lo];
        flat[t] = make_uint2(offsets[cp.x] + (uint32_t)(w / nr), offsets[cp.y] + (uint32_t)(w % nr));
    }
}

// The reference's reductions over one compound pair's primitive pairs, in order. mode 0: intersection_complex — CollisionOnly
// returns the first hit; FullResolution keeps the LAST maximum penetration depth (Iterator::max_by) and panics on a NaN depth
// among two or more hits (partial_cmp(..).unwrap(), :403-405) -> ST_REF_PANIC. mode 1: distance_complex — None at the first
// primitive pair that is not Some (a would-be panic there stays ST_REF_PANIC), else f32::min over all. mode 2:
// intersection_complex_time_of_impact — a would-be panic inside a primitive pair reached before the result is decided ->
// ST_REF_PANIC; CollisionOnly returns the first hit; FullResolution keeps the FIRST minimum time of impact (min_by with
// unwrap_or(Equal)). A primitive pair the device could not decide (ST_CAPACITY; the time-of-impact iteration cap ST_MAX_ITER,
// which the reference does not have) makes the compound result undecided too: its status is reported instead of a contact
// that might not be the reference's.
__global__ void __launch_bounds__(TPB) compound_reduce_kernel(int mode, int strategy, const unsigned long long* __restrict__ seg, uint32_t n_pairs,
                                                              const float4* __restrict__ contacts, const float* __restrict__ ref_value,
                                                              const uint8_t* __restrict__ status, float4* __restrict__ out_contacts,
                                                              float* __restrict__ out_value, uint8_t* __restrict__ out_status) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pairs) return;
    const unsigned long long b = seg[k], e = seg[k + 1];
    uint8_t st = ST_MISS;
    float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0;
    float value = 0.f;
    if (mode == 1) {
        bool have = false;
        float mn = 0.f;
        st = ST_OK;
        for (unsigned long long t = b; t < e; ++t) {
            const uint8_t s1 = status[t];
            if (s1 != ST_OK) { st = (s1 == ST_REF_PANIC) ? ST_REF_PANIC : ST_MISS; break; }
            const float rv = ref_value[t];
            mn = have ? rmin(rv, mn) : rv;
            have = true;
        }
        if (st == ST_OK && !have) st = ST_MISS;
        if (st == ST_OK) value = mn;
    } else {
        long long best = -1;
        int nhits = 0;
        bool nan_depth = false;
        uint8_t undecided = ST_OK;
        for (unsigned long long t = b; t < e; ++t) {
            const uint8_t s1 = status[t];
            if (s1 == ST_OK) {
                if (strategy == 1) { best = (long long)t; break; }  // CollisionOnly: the first hit returns
                ++nhits;
                if (mode == 0) {
                    const float d = contacts[2 * t].w;
                    if (best < 0) best = (long long)t;
                    else {
                        const float a = contacts[2 * (size_t)best].w;
                        if (a != a || d != d) nan_depth = true;
                        if (!(a > d)) best = (long long)t;  // max_by keeps the later element unless the earlier compares Greater
                    }
                } else {
                    if (best < 0 || contacts[2 * t + 1].w < contacts[2 * (size_t)best + 1].w) best = (long long)t;  // min_by: first minimum
                }
            } else if (mode == 2 && s1 == ST_REF_PANIC) {
                best = -2; break;  // the reference panics inside this primitive pair
            } else if (s1 == ST_CAPACITY || (mode == 2 && s1 == ST_MAX_ITER)) {
                if (undecided == ST_OK) undecided = s1;
            }
        }
        if (best == -2) st = ST_REF_PANIC;
        else if (undecided != ST_OK) st = undecided;
        else if (mode == 0 && nan_depth) st = ST_REF_PANIC;
        else if (best >= 0) { st = ST_OK; c0 = contacts[2 * (size_t)best]; c1 = contacts[2 * (size_t)best + 1]; }
        (void)nhits;
    }
    out_status[k] = st;
    if (mode == 1) out_value[k] = value;
    else { out_contacts[2 * (size_t)k] = c0; out_contacts[2 * (size_t)k + 1] = c1; }
}

int launch_compound_compose(cudaStream_t st, const float* d_comp_xform, const float* d_local, const uint32_t* d_owner, uint32_t n_prims, float* d_out) {
    if (n_prims == 0) return 0;
    compound_compose_kernel<<<(n_prims + TPB - 1) / TPB, TPB, 0, st>>>(reinterpret_cast<const float4*>(d_comp_xform), reinterpret_cast<const float4*>(d_local),
                                                                     d_owner, n_prims, reinterpret_cast<float4*>(d_out));
    return 1;
}
int launch_compound_expand(cudaStream_t st, const uint2* d_comp_pairs, const unsigned long long* d_seg, uint32_t n_pairs, uint64_t total,
                           const uint32_t* d_offsets, uint2* d_flat) {
    if (total == 0) return 0;
    uint64_t g = (total + TPB - 1) / TPB;
    if (g > (uint64_t)NUM_SMS * 32) g = (uint64_t)NUM_SMS * 32;
    compound_expand_kernel<<<(unsigned)g, TPB, 0, st>>>(d_comp_pairs, d_seg, n_pairs, d_offsets, d_flat);
    return 1;
}
int launch_compound_reduce(cudaStream_t st, int mode, int strategy, const unsigned long long* d_seg, uint32_t n_pairs, const float4* d_contacts,
                           const float* d_ref, const uint8_t* d_status, float4* d_out_contacts, float* d_out_value, uint8_t* d_out_status) {
    if (n_pairs == 0) return 0;
    compound_reduce_kernel<<<(n_pairs + TPB - 1) / TPB, TPB, 0, st>>>(mode, strategy, d_seg, n_pairs, d_contacts, d_ref, d_status, d_out_contacts, d_out_value,
                                                                    d_out_status);
    return 1;
}

// ------------------------------------------------------------------------------------------------------------
// contact compaction (hits only, tagged with the global pair index) — feeds the multi-GPU gather
// ------------------------------------------------------------------------------------------------------------
struct Contact40 { uint32_t pair_index, status; float4 a, b; };

__global__ void __launch_bounds__(TPB) compact_contacts_kernel(const float4* __restrict__ contacts, const uint8_t* __restrict__ status,
                                                               uint64_t n, uint64_t base, uint32_t* __restrict__ out, uint64_t capacity,
                                                               unsigned long long* __restrict__ count) {
    const uint32_t lane = threadIdx.x & 31u;
    for (uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull; i0 < n; i0 += (uint64_t)gridDim.x * blockDim.x) {
        uint64_t i = i0 + lane;
        bool hit = (i < n) && status[i] == ST_OK;
        uint32_t mask = __ballot_sync(0xFFFFFFFFu, hit);
        if (mask == 0) continue;
        unsigned long long wbase = 0;
        if (lane == 0) wbase = atomicAdd(count, (unsigned long long)__popc(mask));
        wbase = __shfl_sync(0xFFFFFFFFu, wbase, 0);
        if (hit) {
            uint64_t slot = wbase + __popc(mask & ((1u << lane) - 1u));
            if (slot < capacity) {
                float4 a = contacts[2 * i], b = contacts[2 * i + 1];
                uint32_t* o = out + slot * 10;  // 40-byte records
                o[0] = (uint32_t)(base + i); o[1] = ST_OK;
                o[2] = __float_as_uint(a.x); o[3] = __float_as_uint(a.y); o[4] = __float_as_uint(a.z); o[5] = __float_as_uint(a.w);
                o[6] = __float_as_uint(b.x); o[7] = __float_as_uint(b.y); o[8] = __float_as_uint(b.z); o[9] = __float_as_uint(b.w);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------------------
static inline int grid_for(uint64_t items, int per_block, int max_blocks) {
    uint64_t g = (items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > (uint64_t)max_blocks) g = max_blocks;
    return (int)g;
}

int launch_pack_shapes(cudaStream_t st, const uint8_t* d_kind, const float* d_params, const float* d_xform,
                       const uint32_t* d_mesh_id, uint32_t n, uint32_t n_meshes, ShapeRec* recs, uint32_t* meta, uint32_t* d_err, bool affine48) {
    if (n == 0) return 0;
    if (affine48)
        pack_shapes_kernel<true><<<(n + TPB - 1) / TPB, TPB, 0, st>>>(d_kind, reinterpret_cast<const float4*>(d_params),
                                                                      reinterpret_cast<const float4*>(d_xform), d_mesh_id, n, n_meshes, recs, meta, d_err);
    else
        pack_shapes_kernel<false><<<(n + TPB - 1) / TPB, TPB, 0, st>>>(d_kind, reinterpret_cast<const float4*>(d_params),
                                                                       reinterpret_cast<const float4*>(d_xform), d_mesh_id, n, n_meshes, recs, meta, d_err);
    return 1;
}

int launch_mesh_face_planes(cudaStream_t st, const float4* d_verts, uint4* d_faces, uint64_t nfaces, uint32_t* d_err) {
    if (nfaces == 0) return 0;
    mesh_face_planes_kernel<<<(unsigned)((nfaces + TPB - 1) / TPB), TPB, 0, st>>>(d_verts, d_faces, nfaces, d_err);
    return 1;
}

int launch_bin_pairs(cudaStream_t st, const ShapeSetDev& S, const uint2* d_pairs, uint32_t n, const BinScratch& B) {
    cudaMemsetAsync(B.hist, 0, 64 * sizeof(uint32_t), st);
    int g = grid_for(n, TPB, NUM_SMS * 8);
    bin_hist_kernel<<<g, TPB, 0, st>>>(S.meta, S.n, d_pairs, n, B.keys, B.hist);
    bin_scan_kernel<<<1, 32, 0, st>>>(B.hist, B.start, B.cursor, B.ranges, n);
    bin_scatter_kernel<<<g, TPB, 0, st>>>(B.keys, n, B.cursor, B.order);
    return 3;
}

// Thread kernels: grid-stride with enough blocks to fill the machine several times over (iteration counts vary
// per pair, so small blocks and many of them balance better than one fat wave). Warp kernels: one pair per warp.
static inline int thread_grid(uint32_t n) { return grid_for(n, TPB, NUM_SMS * 32); }
static inline int warp_grid(uint32_t n) { return grid_for(n, WARPS_PER_BLOCK, NUM_SMS * 32); }

// run_thread / run_warp select the execution shapes; with a BinScratch both read their ranges from device memory,
// without one the selected kernel covers all n pairs (the thread kernel scans polyhedron vertices sequentially,
// so "thread only" is always correct, just slow for polyhedra).
#define B3_LAUNCH(KERNEL, SMEM_WARP, ...)                                                                       \
    int launched = 0;                                                                                            \
    const uint32_t* order = B ? B->order : nullptr;                                                              \
    const uint32_t* ranges = B ? B->ranges : nullptr;                                                            \
    if (n == 0) return 0;                                                                                        \
    if (run_thread) { KERNEL<false><<<thread_grid(n), TPB, 0, st>>>(__VA_ARGS__); ++launched; }                  \
    if (run_warp) { KERNEL<true><<<warp_grid(n), TPB, SMEM_WARP, st>>>(__VA_ARGS__); ++launched; }               \
    return launched;

size_t intersect_scratch_bytes(uint32_t n) { return 256 + (size_t)n * 4; }

// grid of a persistent kernel: as many blocks as are resident at once (occupancy x SMs)
static int persistent_grid(const void* kernel) {
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, TPB, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
    return per_sm * NUM_SMS;
}
// Function attributes and occupancy answers belong to a device: cache them per device ordinal (a process may drive several).
constexpr int MAX_DEVICES = 64;
static int current_device_slot() { int d = 0; cudaGetDevice(&d); return (d >= 0 && d < MAX_DEVICES) ? d : 0; }
template <class K>
static int persistent_grid_cached(K kernel, int (&cache)[MAX_DEVICES]) {
    int& g = cache[current_device_slot()];
    if (!g) g = persistent_grid((const void*)kernel);
    return g;
}

int launch_intersect(cudaStream_t st, const ShapeSetDev& S, const uint2* d_pairs, uint32_t n, const BinScratch* B, bool run_thread,
                     bool run_warp, const SettingsDev& cfg, void* d_scratch, int refill_min, uint8_t* d_status) {
    if (n == 0) return 0;
    int launched = 0;
    const uint32_t* order = B ? B->order : nullptr;
    const uint32_t* ranges = B ? B->ranges : nullptr;
    if (run_thread) {
#ifdef B3_EXPERIMENTS
        if (d_scratch && refill_min > 0) {
            static int grid_cache[MAX_DEVICES];
            const int grid = persistent_grid_cached(gjk_intersect_refill_kernel, grid_cache);
            cudaMemsetAsync(d_scratch, 0, 4, st);
            gjk_intersect_refill_kernel<<<grid, TPB, 0, st>>>(S, d_pairs, n, order, ranges, cfg, reinterpret_cast<uint32_t*>(d_scratch),
                                                              refill_min > 32 ? 32 : refill_min, d_status);
            ++launched;
        } else if (d_scratch) {
            uint32_t* n_survivors = reinterpret_cast<uint32_t*>(d_scratch);
            uint32_t* survivors = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(d_scratch) + 256);
            cudaMemsetAsync(n_survivors, 0, 4, st);
            gjk_first_test_kernel<<<thread_grid(n), TPB, 0, st>>>(S, d_pairs, n, order, ranges, survivors, n_survivors, d_status);
            gjk_intersect_survivors_kernel<<<thread_grid(n), TPB, 0, st>>>(S, d_pairs, survivors, n_survivors, cfg, d_status);
            launched += 2;
        } else
#endif
        {
            (void)d_scratch; (void)refill_min;
            gjk_intersect_kernel<false><<<thread_grid(n), TPB, 0, st>>>(S, d_pairs, n, order, ranges, cfg, d_status);
            ++launched;
        }
    }
    if (run_warp) { gjk_intersect_kernel<true><<<warp_grid(n), TPB, 0, st>>>(S, d_pairs, n, order, ranges, cfg, d_status); ++launched; }
    return launched;
}
size_t contact_poly_slab_bytes() { return (size_t)NUM_SMS * B3_EPA_BLOCKS * TPB * (EPA_PIPE_SLAB_BYTES > EPA_SLAB_BYTES ? EPA_PIPE_SLAB_BYTES : EPA_SLAB_BYTES); }
size_t contact_overflow_scratch_bytes() { return OVF_HEADER_BYTES + (size_t)OVF_WARPS * EPA_BIG_SLAB_BYTES; }
size_t contact_huge_slab_bytes() { return (size_t)HUGE_BLOCKS * HUGE_SLAB_BYTES; }
// header | n GJK->EPA records | 3 promotion lists of n record indices | sup_a scratch of the shared-memory EPA warps
static size_t queue_lists_bytes(uint32_t n) { return ((size_t)n * 12 + 15) & ~(size_t)15; }
size_t contact_queue_bytes(uint32_t n) {
    return EPAQ_HEADER_BYTES + (size_t)n * EPAQ_REC_F4 * 16 + (n ? queue_lists_bytes(n) + TIER_VA_BYTES : 0);
}

static int epa_refill_grid() {
    static int blocks_per_device[MAX_DEVICES];
    int& blocks = blocks_per_device[current_device_slot()];
    if (!blocks) {
        int per_sm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, epa_refill_kernel<true, false>, TPB, 0) != cudaSuccess || per_sm < 1) per_sm = 2;
        if (per_sm > B3_EPA_BLOCKS) per_sm = B3_EPA_BLOCKS;  // (the slab of the global-polytope variant is sized for this many blocks per SM)
        if (const char* e = getenv("BAM3D_EPA_BLOCKS_PER_SM")) { int v = atoi(e); if (v >= 1 && v < per_sm) per_sm = v; }
        blocks = per_sm * NUM_SMS;
    }
    return blocks;
}

// Consumption order of the queue segments: expected EPA cost per pair, heaviest first, so that the long jobs
// (sphere-like Minkowski sums: ~67 iterations and ~136 faces at the default tolerance) start early, the light
// ones (box-box: ~6 iterations) fill the tail, and the pairs that end in the overflow tier are known early.
static BinOrder epa_bin_order() {
    static BinOrder bo;
    static bool ready = false;
    if (!ready) {
        // roundness: how much of the shape's surface is smooth
        // (cylinder below sphere + cuboid: sphere-cuboid pairs hit the iteration cap ten times as often as cylinder-cylinder ones)
        const int round[7] = {/*particle*/ 0, /*quad*/ 1, /*sphere*/ 10, /*cuboid*/ 2, /*cylinder*/ 5, /*capsule*/ 10, /*polyhedron*/ 0};
        int tier[EPAQ_BINS], score[EPAQ_BINS];
        int rim_boost = 150;  // measured on C2: 0 (after the curved bins) 8.4 ms, 150 (right after the ball bins) 7.8 ms, 250 (first) 9.3 ms
        if (const char* e = getenv("BAM3D_RIM_BOOST")) rim_boost = atoi(e);  // development knob
        for (int b = 0; b < EPAQ_BINS; ++b) {
            const int kl = b / 7, kr = b % 7;
            const bool ball_l = kl == K_SPHERE || kl == K_CAPSULE || kl == K_PARTICLE, ball_r = kr == K_SPHERE || kr == K_CAPSULE || kr == K_PARTICLE;
            const bool flat_l = kl == K_CUBOID || kl == K_QUAD, flat_r = kr == K_CUBOID || kr == K_QUAD;
            const int sum = round[kl] + round[kr];
            // starting tier of the shared-memory EPA: ball-like sums need ~71 vertices (tier 2), one curved side ~20-40
            // (tier 1), box-like pairs < 22 (tier 0); a wrong guess only costs a promotion
            tier[b] = (ball_l && ball_r && sum > 0) ? 2 : (sum >= 12 ? 1 : 0);
            // Cylinder against a flat-faced shape early (right after the ball-like bins): cheap on average, but these
            // bins hold the degenerate pairs (coplanar rim supports) that run to the iteration cap or end in the
            // overflow tier, the longest jobs of the call: the sooner they are found, the sooner the overflow tier
            // (running on the side stream) can start on them.
            const bool rim_vs_flat = (kl == K_CYLINDER && flat_r) || (kr == K_CYLINDER && flat_l);
            score[b] = sum + 100 * tier[b] + (rim_vs_flat ? rim_boost : 0);
            bo.bin[b] = (uint8_t)b;
        }
        for (int i = 1; i < EPAQ_BINS; ++i)  // stable insertion sort, descending score
            for (int j = i; j > 0 && score[bo.bin[j]] > score[bo.bin[j - 1]]; --j) { uint8_t t = bo.bin[j]; bo.bin[j] = bo.bin[j - 1]; bo.bin[j - 1] = t; }
        if (const char* e = getenv("BAM3D_BIN_ORDER")) {  // development knob: comma-separated bins to consume first, in this order
            int pos = 0;
            for (const char* q = e; *q && pos < EPAQ_BINS;) {
                const int b = atoi(q);
                if (b >= 0 && b < EPAQ_BINS) {
                    int at = pos;
                    while (at < EPAQ_BINS && bo.bin[at] != b) ++at;
                    if (at < EPAQ_BINS) { for (int j = at; j > pos; --j) bo.bin[j] = bo.bin[j - 1]; bo.bin[pos++] = (uint8_t)b; }
                }
                while (*q && *q != ',') ++q;
                if (*q == ',') ++q;
            }
        }
        for (int i = 0; i < EPAQ_BINS; ++i) bo.tier[i] = (uint8_t)tier[bo.bin[i]];
        int cohort = 28;  // idle lanes that trigger a fetch in the ball-like segments (1 = refill at once)
        if (const char* e = getenv("BAM3D_EPA_COHORT")) { int v = atoi(e); if (v >= 1 && v <= 32) cohort = v; }  // development knob
        for (int i = 0; i < EPAQ_BINS; ++i) bo.refill[i] = (uint8_t)(bo.tier[i] == 2 ? cohort : 1);
        ready = true;
    }
    return bo;
}

int launch_contact(cudaStream_t st, const ShapeSetDev& S, const uint2* d_pairs, uint32_t n, const BinScratch* B, bool run_thread,
                   bool run_warp, const SettingsDev& cfg, int strategy, uint32_t default_key, int epa_mode, bool epa_pipeline, const ContactSide* side,
                   void* d_poly_slab, void* d_overflow_scratch, void* d_huge_slab, void* d_queue, float4* d_contacts, uint8_t* d_status) {
    unsigned char* ovf = reinterpret_cast<unsigned char*>(d_overflow_scratch);
    unsigned char* queue = reinterpret_cast<unsigned char*>(d_queue);
    if (n == 0) return 0;
    int launched = 0;
    const uint32_t* order = B ? B->order : nullptr;
    const uint32_t* ranges = B ? B->ranges : nullptr;
    static bool smem_opt_in_per_device[MAX_DEVICES];
    bool& smem_opt_in = smem_opt_in_per_device[current_device_slot()];
    if (!smem_opt_in) {
        cudaFuncSetAttribute(gjk_contact_overflow_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EPA_BIG_SMEM_BYTES);
        cudaFuncSetAttribute(gjk_contact_huge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HUGE_SMEM_BYTES);
#ifdef B3_EXPERIMENTS
        cudaFuncSetAttribute(epa_tiers_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaFuncSetAttribute(epa_refill_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)EPA_PIPE_SMEM);
        if (B3_PIPE_A) cudaFuncSetAttribute(epa_refill_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
#endif
        cudaFuncSetAttribute(gjk_contact_overflow_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        smem_opt_in = true;
    }
    cudaMemsetAsync(ovf, 0, 256, st);
    if (strategy == 0) cudaMemsetAsync(ovf + 256, 0xFF, (size_t)OVF_MAX * 4, st);
    // The overflow tier runs next to the thread-per-pair EPA kernel (side stream) when that kernel is the only producer
    // of its list: the pairs that end there are the longest jobs of the call, and the EPA kernel's tail leaves SMs idle.
    // (Not next to epa_tiers_kernel: the overflow blocks would take the shared memory its blocks need to become resident.)
    const bool overlap = strategy == 0 && run_thread && !run_warp && side != nullptr && epa_mode == 0;
    if (run_thread) {
        const uint8_t* keys = B ? B->keys : nullptr;
        const uint32_t* bin_start = B ? B->start : nullptr;
        cudaMemsetAsync(queue, 0, EPAQ_HEADER_BYTES, st);
#ifdef B3_TIER_TRACE
        cudaMemsetAsync(queue + 64, 0xFF, 16, st);
        cudaMemsetAsync(queue + 64 + 5 * 8, 0xFF, 8, st);
#endif
        gjk_phase_kernel<<<thread_grid(n), TPB, 0, st>>>(S, d_pairs, n, order, ranges, keys, default_key, bin_start, cfg, strategy, queue,
                                                         d_contacts, d_status);
        ++launched;
        if (overlap) cudaEventRecord(side->fork, st);
#ifdef B3_EXPERIMENTS
        if (strategy == 0 && epa_mode != 0) {
            cudaMemsetAsync(queue + EPAQ_HEADER_BYTES + (size_t)n * EPAQ_REC_F4 * 16, 0xFF, queue_lists_bytes(n), st);
            epa_tiers_kernel<<<NUM_SMS * TIER_WARPS_PER_SM, 32, TIER_POOL_BYTES, st>>>(S, d_pairs, cfg, bin_start, epa_bin_order(), queue, n, ovf,
                                                                                       d_contacts, d_status);
            ++launched;
#ifdef B3_TIER_TRACE
            {
                unsigned long long h[16];
                uint32_t w[32];
                cudaMemcpyAsync(h, queue + 64, sizeof(h), cudaMemcpyDeviceToHost, st);
                cudaMemcpyAsync(w, queue, sizeof(w), cudaMemcpyDeviceToHost, st);
                cudaStreamSynchronize(st);
                fprintf(stderr, "[tier trace] stage2 end %.3f ms, stage1 end %.3f, stage0 end %.3f, exit %.3f | promoted to t1 %u t2 %u t3 %u, done %u\n",
                        (h[1] - h[0]) * 1e-6, (h[2] - h[0]) * 1e-6, (h[3] - h[0]) * 1e-6, (h[4] - h[0]) * 1e-6, w[QH_PROMO_COUNT + 1],
                        w[QH_PROMO_COUNT + 2], w[QH_PROMO_COUNT + 3], w[QH_DONE]);
            }
#endif
        } else
#endif
        if (strategy == 0) {
            (void)epa_mode; (void)epa_pipeline;
#ifdef B3_EXPERIMENTS
            if (d_poly_slab && epa_mode == 0 && epa_pipeline)
                epa_refill_kernel<true, true><<<epa_refill_grid(), TPB, EPA_PIPE_SMEM, st>>>(S, d_pairs, cfg, bin_start, epa_bin_order(), queue, ovf,
                                                                                             reinterpret_cast<unsigned char*>(d_poly_slab), d_contacts, d_status);
            else
#endif
            if (d_poly_slab)
                epa_refill_kernel<true, false><<<epa_refill_grid(), TPB, 0, st>>>(S, d_pairs, cfg, bin_start, epa_bin_order(), queue, ovf,
                                                                                  reinterpret_cast<unsigned char*>(d_poly_slab), d_contacts, d_status);
            else
                epa_refill_kernel<false, false><<<epa_refill_grid(), TPB, 0, st>>>(S, d_pairs, cfg, bin_start, epa_bin_order(), queue, ovf, nullptr,
                                                                                   d_contacts, d_status);
            ++launched;
#ifdef B3_TIER_TRACE
            {
                unsigned long long h[16];
                cudaMemcpyAsync(h, queue + 64, sizeof(h), cudaMemcpyDeviceToHost, st);
                cudaStreamSynchronize(st);
                fprintf(stderr, "[refill trace] first lane out of work %.3f ms, last lane out of work %.3f ms, first warp exit %.3f, last warp exit %.3f ms\n",
                        (h[1] - h[0]) * 1e-6, (h[2] - h[0]) * 1e-6, (h[5] - h[0]) * 1e-6, (h[3] - h[0]) * 1e-6);
            }
#endif
        }
    }
    if (run_warp) {
        gjk_contact_kernel<true><<<warp_grid(n), TPB, WARP_CONTACT_SMEM, st>>>(S, d_pairs, n, order, ranges, cfg,
                                                                                                     strategy, ovf, d_contacts, d_status);
        ++launched;
    }
    if (strategy == 0) {
        // a polling consumer must never be launched behind a producer that failed to launch: it would wait forever
        const bool producer_ok = cudaPeekAtLastError() == cudaSuccess;
        if (overlap && producer_ok) {
            // launched AFTER the producer (a tool that serialises kernels then still runs them in a terminating order)
            cudaStreamWaitEvent(side->stream, side->fork, 0);
            gjk_contact_overflow_kernel<<<OVF_WARPS, 32, EPA_BIG_SMEM_BYTES, side->stream>>>(S, d_pairs, cfg, ovf, 1, d_contacts, d_status);
            if (d_huge_slab) { gjk_contact_huge_kernel<<<HUGE_BLOCKS, HUGE_THREADS, HUGE_SMEM_BYTES, side->stream>>>(S, d_pairs, cfg, ovf, reinterpret_cast<unsigned char*>(d_huge_slab), d_contacts, d_status); ++launched; }
            cudaEventRecord(side->join, side->stream);
            cudaStreamWaitEvent(st, side->join, 0);
        } else {
            gjk_contact_overflow_kernel<<<OVF_WARPS, 32, EPA_BIG_SMEM_BYTES, st>>>(S, d_pairs, cfg, ovf, 0, d_contacts, d_status);
            if (d_huge_slab) { gjk_contact_huge_kernel<<<HUGE_BLOCKS, HUGE_THREADS, HUGE_SMEM_BYTES, st>>>(S, d_pairs, cfg, ovf, reinterpret_cast<unsigned char*>(d_huge_slab), d_contacts, d_status); ++launched; }
        }
        ++launched;
    }
    return launched;
}
int launch_distance(cudaStream_t st, const ShapeSetDev& S, const uint2* d_pairs, uint32_t n, const BinScratch* B, bool run_thread,
                    bool run_warp, const SettingsDev& cfg, void* d_cursor, int refill_min, float* d_ref, float* d_geo, uint8_t* d_status) {
    if (run_thread && d_cursor && n) {
        const uint32_t* order = B ? B->order : nullptr;
        const uint32_t* ranges = B ? B->ranges : nullptr;
        int launched = 1;
        cudaMemsetAsync(d_cursor, 0, 4, st);
        static int grid_cache[MAX_DEVICES];
        const int grid = persistent_grid_cached(gjk_distance_refill_kernel, grid_cache);
        gjk_distance_refill_kernel<<<grid, TPB, 0, st>>>(S, d_pairs, n, order, ranges, cfg, reinterpret_cast<uint32_t*>(d_cursor),
                                                         refill_min < 1 ? 1 : (refill_min > 32 ? 32 : refill_min), d_ref, d_geo, d_status);
        if (run_warp) { gjk_distance_kernel<true><<<warp_grid(n), TPB, 0, st>>>(S, d_pairs, n, order, ranges, cfg, d_ref, d_geo, d_status); ++launched; }
        return launched;
    }
    B3_LAUNCH(gjk_distance_kernel, 0, S, d_pairs, n, order, ranges, cfg, d_ref, d_geo, d_status)
}
// d_cursor: 4 bytes of device scratch for the lane-refill form of the thread path ("toi_refill"); null = one pair per thread.
int launch_toi(cudaStream_t st, const ShapeSetDev& S0, const ShapeSetDev& S1, const uint2* d_pairs, uint32_t n, const BinScratch* B,
               bool run_thread, bool run_warp, const SettingsDev& cfg, void* d_cursor, int refill_min, float4* d_contacts, uint8_t* d_status) {
    if (run_thread && d_cursor && n) {
        const uint32_t* order = B ? B->order : nullptr;
        const uint32_t* ranges = B ? B->ranges : nullptr;
        int launched = 1;
        cudaMemsetAsync(d_cursor, 0, 4, st);
        static int grid_cache[MAX_DEVICES];
        const int grid = persistent_grid_cached(gjk_toi_refill_kernel, grid_cache);
        gjk_toi_refill_kernel<<<grid, TPB, 0, st>>>(S0, S1, d_pairs, n, order, ranges, cfg, reinterpret_cast<uint32_t*>(d_cursor),
                                                                 refill_min < 1 ? 1 : (refill_min > 32 ? 32 : refill_min), d_contacts, d_status);
        if (run_warp) { gjk_toi_kernel<true><<<warp_grid(n), TPB, 0, st>>>(S0, S1, d_pairs, n, order, ranges, cfg, d_contacts, d_status); ++launched; }
        return launched;
    }
    B3_LAUNCH(gjk_toi_kernel, 0, S0, S1, d_pairs, n, order, ranges, cfg, d_contacts, d_status)
}

// ------------------------------------------------------------------------------------------------------------
// Ray casts against primitives (raycast.cuh): one thread per (ray, shape) cast
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TPB) ray_cast_kernel(ShapeSetDev S, const float* __restrict__ rays, const uint2* __restrict__ casts,
                                                       uint64_t n, int query, float* __restrict__ points, uint8_t* __restrict__ status) {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint2 c = __ldg(casts + i);
        const float* r = rays + 6 * (size_t)c.x;
        Shape s;
        load_pair_shape(s, S, c.y);
        V3 p;
        const uint8_t st = ray_cast_one(s, v3(__ldg(r), __ldg(r + 1), __ldg(r + 2)), v3(__ldg(r + 3), __ldg(r + 4), __ldg(r + 5)), query, p);
        status[i] = st;
        points[3 * i] = p.x; points[3 * i + 1] = p.y; points[3 * i + 2] = p.z;
    }
}

int launch_ray_cast(cudaStream_t st, const ShapeSetDev& S, const float* d_rays, const uint2* d_casts, uint64_t n, int query, float* d_points,
                    uint8_t* d_status) {
    if (n == 0) return 0;
    ray_cast_kernel<<<grid_for(n, TPB, NUM_SMS * 32), TPB, 0, st>>>(S, d_rays, d_casts, n, query, d_points, d_status);
    return 1;
}

int launch_compact_contacts(cudaStream_t st, const float4* d_contacts, const uint8_t* d_status, uint64_t n, uint64_t base, void* d_out,
                            uint64_t capacity, unsigned long long* d_count) {
    cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st);
    if (n == 0) return 0;
    compact_contacts_kernel<<<grid_for(n, TPB, NUM_SMS * 16), TPB, 0, st>>>(d_contacts, d_status, n, base,
                                                                              reinterpret_cast<uint32_t*>(d_out), capacity, d_count);
    return 1;
}

}  // namespace b3
