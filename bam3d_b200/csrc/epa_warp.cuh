// Warp-cooperative EPA for the warp-per-pair kernels (pairs involving a ConvexPolyhedron): the polytope lives in a
// per-warp shared-memory slice and the 32 lanes share the O(faces) work instead of repeating it —
//   * closest face: lanes stride the faces, shuffle arg-min (lower distance, then lower index = first minimum);
//   * support point: lanes stride the polyhedron vertices (support.cuh);
//   * visibility: 32 faces per step, one ballot word each;
//   * the order-dependent swap_remove walk is executed identically by all lanes on the ballot words (registers /
//     shared memory); each horizon edge operation searches the ordered edge list 32 entries per ballot and removes
//     with an order-preserving parallel shift — the same ordered lists as epa/epa3d.rs:121-149,183-190;
//   * the new faces (normalize + dot each) are built one per lane.
// Same results as epa.cuh's epa_process (the thread-per-pair tier) by construction; capacities are identical.
#pragma once
#include "epa.cuh"

namespace b3 {

struct WarpPolytopeSmem {
    float vv[EPA_VCAP * 3];
    float va[EPA_VCAP * 3];
    float4 fn[EPA_FCAP];
    uint32_t fi[EPA_FCAP];
    uint16_t ed[EPA_ECAP];
    uint32_t vm[EPA_FCAP / 32];
};

B3_DEV void warp_make_face(WarpPolytopeSmem& W, int slot, uint32_t a, uint32_t b, uint32_t c) {
    V3 va = ld3(W.vv, a);
    V3 ab = ld3(W.vv, b) - va;
    V3 ac = ld3(W.vv, c) - va;
    V3 n = normalize(cross(ab, ac));
    W.fn[slot] = make_float4(n.x, n.y, n.z, dot(n, va));
    W.fi[slot] = a | (b << 8) | (c << 16);
}

// remove_or_add_edge (epa3d.rs:183-190), all lanes; returns false when the edge list is full.
B3_DEV bool warp_edge_op(uint16_t* ed, int& ne, uint32_t a, uint32_t b, uint32_t lane) {
    const uint16_t rev = (uint16_t)(b | (a << 8));
    int found = -1;
    for (int base = 0; base < ne; base += 32) {
        const int idx = base + lane;
        const uint32_t m = __ballot_sync(0xFFFFFFFFu, idx < ne && ed[idx] == rev);
        if (m) { found = base + __ffs(m) - 1; break; }
    }
    if (found >= 0) {  // Vec::remove(found): shift the tail down by one, 32 entries per step
        for (int base = found; base < ne - 1; base += 32) {
            const int idx = base + lane;
            uint16_t nxt = (idx + 1 < ne) ? ed[idx + 1] : (uint16_t)0;
            __syncwarp();
            if (idx < ne - 1) ed[idx] = nxt;
            __syncwarp();
        }
        --ne;
        return true;
    }
    if (ne >= EPA_ECAP) return false;
    if (lane == 0) ed[ne] = (uint16_t)(a | (b << 8));
    ++ne;
    __syncwarp();
    return true;
}

// EPA3::process (epa3d.rs:16-52); every lane of the warp calls it with identical arguments.
B3_DEV uint8_t epa_process_warp(const Shape& L, const Shape& R, const Simplex<true>& s, float tolerance, uint32_t max_iterations,
                                WarpPolytopeSmem& W, ContactOut& out) {
    const uint32_t lane = threadIdx.x & 31u;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { st3(W.vv, k, s.v[k]); st3(W.va, k, s.a[k]); }
    }
    __syncwarp();
    if (lane == 0) warp_make_face(W, 0, 3, 2, 1);
    if (lane == 1) warp_make_face(W, 1, 3, 1, 0);
    if (lane == 2) warp_make_face(W, 2, 3, 0, 2);
    if (lane == 3) warp_make_face(W, 3, 2, 0, 1);
    __syncwarp();
    int nv = 4, nf = 4;
    uint32_t it = 1;
    for (;;) {
        // closest_face_to_origin (epa3d.rs:111-119). Sequentially: best = 0; for k >= 1: if d[k] < d[best] best = k.
        // If d[0] is NaN nothing compares below it; otherwise: lowest index among the minimum over non-NaN distances.
        const float d0 = W.fn[0].w;
        int best = 0;
        if (d0 == d0) {
            float lb = INFINITY; int li = 0x7FFFFFFF;
            for (int k = lane; k < nf; k += 32) {
                float dk = W.fn[k].w;
                if (dk < lb) { lb = dk; li = k; }
            }
            // minimum by (distance, index) in two integer warp reductions: an order-preserving map of f32 onto u32
            // (-0 folded onto +0, as == would; lb is never NaN), then the lowest index among the lanes that hold it
            const uint32_t bits = __float_as_uint(lb + 0.0f);
            const uint32_t key = bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u);
            const uint32_t kmin = __reduce_min_sync(0xFFFFFFFFu, key);
            li = (int)__reduce_min_sync(0xFFFFFFFFu, key == kmin ? (uint32_t)li : 0x7FFFFFFFu);
            if (li != 0x7FFFFFFF) best = li;
        }
        const float4 bf = W.fn[best];
        const V3 normal = v3(bf.x, bf.y, bf.z);
        V3 pv, pa;
        minkowski_support<true>(L, R, normal, pv, pa);
        const float d = dot(pv, normal);
        if (d - bf.w < tolerance || it >= max_iterations) {
            const uint32_t idx = W.fi[best];
            const uint32_t i0 = idx & 0xFF, i1 = (idx >> 8) & 0xFF, i2 = (idx >> 16) & 0xFF;
            float u, v, w;
            barycentric_vector(normal * bf.w, ld3(W.vv, i0), ld3(W.vv, i1), ld3(W.vv, i2), u, v, w);
            out.normal = normal;
            out.depth = bf.w;
            out.point = ld3(W.va, i0) * u + ld3(W.va, i1) * v + ld3(W.va, i2) * w;
            return ST_OK;
        }
        // Polytope::add (epa3d.rs:121-149). Pass A: visibility, one ballot word per 32 faces.
        const int nwords = (nf + 31) >> 5;
        for (int c = 0; c < nwords; ++c) {
            const int k = (c << 5) + lane;
            bool vis = false;
            if (k < nf) {
                const float4 f = W.fn[k];
                const V3 v0 = ld3(W.vv, W.fi[k] & 0xFF);
                vis = dot(v3(f.x, f.y, f.z), pv - v0) > 0.f;
            }
            const uint32_t wmask = __ballot_sync(0xFFFFFFFFu, vis);
            if (lane == 0) W.vm[c] = wmask;
        }
        __syncwarp();
        // Pass B: the reference's `while i < len` walk with swap_remove, all lanes in step. The horizon edge list lives
        // in registers while it is short — lane i holds entry i: the search for a reversed duplicate is one ballot,
        // Vec::remove one shuffle-down, push one predicated move — and moves to the shared-memory list (warp_edge_op)
        // the moment it would need a 33rd entry.
        int ne = 0;
        bool ok = true;
        bool spilled = false;
        uint32_t my_edge = 0;
        for (int k = 0;;) {
            int c = k >> 5;
            uint32_t w = (c < nwords) ? (W.vm[c] >> (k & 31)) : 0u;
            while (w == 0u && ++c < nwords) { k = c << 5; w = W.vm[c]; }
            if (w == 0u) break;
            k += __ffs(w) - 1;
            const uint32_t idx = W.fi[k];
            --nf;
            const uint32_t last_bit = (W.vm[nf >> 5] >> (nf & 31)) & 1u;
            const float4 lf = W.fn[nf];
            const uint32_t li = W.fi[nf];
            const uint32_t wk = W.vm[k >> 5], wl = W.vm[nf >> 5];
            __syncwarp();
            if (lane == 0) {
                // clear the last slot's bit, then move the last face (with its flag) into slot k
                uint32_t nl = wl & ~(1u << (nf & 31));
                if ((k >> 5) == (nf >> 5)) {
                    if (k != nf) nl = (nl & ~(1u << (k & 31))) | (last_bit << (k & 31));
                    W.vm[nf >> 5] = nl;
                } else {
                    W.vm[nf >> 5] = nl;
                    W.vm[k >> 5] = (wk & ~(1u << (k & 31))) | (last_bit << (k & 31));
                }
                if (k != nf) { W.fn[k] = lf; W.fi[k] = li; }
            }
            __syncwarp();
            const uint32_t vi[4] = {idx & 0xFF, (idx >> 8) & 0xFF, (idx >> 16) & 0xFF, idx & 0xFF};
            // not unrolled on purpose: the warp kernels are instruction-fetch bound (see support.cuh: support_point_outlined)
#pragma unroll 1
            for (int e = 0; e < 3; ++e) {
                const uint32_t a = vi[e], b = vi[e + 1];
                if (!spilled) {
                    const uint32_t rev = b | (a << 8);
                    const uint32_t m = __ballot_sync(0xFFFFFFFFu, (int)lane < ne && my_edge == rev);
                    if (m) {
                        const uint32_t found = __ffs(m) - 1;
                        const uint32_t nxt = __shfl_down_sync(0xFFFFFFFFu, my_edge, 1);
                        if (lane >= found) my_edge = nxt;
                        --ne;
                    } else if (ne < 32) {
                        if ((int)lane == ne) my_edge = a | (b << 8);
                        ++ne;
                    } else {
                        W.ed[lane] = (uint16_t)my_edge;
                        __syncwarp();
                        spilled = true;
                    }
                }
                if (spilled) ok = ok && warp_edge_op(W.ed, ne, a, b, lane);
            }
        }
        if (!ok || nv >= EPA_VCAP || nf + ne > EPA_FCAP) return ST_CAPACITY;
        const int n = nv;
        if (lane == 0) { st3(W.vv, n, pv); st3(W.va, n, pa); }
        ++nv;
        __syncwarp();
        // Pass C: the new faces, one per lane
        if (!spilled) {
            if ((int)lane < ne) warp_make_face(W, nf + lane, (uint32_t)n, my_edge & 0xFF, (my_edge >> 8) & 0xFF);
        } else {
            for (int e = lane; e < ne; e += 32) {
                const uint32_t ab = W.ed[e];
                warp_make_face(W, nf + e, (uint32_t)n, ab & 0xFF, (ab >> 8) & 0xFF);
            }
        }
        nf += ne;
        __syncwarp();
        it += 1;
    }
}

}  // namespace b3
