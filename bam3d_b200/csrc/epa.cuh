// EPA3 on the device. Reference: epa/epa3d.rs:16-190 (EPA3::process, Polytope, Face, remove_or_add_edge) and
// contact()/point() (:70-94).
//
// The reference's tie-breaks are order-dependent — `closest_face_to_origin` keeps the FIRST minimum, `add`
// walks the face list while `swap_remove`-ing, and the horizon is an ordered edge list with order-preserving
// removal — and they decide which of several coplanar faces supplies the contact point. The device polytope
// therefore keeps the same three ordered lists (vertices, faces, horizon edges) and mutates them in the same
// order. Storage is supplied by the caller: per-thread local arrays for the thread-per-pair kernels, a per-warp
// shared-memory slice for the warp-cooperative polyhedron kernels (all lanes then execute this code
// redundantly on identical data and store identical values).
#pragma once
#include "simplex.cuh"
#include "support.cuh"

namespace b3 {

// Capacities. EPA's loop counter starts at 1 and stops at max_iterations (<= 100, enforced by the host layer), so
// at most 99 vertices are added to the 4 of the GJK tetrahedron: VCAP >= 103 always suffices. A closed
// triangulated polytope has F = 2V - 4 <= 202 faces, but near-degenerate input (curved shapes, NaN normals)
// makes the reference build pinched polytopes with many more (63 of 2^20 mixed pairs exceed 256 faces, 44 exceed
// 1024, and those keep growing multiplicatively). Two tiers therefore: the fast tier (per-thread local memory /
// per-warp shared memory) holds EPA_FCAP faces; a pair that outgrows it gets ST_CAPACITY and is redone by the
// warp-per-pair overflow tier (epa_big.cuh) with EPA_BIG_FCAP faces. Only a pair that outgrows that keeps
// ST_CAPACITY (the CPU oracle applies the same limit, oracle/bam3d_oracle.hpp EPA_FACE_LIMIT).
constexpr int EPA_VCAP = 104;
constexpr int EPA_FCAP = 256;
constexpr int EPA_ECAP = 320;

struct PolytopeStore {
    float* vv;        // EPA_VCAP * 3 : SupportPoint.v
    float* va;        // EPA_VCAP * 3 : SupportPoint.sup_a
    float4* fn;       // EPA_FCAP     : face normal xyz + distance
    uint32_t* fi;     // EPA_FCAP     : face vertex indices, 8 bits each (a | b<<8 | c<<16)
    uint16_t* ed;     // ecap entries, `estride` apart : horizon edges (a | b<<8)
    int fcap, ecap;   // capacities of fn/fi and ed
    int estride;      // 1 = contiguous (local memory); blockDim.x = one shared-memory column per thread
};

struct ContactOut { V3 normal; float depth; V3 point; };

B3_DEV V3 ld3(const float* p, int i) { return v3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
B3_DEV void st3(float* p, int i, V3 v) { p[3 * i] = v.x; p[3 * i + 1] = v.y; p[3 * i + 2] = v.z; }

// Face::new_impl (epa3d.rs:160-170)
B3_DEV float make_face(const PolytopeStore& P, int slot, uint32_t a, uint32_t b, uint32_t c) {
    V3 va = ld3(P.vv, a);
    V3 ab = ld3(P.vv, b) - va;
    V3 ac = ld3(P.vv, c) - va;
    V3 n = normalize(cross(ab, ac));
    float dist = dot(n, va);
    P.fn[slot] = make_float4(n.x, n.y, n.z, dist);
    P.fi[slot] = a | (b << 8) | (c << 16);
    return dist;
}

// closest_face_to_origin (epa3d.rs:111-119) is `best = 0; for k >= 1: if d[k] < d[best] best = k`: the FIRST index of
// the minimum over non-NaN distances, except that a NaN d[0] is never displaced. The thread-per-pair tier does not
// rescan the face list for it every iteration: the running (distance, index) is carried through Polytope::add —
// survivors in pass A (strict <, increasing index), faces moved by swap_remove (an equal distance at a lower index
// takes over), new faces in pass C — which yields exactly the index a fresh scan of the final list would.
struct ClosestFace {
    float d;
    int idx;
    B3_DEV void reset() { d = INFINITY; idx = -1; }
    B3_DEV void offer(float fd, int k) { if (fd < d) { d = fd; idx = k; } }                       // in increasing k
    B3_DEV void offer_any(float fd, int k) { if (fd < d || (fd == d && k < idx)) { d = fd; idx = k; } }  // in any order
    B3_DEV void moved(float fd, int k) { if (fd == d && k < idx) idx = k; }                        // survivor moved to a lower slot
    B3_DEV int finish(const PolytopeStore& P) const { float d0 = P.fn[0].w; return (d0 != d0 || idx < 0) ? 0 : idx; }
};

// remove_or_add_edge (epa3d.rs:183-190): a reversed duplicate is removed (order preserving), else push.
// `seen` is a 64-bit Bloom filter over the directed edges pushed during this add(): a reversed duplicate can only be in the
// list if its bit is set, so most pushes (two thirds of all calls on a well-formed polytope) skip the search loop
// (C2: -3 %). (Tombstones instead of the order-preserving shift on removal were measured too: no gain.)
B3_DEV uint32_t edge_bit(uint32_t a, uint32_t b) { return (a * 7u + b * 13u) & 63u; }
B3_DEV bool remove_or_add_edge(uint16_t* ed, int es, int& ne, int ecap, uint32_t a, uint32_t b, unsigned long long& seen) {
    const uint16_t rev = (uint16_t)(b | (a << 8));
    const bool maybe = (seen >> edge_bit(b, a)) & 1ull;
    for (int i = 0; maybe && i < ne; ++i) {
        if (ed[i * es] == rev) {
            for (int k = i; k + 1 < ne; ++k) ed[k * es] = ed[(k + 1) * es];
            --ne;
            return true;
        }
    }
    if (ne >= ecap) return false;
    ed[ne * es] = (uint16_t)(a | (b << 8));
    ++ne;
    seen |= 1ull << edge_bit(a, b);
    return true;
}

// Polytope::new (epa3d.rs:103-109) + Face::new (:172-179) from the 4-point GJK simplex (with sup_a).
template <bool WARP>
B3_DEV void epa_init(const Simplex<true>& s, const PolytopeStore& P, int& nv, int& nf, uint32_t& it, V3& vmax, int& best) {
    vmax = v3(0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        st3(P.vv, k, s.v[k]); st3(P.va, k, s.a[k]);
        vmax = v3(fmaxf(vmax.x, fabsf(s.v[k].x)), fmaxf(vmax.y, fabsf(s.v[k].y)), fmaxf(vmax.z, fabsf(s.v[k].z)));
    }
    if (WARP) __syncwarp();
    ClosestFace cf;
    cf.reset();
    cf.offer(make_face(P, 0, 3, 2, 1), 0);
    cf.offer(make_face(P, 1, 3, 1, 0), 1);
    cf.offer(make_face(P, 2, 3, 0, 2), 2);
    cf.offer(make_face(P, 3, 2, 0, 1), 3);
    if (WARP) __syncwarp();
    best = cf.finish(P);
    nv = 4; nf = 4; it = 1;
}

constexpr uint8_t EPA_CONTINUE = 0xFF;

#ifndef B3_EPA_PREFETCH
#define B3_EPA_PREFETCH 0  // 0 = off (measured: L1/L2 prefetch of the face list costs 0.8 ms on C2 — the caches are over capacity), 1 = prefetch.L1, 2 = prefetch.L2
#endif

// Thread-per-pair tier: pull the whole face list towards the SM while the support function computes, so that pass A
// pays one memory round trip instead of one per 8-face batch. Per-thread local memory is interleaved by lane at word
// granularity — word w of all 32 lanes shares one 128-byte line — so the active lanes split the lines between
// them: the lane of rank r touches words r, r + count, ... of its own face array.
B3_DEV void prefetch_faces(const PolytopeStore& P, int nf) {
    const uint32_t act = __activemask();
    const int cnt = __popc(act);
#ifdef B3_EPA_PREFETCH_MAX_LANES
    // only when the warp is nearly empty (the kernel's tail: the queue has run dry and a few long pairs are left): then
    // L1 is not under pressure and a trip's cost is its chain of dependent memory round trips
    if (cnt > B3_EPA_PREFETCH_MAX_LANES) return;
#endif
    const int rank = __popc(act & ((1u << (threadIdx.x & 31u)) - 1u));
    const int words = 4 * (int)__reduce_max_sync(act, (unsigned)nf);
    const float* base = reinterpret_cast<const float*>(P.fn);
    for (int w = rank; w < words; w += cnt) {
#if B3_EPA_PREFETCH == 2
        asm volatile("prefetch.L2 [%0];" ::"l"(base + w));
#else
        asm volatile("prefetch.L1 [%0];" ::"l"(base + w));
#endif
    }
}

// One trip of EPA3::process's loop (epa3d.rs:33-51): closest face, support, convergence test, Polytope::add.
// Returns EPA_CONTINUE, or a final Status (ST_OK with `out` filled, ST_CAPACITY).
template <bool WARP>
B3_DEV uint8_t epa_iterate(const Shape& L, const Shape& R, float tolerance, uint32_t max_iterations, const PolytopeStore& P,
                           int& nv, int& nf, uint32_t& it, V3& vmax, int& best, ContactOut& out) {
    float4 bf = P.fn[best];
    V3 normal = v3(bf.x, bf.y, bf.z);
    V3 pv, pa;
#if B3_EPA_PREFETCH
    if (!WARP) prefetch_faces(P, nf);
#endif
    minkowski_support<WARP>(L, R, normal, pv, pa);
    float d = dot(pv, normal);
    if (d - bf.w < tolerance || it >= max_iterations) {
        // contact() / point() (epa3d.rs:70-94)
        uint32_t idx = P.fi[best];
        uint32_t i0 = idx & 0xFF, i1 = (idx >> 8) & 0xFF, i2 = (idx >> 16) & 0xFF;
        float u, v, w;
        barycentric_vector(normal * bf.w, ld3(P.vv, i0), ld3(P.vv, i1), ld3(P.vv, i2), u, v, w);
        out.normal = normal;
        out.depth = bf.w;
        out.point = ld3(P.va, i0) * u + ld3(P.va, i1) * v + ld3(P.va, i2) * w;
        return ST_OK;
    }

    // Polytope::add (epa3d.rs:121-149), in three passes so that the lanes of a thread-per-pair warp stay together:
    //  A. visibility flag of every face (uniform loop body);
    //  B. the order-dependent walk: visit the flagged faces exactly as the reference's `while i < len` loop with
    //     swap_remove does (the face moved into slot k is examined next), feeding the horizon edge list;
    //  C. the new faces.
    // Pass A needs dot(n, p - v0) > 0 with v0 = the face's first vertex (a divergent gather). face.distance is
    // fl(dot(n, v0)) (Face::new_impl), so a2 = fl(dot(n, p)) - distance equals the reference's value up to
    // 8.02 u S, S = sum |n_i| (|p_i| + |v0_i|), u = 2^-24 (standard dot-product error bounds). Whenever
    // |a2| > 2^-20 S (S bounded with the running max |vertex coordinate| and |n_i| <= 1) the sign is decided without the gather;
    // otherwise — and for any NaN/inf — the reference's expression is evaluated as written. Exact either way.
    // The flags live in a bitmask (32 faces per word, built in a register chunk by chunk) so that pass A is a pure
    // load + arithmetic loop the compiler can software-pipeline, and pass B finds the next flagged face with ffs.
    uint32_t vm[EPA_FCAP / 32];
    ClosestFace cf;
    cf.reset();
    // |n_i| <= 1 (+ an ulp) for a normalised face normal, so S <= sum_i (|p_i| + max|v_i|): one face-independent bound per
    // iteration instead of six more operations per face (a NaN / inf normal makes a2 NaN and lands in the exact path)
    const float sure_bound = ((fabsf(pv.x) + vmax.x) + (fabsf(pv.y) + vmax.y) + (fabsf(pv.z) + vmax.z)) * 9.5367431640625e-07f + 1e-30f;
    const int nchunks = (nf + 31) >> 5;
    for (int c = 0; c < nchunks; ++c) {
        uint32_t m = 0, unsure = 0;
        const int k0 = c << 5;
        const int k1 = min(nf, k0 + 32);
        for (int kb = k0; kb < k1; kb += 8) {
            float4 fb[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) fb[j] = (kb + j < k1) ? P.fn[kb + j] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = kb + j;
                if (k < k1) {
                    const float4 f = fb[j];
                    float a2 = dot(v3(f.x, f.y, f.z), pv) - f.w;
                    const bool sure = fabsf(a2) > sure_bound;  // false for any NaN/inf too
                    const bool vis = a2 > 0.f;
                    m |= ((sure && vis) ? 1u : 0u) << (k - k0);
                    unsure |= (sure ? 0u : 1u) << (k - k0);
                    if (sure && !vis) cf.offer(f.w, k);
                }
            }
        }
        // the few faces the filter could not decide: the reference's expression as written (kept out of the loop
        // above so that its gather does not cost issue slots there)
        while (unsure) {
            const int b = __ffs(unsure) - 1;
            unsure &= unsure - 1;
            const int k = k0 + b;
            const float4 f = P.fn[k];
            V3 v0 = ld3(P.vv, P.fi[k] & 0xFF);
            if (dot(v3(f.x, f.y, f.z), pv - v0) > 0.f) m |= 1u << b;
            else cf.offer_any(f.w, k);
        }
        vm[c] = m;
    }
    int ne = 0;
    bool ok = true;
    unsigned long long seen = 0ull;
    for (int k = 0;;) {
        // next flagged slot >= k (bits at or beyond nf are always clear)
        int c = k >> 5;
        uint32_t w = (c < nchunks) ? (vm[c] >> (k & 31)) : 0u;
        while (w == 0u && ++c < nchunks) { k = c << 5; w = vm[c]; }
        if (w == 0u) break;
        k += __ffs(w) - 1;
        const uint32_t idx = P.fi[k];
        --nf;                             // swap_remove(k): the last face (with its flag) moves into slot k
        const uint32_t last_bit = (vm[nf >> 5] >> (nf & 31)) & 1u;
        vm[nf >> 5] &= ~(1u << (nf & 31));
        if (k != nf) {
            float4 lf = P.fn[nf];
            uint32_t li = P.fi[nf];
            if (WARP) __syncwarp();
            P.fn[k] = lf;
            P.fi[k] = li;
            vm[k >> 5] = (vm[k >> 5] & ~(1u << (k & 31))) | (last_bit << (k & 31));
            if (!last_bit) cf.moved(lf.w, k);
        }
        uint32_t i0 = idx & 0xFF, i1 = (idx >> 8) & 0xFF, i2 = (idx >> 16) & 0xFF;
        ok = ok && remove_or_add_edge(P.ed, P.estride, ne, P.ecap, i0, i1, seen);
        ok = ok && remove_or_add_edge(P.ed, P.estride, ne, P.ecap, i1, i2, seen);
        ok = ok && remove_or_add_edge(P.ed, P.estride, ne, P.ecap, i2, i0, seen);
        if (WARP) __syncwarp();
    }
    if (!ok || nv >= EPA_VCAP || nf + ne > P.fcap) return ST_CAPACITY;
    const int n = nv;
    st3(P.vv, n, pv);
    st3(P.va, n, pa);
    // running max |coordinate| over the polytope's vertices; a NaN coordinate poisons it so the filter never fires
    vmax = v3(pv.x != pv.x ? pv.x : fmaxf(vmax.x, fabsf(pv.x)), pv.y != pv.y ? pv.y : fmaxf(vmax.y, fabsf(pv.y)),
              pv.z != pv.z ? pv.z : fmaxf(vmax.z, fabsf(pv.z)));
    ++nv;
    if (WARP) __syncwarp();
    for (int e = 0; e < ne; ++e) {
        uint32_t ab = P.ed[e * P.estride];
        cf.offer(make_face(P, nf + e, (uint32_t)n, ab & 0xFF, (ab >> 8) & 0xFF), nf + e);
    }
    nf += ne;
    if (WARP) __syncwarp();
    best = cf.finish(P);
    it += 1;
    return EPA_CONTINUE;
}

// EPA3::process (epa3d.rs:16-52). `s` is the 4-point GJK simplex (with sup_a). Returns a Status.
template <bool WARP>
B3_DEV uint8_t epa_process(const Shape& L, const Shape& R, const Simplex<true>& s, float tolerance, uint32_t max_iterations,
                           const PolytopeStore& P, ContactOut& out) {
    int nv, nf;
    uint32_t it;
    V3 vmax;
    int best;
    epa_init<WARP>(s, P, nv, nf, it, vmax, best);
    for (;;) {
        uint8_t st = epa_iterate<WARP>(L, R, tolerance, max_iterations, P, nv, nf, it, vmax, best, out);
        if (st != EPA_CONTINUE) return st;
    }
}

}  // namespace b3
