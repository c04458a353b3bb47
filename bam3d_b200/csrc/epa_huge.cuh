// Third EPA tier: one BLOCK per pair, polytope and all scratch in a per-block global-memory slab, for the pairs whose polytope
// outgrows the overflow tier (epa_big.cuh, 1024 faces). The reference has no face limit: on near-degenerate input (a Cylinder
// cap against a flat face: coplanar rim supports, visibility is rounding noise) its polytope pinches and the face list grows
// multiplicatively — 44 of 2^20 mixed pairs pass 1024 faces, the largest reaches 121 962 — and it still returns a Contact
// after at most max_iterations trips (tests/test_host_cpu.py runs the oracle without a limit on exactly those pairs). To
// return the same Contact the device has to build the same lists, and a sequential walk over 10^5 faces per trip is out of
// the question on one lane, so Polytope::add (epa3d.rs:121-149) is evaluated here WITHOUT a walk:
//   * visibility flags, then exclusive prefix counts of them (block scan); S = survivors;
//   * `while i < len { if visible(i) { swap_remove(i) } else { i += 1 } }` is a two-pointer process: the holes (visible
//     positions below S, ascending h_1 < h_2 < ...) are filled by the tail survivors (non-visible positions >= S) taken from
//     the back (t_1 > t_2 > ...), so the surviving list is "positions below S, holes replaced by t_k" — one gather;
//   * the faces are REMOVED in the order h_1, [visible positions in (t_1, F) descending], h_2, [visible positions in
//     (t_2, t_1) descending], ..., and at the end, if positions remain between S and the last tail survivor taken, S itself
//     followed by the rest of them descending: every visible face computes its own place in that order from the prefix counts;
//   * the horizon is that sequence of directed edges under remove_or_add_edge (:183-190): per undirected edge a
//     first-in-first-out matching of the two directions. Events are bucketed by directed edge (vertex indices < 104: a
//     16384-entry key table), each undirected edge is matched by one thread over its two short lists, and the surviving
//     events are compacted in sequence order — the reference's edge list, in the reference's order;
//   * new faces one per thread; closest face by a block reduction with the reference's first-minimum rule.
// oracle/bam3d_oracle.hpp: Polytope::add_in_phases is this formulation on the CPU, checked against the sequential add().
#pragma once
#include "epa.cuh"

namespace b3 {

constexpr int HUGE_THREADS = 256;
constexpr uint32_t HUGE_FCAP = 1u << 18;   // faces per polytope (== orc::EPA_FACE_LIMIT); beyond: ST_CAPACITY
constexpr uint32_t HUGE_KEYS = 128 * 128;  // directed-edge keys a | b << 7
constexpr int HUGE_BLOCKS = 32;

struct HugeStore {
    float4* fn;        // FCAP
    uint32_t* fi;      // FCAP
    uint8_t* vis;      // FCAP
    uint32_t* before;  // FCAP + 1 : visible positions below p
    uint32_t* tsrc;    // FCAP + 2 : tsrc[k] = position of the k-th tail survivor from the back, tsrc[0] = F
    uint16_t* evkey;   // 3 FCAP  : directed edge of event e (3 per removed face, in removal order)
    uint8_t* keep;     // 3 FCAP  : event e is in the final horizon list
    uint32_t* sorted;  // 3 FCAP  : event ids bucketed by key
    uint32_t* hpos;    // 3 FCAP + 1 : exclusive prefix of keep
    uint32_t* cnt;     // KEYS
    uint32_t* start;   // KEYS + 1
    uint32_t* fill;    // KEYS
    uint16_t* hor;     // FCAP    : horizon edges in order
};
constexpr size_t HUGE_SLAB_BYTES = (size_t)HUGE_FCAP * (16 + 4 + 1 + 4 + 4 + 6 + 3 + 12 + 12 + 2) + (size_t)HUGE_KEYS * 12 + 4096;

B3_DEV HugeStore carve_huge_store(unsigned char* p) {
    HugeStore H;
    auto take = [&](size_t bytes) { unsigned char* r = p; p += (bytes + 255) & ~(size_t)255; return r; };
    H.fn = reinterpret_cast<float4*>(take((size_t)HUGE_FCAP * 16));
    H.fi = reinterpret_cast<uint32_t*>(take((size_t)HUGE_FCAP * 4));
    H.before = reinterpret_cast<uint32_t*>(take(((size_t)HUGE_FCAP + 1) * 4));
    H.tsrc = reinterpret_cast<uint32_t*>(take(((size_t)HUGE_FCAP + 2) * 4));
    H.sorted = reinterpret_cast<uint32_t*>(take((size_t)HUGE_FCAP * 12));
    H.hpos = reinterpret_cast<uint32_t*>(take(((size_t)HUGE_FCAP * 3 + 1) * 4));
    H.cnt = reinterpret_cast<uint32_t*>(take((size_t)HUGE_KEYS * 4));
    H.start = reinterpret_cast<uint32_t*>(take(((size_t)HUGE_KEYS + 1) * 4));
    H.fill = reinterpret_cast<uint32_t*>(take((size_t)HUGE_KEYS * 4));
    H.evkey = reinterpret_cast<uint16_t*>(take((size_t)HUGE_FCAP * 6));
    H.hor = reinterpret_cast<uint16_t*>(take((size_t)HUGE_FCAP * 2));
    H.vis = take((size_t)HUGE_FCAP);
    H.keep = take((size_t)HUGE_FCAP * 3);
    return H;
}

// Exclusive prefix sums of in[0..n) into out[0..n], out[n] = total; returns the total to every thread. s_warp: 8 words of
// shared memory. Ends with a barrier: out[] is visible to the whole block on return.
template <class T>
B3_DEV uint32_t huge_block_scan(const T* in, uint32_t* out, uint32_t n, uint32_t* s_warp) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n; base += HUGE_THREADS) {
        const uint32_t i = base + tid;
        const uint32_t v = i < n ? (uint32_t)in[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, off);
            if ((int)lane >= off) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        uint32_t wp = 0, tot = 0;
#pragma unroll
        for (int w = 0; w < HUGE_THREADS / 32; ++w) { const uint32_t sw = s_warp[w]; if (w < (int)warp) wp += sw; tot += sw; }
        if (i < n) out[i] = carry + wp + x - v;
        carry += tot;
        __syncthreads();
    }
    if (tid == 0) out[n] = carry;
    __syncthreads();
    return carry;
}

struct HugeShared {
    float vv[EPA_VCAP * 3];
    float va[EPA_VCAP * 3];
    uint32_t warp_scratch[HUGE_THREADS / 32];
    unsigned long long red[HUGE_THREADS / 32];
    V3 pv, pa;
    int best;
    uint32_t ok;
};

// closest_face_to_origin (epa3d.rs:111-119) over faces [0, F): the first index of the minimum distance over non-NaN distances;
// face 0 when its own distance is NaN (nothing ever compares below NaN). Result in sh.best after the closing barrier.
B3_DEV void huge_closest_face(const HugeStore& H, uint32_t F, HugeShared& sh) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    unsigned long long key = ~0ull;
    for (uint32_t k = tid; k < F; k += HUGE_THREADS) {
        const float d = H.fn[k].w;
        if (d == d) {
            const uint32_t bits = __float_as_uint(d + 0.0f);  // -0 folded onto +0, as < would
            const uint32_t ord = bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u);
            const unsigned long long kk = ((unsigned long long)ord << 32) | k;
            key = kk < key ? kk : key;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, key, off);
        key = o < key ? o : key;
    }
    if (lane == 0) sh.red[warp] = key;
    __syncthreads();
    if (tid == 0) {
        unsigned long long m = ~0ull;
        for (int w = 0; w < HUGE_THREADS / 32; ++w) m = sh.red[w] < m ? sh.red[w] : m;
        const float d0 = H.fn[0].w;
        sh.best = (d0 != d0 || m == ~0ull) ? 0 : (int)(uint32_t)(m & 0xFFFFFFFFull);
    }
    __syncthreads();
}

B3_DEV void huge_make_face(const HugeStore& H, const HugeShared& sh, uint32_t slot, uint32_t a, uint32_t b, uint32_t c) {
    const V3 va = ld3(sh.vv, a);
    const V3 ab = ld3(sh.vv, b) - va;
    const V3 ac = ld3(sh.vv, c) - va;
    const V3 n = normalize(cross(ab, ac));
    H.fn[slot] = make_float4(n.x, n.y, n.z, dot(n, va));
    H.fi[slot] = a | (b << 8) | (c << 16);
}

// remove_or_add_edge (epa3d.rs:183-190) for ONE undirected edge {a, b}, a < b, over all its events of this add(): l1 / l2 =
// the ids (sequence numbers) of the events a->b / b->a. Both lists are put in ascending order first (the bucketing does not
// keep it). An event pushes unless a copy of the opposite direction is in the list, in which case it removes the OLDEST such
// copy; pushed events of one direction are a subsequence of that direction's list, so "the oldest live copy" is found by a head
// pointer that only moves forward.
B3_DEV void huge_match_edge(uint32_t* l1, uint32_t n1, uint32_t* l2, uint32_t n2, uint8_t* keep) {
    for (uint32_t i = 1; i < n1; ++i) { const uint32_t v = l1[i]; uint32_t j = i; while (j > 0 && l1[j - 1] > v) { l1[j] = l1[j - 1]; --j; } l1[j] = v; }
    for (uint32_t i = 1; i < n2; ++i) { const uint32_t v = l2[i]; uint32_t j = i; while (j > 0 && l2[j - 1] > v) { l2[j] = l2[j - 1]; --j; } l2[j] = v; }
    uint32_t i1 = 0, i2 = 0, h1 = 0, h2 = 0, live1 = 0, live2 = 0;  // live_k = copies of direction k in the list now
    while (i1 < n1 || i2 < n2) {
        const bool take1 = i2 >= n2 || (i1 < n1 && l1[i1] < l2[i2]);
        if (take1) {
            const uint32_t e = l1[i1++];
            if (live2) { while (!keep[l2[h2]]) ++h2; keep[l2[h2++]] = 0; --live2; }
            else { keep[e] = 1; ++live1; }
        } else {
            const uint32_t e = l2[i2++];
            if (live1) { while (!keep[l1[h1]]) ++h1; keep[l1[h1++]] = 0; --live1; }
            else { keep[e] = 1; ++live2; }
        }
    }
}

// EPA3::process (epa3d.rs:16-52) for one pair, executed by all HUGE_THREADS threads of the block (convergent; every value
// that steers control flow is block-uniform). `s` must be valid in warp 0 (the GJK tetrahedron with sup_a).
B3_DEV uint8_t epa_process_huge(const Shape& L, const Shape& R, const Simplex<true>& s, float tolerance, uint32_t max_iterations, const HugeStore& H,
                                HugeShared& sh, ContactOut& out) {
    const uint32_t tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { st3(sh.vv, k, s.v[k]); st3(sh.va, k, s.a[k]); }
    }
    __syncthreads();
    if (tid == 0) {
        huge_make_face(H, sh, 0, 3, 2, 1);
        huge_make_face(H, sh, 1, 3, 1, 0);
        huge_make_face(H, sh, 2, 3, 0, 2);
        huge_make_face(H, sh, 3, 2, 0, 1);
    }
    __syncthreads();
    uint32_t nv = 4, F = 4, it = 1;
    huge_closest_face(H, F, sh);
    for (;;) {
        const int best = sh.best;
        const float4 bf = H.fn[best];
        const V3 normal = v3(bf.x, bf.y, bf.z);
        if (tid < 32) {  // the support functions are warp-convergent (polyhedron vertex scans): warp 0 computes, the block reads
            V3 pv, pa;
            minkowski_support<true>(L, R, normal, pv, pa);
            if (tid == 0) { sh.pv = pv; sh.pa = pa; }
        }
        __syncthreads();
        const V3 pv = sh.pv;
        const float d = dot(pv, normal);
        if (d - bf.w < tolerance || it >= max_iterations) {
            const uint32_t idx = H.fi[best];
            const uint32_t i0 = idx & 0xFF, i1 = (idx >> 8) & 0xFF, i2 = (idx >> 16) & 0xFF;
            float u, v, w;
            barycentric_vector(normal * bf.w, ld3(sh.vv, i0), ld3(sh.vv, i1), ld3(sh.vv, i2), u, v, w);
            out.normal = normal;
            out.depth = bf.w;
            out.point = ld3(sh.va, i0) * u + ld3(sh.va, i1) * v + ld3(sh.va, i2) * w;
            __syncthreads();
            return ST_OK;
        }
        // visibility (epa3d.rs:126-129), as the reference writes it
        for (uint32_t p = tid; p < F; p += HUGE_THREADS) {
            const float4 f = H.fn[p];
            const V3 v0 = ld3(sh.vv, H.fi[p] & 0xFF);
            H.vis[p] = dot(v3(f.x, f.y, f.z), pv - v0) > 0.f ? 1 : 0;
        }
        __syncthreads();
        const uint32_t V = huge_block_scan(H.vis, H.before, F, sh.warp_scratch);
        const uint32_t S = F - V;
        const uint32_t Hn = H.before[S];  // holes below S == tail survivors
        auto vis_after = [&](uint32_t p) { return V - H.before[p] - (uint32_t)H.vis[p]; };      // visible positions above p
        auto surv_after = [&](uint32_t p) { return (F - 1 - p) - vis_after(p); };               // survivors above p
        for (uint32_t p = S + tid; p < F; p += HUGE_THREADS)
            if (!H.vis[p]) H.tsrc[surv_after(p) + 1] = p;
        if (tid == 0) H.tsrc[0] = F;
        __syncthreads();
        const uint32_t t_last = Hn ? H.tsrc[Hn] : F;
        // the three directed edges of every removed face, at its place in the removal order
        for (uint32_t p = tid; p < F; p += HUGE_THREADS) {
            if (!H.vis[p]) continue;
            uint32_t idx;
            if (p < S) {
                const uint32_t k = H.before[p] + 1;                       // the k-th hole
                idx = (k - 1) + (k >= 2 ? vis_after(H.tsrc[k - 1]) : 0u);
            } else {
                const uint32_t sv = surv_after(p);
                if (sv < Hn) idx = (sv + 1) + vis_after(p);               // pulled in while hole sv + 1 was being filled
                else if (p == S) idx = Hn + (Hn ? vis_after(t_last) : 0u);
                else idx = Hn + 1 + vis_after(p);
            }
            const uint32_t f = H.fi[p];
            const uint32_t a = f & 0xFF, b = (f >> 8) & 0xFF, c = (f >> 16) & 0xFF;
            H.evkey[3 * (size_t)idx] = (uint16_t)(a | (b << 7));
            H.evkey[3 * (size_t)idx + 1] = (uint16_t)(b | (c << 7));
            H.evkey[3 * (size_t)idx + 2] = (uint16_t)(c | (a << 7));
        }
        for (uint32_t k = tid; k < HUGE_KEYS; k += HUGE_THREADS) { H.cnt[k] = 0; H.fill[k] = 0; }
        __syncthreads();
        const uint32_t E = 3 * V;
        for (uint32_t e = tid; e < E; e += HUGE_THREADS) atomicAdd(&H.cnt[H.evkey[e]], 1u);
        __syncthreads();
        huge_block_scan(H.cnt, H.start, HUGE_KEYS, sh.warp_scratch);
        for (uint32_t e = tid; e < E; e += HUGE_THREADS) {
            const uint32_t key = H.evkey[e];
            H.sorted[H.start[key] + atomicAdd(&H.fill[key], 1u)] = e;
            H.keep[e] = 0;
        }
        __syncthreads();
        for (uint32_t u = tid; u < HUGE_KEYS; u += HUGE_THREADS) {
            const uint32_t a = u & 127u, b = u >> 7;
            if (a >= b) continue;
            const uint32_t k1 = a | (b << 7), k2 = b | (a << 7);
            const uint32_t n1 = H.cnt[k1], n2 = H.cnt[k2];
            if (n1 + n2) huge_match_edge(H.sorted + H.start[k1], n1, H.sorted + H.start[k2], n2, H.keep);
        }
        __syncthreads();
        const uint32_t ne = huge_block_scan(H.keep, H.hpos, E, sh.warp_scratch);
        if (nv >= (uint32_t)EPA_VCAP || (unsigned long long)S + ne > HUGE_FCAP) { __syncthreads(); return ST_CAPACITY; }
        for (uint32_t e = tid; e < E; e += HUGE_THREADS)
            if (H.keep[e]) H.hor[H.hpos[e]] = H.evkey[e];
        // the surviving list: holes below S take the tail survivors from the back (sources >= S, destinations < S)
        for (uint32_t p = tid; p < S; p += HUGE_THREADS) {
            if (H.vis[p]) {
                const uint32_t src = H.tsrc[H.before[p] + 1];
                H.fn[p] = H.fn[src];
                H.fi[p] = H.fi[src];
            }
        }
        const uint32_t n = nv;
        if (tid == 0) { st3(sh.vv, n, pv); st3(sh.va, n, sh.pa); }
        ++nv;
        __syncthreads();
        for (uint32_t j = tid; j < ne; j += HUGE_THREADS) {
            const uint32_t key = H.hor[j];
            huge_make_face(H, sh, S + j, n, key & 127u, key >> 7);
        }
        F = S + ne;
        __syncthreads();
        huge_closest_face(H, F, sh);
        ++it;
    }
}

}  // namespace b3
