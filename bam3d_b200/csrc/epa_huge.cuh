// Third EPA tier: one BLOCK (1024 threads) per pair, polytope and scratch in a per-block global-memory slab, for the pairs whose
// polytope outgrows the overflow tier (epa_big.cuh, 1024 faces). The reference has no face limit: on near-degenerate input (a
// Cylinder cap against a flat face: coplanar rim supports, visibility is rounding noise) its polytope pinches and the face list
// grows multiplicatively — 44 of 2^20 mixed pairs pass 1024 faces, the largest reaches 121 962 — and it still returns a Contact
// after at most max_iterations trips (tests/test_host_cpu.py runs the oracle without a limit on exactly those pairs). To
// return the same Contact the device has to build the same lists, and a sequential walk over 10^5 faces per trip is out of
// the question on one lane, so Polytope::add (epa3d.rs:121-149) is evaluated here WITHOUT a walk:
//   * one pass over the faces compacts the visible positions into a list (block scan); V of them, S = F - V survivors;
//   * `while i < len { if visible(i) { swap_remove(i) } else { i += 1 } }` is a two-pointer process: the holes (visible
//     positions below S, ascending h_1 < h_2 < ...) are filled by the tail survivors (non-visible positions >= S) taken from
//     the back (t_1 > t_2 > ...), so the surviving list is "positions below S, holes replaced by t_k" — one gather over the
//     holes; everything about the tail lives in the V positions [S, F), a short second scan;
//   * the faces are REMOVED in the order h_1, [visible positions in (t_1, F) descending], h_2, [visible positions in
//     (t_2, t_1) descending], ..., and at the end, if positions remain between S and the last tail survivor taken, S itself
//     followed by the rest of them descending: every visible face computes its own place in that order in closed form;
//   * the horizon is that sequence of directed edges under remove_or_add_edge (:183-190). Per undirected edge {a, b} the list only
//     ever holds copies of ONE direction (an arriving edge removes the oldest copy of its reverse if there is one, else it is
//     pushed), so its state is the running balance #(a->b) - #(b->a), removal is first-in-first-out, and what is left at the end
//     are the LAST |balance| events of the majority direction: event e of direction k survives iff its rank among the events
//     with the same directed edge, in sequence order, is >= the number of events of the reverse direction. Counts per directed
//     edge come from shared-memory atomics (vertex indices < 104: two 16384-entry tables), the ranks from one warp that walks
//     the events 32 at a time in sequence order (match_any + a running count per edge), and the survivors are compacted in
//     sequence order — the reference's edge list, in the reference's order. (Pinched polytopes share one edge among
//     thousands of faces: per-edge lists sorted by one thread, the first version of this step, cost 10^9 cycles.)
//   * new faces one per thread; closest face by a block reduction with the reference's first-minimum rule.
// Only the first and the last step touch all F faces; the rest is proportional to the V removed faces (about 1 % of F).
// oracle/bam3d_oracle.hpp: Polytope::add_in_phases is this formulation on the CPU, checked against the sequential add().
#pragma once
#include "epa.cuh"

namespace b3 {

constexpr int HUGE_THREADS = 1024;
constexpr int HUGE_PER_THREAD = 4;                         // faces per thread per tile of the two passes over all faces
constexpr uint32_t HUGE_TILE = HUGE_THREADS * HUGE_PER_THREAD;
constexpr uint32_t HUGE_FCAP = 1u << 18;                   // faces per polytope (== orc::EPA_FACE_LIMIT); beyond: ST_CAPACITY
constexpr uint32_t HUGE_KEYS = 128 * 128;                  // directed-edge keys a | b << 7
constexpr uint32_t HUGE_NIL = 0xFFFFFFFFu;
constexpr int HUGE_BLOCKS = 48;
constexpr uint32_t HUGE_SV = 2048;                         // removed faces per add() whose lists fit in shared memory (nearly every add())
// dynamic shared memory: events per directed edge + the running count of the ranking pass, then the short lists of one add()
constexpr size_t HUGE_SMEM_BYTES = (size_t)HUGE_KEYS * 8 + (size_t)HUGE_SV * (4 + 4 + 4 + 6 + 6 + 1 + 3) + 64;

struct HugeStore {
    float4* fn;         // FCAP      : face normal + distance
    uint32_t* fi;       // FCAP      : face vertex indices
    uint32_t* vlist;    // FCAP      : visible positions, ascending
    uint32_t* tbefore;  // FCAP + 1  : tail position S + x: visible tail positions below it
    uint32_t* tsrc;     // FCAP + 2  : tsrc[k] = position of the k-th tail survivor from the back
    uint32_t* tva;      // FCAP + 2  : tva[k] = visible positions above tsrc[k]
    uint16_t* evkey;    // 3 FCAP    : directed edge of event e (3 per removed face, in removal order)
    uint16_t* hor;      // FCAP      : horizon edges in order
    uint8_t* tflag;     // FCAP      : tail position S + x is visible
    uint8_t* keep;      // 3 FCAP    : event e is in the final horizon list
};
constexpr size_t HUGE_SLAB_BYTES = (size_t)HUGE_FCAP * (16 + 4 + 4 + 4 + 4 + 4 + 6 + 2 + 1 + 3) + 16 * 256;

// The lists whose length is the number of REMOVED faces (vlist .. keep above): in shared memory when they fit (HUGE_SV), else in
// the slab. A global-memory round trip per step of a ten-step chain was most of an add()'s time on the ordinary pinched pairs.
struct HugeLists { uint32_t* vlist; uint32_t* tsrc; uint32_t* tva; uint16_t* evkey; uint16_t* hor; uint8_t* tflag; uint8_t* keep; };
B3_DEV HugeLists carve_huge_lists(unsigned char* p) {
    HugeLists S;
    S.vlist = reinterpret_cast<uint32_t*>(p); p += (size_t)HUGE_SV * 4;
    S.tsrc = reinterpret_cast<uint32_t*>(p); p += ((size_t)HUGE_SV + 2) * 4;
    S.tva = reinterpret_cast<uint32_t*>(p); p += ((size_t)HUGE_SV + 2) * 4;
    S.evkey = reinterpret_cast<uint16_t*>(p); p += (size_t)HUGE_SV * 6;
    S.hor = reinterpret_cast<uint16_t*>(p); p += (size_t)HUGE_SV * 6;
    S.tflag = p; p += HUGE_SV;
    S.keep = p;
    return S;
}

B3_DEV HugeStore carve_huge_store(unsigned char* p) {
    HugeStore H;
    auto take = [&](size_t bytes) { unsigned char* r = p; p += (bytes + 255) & ~(size_t)255; return r; };
    H.fn = reinterpret_cast<float4*>(take((size_t)HUGE_FCAP * 16));
    H.fi = reinterpret_cast<uint32_t*>(take((size_t)HUGE_FCAP * 4));
    H.vlist = reinterpret_cast<uint32_t*>(take((size_t)HUGE_FCAP * 4));
    H.tbefore = reinterpret_cast<uint32_t*>(take(((size_t)HUGE_FCAP + 1) * 4));
    H.tsrc = reinterpret_cast<uint32_t*>(take(((size_t)HUGE_FCAP + 2) * 4));
    H.tva = reinterpret_cast<uint32_t*>(take(((size_t)HUGE_FCAP + 2) * 4));
    H.evkey = reinterpret_cast<uint16_t*>(take((size_t)HUGE_FCAP * 6));
    H.hor = reinterpret_cast<uint16_t*>(take((size_t)HUGE_FCAP * 2));
    H.tflag = take((size_t)HUGE_FCAP);
    H.keep = take((size_t)HUGE_FCAP * 3);
    return H;
}

struct HugeShared {
    float vv[EPA_VCAP * 3];
    float va[EPA_VCAP * 3];
    uint32_t scan[HUGE_THREADS / 32];
    unsigned long long red[HUGE_THREADS / 32];
    V3 pv, pa;
    int best;
    uint32_t cursor;
};

// Exclusive prefix of `v` over the block's threads (thread order) and the block total. Two barriers; sh.scan is free again on return.
B3_DEV uint32_t huge_scan(uint32_t v, HugeShared& sh, uint32_t& total) {
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, off);
        if ((int)lane >= off) x += y;
    }
    if (lane == 31) sh.scan[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t y = sh.scan[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t z = __shfl_up_sync(0xFFFFFFFFu, y, off);
            if ((int)lane >= off) y += z;
        }
        sh.scan[lane] = y;
    }
    __syncthreads();
    const uint32_t wp = warp ? sh.scan[warp - 1] : 0u;
    total = sh.scan[HUGE_THREADS / 32 - 1];
    __syncthreads();
    return wp + x - v;
}

// closest_face_to_origin (epa3d.rs:111-119) over faces [0, F): the first index of the minimum distance over non-NaN distances;
// face 0 when its own distance is NaN (nothing ever compares below NaN). Result in sh.best after the closing barrier.
B3_DEV void huge_closest_face(const HugeStore& H, uint32_t F, HugeShared& sh) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    unsigned long long key = ~0ull;
    for (uint32_t base = 0; base < F; base += HUGE_TILE) {
        float d[HUGE_PER_THREAD];
#pragma unroll
        for (int j = 0; j < HUGE_PER_THREAD; ++j) { const uint32_t k = base + tid + j * HUGE_THREADS; d[j] = k < F ? H.fn[k].w : NAN; }
#pragma unroll
        for (int j = 0; j < HUGE_PER_THREAD; ++j) {
            if (d[j] == d[j]) {
                const uint32_t k = base + tid + j * HUGE_THREADS;
                const uint32_t bits = __float_as_uint(d[j] + 0.0f);  // -0 folded onto +0, as < would
                const uint32_t ord = bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u);
                const unsigned long long kk = ((unsigned long long)ord << 32) | k;
                key = kk < key ? kk : key;
            }
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

#ifdef B3_HUGE_TRACE
#define HT_DECL long long ht_[10] = {0,0,0,0,0,0,0,0,0,0}; long long ht_t = clock64(); unsigned long long ht_V = 0, ht_chain = 0;
#define HT(i) do { __syncthreads(); const long long n_ = clock64(); ht_[i] += n_ - ht_t; ht_t = n_; } while (0)
#define HT_DUMP(F) do { if (threadIdx.x == 0) printf("[huge] F=%u sumV=%llu maxchain=%llu cycles: sup %lld vis %lld holes %lld tail %lld ev %lld chain %lld match %lld hor %lld moves+faces %lld closest %lld\n", F, ht_V, ht_chain, ht_[0], ht_[1], ht_[2], ht_[3], ht_[4], ht_[5], ht_[6], ht_[7], ht_[8], ht_[9]); } while (0)
#else
#define HT_DECL
#define HT(i) do { } while (0)
#define HT_DUMP(F) do { } while (0)
#endif

B3_DEV uint8_t epa_process_huge(const Shape& L, const Shape& R, const Simplex<true>& s, float tolerance, uint32_t max_iterations, const HugeStore& H,
                                HugeShared& sh, uint32_t* head, ContactOut& out) {
    const uint32_t tid = threadIdx.x;
    const HugeLists SL = carve_huge_lists(reinterpret_cast<unsigned char*>(head + 2 * HUGE_KEYS));
    HT_DECL
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
            if (tid == 0) { sh.pv = pv; sh.pa = pa; sh.cursor = 0; }
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
            HT_DUMP(F);
            return ST_OK;
        }
        HT(0);
        // ---- 1. visibility (epa3d.rs:126-129, as the reference writes it) and the ascending list of visible positions ----
        uint32_t V = 0;
        for (uint32_t base = 0; base < F; base += HUGE_TILE) {
            const uint32_t p0 = base + HUGE_PER_THREAD * tid;  // four consecutive faces per thread: the list comes out ascending
            float4 f[HUGE_PER_THREAD];
            uint32_t fi[HUGE_PER_THREAD];
#pragma unroll
            for (int j = 0; j < HUGE_PER_THREAD; ++j)
                if (p0 + j < F) { f[j] = H.fn[p0 + j]; fi[j] = H.fi[p0 + j]; }
            uint32_t bits = 0;
#pragma unroll
            for (int j = 0; j < HUGE_PER_THREAD; ++j) {
                if (p0 + j < F) {
                    const V3 v0 = ld3(sh.vv, fi[j] & 0xFF);
                    if (dot(v3(f[j].x, f[j].y, f[j].z), pv - v0) > 0.f) bits |= 1u << j;
                }
            }
            uint32_t tot;
            uint32_t w = V + huge_scan(__popc(bits), sh, tot);
#pragma unroll
            for (int j = 0; j < HUGE_PER_THREAD; ++j)
                if (bits & (1u << j)) { H.vlist[w] = p0 + j; if (w < HUGE_SV) SL.vlist[w] = p0 + j; ++w; }
            V += tot;
        }
        __syncthreads();
        HT(1);
        const uint32_t S = F - V;
        // the lists of this add(): shared memory unless it removes more than HUGE_SV faces
        const bool small = V <= HUGE_SV;
        uint32_t* const vlist = small ? SL.vlist : H.vlist;
        uint32_t* const tsrc = small ? SL.tsrc : H.tsrc;
        uint32_t* const tva = small ? SL.tva : H.tva;
        uint16_t* const evkey = small ? SL.evkey : H.evkey;
        uint16_t* const hor = small ? SL.hor : H.hor;
        uint8_t* const tflag = small ? SL.tflag : H.tflag;
        uint8_t* const keep = small ? SL.keep : H.keep;
        // holes = visible positions below S = the first Hn entries of the list
        uint32_t cnt = 0;
        for (uint32_t i = tid; i < V; i += HUGE_THREADS) cnt += vlist[i] < S ? 1u : 0u;
        uint32_t Hn;
        huge_scan(cnt, sh, Hn);
        const uint32_t m = V - Hn;  // visible tail positions
        HT(2);
        // ---- 2. the tail [S, F): V positions, m of them visible; survivors numbered from the back ----
        for (uint32_t x = tid; x < V; x += HUGE_THREADS) tflag[x] = 0;
        __syncthreads();
        for (uint32_t i = Hn + tid; i < V; i += HUGE_THREADS) tflag[vlist[i] - S] = 1;
        __syncthreads();
        {
            uint32_t carry = 0;
            for (uint32_t base = 0; base < V; base += HUGE_THREADS) {
                const uint32_t x = base + tid;
                const uint32_t fl = x < V ? (uint32_t)tflag[x] : 0u;
                uint32_t tot;
                const uint32_t before = carry + huge_scan(fl, sh, tot);
                if (x < V && !fl) {
                    const uint32_t vis_after = m - before;                  // visible positions above S + x
                    const uint32_t k = ((V - 1 - x) - vis_after) + 1;       // survivors above it, plus one
                    tsrc[k] = S + x;
                    tva[k] = vis_after;
                }
                carry += tot;
            }
        }
        __syncthreads();
        HT(3);
        // ---- 3. the three directed edges of every removed face, at its place in the removal order ----
        const uint32_t tva_last = Hn ? tva[Hn] : 0u;
        for (uint32_t i = tid; i < V; i += HUGE_THREADS) {
            const uint32_t p = vlist[i];
            uint32_t idx;
            if (i < Hn) {
                idx = i + (i >= 1 ? tva[i] : 0u);                        // hole i + 1: after holes 1..i and the tail faces pulled in so far
            } else {
                const uint32_t va = m - 1 - (i - Hn);                      // visible positions above p
                const uint32_t sv = (F - 1 - p) - va;                      // survivors above p
                if (sv < Hn) idx = (sv + 1) + va;                          // pulled in while hole sv + 1 was being filled
                else if (p == S) idx = Hn + tva_last;
                else idx = Hn + 1 + va;
            }
            const uint32_t f = H.fi[p];
            const uint32_t a = f & 0xFF, b = (f >> 8) & 0xFF, c = (f >> 16) & 0xFF;
            evkey[3 * (size_t)idx] = (uint16_t)(a | (b << 7));
            evkey[3 * (size_t)idx + 1] = (uint16_t)(b | (c << 7));
            evkey[3 * (size_t)idx + 2] = (uint16_t)(c | (a << 7));
        }
        __syncthreads();
        HT(4);
        // ---- 4. first-in-first-out matching per undirected edge ----
        const uint32_t E = 3 * V;
#ifdef B3_HUGE_TRACE
        ht_V += V;
#endif
        uint32_t* ecnt = head;              // events per directed edge
        uint32_t* erun = head + HUGE_KEYS;  // same-edge events seen so far by the ranking pass
        for (uint32_t e = tid; e < E; e += HUGE_THREADS) atomicAdd(&ecnt[evkey[e]], 1u);
        __syncthreads();
        HT(5);
        if (tid < 32) {
            for (uint32_t base = 0; base < E; base += 32) {
                const uint32_t e = base + tid;
                const bool valid = e < E;
                const uint32_t key = valid ? (uint32_t)evkey[e] : (0x10000u + tid);  // idle lanes: keys of their own
                const uint32_t peers = __match_any_sync(0xFFFFFFFFu, key);
                if (valid) {
                    const uint32_t rev = ((key & 127u) << 7) | (key >> 7);
                    const uint32_t n_rev = ecnt[rev];
                    const uint32_t rank = erun[key] + __popc(peers & ((1u << tid) - 1u));
                    keep[e] = (ecnt[key] > n_rev && rank >= n_rev) ? 1 : 0;
                }
                __syncwarp();
                if (valid && (peers >> tid) <= 1u) erun[key] += __popc(peers);  // the highest lane of each group
                __syncwarp();
            }
        }
        __syncthreads();
        HT(6);
        for (uint32_t e = tid; e < E; e += HUGE_THREADS) { const uint32_t key = evkey[e]; ecnt[key] = 0; erun[key] = 0; }
        // ---- 5. the horizon: surviving events in sequence order ----
        uint32_t ne = 0;
        for (uint32_t base = 0; base < E; base += HUGE_TILE) {
            const uint32_t e0 = base + HUGE_PER_THREAD * tid;
            uint32_t bits = 0;
#pragma unroll
            for (int j = 0; j < HUGE_PER_THREAD; ++j)
                if (e0 + j < E && keep[e0 + j]) bits |= 1u << j;
            uint32_t tot;
            uint32_t w = ne + huge_scan(__popc(bits), sh, tot);
#pragma unroll
            for (int j = 0; j < HUGE_PER_THREAD; ++j)
                if ((bits & (1u << j)) && w < HUGE_FCAP) hor[w++] = evkey[e0 + j];
            ne += tot;
        }
        HT(7);
        if (nv >= (uint32_t)EPA_VCAP || (unsigned long long)S + ne > HUGE_FCAP) { __syncthreads(); return ST_CAPACITY; }
        // ---- 6. the surviving list: hole k takes the k-th tail survivor from the back (sources >= S, destinations < S) ----
        for (uint32_t i = tid; i < Hn; i += HUGE_THREADS) {
            const uint32_t dst = vlist[i], src = tsrc[i + 1];
            H.fn[dst] = H.fn[src];
            H.fi[dst] = H.fi[src];
        }
        const uint32_t n = nv;
        if (tid == 0) { st3(sh.vv, n, pv); st3(sh.va, n, sh.pa); sh.cursor = 0; }
        ++nv;
        __syncthreads();
        for (uint32_t j = tid; j < ne; j += HUGE_THREADS) {
            const uint32_t key = hor[j];
            huge_make_face(H, sh, S + j, n, key & 127u, key >> 7);
        }
        F = S + ne;
        __syncthreads();
        HT(8);
        huge_closest_face(H, F, sh);
        HT(9);
        ++it;
    }
}

}  // namespace b3
