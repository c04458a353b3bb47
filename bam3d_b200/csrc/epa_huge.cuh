// Third EPA tier: one BLOCK (1024 threads) per pair, polytope and scratch in a per-block global-memory slab, for the pairs whose
// polytope outgrows the overflow tier (epa_big.cuh, 1024 faces). The reference has no face limit: on near-degenerate input (a
// Cylinder cap against a flat face: coplanar rim supports, visibility is rounding noise) its polytope pinches and the face list
// grows multiplicatively — 44 of 2^20 mixed pairs pass 1024 faces, the largest reaches 121 962 — and it still returns a Contact
// after at most max_iterations trips (tests/test_host_cpu.py runs the oracle without a limit on exactly those pairs). To
// return the same Contact the device has to build the same lists, and a sequential walk over 10^5 faces per trip is out of
// the question on one lane, so Polytope::add (epa3d.rs:121-149) is evaluated here WITHOUT a walk:
//   * one pass over the faces marks the visible ones; every warp owns a contiguous range of positions, keeps one ballot word per
//     32 faces in shared memory and a count, and after ONE barrier writes its visible positions at its offset: the ascending
//     list of the V visible positions; S = F - V faces survive. The same pass carries the minimum (distance, position) of the
//     faces that stay where they are, so the closest face of the next trip needs no second pass over the list;
//   * `while i < len { if visible(i) { swap_remove(i) } else { i += 1 } }` is a two-pointer process: the holes (visible
//     positions below S, ascending h_1 < h_2 < ...) are filled by the tail survivors (non-visible positions >= S) taken from
//     the back (t_1 > t_2 > ...), so the surviving list is "positions below S, holes replaced by t_k" — one gather over the
//     holes; everything about the tail lives in the V positions [S, F): a thread per tail position finds the visible tail
//     positions below it by binary search in the list;
//   * the faces are REMOVED in the order h_1, [visible positions in (t_1, F) descending], h_2, [visible positions in
//     (t_2, t_1) descending], ..., and at the end, if positions remain between S and the last tail survivor taken, S itself
//     followed by the rest of them descending: every visible face computes its own place in that order in closed form;
//   * the horizon is that sequence of directed edges under remove_or_add_edge (:183-190). Per undirected edge {a, b} the list only
//     ever holds copies of ONE direction (an arriving edge removes the oldest copy of its reverse if there is one, else it is
//     pushed), so its state is the running balance #(a->b) - #(b->a), removal is first-in-first-out, and what is left at the end
//     are the LAST |balance| events of the majority direction: event e of direction k survives iff its rank among the events
//     with the same directed edge, in sequence order, is >= the number of events of the reverse direction. Counts per directed
//     edge come from shared-memory atomics (vertex indices < 104: one 16384-entry table). An event whose reverse never occurs
//     survives, one of the minority (or balanced) direction does not: no rank needed. The directed edges that do need ranks are
//     few (at most 34 per add() over the 44 pinched pairs of the C2 scene, none in half of the add()s, 9 % of the events): each
//     gets a slot (64 per round), every warp ranks the events of its contiguous range against a per-(slot, warp) counter
//     (match_any), one prefix over the warps per slot makes the ranks global, and the survivors are compacted in sequence order
//     like the visible positions — the reference's edge list, in the reference's order. (One warp walking all events in order,
//     the previous version of this step, was half of the largest pair's time; per-edge lists sorted by one thread, the first
//     version, cost 10^9 cycles.)
//   * new faces one per thread; closest face = minimum over (faces that stayed, faces moved into holes, new faces) with the
//     reference's first-minimum rule; a full pass only when the minimum of the first set sits in the tail that moved.
// Eleven block barriers per add() (nine when no edge needs a rank). Only the first pass touches all F faces; the rest is
// proportional to the V removed faces (about 1 % of F).
// oracle/bam3d_oracle.hpp: Polytope::add_in_phases is this formulation on the CPU, checked against the sequential add().
#pragma once
#include "epa.cuh"

namespace b3 {

#ifndef B3_HUGE_THREADS
#define B3_HUGE_THREADS 512
#endif
constexpr int HUGE_THREADS = B3_HUGE_THREADS;              // 512: 128 registers per thread, no spills (1024 spilled 140 bytes into a 28 KB L1)
constexpr uint32_t HUGE_WARPS = HUGE_THREADS / 32;
constexpr int HUGE_UNROLL = HUGE_THREADS >= 1024 ? 4 : 8;   // 32-face chunks a warp has in flight in the pass over all faces
constexpr uint32_t HUGE_FCAP = 1u << 18;                   // faces per polytope (== orc::EPA_FACE_LIMIT); beyond: ST_CAPACITY
constexpr uint32_t HUGE_KEYS = 128 * 128;                  // directed-edge keys a | b << 7
constexpr int HUGE_BLOCKS = 48;
constexpr uint32_t HUGE_SV = 2048;                         // removed faces per add() whose lists fit in shared memory (99 % of the add()s)
constexpr uint32_t HUGE_SLOTS = 64;                        // directed edges ranked per round
constexpr uint16_t HUGE_LOCK = 0xFFFFu;

struct HugeStore {
    float4* fn;         // FCAP      : face normal + distance
    uint32_t* fi;       // FCAP      : face vertex indices
    uint32_t* vlist;    // FCAP      : visible positions, ascending
    uint32_t* tsrc;     // FCAP + 2  : tsrc[k] = position of the k-th tail survivor from the back
    uint32_t* tva;      // FCAP + 2  : tva[k] = visible positions above tsrc[k]
    uint16_t* evkey;    // 3 FCAP    : directed edge of event e (3 per removed face, in removal order)
    uint16_t* lrank;    // 3 FCAP    : rank of event e among the same-edge events of its warp's range
    uint16_t* hor;      // FCAP      : horizon edges in order
    uint8_t* keep;      // 3 FCAP    : 0 dropped, 1 in the final horizon list, 2 needs its rank
};
constexpr size_t HUGE_SLAB_BYTES = (size_t)HUGE_FCAP * (16 + 4 + 4 + 4 + 4 + 6 + 6 + 2 + 3) + 16 * 256;

B3_DEV HugeStore carve_huge_store(unsigned char* p) {
    HugeStore H;
    auto take = [&](size_t bytes) { unsigned char* r = p; p += (bytes + 255) & ~(size_t)255; return r; };
    H.fn = reinterpret_cast<float4*>(take((size_t)HUGE_FCAP * 16));
    H.fi = reinterpret_cast<uint32_t*>(take((size_t)HUGE_FCAP * 4));
    H.vlist = reinterpret_cast<uint32_t*>(take((size_t)HUGE_FCAP * 4));
    H.tsrc = reinterpret_cast<uint32_t*>(take(((size_t)HUGE_FCAP + 2) * 4));
    H.tva = reinterpret_cast<uint32_t*>(take(((size_t)HUGE_FCAP + 2) * 4));
    H.evkey = reinterpret_cast<uint16_t*>(take((size_t)HUGE_FCAP * 6));
    H.lrank = reinterpret_cast<uint16_t*>(take((size_t)HUGE_FCAP * 6));
    H.hor = reinterpret_cast<uint16_t*>(take((size_t)HUGE_FCAP * 2));
    H.keep = take((size_t)HUGE_FCAP * 3);
    return H;
}

struct HugeShared {
    float vv[EPA_VCAP * 3];
    float va[EPA_VCAP * 3];
    Shape L, R;        // the pair (read by warp 0: GJK and the support points)
    uint32_t scan[32];            // per warp; entries from HUGE_WARPS on stay 0
    unsigned long long red[32];   // per warp; entries from HUGE_WARPS on stay HUGE_NOKEY
    V3 pv, pa;
    float d0;          // distance of the face at position 0 (NaN there decides closest_face_to_origin)
    uint32_t hn;       // holes of the add() in flight
    uint32_t nslots;   // directed edges that need ranks in the add() in flight
};

// Dynamic shared memory of the block: the per-edge tables, then the lists whose length is the number of REMOVED faces (in shared
// memory when V <= HUGE_SV, else in the slab: a global-memory round trip per step of the chain is most of a small add()).
struct HugeTables {
    uint32_t* ecnt;     // KEYS           : events per directed edge (zero between add()s)
    uint32_t* segcnt;   // SLOTS x 32     : ranked events of slot s in warp w's range (zero between add()s)
    uint32_t* vbits;    // FCAP / 32      : visibility ballots
    uint16_t* slot_of;  // KEYS           : 1 + slot of a directed edge that needs ranks (zero between add()s)
    uint32_t* vlist; uint32_t* tsrc; uint32_t* tva; uint16_t* evkey; uint16_t* lrank; uint16_t* hor; uint8_t* keep;
};
constexpr size_t HUGE_ZERO_WORDS = (size_t)HUGE_KEYS + HUGE_SLOTS * 32;  // ecnt + segcnt, zeroed by the kernel (slot_of separately)
constexpr size_t HUGE_SMEM_BYTES = (size_t)HUGE_KEYS * 4 + HUGE_SLOTS * 32 * 4 + (HUGE_FCAP / 32) * 4 + (size_t)HUGE_KEYS * 2 +
                                   (size_t)HUGE_SV * (4 + 4 + 4 + 6 + 6 + 6 + 3) + 64;
B3_DEV HugeTables carve_huge_tables(unsigned char* p) {
    HugeTables T;
    T.ecnt = reinterpret_cast<uint32_t*>(p); p += (size_t)HUGE_KEYS * 4;
    T.segcnt = reinterpret_cast<uint32_t*>(p); p += (size_t)HUGE_SLOTS * 32 * 4;
    T.vbits = reinterpret_cast<uint32_t*>(p); p += (size_t)(HUGE_FCAP / 32) * 4;
    T.slot_of = reinterpret_cast<uint16_t*>(p); p += (size_t)HUGE_KEYS * 2;
    T.vlist = reinterpret_cast<uint32_t*>(p); p += (size_t)HUGE_SV * 4;
    T.tsrc = reinterpret_cast<uint32_t*>(p); p += ((size_t)HUGE_SV + 2) * 4;
    T.tva = reinterpret_cast<uint32_t*>(p); p += ((size_t)HUGE_SV + 2) * 4;
    T.evkey = reinterpret_cast<uint16_t*>(p); p += (size_t)HUGE_SV * 6;
    T.lrank = reinterpret_cast<uint16_t*>(p); p += (size_t)HUGE_SV * 6;
    T.hor = reinterpret_cast<uint16_t*>(p); p += (size_t)HUGE_SV * 6;
    T.keep = p;
    return T;
}
// the kernel calls this once; epa_process_huge leaves the tables zero again on every return
B3_DEV void huge_zero_tables(unsigned char* smem, HugeShared& sh) {
    const HugeTables T = carve_huge_tables(smem);
    if (threadIdx.x < 32) { sh.scan[threadIdx.x] = 0u; sh.red[threadIdx.x] = ~0ull; }
    for (uint32_t k = threadIdx.x; k < HUGE_ZERO_WORDS; k += HUGE_THREADS) T.ecnt[k] = 0u;  // ecnt and segcnt are adjacent
    for (uint32_t k = threadIdx.x; k < HUGE_KEYS; k += HUGE_THREADS) T.slot_of[k] = 0;
    __syncthreads();
}


constexpr uint32_t HUGE_FULL = 0xFFFFFFFFu;
constexpr unsigned long long HUGE_NOKEY = ~0ull;

// (distance, position) as one ordered key: `<` on the distances (-0 folded onto +0), ties to the lower position
B3_DEV unsigned long long huge_key(float d, uint32_t pos) {
    const uint32_t bits = __float_as_uint(d + 0.0f);
    const uint32_t ord = bits ^ ((bits >> 31) ? 0xFFFFFFFFu : 0x80000000u);
    return ((unsigned long long)ord << 32) | pos;
}
B3_DEV unsigned long long huge_warp_min(unsigned long long key) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long o = __shfl_xor_sync(HUGE_FULL, key, off);
        key = o < key ? o : key;
    }
    return key;
}
// minimum over the block, in every thread. One barrier; sh.red may be rewritten after the caller's next barrier.
B3_DEV unsigned long long huge_block_min(unsigned long long key, HugeShared& sh) {
    key = huge_warp_min(key);
    if ((threadIdx.x & 31u) == 0) sh.red[threadIdx.x >> 5] = key;
    __syncthreads();
    return huge_warp_min(sh.red[threadIdx.x & 31u]);
}
// sh.scan[w] = items warp w keeps, published by a barrier: -> this warp's offset, and the total
B3_DEV uint32_t huge_offsets(const HugeShared& sh, uint32_t& total) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t x = sh.scan[lane];
    uint32_t incl = x;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t y = __shfl_up_sync(HUGE_FULL, incl, off);
        if ((int)lane >= off) incl += y;
    }
    total = __shfl_sync(HUGE_FULL, incl, 31);
    return __shfl_sync(HUGE_FULL, incl - x, threadIdx.x >> 5);
}
// every warp's contiguous range of [0, n): whole 32-item chunks
B3_DEV void huge_range(uint32_t n, uint32_t& chunk, uint32_t& lo, uint32_t& hi) {
    chunk = ((n + (uint32_t)HUGE_THREADS - 1u) / (uint32_t)HUGE_THREADS) << 5;
    lo = min(n, (threadIdx.x >> 5) * chunk);
    hi = min(n, lo + chunk);
}

// closest_face_to_origin (epa3d.rs:111-119) over faces [0, F): the first index of the minimum distance over non-NaN distances;
// face 0 when its own distance is NaN (nothing ever compares below NaN) — the caller applies that with sh.d0.
B3_DEV unsigned long long huge_closest_scan(const HugeStore& H, uint32_t F, HugeShared& sh) {
    unsigned long long key = HUGE_NOKEY;
    for (uint32_t base = 0; base < F; base += 4 * HUGE_THREADS) {
        float d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { const uint32_t k = base + threadIdx.x + j * HUGE_THREADS; d[j] = k < F ? H.fn[k].w : NAN; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (d[j] == d[j]) {
                const unsigned long long kk = huge_key(d[j], base + threadIdx.x + j * HUGE_THREADS);
                key = kk < key ? kk : key;
            }
        }
    }
    return huge_block_min(key, sh);
}

// Face::new (epa3d.rs:159-172) of (a, b, c) at `slot`; -> its distance
B3_DEV float huge_make_face(const HugeStore& H, uint32_t slot, uint32_t a, uint32_t b, uint32_t c, const V3& va, const V3& vb, const V3& vc) {
    const V3 ab = vb - va;
    const V3 ac = vc - va;
    const V3 n = normalize(cross(ab, ac));
    const float d = dot(n, va);
    H.fn[slot] = make_float4(n.x, n.y, n.z, d);
    H.fi[slot] = a | (b << 8) | (c << 16);
    return d;
}

#ifdef B3_HUGE_TRACE
#define HT_DECL long long ht_[8] = {0,0,0,0,0,0,0,0}; long long ht_t = clock64(); unsigned long long ht_V = 0, ht_ranked = 0;
#define HT(i) do { __syncthreads(); const long long n_ = clock64(); ht_[i] += n_ - ht_t; ht_t = n_; } while (0)
#define HT_DUMP(F) do { if (threadIdx.x == 0) printf("[huge] F=%u sumV=%llu ranked_adds=%llu cycles: sup %lld vis %lld tail %lld ev %lld rank %lld hor %lld moves %lld faces+min %lld\n", F, ht_V, ht_ranked, ht_[0], ht_[1], ht_[2], ht_[3], ht_[4], ht_[5], ht_[6], ht_[7]); } while (0)
#else
#define HT_DECL
#define HT(i) do { } while (0)
#define HT_DUMP(F) do { } while (0)
#endif

// The pair is in sh.L / sh.R and the GJK simplex (4 points) in sh.vv / sh.va [0, 4).
B3_DEV uint8_t epa_process_huge(float tolerance, uint32_t max_iterations, const HugeStore& H, HugeShared& sh, unsigned char* smem, ContactOut& out) {
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;
    const HugeTables T = carve_huge_tables(smem);
    HT_DECL
    if (tid == 0) {
        // Polytope::new (epa3d.rs:100-109)
        const V3 s0 = ld3(sh.vv, 0), s1 = ld3(sh.vv, 1), s2 = ld3(sh.vv, 2), s3 = ld3(sh.vv, 3);
        const float d0 = huge_make_face(H, 0, 3, 2, 1, s3, s2, s1);
        huge_make_face(H, 1, 3, 1, 0, s3, s1, s0);
        huge_make_face(H, 2, 3, 0, 2, s3, s0, s2);
        huge_make_face(H, 3, 2, 0, 1, s2, s0, s1);
        sh.d0 = d0;
        sh.hn = 0;
        sh.nslots = 0;
    }
    __syncthreads();
    uint32_t nv = 4, F = 4, it = 1;
    uint32_t best;
    {
        const unsigned long long k = huge_closest_scan(H, F, sh);
        const float d0 = sh.d0;
        best = (d0 != d0 || k == HUGE_NOKEY) ? 0u : (uint32_t)k;
    }
    for (;;) {
        const float4 bf = H.fn[best];
        const V3 normal = v3(bf.x, bf.y, bf.z);
        if (tid < 32) {  // the support functions are warp-convergent (polyhedron vertex scans): warp 0 computes, the block reads
            V3 pv, pa;
            minkowski_support<true>(sh.L, sh.R, normal, pv, pa);
            if (tid == 0) { sh.pv = pv; sh.pa = pa; }
        }
        __syncthreads();  // (1)
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
        // ---- 1. visibility (epa3d.rs:126-129, as the reference writes it): ballots per warp range, then the ascending list ----
        uint32_t fchunk, flo, fhi;
        huge_range(F, fchunk, flo, fhi);
        unsigned long long k_stay = HUGE_NOKEY;  // closest of the faces that are not removed (those in the tail will move: checked below)
        {
            uint32_t cnt = 0;
            for (uint32_t base = flo; base < fhi; base += 32 * HUGE_UNROLL) {
                float4 f[HUGE_UNROLL];
                uint32_t i0[HUGE_UNROLL];
#pragma unroll
                for (int j = 0; j < HUGE_UNROLL; ++j) {
                    const uint32_t p = base + 32 * j + lane;
                    if (p < fhi) { f[j] = __ldcg(H.fn + p); i0[j] = __ldcg(H.fi + p) & 0xFF; }  // streamed: L1 is small and holds the stack
                }
#pragma unroll
                for (int j = 0; j < HUGE_UNROLL; ++j) {
                    const uint32_t p = base + 32 * j + lane;
                    bool vis = false;
                    if (p < fhi) {
                        vis = dot(v3(f[j].x, f[j].y, f[j].z), pv - ld3(sh.vv, i0[j])) > 0.f;
                        if (!vis && f[j].w == f[j].w) { const unsigned long long kk = huge_key(f[j].w, p); k_stay = kk < k_stay ? kk : k_stay; }
                    }
                    const uint32_t b = __ballot_sync(HUGE_FULL, vis);
                    if (base + 32 * j < fhi) {
                        if (lane == 0) T.vbits[(base + 32 * j) >> 5] = b;
                        cnt += __popc(b);
                    }
                }
            }
            k_stay = huge_warp_min(k_stay);
            if (lane == 0) { sh.scan[warp] = cnt; sh.red[warp] = k_stay; }
        }
        __syncthreads();  // (2)
        uint32_t V;
        const uint32_t voff = huge_offsets(sh, V);
        k_stay = huge_warp_min(sh.red[lane]);
        const uint32_t S = F - V;
        // the lists of this add(): shared memory unless it removes more than HUGE_SV faces
        const bool small = V <= HUGE_SV;
        uint32_t* const vlist = small ? T.vlist : H.vlist;
        uint32_t* const tsrc = small ? T.tsrc : H.tsrc;
        uint32_t* const tva = small ? T.tva : H.tva;
        uint16_t* const evkey = small ? T.evkey : H.evkey;
        uint16_t* const lrank = small ? T.lrank : H.lrank;
        uint16_t* const hor = small ? T.hor : H.hor;
        uint8_t* const keep = small ? T.keep : H.keep;
        {
            uint32_t w = voff, below = 0;
            for (uint32_t base = flo; base < fhi; base += 32) {
                const uint32_t b = T.vbits[base >> 5];
                if ((b >> lane) & 1u) vlist[w + __popc(b & lt)] = base + lane;
                w += __popc(b);
                if (base < S) below += __popc(S - base >= 32u ? b : (b & ((1u << (S - base)) - 1u)));  // holes: visible positions below S
            }
            if (lane == 0 && below) atomicAdd(&sh.hn, below);
        }
        __syncthreads();  // (3)
        HT(1);
        const uint32_t Hn = sh.hn;
        const uint32_t m = V - Hn;  // visible tail positions
        // ---- 2. the tail [S, F): V positions, m of them visible (the list from Hn on); survivors numbered from the back ----
        {
            const uint32_t* tv = vlist + Hn;
            for (uint32_t x = tid; x < V; x += HUGE_THREADS) {
                const uint32_t p = S + x;
                uint32_t l = 0, r = m;
                while (l < r) {
                    const uint32_t mid = (l + r) >> 1;
                    if (tv[mid] < p) l = mid + 1; else r = mid;
                }
                if (!(l < m && tv[l] == p)) {
                    const uint32_t vis_after = m - l;                       // visible positions above p
                    const uint32_t k = ((V - 1 - x) - vis_after) + 1;       // survivors above it, plus one
                    tsrc[k] = p;
                    tva[k] = vis_after;
                }
            }
        }
        __syncthreads();  // (4)
        HT(2);
        // ---- 3. the three directed edges of every removed face, at its place in the removal order; events per directed edge ----
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
            const uint32_t k0 = a | (b << 7), k1 = b | (c << 7), k2 = c | (a << 7);
            evkey[3 * (size_t)idx] = (uint16_t)k0;
            evkey[3 * (size_t)idx + 1] = (uint16_t)k1;
            evkey[3 * (size_t)idx + 2] = (uint16_t)k2;
            atomicAdd(&T.ecnt[k0], 1u);
            atomicAdd(&T.ecnt[k1], 1u);
            atomicAdd(&T.ecnt[k2], 1u);
        }
        __syncthreads();  // (5)
        HT(3);
        // ---- 4. first-in-first-out matching per undirected edge: decide what needs no rank, give the other directed edges a slot ----
        const uint32_t E = 3 * V;
#ifdef B3_HUGE_TRACE
        ht_V += V;
#endif
        for (uint32_t e0 = tid; e0 < E; e0 += 4 * HUGE_THREADS) {  // four events per thread in flight (the lists of a large add() are in global memory)
            uint32_t keys[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const uint32_t e = e0 + j * HUGE_THREADS; keys[j] = e < E ? (uint32_t)evkey[e] : 0u; }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t e = e0 + j * HUGE_THREADS;
                if (e >= E) break;
                const uint32_t key = keys[j];
                const uint32_t rev = ((key & 127u) << 7) | (key >> 7);
                const uint32_t n_same = T.ecnt[key], n_rev = T.ecnt[rev];
                const uint8_t k = n_same > n_rev ? (n_rev ? 2 : 1) : 0;
                keep[e] = k;
                if (k == 2) {
                    const unsigned short old = atomicCAS(reinterpret_cast<unsigned short*>(&T.slot_of[key]), (unsigned short)0, (unsigned short)HUGE_LOCK);
                    if (old == 0) *reinterpret_cast<volatile uint16_t*>(&T.slot_of[key]) = (uint16_t)(atomicAdd(&sh.nslots, 1u) + 1u);
                }
            }
        }
        __syncthreads();  // (6)
        const uint32_t nslots = sh.nslots;
        uint32_t echunk, elo, ehi;
        huge_range(E, echunk, elo, ehi);
#ifdef B3_HUGE_TRACE
        ht_ranked += nslots ? 1 : 0;
#endif
        for (uint32_t r0 = 0; r0 < nslots; r0 += HUGE_SLOTS) {
            // ranks inside every warp's range, against the warp's own column of the slot counters
            for (uint32_t base0 = elo; base0 < ehi; base0 += 128) {
                uint32_t kp[4], ky[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t e = base0 + 32 * j + lane;
                    kp[j] = e < ehi ? (uint32_t)keep[e] : 0u;
                    ky[j] = e < ehi ? (uint32_t)evkey[e] : 0u;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t e = base0 + 32 * j + lane;
                    bool need = false;
                    const uint32_t key = ky[j];
                    uint32_t slot = 0;
                    if (kp[j] == 2) {
                        slot = (uint32_t)T.slot_of[key] - 1u - r0;
                        need = slot < HUGE_SLOTS;
                    }
                    if (!__any_sync(HUGE_FULL, need)) continue;
                    const uint32_t peers = __match_any_sync(HUGE_FULL, need ? key : (0x10000u + lane));  // other lanes: keys of their own
                    uint32_t* const cell = T.segcnt + (need ? slot : 0u) * 32u + warp;
                    if (need) lrank[e] = (uint16_t)(*cell + __popc(peers & lt));
                    __syncwarp();
                    if (need && (peers >> lane) <= 1u) *cell += __popc(peers);  // the highest lane of each group
                    __syncwarp();
                }
            }
            __syncthreads();  // (7)
            const uint32_t ns = min(HUGE_SLOTS, nslots - r0);
            for (uint32_t sl = warp; sl < ns; sl += HUGE_WARPS) {  // exclusive prefix over the warps: ranked events of the slot in earlier ranges
                const uint32_t x = T.segcnt[sl * 32u + lane];
                uint32_t incl = x;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) {
                    const uint32_t y = __shfl_up_sync(HUGE_FULL, incl, off);
                    if ((int)lane >= off) incl += y;
                }
                T.segcnt[sl * 32u + lane] = incl - x;
            }
            __syncthreads();  // (8)
            for (uint32_t e0 = tid; e0 < E; e0 += 4 * HUGE_THREADS) {
                uint32_t kp[4], ky[4], lr[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t e = e0 + j * HUGE_THREADS;
                    kp[j] = e < E ? (uint32_t)keep[e] : 0u;
                    ky[j] = e < E ? (uint32_t)evkey[e] : 0u;
                    lr[j] = e < E ? (uint32_t)lrank[e] : 0u;  // only meaningful where keep is 2
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (kp[j] != 2) continue;
                    const uint32_t e = e0 + j * HUGE_THREADS;
                    const uint32_t key = ky[j];
                    const uint32_t slot = (uint32_t)T.slot_of[key] - 1u - r0;
                    if (slot >= HUGE_SLOTS) continue;
                    const uint32_t rev = ((key & 127u) << 7) | (key >> 7);
                    const uint32_t rank = T.segcnt[slot * 32u + e / echunk] + lr[j];
                    keep[e] = rank >= T.ecnt[rev] ? 1 : 0;
                }
            }
            __syncthreads();  // (9)
            for (uint32_t k = tid; k < ns * 32u; k += HUGE_THREADS) T.segcnt[k] = 0u;
            if (r0 + HUGE_SLOTS < nslots) __syncthreads();
        }
        HT(4);
        // ---- 5. the horizon: surviving events in sequence order ----
        {
            uint32_t cnt = 0;
            for (uint32_t base0 = elo; base0 < ehi; base0 += 128) {
                uint32_t kp[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { const uint32_t e = base0 + 32 * j + lane; kp[j] = e < ehi ? (uint32_t)keep[e] : 0u; }
#pragma unroll
                for (int j = 0; j < 4; ++j) cnt += __popc(__ballot_sync(HUGE_FULL, kp[j] == 1));
            }
            if (lane == 0) sh.scan[warp] = cnt;
        }
        __syncthreads();  // (10)
        uint32_t ne;
        {
            uint32_t w = huge_offsets(sh, ne);
            for (uint32_t base0 = elo; base0 < ehi; base0 += 128) {
                uint32_t kp[4], ky[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint32_t e = base0 + 32 * j + lane;
                    kp[j] = e < ehi ? (uint32_t)keep[e] : 0u;
                    ky[j] = e < ehi ? (uint32_t)evkey[e] : 0u;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const bool k = kp[j] == 1;
                    const uint32_t b = __ballot_sync(HUGE_FULL, k);
                    const uint32_t pos = w + __popc(b & lt);
                    if (k && pos < HUGE_FCAP) hor[pos] = (uint16_t)ky[j];
                    w += __popc(b);
                }
            }
        }
        // the tables are zero again for the next add()
        for (uint32_t e = tid; e < E; e += HUGE_THREADS) { const uint32_t key = evkey[e]; T.ecnt[key] = 0u; T.slot_of[key] = 0; }
        if (tid == 0) { sh.hn = 0; sh.nslots = 0; }
        if (nv >= (uint32_t)EPA_VCAP || (unsigned long long)S + ne > HUGE_FCAP) { __syncthreads(); return ST_CAPACITY; }
        // ---- 6. the surviving list: hole k takes the k-th tail survivor from the back (sources >= S, destinations < S) ----
        unsigned long long k_new = HUGE_NOKEY;
        for (uint32_t i = tid; i < Hn; i += HUGE_THREADS) {
            const uint32_t dst = vlist[i], src = tsrc[i + 1];
            const float4 f = H.fn[src];
            H.fn[dst] = f;
            H.fi[dst] = H.fi[src];
            if (f.w == f.w) { const unsigned long long kk = huge_key(f.w, dst); k_new = kk < k_new ? kk : k_new; }
            if (dst == 0) sh.d0 = f.w;
        }
        HT(5);
        __syncthreads();  // (11) the tail has been read: the new faces may overwrite it; hor is complete
        const uint32_t n = nv;
        if (tid == 0) { st3(sh.vv, n, pv); st3(sh.va, n, sh.pa); }
        ++nv;
        for (uint32_t j = tid; j < ne; j += HUGE_THREADS) {
            const uint32_t key = hor[j];
            const uint32_t a = key & 127u, b = key >> 7;
            const float dj = huge_make_face(H, S + j, n, a, b, pv, ld3(sh.vv, a), ld3(sh.vv, b));
            if (dj == dj) { const unsigned long long kk = huge_key(dj, S + j); k_new = kk < k_new ? kk : k_new; }
            if (S + j == 0) sh.d0 = dj;
        }
        HT(6);
        F = S + ne;
        // ---- closest face of the new list ----
        k_new = huge_block_min(k_new, sh);  // (12) also publishes the new vertex, the faces and sh.d0
        unsigned long long kmin;
        if (k_stay != HUGE_NOKEY && (uint32_t)k_stay >= S) {
            __syncthreads();                 // sh.red is read above
            kmin = huge_closest_scan(H, F, sh);  // the closest staying face sat in the tail: it moved, its position with it
        } else {
            kmin = k_stay < k_new ? k_stay : k_new;
        }
        {
            const float d0 = sh.d0;
            best = (d0 != d0 || kmin == HUGE_NOKEY) ? 0u : (uint32_t)kmin;
        }
        HT(7);
        ++it;
    }
}

}  // namespace b3
