// EPA3 with the polytope in SHARED memory, in size tiers. Reference: epa/epa3d.rs:16-190 (same ordered lists and
// tie-breaks as epa.cuh, which documents them; this file reuses its Face/edge/closest-face helpers).
//
// Why: with one polytope per thread in local memory (epa.cuh + epa_refill_kernel) the working set of the resident
// threads (~75k x 1-4 KB) is beyond L1 and L2, every face scan streams from DRAM, lanes of a warp hold polytopes
// of different sizes (12 of 32 lanes active, every 32-byte sector fetched for a few useful bytes) and ncu shows the
// kernel at 1.1 IPC/SM with 59 % of the stall samples on the long scoreboard. Here every polytope that is being
// expanded lives in the SM's shared memory, so an EPA iteration never waits on L2/DRAM.
//
// Layout: one warp per block, six blocks per SM, each warp owns a pool of TIER_POOL_BYTES of shared memory cut
// into G = 32 / LPS equal slots; LPS lanes cooperate on the pair in a slot ("lanes per slot"):
//   tier 0  LPS 1   32 pairs/warp   22 vertices,  40 faces   box-like pairs (<= ~18 EPA iterations)
//   tier 1  LPS 2   16 pairs/warp   42 vertices,  86 faces   one curved side
//   tier 2  LPS 4    8 pairs/warp   83 vertices, 178 faces   ball-like Minkowski sums (~67 iterations, ~136 faces)
//   tier 3  LPS 8    4 pairs/warp  104 vertices, 256 faces   whatever outgrew tier 2
// A pair starts in the tier its (kindL, kindR) bin predicts; one that outgrows its slot is appended to the next
// tier's promotion list and restarted there from its GJK simplex (EPA is deterministic, so the restart replays the
// same sequence); one that outgrows tier 3 goes to the overflow tier (epa_big.cuh) exactly as before.
// Within a slot the first lane ("owner") runs the serial parts — support point, convergence test, the
// order-dependent swap_remove walk — and all LPS lanes share the two O(faces) parts: the visibility scan
// (8 faces per lane per step, one byte of the visibility bitmask each) and the construction of the new faces.
// The closest face is carried through Polytope::add as in epa.cuh and merged across the lanes of a slot as the
// minimum by (distance, index), which is the FIRST minimum of the final face list.
// sup_a (read only by the final contact point) is kept in a global scratch column per lane, not in the slot.
#pragma once
#include "epa.cuh"

namespace b3 {

constexpr int TIER_POOL_BYTES = 37376;   // 6 x (pool + 400 B of tables + 1 KB system reservation) <= 228 KB per SM
constexpr int TIER_WARPS_PER_SM = 6;
constexpr uint32_t FULL = 0xFFFFFFFFu;

template <int LPS> struct TierCaps;
template <> struct TierCaps<1> { static constexpr int VCAP = 22, FCAP = 40, ECAP = 24, VM_BYTES = 16; };
template <> struct TierCaps<2> { static constexpr int VCAP = 42, FCAP = 86, ECAP = 24, VM_BYTES = 16; };
template <> struct TierCaps<4> { static constexpr int VCAP = 83, FCAP = 178, ECAP = 24, VM_BYTES = 32; };
template <> struct TierCaps<8> { static constexpr int VCAP = EPA_VCAP, FCAP = EPA_FCAP, ECAP = 64, VM_BYTES = 32; };

// Byte offsets inside a slot. Header: pv.xyz (the new support point), |pv| + running max |vertex| per axis (the
// visibility filter's magnitude bound), 2 spare words.
template <int LPS>
struct TierLayout {
    using C = TierCaps<LPS>;
    static constexpr int SLOT = TIER_POOL_BYTES / (32 / LPS);
    static constexpr int OFF_VM = 32;
    static constexpr int OFF_ED = OFF_VM + C::VM_BYTES;
    static constexpr int OFF_FN = OFF_ED + ((2 * C::ECAP + 15) / 16) * 16;
    static constexpr int OFF_FI = OFF_FN + 16 * C::FCAP;
    static constexpr int OFF_VV = OFF_FI + 4 * C::FCAP;
    static constexpr int END = OFF_VV + 12 * C::VCAP;
    static_assert(END <= SLOT, "tier slot overflows its share of the pool");
    static_assert(SLOT % 16 == 0, "slots must keep float4 alignment");
    static_assert(C::VM_BYTES * 8 >= ((C::FCAP + 31) / 32) * 32, "visibility mask too small");
};

// Where a stage's work comes from. Static stage: positions [lo, hi) of the bin-ordered queue (segment tables in
// shared memory map a position to a record). Promotion stage: a list of record indices that other warps are still
// appending to; entries are preset to 0xFFFFFFFF and written after their position was reserved.
struct TierSource {
    bool promo;
    uint32_t lo, hi;                 // static: position range
    uint32_t* cursor;                // static: relative cursor; promo: absolute cursor
    const uint32_t* count;           // promo: entries appended so far
    const uint32_t* list;            // promo: record indices
    const uint32_t* s_pref;          // static: segment tables (shared memory)
    const uint32_t* s_base;
};

struct TierShared {
    const ShapeSetDev* S;
    const uint2* pairs;
    SettingsDev cfg;
    const float4* q_recs;            // GJK -> EPA records
    uint32_t* promo_count;           // [4]: entries appended to the promotion list of tier t (t = 1..3)
    uint32_t* promo_list[4];
    uint32_t* done;                  // pairs finished (final status written) by this kernel
    float4* va;                      // this warp's sup_a scratch: [vertex][lane]
    unsigned char* ovf_scratch;
    float4* contacts;
    uint8_t* status;
    unsigned char* pool;             // this warp's shared-memory pool
};


// Runs one stage with LPS lanes per slot until the source has nothing more to hand out and every slot of this warp
// is finished. Returns true if the warp processed at least one pair. All 32 lanes call this convergently.
template <int LPS>
B3_DEV bool tier_run(const TierSource& src, const TierShared& sh) {
    using C = TierCaps<LPS>;
    using Lay = TierLayout<LPS>;
    constexpr int TIER = LPS == 1 ? 0 : LPS == 2 ? 1 : LPS == 4 ? 2 : 3;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t sub = lane % LPS;
    const uint32_t own_lane = lane - sub;
    const bool owner = sub == 0;
    unsigned char* slot = sh.pool + (lane / LPS) * Lay::SLOT;
    float* hdr = reinterpret_cast<float*>(slot);
    uint8_t* vm8 = slot + Lay::OFF_VM;
    uint32_t* vm32 = reinterpret_cast<uint32_t*>(slot + Lay::OFF_VM);
    PolytopeStore P;
    P.vv = reinterpret_cast<float*>(slot + Lay::OFF_VV);
    P.va = nullptr;
    P.fn = reinterpret_cast<float4*>(slot + Lay::OFF_FN);
    P.fi = reinterpret_cast<uint32_t*>(slot + Lay::OFF_FI);
    P.ed = reinterpret_cast<uint16_t*>(slot + Lay::OFF_ED);
    P.fcap = C::FCAP; P.ecap = C::ECAP; P.estride = 1;
    float4* va = sh.va + lane;  // vertex i of this lane's pair: va[i * 32]

    bool active = false, exhausted = false, any_work = false;
    Shape L, R;
    int nv = 0, nf = 0, best = 0;
    uint32_t it = 0, p = 0, rec = 0;
    V3 vmax = v3(0.f, 0.f, 0.f);
    const float tolerance = sh.cfg.epa_tolerance;
    const uint32_t max_iterations = sh.cfg.max_iterations;

    for (;;) {
        // ---- refill the idle slots (one cursor update per warp per trip) ----------------------------------
        const uint32_t need = __ballot_sync(FULL, owner && !active && !exhausted);
        if (need) {
            const uint32_t want = (uint32_t)__popc(need);
            const uint32_t leader = __ffs(need) - 1;
            uint32_t base = 0, got = 0;
            if (lane == leader) {
                if (!src.promo) {
                    base = atomicAdd(src.cursor, want);
                    const uint32_t span = src.hi - src.lo;
                    got = base >= span ? 0u : min(want, span - base);
                } else {
                    // never reserve positions that have not been appended yet: they would be skipped for good
                    for (;;) {
                        const uint32_t cur = ld_volatile_u32(src.cursor);
                        const uint32_t cnt = ld_volatile_u32(src.count);
                        got = cnt > cur ? min(want, cnt - cur) : 0u;
                        if (got == 0u) break;
                        if (atomicCAS(src.cursor, cur, cur + got) == cur) { base = cur; break; }
                    }
                }
            }
            base = __shfl_sync(FULL, base, leader);
            got = __shfl_sync(FULL, got, leader);
            if (owner && !active && !exhausted) {
                const uint32_t rank = (uint32_t)__popc(need & ((1u << lane) - 1u));
                if (rank < got) {
                    if (!src.promo) {
                        const uint32_t pos = src.lo + base + rank;
                        int seg = 0;
                        while (pos >= src.s_pref[seg + 1]) ++seg;
                        rec = src.s_base[seg] + (pos - src.s_pref[seg]);
                    } else {
                        const uint32_t* e = src.list + base + rank;
                        while ((rec = ld_volatile_u32(e)) == 0xFFFFFFFFu) __nanosleep(100);
                    }
                    const float4* r = sh.q_recs + (size_t)rec * 7;
                    float4 a = __ldcg(r), b = __ldcg(r + 1), c = __ldcg(r + 2), d = __ldcg(r + 3), e = __ldcg(r + 4), f = __ldcg(r + 5), g = __ldcg(r + 6);
                    p = __float_as_uint(g.x);
                    uint2 pr = __ldg(sh.pairs + p);
                    load_pair_shape(L, *sh.S, pr.x);
                    load_pair_shape(R, *sh.S, pr.y);
                    // Polytope::new (epa3d.rs:103-109) + Face::new (:172-179)
                    const V3 sv0 = v3(a.x, a.y, a.z), sv1 = v3(b.z, b.w, c.x), sv2 = v3(d.x, d.y, d.z), sv3 = v3(e.z, e.w, f.x);
                    st3(P.vv, 0, sv0); st3(P.vv, 1, sv1); st3(P.vv, 2, sv2); st3(P.vv, 3, sv3);
                    va[0 * 32] = make_float4(a.w, b.x, b.y, 0.f);
                    va[1 * 32] = make_float4(c.y, c.z, c.w, 0.f);
                    va[2 * 32] = make_float4(d.w, e.x, e.y, 0.f);
                    va[3 * 32] = make_float4(f.y, f.z, f.w, 0.f);
                    vmax = v3(fmaxf(fmaxf(fabsf(sv0.x), fabsf(sv1.x)), fmaxf(fabsf(sv2.x), fabsf(sv3.x))),
                              fmaxf(fmaxf(fabsf(sv0.y), fabsf(sv1.y)), fmaxf(fabsf(sv2.y), fabsf(sv3.y))),
                              fmaxf(fmaxf(fabsf(sv0.z), fabsf(sv1.z)), fmaxf(fabsf(sv2.z), fabsf(sv3.z))));
                    ClosestFace c0;
                    c0.reset();
                    c0.offer(make_face(P, 0, 3, 2, 1), 0);
                    c0.offer(make_face(P, 1, 3, 1, 0), 1);
                    c0.offer(make_face(P, 2, 3, 0, 2), 2);
                    c0.offer(make_face(P, 3, 2, 0, 1), 3);
                    best = c0.finish(P);
                    nv = 4; nf = 4; it = 1;
                    active = true;
                    any_work = true;
                } else {
                    exhausted = true;
                }
            }
        }
        if (!__any_sync(FULL, active)) break;

        // ---- owner: closest face, support point, convergence test (epa3d.rs:33-51) -------------------------
        bool scanning = false, finished = false, promote = false;
        V3 pv = v3(0.f, 0.f, 0.f);
        if (owner && active) {
            const float4 bf = P.fn[best];
            const V3 normal = v3(bf.x, bf.y, bf.z);
            V3 pa;
            minkowski_support<false>(L, R, normal, pv, pa);
            const float d = dot(pv, normal);
            if (d - bf.w < tolerance || it >= max_iterations) {
                // contact() / point() (epa3d.rs:70-94)
                const uint32_t idx = P.fi[best];
                const uint32_t i0 = idx & 0xFF, i1 = (idx >> 8) & 0xFF, i2 = (idx >> 16) & 0xFF;
                float u, v, w;
                barycentric_vector(normal * bf.w, ld3(P.vv, i0), ld3(P.vv, i1), ld3(P.vv, i2), u, v, w);
                const float4 a0 = va[i0 * 32], a1 = va[i1 * 32], a2 = va[i2 * 32];
                const V3 point = v3(a0.x, a0.y, a0.z) * u + v3(a1.x, a1.y, a1.z) * v + v3(a2.x, a2.y, a2.z) * w;
                sh.contacts[2 * (size_t)p] = make_float4(normal.x, normal.y, normal.z, bf.w);
                sh.contacts[2 * (size_t)p + 1] = make_float4(point.x, point.y, point.z, 0.f);
                sh.status[p] = ST_OK;
                active = false;
                finished = true;
            } else if (nv >= C::VCAP) {
                active = false;
                promote = true;
            } else {
                hdr[0] = pv.x; hdr[1] = pv.y; hdr[2] = pv.z;
                hdr[3] = fabsf(pv.x) + vmax.x; hdr[4] = fabsf(pv.y) + vmax.y; hdr[5] = fabsf(pv.z) + vmax.z;
                st3(P.vv, nv, pv);
                va[nv * 32] = make_float4(pa.x, pa.y, pa.z, 0.f);
                scanning = true;
            }
        }
        __syncwarp();

        // ---- all lanes of the slot: visibility of every face (pass A of epa.cuh), 8 faces per lane per step ----
        const bool g_scan = __shfl_sync(FULL, scanning ? 1 : 0, own_lane) != 0;
        const int g_nf = __shfl_sync(FULL, nf, own_lane);
        ClosestFace cf;
        cf.reset();
        if (g_scan) {
            const V3 gp = v3(hdr[0], hdr[1], hdr[2]);
            const V3 ap = v3(hdr[3], hdr[4], hdr[5]);
            const int nbytes = ((g_nf + 31) >> 5) << 2;  // whole mask words: bits at or beyond nf must read 0
            for (int b = (int)sub; b < nbytes; b += LPS) {
                const int k0 = b << 3;
                uint32_t m = 0;
                if (k0 < g_nf) {
                    float4 fb[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) fb[j] = (k0 + j < g_nf) ? P.fn[k0 + j] : make_float4(0.f, 0.f, 0.f, 0.f);
                    uint32_t unsure = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int k = k0 + j;
                        if (k < g_nf) {
                            const float4 f = fb[j];
                            const float a2 = dot(v3(f.x, f.y, f.z), gp) - f.w;
                            const float S = (fabsf(f.x) * ap.x + fabsf(f.y) * ap.y) + fabsf(f.z) * ap.z;
                            const bool sure = fabsf(a2) > S * 9.5367431640625e-07f + 1e-30f;  // false for NaN/inf too
                            const bool vis = a2 > 0.f;
                            m |= ((sure && vis) ? 1u : 0u) << j;
                            unsure |= (sure ? 0u : 1u) << j;
                            if (sure && !vis) cf.offer(f.w, k);
                        }
                    }
                    while (unsure) {  // the reference's expression as written (see epa.cuh)
                        const int j = __ffs(unsure) - 1;
                        unsure &= unsure - 1;
                        const int k = k0 + j;
                        const float4 f = P.fn[k];
                        const V3 v0 = ld3(P.vv, P.fi[k] & 0xFF);
                        if (dot(v3(f.x, f.y, f.z), gp - v0) > 0.f) m |= 1u << j;
                        else cf.offer_any(f.w, k);
                    }
                }
                vm8[b] = (uint8_t)m;
            }
        }
        __syncwarp();
#pragma unroll
        for (int o = LPS / 2; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(FULL, cf.d, o);
            const int oi = __shfl_xor_sync(FULL, cf.idx, o);
            cf.offer_any(od, oi);
        }

        // ---- owner: the order-dependent swap_remove walk (pass B of epa.cuh) --------------------------------
        int ne = 0;
        if (scanning) {
            bool ok = true;
            unsigned long long seen = 0ull;
            const int nchunks = (nf + 31) >> 5;
            for (int k = 0;;) {
                int c = k >> 5;
                uint32_t w = (c < nchunks) ? (vm32[c] >> (k & 31)) : 0u;
                while (w == 0u && ++c < nchunks) { k = c << 5; w = vm32[c]; }
                if (w == 0u) break;
                k += __ffs(w) - 1;
                const uint32_t idx = P.fi[k];
                --nf;  // swap_remove(k): the last face (with its flag) moves into slot k
                const uint32_t tail_word = vm32[nf >> 5];
                const uint32_t last_bit = (tail_word >> (nf & 31)) & 1u;
                vm32[nf >> 5] = tail_word & ~(1u << (nf & 31));
                if (k != nf) {
                    const float4 lf = P.fn[nf];
                    const uint32_t li = P.fi[nf];
                    P.fn[k] = lf;
                    P.fi[k] = li;
                    vm32[k >> 5] = (vm32[k >> 5] & ~(1u << (k & 31))) | (last_bit << (k & 31));
                    if (!last_bit) cf.moved(lf.w, k);
                }
                const uint32_t i0 = idx & 0xFF, i1 = (idx >> 8) & 0xFF, i2 = (idx >> 16) & 0xFF;
                ok = ok && remove_or_add_edge(P.ed, 1, ne, C::ECAP, i0, i1, seen);
                ok = ok && remove_or_add_edge(P.ed, 1, ne, C::ECAP, i1, i2, seen);
                ok = ok && remove_or_add_edge(P.ed, 1, ne, C::ECAP, i2, i0, seen);
            }
            if (!ok || nf + ne > C::FCAP) {
                scanning = false;
                active = false;
                promote = true;
            }
        }
        __syncwarp();

        // ---- all lanes of the slot: the new faces (pass C of epa.cuh) ---------------------------------------
        const bool g_add = __shfl_sync(FULL, scanning ? 1 : 0, own_lane) != 0;
        const int g_ne = __shfl_sync(FULL, ne, own_lane);
        const int g_base = __shfl_sync(FULL, nf, own_lane);
        const uint32_t g_n = (uint32_t)__shfl_sync(FULL, nv, own_lane);
        if (!owner) cf.reset();  // the owner's running minimum already contains every lane's pass-A candidates
        if (g_add) {
            for (int e = (int)sub; e < g_ne; e += LPS) {
                const uint32_t ab = P.ed[e];
                cf.offer(make_face(P, g_base + e, g_n, ab & 0xFF, (ab >> 8) & 0xFF), g_base + e);
            }
        }
        __syncwarp();
#pragma unroll
        for (int o = LPS / 2; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(FULL, cf.d, o);
            const int oi = __shfl_xor_sync(FULL, cf.idx, o);
            cf.offer_any(od, oi);
        }
        if (scanning) {
            nf += ne;
            // running max |coordinate| over the polytope's vertices; a NaN coordinate poisons it so the filter never fires
            vmax = v3(pv.x != pv.x ? pv.x : fmaxf(vmax.x, fabsf(pv.x)), pv.y != pv.y ? pv.y : fmaxf(vmax.y, fabsf(pv.y)),
                      pv.z != pv.z ? pv.z : fmaxf(vmax.z, fabsf(pv.z)));
            ++nv;
            best = cf.finish(P);
            it += 1;
        }

        // ---- promotions and the finished-pair counter ---------------------------------------------------------
        if (promote) {
            if (TIER < 3) {
                const uint32_t pos = atomicAdd(sh.promo_count + TIER + 1, 1u);
                sh.promo_list[TIER + 1][pos] = rec;
                __threadfence();
            } else {
                // beyond the last shared-memory tier: the overflow tier redoes the pair (epa_big.cuh)
                sh.contacts[2 * (size_t)p] = make_float4(0.f, 0.f, 0.f, 0.f);
                sh.contacts[2 * (size_t)p + 1] = make_float4(0.f, 0.f, 0.f, 0.f);
                sh.status[p] = ST_CAPACITY;
                overflow_append(sh.ovf_scratch, p);
                finished = true;
            }
        }
        const uint32_t fin = __ballot_sync(FULL, finished);
        if (fin && lane == 0) atomicAdd(sh.done, (uint32_t)__popc(fin));
    }
    return __any_sync(FULL, any_work);
}

}  // namespace b3
