// Overflow tier of EPA: one WARP per pair, polytope in a per-warp global-memory slab, for the rare pairs whose
// polytope outgrows the fast tier (epa.cuh). Near-degenerate input (typically a cylinder cap: its rim supports are
// coplanar, so face visibility is rounding noise) makes the reference build pinched polytopes whose face count
// grows multiplicatively; such a pair costs the reference O(faces^2) per iteration in remove_or_add_edge's
// linear search. Here the same ordered-list semantics (epa/epa3d.rs:121-149, :183-190) are kept exactly, but
//   * the O(faces) passes (closest face, visibility) are strided over the 32 lanes,
//   * the horizon list uses per-key FIFO queues (key = directed edge) so "find the first reversed duplicate" is
//     O(1): entries with one key are appended in list order, so the queue head IS the first match; removed entries
//     are tombstoned and the survivors compacted in list order afterwards — the same list the reference builds,
//   * the order-dependent swap_remove walk runs on lane 0 over precomputed visibility flags.
// A polytope that would exceed EPA_BIG_FCAP faces ends with ST_CAPACITY (the CPU oracle applies the same limit).
#pragma once
#include "epa.cuh"

namespace b3 {

constexpr int EPA_BIG_FCAP = 1024;             // == orc::EPA_FACE_LIMIT
constexpr int EPA_BIG_EAPP = 3 * EPA_BIG_FCAP;  // most edges one add() can append (3 per removed face)
constexpr int EPA_KEYS = 128 * 128;             // directed-edge key space: a | b << 7, vertex indices < 104
constexpr uint16_t EPA_NIL = 0xFFFFu;

struct BigStore {
    float4* fn;      // EPA_BIG_FCAP
    uint32_t* fi;    // EPA_BIG_FCAP
    uint32_t* vm;    // EPA_BIG_FCAP / 32 : visibility bitmask
    float* vv;       // EPA_VCAP * 3
    float* va;       // EPA_VCAP * 3
    uint16_t* ed;    // EPA_BIG_EAPP : appended edges (key), tombstoned via live[]
    uint16_t* nxt;   // EPA_BIG_EAPP : next entry with the same key
    uint8_t* live;   // EPA_BIG_EAPP
    uint16_t* head;  // EPA_KEYS
    uint16_t* tail;  // EPA_KEYS
};

// Faces live in a per-warp global-memory slab; everything the serial part of add() chases pointers through (edge
// queues, key tables, visibility words, vertices) lives in the block's shared memory (one warp per block).
constexpr size_t EPA_BIG_SLAB_BYTES = (size_t)EPA_BIG_FCAP * (16 + 4);
constexpr size_t EPA_BIG_SMEM_BYTES = (size_t)EPA_VCAP * 3 * 4 * 2 + (size_t)EPA_BIG_EAPP * (2 + 2 + 1) + (size_t)EPA_KEYS * 2 * 2 +
                                      (size_t)EPA_BIG_FCAP / 8 + 64;

B3_DEV BigStore carve_big_store(unsigned char* slab, unsigned char* smem) {
    BigStore B;
    B.fn = reinterpret_cast<float4*>(slab);   slab += (size_t)EPA_BIG_FCAP * 16;
    B.fi = reinterpret_cast<uint32_t*>(slab);
    B.vv = reinterpret_cast<float*>(smem);    smem += (size_t)EPA_VCAP * 12;
    B.va = reinterpret_cast<float*>(smem);    smem += (size_t)EPA_VCAP * 12;
    B.vm = reinterpret_cast<uint32_t*>(smem); smem += (size_t)EPA_BIG_FCAP / 8;
    B.ed = reinterpret_cast<uint16_t*>(smem); smem += (size_t)EPA_BIG_EAPP * 2;
    B.nxt = reinterpret_cast<uint16_t*>(smem); smem += (size_t)EPA_BIG_EAPP * 2;
    B.head = reinterpret_cast<uint16_t*>(smem); smem += (size_t)EPA_KEYS * 2;
    B.tail = reinterpret_cast<uint16_t*>(smem); smem += (size_t)EPA_KEYS * 2;
    B.live = smem;
    return B;
}

B3_DEV void big_make_face(const BigStore& B, int slot, uint32_t a, uint32_t b, uint32_t c) {
    V3 va = ld3(B.vv, a);
    V3 ab = ld3(B.vv, b) - va;
    V3 ac = ld3(B.vv, c) - va;
    V3 n = normalize(cross(ab, ac));
    B.fn[slot] = make_float4(n.x, n.y, n.z, dot(n, va));
    B.fi[slot] = a | (b << 8) | (c << 16);
}

// remove_or_add_edge with FIFO queues per key (lane 0 only). Returns false on append overflow (cannot happen
// while faces <= EPA_BIG_FCAP).
B3_DEV bool big_edge_op(const BigStore& B, int& napp, uint32_t a, uint32_t b) {
    const uint32_t rev = b | (a << 7);
    uint16_t h = B.head[rev];
    if (h != EPA_NIL) {                       // first (oldest) live entry equal to (b, a): remove it
        B.live[h] = 0;
        uint16_t nx = B.nxt[h];
        B.head[rev] = nx;
        if (nx == EPA_NIL) B.tail[rev] = EPA_NIL;
        return true;
    }
    if (napp >= EPA_BIG_EAPP) return false;
    const uint32_t key = a | (b << 7);
    const uint16_t pos = (uint16_t)napp++;
    B.ed[pos] = (uint16_t)key;
    B.nxt[pos] = EPA_NIL;
    B.live[pos] = 1;
    uint16_t t = B.tail[key];
    if (t == EPA_NIL) B.head[key] = pos; else B.nxt[t] = pos;
    B.tail[key] = pos;
    return true;
}

// EPA3::process for one pair, executed by all 32 lanes of a warp (convergent). head/tail must be all-NIL on entry
// and are all-NIL again on return.
B3_DEV uint8_t epa_process_big(const Shape& L, const Shape& R, const Simplex<true>& s, float tolerance, uint32_t max_iterations,
                               const BigStore& B, ContactOut& out) {
    const uint32_t lane = threadIdx.x & 31u;
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { st3(B.vv, k, s.v[k]); st3(B.va, k, s.a[k]); }
    }
    __syncwarp();
    if (lane == 0) {
        big_make_face(B, 0, 3, 2, 1);
        big_make_face(B, 1, 3, 1, 0);
        big_make_face(B, 2, 3, 0, 2);
        big_make_face(B, 3, 2, 0, 1);
    }
    __syncwarp();
    int nv = 4, nf = 4;
    uint32_t i = 1;
    for (;;) {
        // closest_face_to_origin: sequentially `best = 0; for k >= 1: if d[k] < d[best] best = k`. If d[0] is NaN
        // nothing ever compares below it and face 0 stays; otherwise it is the lowest-index minimum over non-NaN d.
        float d0 = B.fn[0].w;
        int best = 0;
        if (d0 == d0) {
            float lb = INFINITY; int li = 0x7FFFFFFF;
            for (int k = lane; k < nf; k += 32) {
                float dk = B.fn[k].w;
                if (dk < lb) { lb = dk; li = k; }
            }
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                float ob = __shfl_xor_sync(0xFFFFFFFFu, lb, off);
                int oi = __shfl_xor_sync(0xFFFFFFFFu, li, off);
                if (ob < lb || (ob == lb && oi < li)) { lb = ob; li = oi; }
            }
            // d[0] is not NaN, so it took part (lane 0, k = 0) unless it is +inf; an all-+inf list keeps face 0
            if (li != 0x7FFFFFFF) best = li;
        }
        float4 bf = B.fn[best];
        V3 normal = v3(bf.x, bf.y, bf.z);
        V3 pv, pa;
        minkowski_support<true>(L, R, normal, pv, pa);
        float d = dot(pv, normal);
        if (d - bf.w < tolerance || i >= max_iterations) {
            uint32_t idx = B.fi[best];
            uint32_t i0 = idx & 0xFF, i1 = (idx >> 8) & 0xFF, i2 = (idx >> 16) & 0xFF;
            float u, v, w;
            barycentric_vector(normal * bf.w, ld3(B.vv, i0), ld3(B.vv, i1), ld3(B.vv, i2), u, v, w);
            out.normal = normal;
            out.depth = bf.w;
            out.point = ld3(B.va, i0) * u + ld3(B.va, i1) * v + ld3(B.va, i2) * w;
            return ST_OK;
        }
        // visibility of every face w.r.t. the new point (epa3d.rs:126-129): 32 faces per step, one ballot word each
        const int nwords = (nf + 31) >> 5;
        for (int c = 0; c < nwords; ++c) {
            const int k = (c << 5) + lane;
            bool vis = false;
            if (k < nf) {
                float4 f = B.fn[k];
                V3 v0 = ld3(B.vv, B.fi[k] & 0xFF);
                vis = dot(v3(f.x, f.y, f.z), pv - v0) > 0.f;
            }
            const uint32_t w = __ballot_sync(0xFFFFFFFFu, vis);
            if (lane == 0) B.vm[c] = w;
        }
        __syncwarp();
        // lane 0: the order-dependent part — swap_remove walk (driven by ffs over the mask) + horizon edge list
        int napp = 0, nlive = 0, ok = 1;
        if (lane == 0) {
            for (int k = 0;;) {
                int c = k >> 5;
                uint32_t w = (c < nwords) ? (B.vm[c] >> (k & 31)) : 0u;
                while (w == 0u && ++c < nwords) { k = c << 5; w = B.vm[c]; }
                if (w == 0u) break;
                k += __ffs(w) - 1;
                const uint32_t idx = B.fi[k];
                --nf;  // swap_remove(k): the last face (with its flag) moves into slot k and is examined next
                const uint32_t last_bit = (B.vm[nf >> 5] >> (nf & 31)) & 1u;
                B.vm[nf >> 5] &= ~(1u << (nf & 31));
                if (k != nf) {
                    B.fn[k] = B.fn[nf]; B.fi[k] = B.fi[nf];
                    B.vm[k >> 5] = (B.vm[k >> 5] & ~(1u << (k & 31))) | (last_bit << (k & 31));
                }
                uint32_t i0 = idx & 0xFF, i1 = (idx >> 8) & 0xFF, i2 = (idx >> 16) & 0xFF;
                ok = ok && big_edge_op(B, napp, i0, i1);
                ok = ok && big_edge_op(B, napp, i1, i2);
                ok = ok && big_edge_op(B, napp, i2, i0);
            }
            // compact the surviving edges in list order; reset the queues of every key that was touched
            for (int pos = 0; pos < napp; ++pos) {
                uint16_t key = B.ed[pos];
                B.head[key] = EPA_NIL; B.tail[key] = EPA_NIL;
                if (B.live[pos]) B.ed[nlive++] = key;
            }
        }
        nf = __shfl_sync(0xFFFFFFFFu, nf, 0);
        nlive = __shfl_sync(0xFFFFFFFFu, nlive, 0);
        ok = __shfl_sync(0xFFFFFFFFu, ok, 0);
        __syncwarp();
        if (!ok || nv >= EPA_VCAP || nf + nlive > EPA_BIG_FCAP) { out.depth = (float)nv; return ST_CAPACITY; }  // depth: how early it outgrew this tier (the next tier's schedule)
        const int n = nv;
        if (lane == 0) { st3(B.vv, n, pv); st3(B.va, n, pa); }
        ++nv;
        __syncwarp();
        for (int e = lane; e < nlive; e += 32) {
            uint32_t key = B.ed[e];
            big_make_face(B, nf + e, (uint32_t)n, key & 0x7F, (key >> 7) & 0x7F);
        }
        nf += nlive;
        __syncwarp();
        i += 1;
    }
}

}  // namespace b3
