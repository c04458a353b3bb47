// Thread-per-pair EPA iteration with the memory round trips taken off the critical path ("epa_pipeline").
//
// What epa_iterate (epa.cuh) costs is not arithmetic: ncu shows the persistent EPA kernel issuing on 30 % of the cycles with
// half of every warp's stall time on `long_scoreboard` (profiles/r1c_c2_ncu_full.txt). One trip of EPA3::process chases about
// forty DEPENDENT global-memory round trips per lane — one per 8-face batch of the visibility scan (a lane's batch is one
// 128-byte line of its own record), three per face the walk removes (vertex indices, then the face moved in from the end of
// the list, then the stores the next face's loads may alias), one per new face — and with 16 warps per SM and records that do
// not all fit in L2 (the ball-like bins: 75 776 lanes x 2-4 KB) a trip takes ~100k cycles. The arithmetic, the three ordered
// lists and every tie-break are exactly epa_iterate's (same device functions, same order per pair: bit-identical results);
// what changes is WHEN memory is touched:
//   pass A  (optional, B3_PIPE_A) the face list streams through a per-lane shared-memory ring filled by cp.async;
//   pass B  the order-dependent swap_remove walk (epa3d.rs:121-149) runs on the visibility bitmask ALONE and only writes
//           two short shared-memory lists — the original slots of the removed faces in removal order, and the final
//           (destination <- source) moves. A tail slot is always still original when it is taken as a source and a hole is
//           examined right after it is filled, so identities are known without touching the face arrays. Then the vertex
//           indices of the removed faces are loaded four at a time (independent loads, parked in shared memory) for the
//           horizon edge list, and the moves are done four at a time (loads first, then stores: sources lie beyond the
//           final length, destinations below it);
//   pass C  the old vertices of four new faces are loaded together (parked in shared memory), then the faces are computed
//           one by one; the new vertex comes from registers. Vertices are float4 here: one load each instead of three.
// A warp issues in order, so loads only overlap when they are issued back to back before the first use: that is what the
// four-wide load groups are for, while everything that consumes them stays in rolled loops — the first version of this
// file unrolled the consumers too, grew the kernel from 3 900 to 5 300 instructions and lost more to instruction fetch
// (ncu: stall_no_instruction 0.2 -> 4.9 per issue) than the shorter memory chains gained (long_scoreboard 6.1 -> 3.8).
#pragma once
#include "epa.cuh"

namespace b3 {

B3_DEV void cp_async16(void* smem, const void* gmem) {
    const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
B3_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
B3_DEV void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#ifndef B3_PIPE_A
#define B3_PIPE_A 0  // pass A through the cp.async ring (0: epa_iterate's direct 8-face batches)
#endif
#ifndef B3_EPA_RING
#define B3_EPA_RING 2  // lines of 8 faces in flight per lane (shared memory: RING x 8 x 16 B per thread)
#endif
constexpr int EPA_RING = B3_EPA_RING;
constexpr int EPA_PCAP = 24;  // removed faces / moves per Polytope::add the shared-memory lists hold; more -> overflow tier

// slab record of the pipelined form: faces | face vertex indices | vertices (float4) | sup_a (float4)
constexpr size_t EPA_PIPE_SLAB_BYTES = (size_t)EPA_FCAP * 16 + (size_t)EPA_FCAP * 4 + 2 * (size_t)EPA_VCAP * 16;

struct PipeStore {
    float4* vv4;     // global: SupportPoint.v of the polytope's vertices (xyz, w unused)
    float4* va4;     // global: SupportPoint.sup_a
    float4* ring;    // shared (B3_PIPE_A): this thread's ring, line slot s, face j at ring[(s * 8 + j) * stride]
    uint32_t* proc;  // shared: original slots of the removed faces in removal order, then their vertex indices; entry i at proc[i * stride]
    uint16_t* mov;   // shared: moves dst | src << 8; entry i at mov[i * stride]
    float* stage;    // shared: 4 x 6 floats, the old vertices of the next four new faces; word w at stage[w * stride]
    int stride;      // blockDim.x
};

B3_DEV V3 ldv(const float4* p, int i) { const float4 q = p[i]; return v3(q.x, q.y, q.z); }
B3_DEV void stv(float4* p, int i, V3 v) { p[i] = make_float4(v.x, v.y, v.z, 0.f); }

B3_DEV float make_face_pipe(const PolytopeStore& P, const PipeStore& Q, int slot, uint32_t a, uint32_t b, uint32_t c) {
    const V3 va = ldv(Q.vv4, a);
    const V3 ab = ldv(Q.vv4, b) - va;
    const V3 ac = ldv(Q.vv4, c) - va;
    const V3 n = normalize(cross(ab, ac));
    const float dist = dot(n, va);
    P.fn[slot] = make_float4(n.x, n.y, n.z, dist);
    P.fi[slot] = a | (b << 8) | (c << 16);
    return dist;
}

// Polytope::new (epa3d.rs:103-109): epa_init with float4 vertices
B3_DEV void epa_init_pipe(const Simplex<true>& s, const PolytopeStore& P, const PipeStore& Q, int& nv, int& nf, uint32_t& it, V3& vmax, int& best) {
    vmax = v3(0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        stv(Q.vv4, k, s.v[k]); stv(Q.va4, k, s.a[k]);
        vmax = v3(fmaxf(vmax.x, fabsf(s.v[k].x)), fmaxf(vmax.y, fabsf(s.v[k].y)), fmaxf(vmax.z, fabsf(s.v[k].z)));
    }
    ClosestFace cf;
    cf.reset();
    cf.offer(make_face_pipe(P, Q, 0, 3, 2, 1), 0);
    cf.offer(make_face_pipe(P, Q, 1, 3, 1, 0), 1);
    cf.offer(make_face_pipe(P, Q, 2, 3, 0, 2), 2);
    cf.offer(make_face_pipe(P, Q, 3, 2, 0, 1), 3);
    best = cf.finish(P);
    nv = 4; nf = 4; it = 1;
}

#if B3_PIPE_A
// request line `l` (faces 8l .. 8l+7 of this lane's record) into its ring slot; one commit group per line
B3_DEV void epa_request_line(const PolytopeStore& P, const PipeStore& Q, int l) {
    const float4* src = P.fn + 8 * l;
    float4* dst = Q.ring + (size_t)((l % EPA_RING) * 8) * Q.stride;
#pragma unroll
    for (int j = 0; j < 8; ++j) cp_async16(dst + (size_t)j * Q.stride, src + j);
    cp_async_commit();
}
#endif

// One trip of EPA3::process's loop; same contract as epa_iterate<false>.
B3_DEV uint8_t epa_iterate_pipe(const Shape& L, const Shape& R, float tolerance, uint32_t max_iterations, const PolytopeStore& P,
                                const PipeStore& Q, int& nv, int& nf, uint32_t& it, V3& vmax, int& best, ContactOut& out) {
#if B3_PIPE_A
    const int nlines = (nf + 7) >> 3;
    int issued = 0;
#pragma unroll
    for (int l = 0; l < EPA_RING; ++l)
        if (l < nlines) { epa_request_line(P, Q, l); ++issued; }
#endif
    float4 bf = P.fn[best];
    V3 normal = v3(bf.x, bf.y, bf.z);
    V3 pv, pa;
    minkowski_support<false>(L, R, normal, pv, pa);
    float d = dot(pv, normal);
    if (d - bf.w < tolerance || it >= max_iterations) {
#if B3_PIPE_A
        cp_async_wait<0>();  // nothing may still be landing in the ring when the lane's next pair reuses it
#endif
        uint32_t idx = P.fi[best];
        uint32_t i0 = idx & 0xFF, i1 = (idx >> 8) & 0xFF, i2 = (idx >> 16) & 0xFF;
        float u, v, w;
        barycentric_vector(normal * bf.w, ldv(Q.vv4, i0), ldv(Q.vv4, i1), ldv(Q.vv4, i2), u, v, w);
        out.normal = normal;
        out.depth = bf.w;
        out.point = ldv(Q.va4, i0) * u + ldv(Q.va4, i1) * v + ldv(Q.va4, i2) * w;
        return ST_OK;
    }

    // ---- pass A: visibility bitmask + running closest survivor (see epa_iterate for the exactness argument) ----
    uint32_t vm[EPA_FCAP / 32];
    ClosestFace cf;
    cf.reset();
    const float sure_bound = ((fabsf(pv.x) + vmax.x) + (fabsf(pv.y) + vmax.y) + (fabsf(pv.z) + vmax.z)) * 9.5367431640625e-07f + 1e-30f;
    const int nchunks = (nf + 31) >> 5;
    for (int c = 0; c < nchunks; ++c) {
        uint32_t m = 0, unsure = 0;
        const int k0 = c << 5;
        const int k1 = min(nf, k0 + 32);
        for (int kb = k0; kb < k1; kb += 8) {
            float4 fb[8];
#if B3_PIPE_A
            const int l = kb >> 3;
            // line l must have landed: at most (issued - l - 1) younger groups may still be in flight
            if (issued - l - 1 >= EPA_RING - 1) cp_async_wait<EPA_RING - 1>();
            else cp_async_wait<0>();
            const float4* line = Q.ring + (size_t)((l % EPA_RING) * 8) * Q.stride;
#pragma unroll
            for (int j = 0; j < 8; ++j) fb[j] = line[(size_t)j * Q.stride];
            if (issued < nlines) { epa_request_line(P, Q, issued); ++issued; }  // the slot just read is free again
#else
#pragma unroll
            for (int j = 0; j < 8; ++j) fb[j] = (kb + j < k1) ? P.fn[kb + j] : make_float4(0.f, 0.f, 0.f, 0.f);
#endif
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
        while (unsure) {
            const int b = __ffs(unsure) - 1;
            unsure &= unsure - 1;
            const int k = k0 + b;
            const float4 f = P.fn[k];
            V3 v0 = ldv(Q.vv4, P.fi[k] & 0xFF);
            if (dot(v3(f.x, f.y, f.z), pv - v0) > 0.f) m |= 1u << b;
            else cf.offer_any(f.w, k);
        }
        vm[c] = m;
    }

    // ---- pass B1: the walk, on the bitmask alone ----
    int nproc = 0, nmov = 0;
    int pend_dst = -1, pend_src = 0;
    for (int k = 0;;) {
        int c = k >> 5;
        uint32_t w = (c < nchunks) ? (vm[c] >> (k & 31)) : 0u;
        while (w == 0u && ++c < nchunks) { k = c << 5; w = vm[c]; }
        if (w == 0u) break;
        k += __ffs(w) - 1;
        // the face in slot k: the one a removal just moved in from the tail, or the slot's own
        const int orig = (k == pend_dst) ? pend_src : k;
        if (nproc < EPA_PCAP) Q.proc[(size_t)nproc * Q.stride] = (uint32_t)orig;
        ++nproc;
        --nf;  // swap_remove(k): the last face (with its flag) moves into slot k
        const uint32_t last_bit = (vm[nf >> 5] >> (nf & 31)) & 1u;
        vm[nf >> 5] &= ~(1u << (nf & 31));
        pend_dst = -1;
        if (k != nf) {
            vm[k >> 5] = (vm[k >> 5] & ~(1u << (k & 31))) | (last_bit << (k & 31));
            if (last_bit) { pend_dst = k; pend_src = nf; }  // visible too: examined (and removed) next, from slot k
            else {
                if (nmov < EPA_PCAP) Q.mov[(size_t)nmov * Q.stride] = (uint16_t)(k | (nf << 8));
                ++nmov;
            }
        }
    }
    if (nproc > EPA_PCAP || nmov > EPA_PCAP) return ST_CAPACITY;

    // ---- pass B2: vertex indices of the removed faces (four independent loads at a time), then their horizon edges ----
    for (int j0 = 0; j0 < nproc; j0 += 4) {
        uint32_t idx[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) idx[j] = (j0 + j < nproc) ? P.fi[Q.proc[(size_t)(j0 + j) * Q.stride]] : 0u;
#pragma unroll
        for (int j = 0; j < 4; ++j) if (j0 + j < nproc) Q.proc[(size_t)(j0 + j) * Q.stride] = idx[j];
    }
    int ne = 0;
    bool ok = true;
    unsigned long long seen = 0ull;
#pragma unroll 1
    for (int j = 0; j < nproc; ++j) {
        uint32_t idx = Q.proc[(size_t)j * Q.stride];
        // edges (i0,i1), (i1,i2), (i2,i0): one copy of the edge operation, the three index bytes rotated between calls
#pragma unroll 1
        for (int e = 0; e < 3; ++e) {
            const uint32_t a = idx & 0xFF, b = (idx >> 8) & 0xFF;
            ok = ok && remove_or_add_edge(P.ed, P.estride, ne, P.ecap, a, b, seen);
            idx = ((idx >> 8) & 0xFFFFu) | (a << 16);
        }
    }
    if (!ok || nv >= EPA_VCAP || nf + ne > P.fcap) return ST_CAPACITY;

    // ---- pass B3: the moves (sources >= the new length, destinations below it) ----
#pragma unroll 1
    for (int j0 = 0; j0 < nmov; j0 += 2) {
        float4 lf[2];
        uint32_t li[2];
        uint32_t mv[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            mv[j] = (j0 + j < nmov) ? (uint32_t)Q.mov[(size_t)(j0 + j) * Q.stride] : 0u;
            if (j0 + j < nmov) { lf[j] = P.fn[mv[j] >> 8]; li[j] = P.fi[mv[j] >> 8]; }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (j0 + j < nmov) {
                const int dst = (int)(mv[j] & 0xFF);
                P.fn[dst] = lf[j];
                P.fi[dst] = li[j];
                cf.moved(lf[j].w, dst);
            }
        }
    }

    // ---- pass C: the new vertex and the new faces ----
    const int n = nv;
    stv(Q.vv4, n, pv);
    stv(Q.va4, n, pa);
    vmax = v3(pv.x != pv.x ? pv.x : fmaxf(vmax.x, fabsf(pv.x)), pv.y != pv.y ? pv.y : fmaxf(vmax.y, fabsf(pv.y)),
              pv.z != pv.z ? pv.z : fmaxf(vmax.z, fabsf(pv.z)));
    ++nv;
#pragma unroll 1
    for (int e0 = 0; e0 < ne; e0 += 4) {
        {   // the two old vertices of up to four new faces: eight independent loads, parked in shared memory
            float4 pb[4], pc[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (e0 + j < ne) {
                    const uint32_t ab = P.ed[(e0 + j) * P.estride];
                    pb[j] = Q.vv4[ab & 0xFF];
                    pc[j] = Q.vv4[(ab >> 8) & 0xFF];
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (e0 + j < ne) {
                    float* sg = Q.stage + (size_t)(6 * j) * Q.stride;
                    sg[0] = pb[j].x; sg[Q.stride] = pb[j].y; sg[2 * Q.stride] = pb[j].z;
                    sg[3 * Q.stride] = pc[j].x; sg[4 * Q.stride] = pc[j].y; sg[5 * Q.stride] = pc[j].z;
                }
            }
        }
        const int e1 = min(ne, e0 + 4);
#pragma unroll 1
        for (int e = e0; e < e1; ++e) {
            const float* sg = Q.stage + (size_t)(6 * (e - e0)) * Q.stride;
            const uint32_t ab = P.ed[e * P.estride];
            // Face::new_impl (epa3d.rs:160-170) with the new vertex (index n) as the first: make_face's arithmetic
            const V3 fab = v3(sg[0], sg[Q.stride], sg[2 * Q.stride]) - pv;
            const V3 fac = v3(sg[3 * Q.stride], sg[4 * Q.stride], sg[5 * Q.stride]) - pv;
            const V3 fnrm = normalize(cross(fab, fac));
            const float dist = dot(fnrm, pv);
            const int slot = nf + e;
            P.fn[slot] = make_float4(fnrm.x, fnrm.y, fnrm.z, dist);
            P.fi[slot] = (uint32_t)n | ((ab & 0xFF) << 8) | (((ab >> 8) & 0xFF) << 16);
            cf.offer(dist, slot);
        }
    }
    nf += ne;
    best = cf.finish(P);
    it += 1;
    return EPA_CONTINUE;
}

}  // namespace b3
