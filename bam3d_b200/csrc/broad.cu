// Broad-phase kernels for sm_100a.
//  * SweepAndPrune3::find_collider_pairs (algorithm/broad_phase/sweep_prune.rs:69-128): 64-bit (min, max) keys on the
//    sweep axis, stable radix sort, a shared-memory tiled forward sweep with warp-aggregated compaction of the
//    overlapping pairs, a final sort of the pair list into the reference's emission order, and Variance3 (:166-216).
//  * A static LBVH that answers the DynamicBoundingVolumeTree queries (dbvt/mod.rs:481-515 with the visitors of
//    dbvt/visitor.rs and dbvt/util.rs). Query result SETS do not depend on the tree shape (every branch bound
//    contains its descendants and every visitor is monotone in the bound), and the reference's own tree shape is
//    randomised (rand::thread_rng, dbvt/mod.rs:901), so the device builds its own tree from the values' real bounds.
//  * ComputeBound<Aabb3> + Aabb::transform for packed shapes (world AABBs feeding the broad phase).
// HBM-bound integer/compare work: no tensor cores. Compiled with -fmad=false like the narrow phase.
#include <cub/cub.cuh>

#include "broad.h"
#include "narrow.h"
#include "support.cuh"

namespace b3 {

constexpr int BT = 128;  // threads per block
constexpr int NSM = 148;

B3_DEV float comp(float4 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }

// order-preserving f32 -> u32 (after canonicalising -0.0 to +0.0: partial_cmp treats them as Equal)
B3_DEV uint32_t ord_key(float f) {
    f = f + 0.0f;
    uint32_t b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// ------------------------------------------------------------------------------------------------------------
// sweep and prune
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT) sap_keys_kernel(const float* __restrict__ aabbs, uint32_t n, int axis,
                                                      unsigned long long* __restrict__ keys, uint32_t* __restrict__ vals) {
    uint32_t i = blockIdx.x * BT + threadIdx.x;
    if (i >= n) return;
    float mn = aabbs[6 * (size_t)i + axis], mx = aabbs[6 * (size_t)i + 3 + axis];
    // sort_by comparator (sweep_prune.rs:80-90): min on the sweep axis, then max; stable
    keys[i] = ((unsigned long long)ord_key(mn) << 32) | ord_key(mx);
    vals[i] = i;
}

__global__ void __launch_bounds__(BT) sap_gather_kernel(const float* __restrict__ aabbs, const uint32_t* __restrict__ perm, uint32_t n,
                                                        float4* __restrict__ lo, float4* __restrict__ hi, uint32_t* __restrict__ out_perm,
                                                        float* __restrict__ sorted6) {
    uint32_t k = blockIdx.x * BT + threadIdx.x;
    if (k >= n) return;
    uint32_t src = perm[k];
    const float* b = aabbs + 6 * (size_t)src;
    lo[k] = make_float4(b[0], b[1], b[2], 0.f);
    hi[k] = make_float4(b[3], b[4], b[5], 0.f);
    if (out_perm) out_perm[k] = src;
    if (sorted6) {
        float* o = sorted6 + 6 * (size_t)k;
        o[0] = b[0]; o[1] = b[1]; o[2] = b[2]; o[3] = b[3]; o[4] = b[4]; o[5] = b[5];
    }
}

// Number of (i, j) candidate tests the forward sweep would make: for every sorted slot i, the slots j > i whose
// sweep-axis min is below i's max (binary search in the sorted mins). Decides sweep vs BVH strategy.
__global__ void __launch_bounds__(BT) sap_candidates_kernel(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n, int axis,
                                                            unsigned long long* __restrict__ total) {
    uint32_t i = blockIdx.x * BT + threadIdx.x;
    unsigned long long c = 0;
    if (i < n) {
        const float amax = comp(__ldg(hi + i), axis);
        uint32_t a = i + 1, b = n;  // first j in (i, n) with min_j >= amax
        while (a < b) {
            uint32_t m = a + ((b - a) >> 1);
            if (comp(__ldg(lo + m), axis) < amax) a = m + 1; else b = m;
        }
        c = a - (i + 1);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, off);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(total, c);
}

// Forward sweep. Block b owns sorted slots i in [128 b, 128 b + 128); it streams the following slots j through shared
// memory one tile at a time (sweep-axis min, and the second axis' extent as a cheap filter) until the tile starts at
// or beyond the largest sweep-axis max of the block: with mins sorted, no later box can overlap any of the block's
// boxes on the sweep axis (Aabb3 overlap is strict: a.max > b.min, volume/aabb/aabb3.rs:237-244). The retain test of
// the reference's active list (max_i >= min_j, sweep_prune.rs:102-105) is implied by that strict overlap, so the pair
// set is exactly "all strictly overlapping (i < j)". Survivors of the filter run the full 3-axis test and are
// appended through one atomic per warp.
__global__ void __launch_bounds__(BT) sap_sweep_kernel(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n, int axis,
                                                       uint32_t j_begin, uint32_t j_end, int key_shift,
                                                       unsigned long long* __restrict__ pair_keys, unsigned long long capacity,
                                                       unsigned long long* __restrict__ count) {
    __shared__ float s_alo[BT], s_blo[BT], s_bhi[BT];
    __shared__ float s_red[BT / 32];
    const int a = axis, b = (axis + 1) % 3;
    const uint32_t i0 = blockIdx.x * BT, i = i0 + threadIdx.x;
    if (i0 + 1 >= j_end) return;  // shard [j_begin, j_end) of the pairs' higher slot: nothing above this block is in it
    const bool valid = i < n;
    float4 li = make_float4(0, 0, 0, 0), hi_i = make_float4(0, 0, 0, 0);
    if (valid) { li = __ldg(lo + i); hi_i = __ldg(hi + i); }
    const float amax = valid ? comp(hi_i, a) : -INFINITY;
    const float bmin = comp(li, b), bmax = comp(hi_i, b);
    // block-wide max of amax
    float m = amax;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = m;
    __syncthreads();
    float block_amax = s_red[0];
#pragma unroll
    for (int w = 1; w < BT / 32; ++w) block_amax = fmaxf(block_amax, s_red[w]);

    for (uint32_t j0 = max(i0 + 1, j_begin); j0 < j_end; j0 += BT) {
        const uint32_t j = j0 + threadIdx.x;
        if (j < j_end) {
            float4 lj = __ldg(lo + j), hj = __ldg(hi + j);
            s_alo[threadIdx.x] = comp(lj, a); s_blo[threadIdx.x] = comp(lj, b); s_bhi[threadIdx.x] = comp(hj, b);
        } else {
            s_alo[threadIdx.x] = INFINITY; s_blo[threadIdx.x] = INFINITY; s_bhi[threadIdx.x] = -INFINITY;
        }
        __syncthreads();
        if (!(s_alo[0] < block_amax)) break;  // uniform: sorted mins, nothing further can overlap
        if (valid) {
#pragma unroll 8
            for (int t = 0; t < BT; ++t) {
                const uint32_t jj = j0 + t;
                if (jj > i && s_alo[t] < amax && s_blo[t] < bmax && s_bhi[t] > bmin) {
                    float4 lj = __ldg(lo + jj), hj = __ldg(hi + jj);
                    // Discrete<Aabb3> for Aabb3 (aabb3.rs:237-244), self = box i, other = box j
                    bool hit = hi_i.x > lj.x && li.x < hj.x && hi_i.y > lj.y && li.y < hj.y && hi_i.z > lj.z && li.z < hj.z;
                    if (hit) {
                        const uint32_t act = __activemask();
                        const uint32_t lane = threadIdx.x & 31u;
                        const uint32_t leader = __ffs(act) - 1;
                        unsigned long long base = 0;
                        if (lane == leader) base = atomicAdd(count, (unsigned long long)__popc(act));
                        base = __shfl_sync(act, base, leader);
                        const unsigned long long slot = base + __popc(act & ((1u << lane) - 1u));
                        if (slot < capacity) pair_keys[slot] = ((unsigned long long)jj << key_shift) | i;
                    }
                }
            }
        }
        __syncthreads();
    }
}

// Variance3 (sweep_prune.rs:166-216): c = (min + max) / 2; csum += c; csumsq += c * c, accumulated SEQUENTIALLY in
// sorted order exactly as the reference's loop does (f32 addition is not associative, and the three axes of a
// uniform scene are near-ties, so a tree reduction could pick a different axis). The six running sums are six
// independent chains: block c of six walks the boxes in order for chain c (0..2 = csum.xyz, 3..5 = csumsq.xyz).
//
// A chain of n dependent FADDs is latency bound (4 cycles each: 14 ms for 4M boxes). The same bits come out of integer
// arithmetic, which IS associative: while the running sum `acc` stays inside one binade [2^e, 2^(e+1)) its ulp u is
// constant, acc = A * u with A an integer in [2^23, 2^24), and fl(acc + c) = (A + r) * u where r = c / u rounded to the
// nearest integer — unless c / u is exactly half-way (then the tie goes to the even RESULT, which depends on A) or the
// sum leaves the binade. So a tile of values is handled in parallel: every thread rounds its values to r_i with exact
// integer shifts (mantissa and exponent of c), the block adds up S = sum r_i and B = sum |r_i|, and if no value was a tie
// and A - B >= 2^23 + 1 and A + B <= 2^24 - 1 (every prefix stays inside the binade, with the half-ulp of slack that
// keeps the un-rounded prefix inside too) the new sum is (A + S) * u exactly. Any other tile — the first few, the ~25
// binade crossings, an exact tie (probability ~2^-22 per value), non-positive / non-finite sums — is replayed
// sequentially by one thread, as before.
constexpr int VAR_TILE = 4096;
constexpr int VAR_THREADS = 1024;  // 4 values per thread per tile: one LDG.128

// c / 2^k rounded to the nearest integer (round-half flagged as `tie`, not resolved). false: not representable here.
B3_DEV bool round_scaled(float c, int k, int& r, bool& tie) {
    const uint32_t b = __float_as_uint(c);
    const uint32_t ex = (b >> 23) & 0xFFu;
    uint32_t mant = b & 0x7FFFFFu;
    tie = false;
    r = 0;
    if (ex == 0xFFu) return false;  // inf / NaN
    int E;
    if (ex == 0) E = -149; else { mant |= 0x800000u; E = (int)ex - 150; }  // |c| = mant * 2^E
    if (mant == 0) return true;
    const int sh = E - k;
    uint32_t q;
    if (sh >= 0) {
        if (sh > 6) return false;  // |c| >= 2^30 ulps: certainly leaves the binade
        q = mant << sh;
    } else {
        const int s = -sh;
        if (s > 25) q = 0;
        else {
            q = mant >> s;
            const uint32_t rem = mant & ((1u << s) - 1u), half = 1u << (s - 1);
            if (rem > half) ++q;
            else if (rem == half) tie = true;
        }
    }
    r = (b >> 31) ? -(int)q : (int)q;
    return true;
}

// The six value streams of Variance3, planar: vals[chain * stride + i] = c (chains 0..2) or c * c (chains 3..5) of box i.
__global__ void __launch_bounds__(BT) sap_centres_kernel(const float4* __restrict__ lo, const float4* __restrict__ hi, uint32_t n, size_t stride,
                                                         float* __restrict__ vals) {
    const uint32_t i = blockIdx.x * BT + threadIdx.x;
    if (i >= n) return;
    const float4 a = __ldg(lo + i), b = __ldg(hi + i);
    const float cx = (a.x + b.x) / 2.0f, cy = (a.y + b.y) / 2.0f, cz = (a.z + b.z) / 2.0f;
    vals[i] = cx; vals[stride + i] = cy; vals[2 * stride + i] = cz;
    vals[3 * stride + i] = cx * cx; vals[4 * stride + i] = cy * cy; vals[5 * stride + i] = cz * cz;
}

__global__ void __launch_bounds__(VAR_THREADS) sap_variance_chain_kernel(const float* __restrict__ vals, uint32_t n, size_t stride,
                                                                         float* __restrict__ sums) {
    __shared__ __align__(16) float cs[VAR_TILE];
    __shared__ int s_S[VAR_THREADS / 32], s_B[VAR_THREADS / 32], s_bad[VAR_THREADS / 32];
    __shared__ float s_acc;
    const int chain = blockIdx.x;
    const float* __restrict__ src = vals + (size_t)chain * stride;  // stride is a multiple of 4 floats: float4 loads stay aligned
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_acc = 0.f;
    auto load_tile = [&](uint32_t base) {
        // values past n read as 0 (a no-op for the sum)
        const uint32_t t = base + 4 * threadIdx.x;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (t + 3 < n) v = __ldg(reinterpret_cast<const float4*>(src + t));
        else {
            if (t < n) v.x = __ldg(src + t);
            if (t + 1 < n) v.y = __ldg(src + t + 1);
            if (t + 2 < n) v.z = __ldg(src + t + 2);
        }
        return v;
    };
    float4 nxt = load_tile(0);
    __syncthreads();
    for (uint32_t base = 0; base < n; base += VAR_TILE) {
        const uint32_t cnt = min((uint32_t)VAR_TILE, n - base);
        const float4 v = nxt;
        if (base + VAR_TILE < n) nxt = load_tile(base + VAR_TILE);  // in flight while this tile is processed
        reinterpret_cast<float4*>(cs)[threadIdx.x] = v;
        const float acc = s_acc;
        const uint32_t ab = __float_as_uint(acc);
        const uint32_t aex = (ab >> 23) & 0xFFu;
        const bool acc_ok = (ab >> 31) == 0 && aex != 0 && aex != 0xFFu;  // positive, normal
        const int k = (int)aex - 150;                                      // ulp of acc = 2^k
        const int A = (int)((ab & 0x7FFFFFu) | 0x800000u);
        int S = 0, B = 0, bad = acc_ok ? 0 : 1;
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int r; bool tie;
            if (!round_scaled(vv[j], k, r, tie) || tie) bad = 1;
            else { S += r; B += min(abs(r), 1 << 24); }
        }
        if (B >= (1 << 24)) bad = 1;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            S += __shfl_xor_sync(0xFFFFFFFFu, S, off);
            B += __shfl_xor_sync(0xFFFFFFFFu, B, off);
            bad |= __shfl_xor_sync(0xFFFFFFFFu, bad, off);
        }
        if (lane == 0) { s_S[warp] = S; s_B[warp] = B; s_bad[warp] = bad; }
        __syncthreads();
        if (warp == 0) {
            long long St = s_S[lane], Bt = s_B[lane];
            int badt = s_bad[lane];
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) {
                St += __shfl_xor_sync(0xFFFFFFFFu, St, off);
                Bt += __shfl_xor_sync(0xFFFFFFFFu, Bt, off);
                badt |= __shfl_xor_sync(0xFFFFFFFFu, badt, off);
            }
            if (lane == 0) {
                float a = acc;
                if (!badt && (long long)A - Bt >= (1ll << 23) + 1 && (long long)A + Bt <= (1ll << 24) - 1) {
                    a = __uint_as_float((aex << 23) | ((uint32_t)(A + (int)St) & 0x7FFFFFu));
                } else {
#pragma unroll 16
                    for (uint32_t t = 0; t < cnt; ++t) a = a + cs[t];
#ifdef B3_VAR_TRACE
                    atomicAdd(reinterpret_cast<int*>(sums) + 16 + chain, 1);
#endif
                }
                s_acc = a;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) sums[chain] = s_acc;
}

// variance = csumsq * (csum * csum / n) — a product, not a difference, as in the reference (:203-204); strict > from axis 0
__global__ void sap_variance_axis_kernel(const float* __restrict__ sums, uint32_t n, int* __restrict__ next_axis) {
    if (threadIdx.x == 0) {
        const float fn = (float)n;
        float vx = sums[3] * (sums[0] * sums[0] / fn), vy = sums[4] * (sums[1] * sums[1] / fn), vz = sums[5] * (sums[2] * sums[2] / fn);
        int ax = 0; float best = vx;
        if (vy > best) { ax = 1; best = vy; }
        if (vz > best) { ax = 2; best = vz; }
        *next_axis = ax;
    }
}

__global__ void __launch_bounds__(BT) unpack_pairs_kernel(const unsigned long long* __restrict__ keys, unsigned long long n, int shift,
                                                          uint2* __restrict__ pairs, int hi_first) {
    unsigned long long k = (unsigned long long)blockIdx.x * BT + threadIdx.x;
    if (k >= n) return;
    unsigned long long key = keys[k];
    uint32_t hi = (uint32_t)(key >> shift), lo = (uint32_t)(key & ((1ull << shift) - 1ull));
    pairs[k] = hi_first ? make_uint2(hi, lo) : make_uint2(lo, hi);
}

int sap_bits_for(uint32_t n) { int b = 1; while (b < 32 && (1ull << b) < n) ++b; return b; }

size_t sap_cub_tmp_bytes(uint32_t n, uint64_t pair_capacity) {
    size_t a = 0, b = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    cub::DeviceRadixSort::SortKeys(nullptr, b, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (long long)pair_capacity);
    return (a > b ? a : b) + 256;
}

int launch_sap_sort(cudaStream_t st, const float* d_aabbs, uint32_t n, int axis, const SapScratch& S, uint32_t* d_perm, float* d_sorted6,
                    unsigned long long* d_candidates) {
    int launched = 0;
    const int g = (n + BT - 1) / BT;
    sap_keys_kernel<<<g, BT, 0, st>>>(d_aabbs, n, axis, S.keys_in, S.vals_in); ++launched;
    size_t tmp = S.cub_tmp_bytes;
    cub::DeviceRadixSort::SortPairs(S.cub_tmp, tmp, S.keys_in, S.keys_out, S.vals_in, S.vals_out, (int)n, 0, 64, st); launched += 10;
    sap_gather_kernel<<<g, BT, 0, st>>>(d_aabbs, S.vals_out, n, S.lo, S.hi, d_perm, d_sorted6); ++launched;
    cudaMemsetAsync(d_candidates, 0, sizeof(unsigned long long), st);
    sap_candidates_kernel<<<g, BT, 0, st>>>(S.lo, S.hi, n, axis, d_candidates); ++launched;
    return launched;
}

__global__ void __launch_bounds__(BT) boxes_to_lohi_kernel(const float* __restrict__ aabbs, uint32_t n, float4* __restrict__ lo, float4* __restrict__ hi) {
    uint32_t k = blockIdx.x * BT + threadIdx.x;
    if (k >= n) return;
    const float* b = aabbs + 6 * (size_t)k;
    lo[k] = make_float4(b[0], b[1], b[2], 0.f);
    hi[k] = make_float4(b[3], b[4], b[5], 0.f);
}
int launch_boxes_to_lohi(cudaStream_t st, const float* d_aabbs, uint32_t n, float4* lo, float4* hi) {
    if (n == 0) return 0;
    boxes_to_lohi_kernel<<<(n + BT - 1) / BT, BT, 0, st>>>(d_aabbs, n, lo, hi);
    return 1;
}

size_t sap_variance_scratch_bytes(uint32_t n) { return 6 * (((size_t)n + 3) & ~(size_t)3) * sizeof(float); }

int launch_sap_variance(cudaStream_t st, const SapScratch& S, uint32_t n, void* d_scratch, float* d_sums, int* d_next_axis) {
    const size_t stride = ((size_t)n + 3) & ~(size_t)3;
    float* vals = reinterpret_cast<float*>(d_scratch);
    if (n) sap_centres_kernel<<<(n + BT - 1) / BT, BT, 0, st>>>(S.lo, S.hi, n, stride, vals);
    sap_variance_chain_kernel<<<6, VAR_THREADS, 0, st>>>(vals, n, stride, d_sums);
    sap_variance_axis_kernel<<<1, 32, 0, st>>>(d_sums, n, d_next_axis);
    return 3;
}

int launch_sap_sweep(cudaStream_t st, uint32_t n, int axis, uint32_t j_begin, uint32_t j_end, const SapScratch& S, uint64_t capacity,
                     unsigned long long* d_count) {
    cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st);
    sap_sweep_kernel<<<(n + BT - 1) / BT, BT, 0, st>>>(S.lo, S.hi, n, axis, j_begin, j_end, sap_bits_for(n), S.pair_keys, capacity, d_count);
    return 1;
}

int launch_sap_finish(cudaStream_t st, const SapScratch& S, uint64_t count, uint2* d_pairs, uint32_t n) {
    if (count == 0) return 0;
    const int shift = sap_bits_for(n);
    size_t tmp = S.cub_tmp_bytes;
    // reference emission order: shape_index (j) ascending, then active index (i) ascending (sweep_prune.rs:98-113)
    cub::DeviceRadixSort::SortKeys(S.cub_tmp, tmp, S.pair_keys, S.pair_keys_alt, (long long)count, 0, 2 * shift, st);
    unpack_pairs_kernel<<<(unsigned)((count + BT - 1) / BT), BT, 0, st>>>(S.pair_keys_alt, count, shift, d_pairs, 0);
    return 2 + (2 * shift + 7) / 8 + 1;
}

// ------------------------------------------------------------------------------------------------------------
// world AABBs: ComputeBound<Aabb3> per primitive, then Aabb::transform (aabb3.rs:89-96: 8 corners, grown in order)
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT) world_aabb_kernel(ShapeSetDev S, float* __restrict__ out) {
    uint32_t i = blockIdx.x * BT + threadIdx.x;
    if (i >= S.n) return;
    Shape s;
    load_shape(s, S.recs, i);
    uint32_t meta = S.meta[i];
    uint32_t kind = meta & 0xFF;
    V3 mn = v3(0.f, 0.f, 0.f), mx = v3(0.f, 0.f, 0.f);
    switch (kind) {
        case K_PARTICLE: break;                                                               // primitive3.rs:86 Aabb3::zero()
        case K_QUAD: mn = v3(-s.p0, -s.p1, 0.f); mx = v3(s.p0, s.p1, 0.f); break;              // quad.rs:62-69
        case K_SPHERE: mn = v3(-s.p0, -s.p0, -s.p0); mx = v3(s.p0, s.p0, s.p0); break;         // sphere.rs:27-34
        case K_CUBOID: mn = v3(-s.p0, -s.p1, -s.p2); mx = v3(s.p0, s.p1, s.p2); break;         // cuboid.rs:66-73
        case K_CYLINDER: mn = v3(-s.p1, -s.p0, -s.p1); mx = v3(s.p1, s.p0, s.p1); break;       // cylinder.rs:64-71
        case K_CAPSULE: mn = v3(-s.p1, -s.p0 - s.p1, -s.p1); mx = v3(s.p1, s.p0 + s.p1, s.p1); break;  // capsule.rs:51-58
        default: {  // polyhedron.rs:84-86: fold from Aabb3::zero() growing by every vertex
            uint32_t mid = meta >> 8;
            uint32_t b = S.mesh_offsets[mid] + MESH_HEADER_SLOTS, e = S.mesh_offsets[mid + 1];  // skip the mesh header slots
            for (uint32_t k = b; k < e; ++k) {
                float4 p = S.mesh_verts[k];
                mn = v3(rmin(mn.x, p.x), rmin(mn.y, p.y), rmin(mn.z, p.z));
                mx = v3(rmax(mx.x, p.x), rmax(mx.y, p.y), rmax(mx.z, p.z));
            }
        }
    }
    // Aabb3::new sorts the two corner points component-wise (aabb3.rs:26-31)
    V3 lo = v3(rmin(mn.x, mx.x), rmin(mn.y, mx.y), rmin(mn.z, mx.z)), hi = v3(rmax(mn.x, mx.x), rmax(mn.y, mx.y), rmax(mn.z, mx.z));
    // to_corners order (aabb3.rs:35-46)
    V3 c[8] = {lo, v3(hi.x, lo.y, lo.z), v3(lo.x, hi.y, lo.z), v3(hi.x, hi.y, lo.z), v3(lo.x, lo.y, hi.z), v3(hi.x, lo.y, hi.z),
               v3(lo.x, hi.y, hi.z), hi};
    V3 f = transform_point3(s, c[0]);
    V3 wmin = f, wmax = f;
#pragma unroll
    for (int k = 1; k < 8; ++k) {
        V3 p = transform_point3(s, c[k]);
        // grow = Aabb::new(min(min, p), max(max, p)), and new() re-sorts min/max (no-op here)
        wmin = v3(rmin(wmin.x, p.x), rmin(wmin.y, p.y), rmin(wmin.z, p.z));
        wmax = v3(rmax(wmax.x, p.x), rmax(wmax.y, p.y), rmax(wmax.z, p.z));
    }
    float* o = out + 6 * (size_t)i;
    o[0] = wmin.x; o[1] = wmin.y; o[2] = wmin.z; o[3] = wmax.x; o[4] = wmax.y; o[5] = wmax.z;
}

int launch_world_aabbs(cudaStream_t st, const ShapeSetDev& S, float* d_out6) {
    if (S.n == 0) return 0;
    world_aabb_kernel<<<(S.n + BT - 1) / BT, BT, 0, st>>>(S, d_out6);
    return 1;
}

// ------------------------------------------------------------------------------------------------------------
// world bounding spheres: ComputeBound<volume::Sphere> per primitive (primitive3.rs:98-114: particle r = 0, quad.rs:71-78,
// sphere.rs:36-43, cuboid.rs:75-83 [the largest half extent, as written], cylinder.rs:73-80, capsule.rs:60-67,
// polyhedron.rs:87-92,363-370 [largest vertex norm]), then Bound::transform_volume (volume/sphere.rs:34-39: the centre
// moves, the radius stays). out = n x (centre.xyz, radius).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT) world_sphere_kernel(ShapeSetDev S, float4* __restrict__ out) {
    uint32_t i = blockIdx.x * BT + threadIdx.x;
    if (i >= S.n) return;
    Shape s;
    load_shape(s, S.recs, i);
    uint32_t meta = S.meta[i];
    float r = 0.f;
    switch (meta & 0xFF) {
        case K_PARTICLE: break;
        case K_QUAD: r = rmax(s.p0, s.p1); break;
        case K_SPHERE: r = s.p0; break;
        case K_CUBOID: r = rmax(rmax(s.p0, s.p1), s.p2); break;
        case K_CYLINDER: r = sqrtf((s.p1 * s.p1) + (s.p0 * s.p0)); break;   // p0 = half_height, p1 = radius
        case K_CAPSULE: r = s.p0 + s.p1; break;
        default: {
            uint32_t mid = meta >> 8;
            uint32_t b = S.mesh_offsets[mid] + MESH_HEADER_SLOTS, e = S.mesh_offsets[mid + 1];
            bool any = false;
            for (uint32_t k = b; k < e; ++k) {
                float4 p = S.mesh_verts[k];
                float len = sqrtf(dot(v3(p.x, p.y, p.z), v3(p.x, p.y, p.z)));
                // Iterator::max_by keeps the later element unless the earlier one compares Greater (NaN compares Equal)
                if (!any || !(r > len)) r = len;
                any = true;
            }
        }
    }
    V3 c = transform_point3(s, v3(0.f, 0.f, 0.f));
    out[i] = make_float4(c.x, c.y, c.z, r);
}

int launch_world_spheres(cudaStream_t st, const ShapeSetDev& S, float4* d_out) {
    if (S.n == 0) return 0;
    world_sphere_kernel<<<(S.n + BT - 1) / BT, BT, 0, st>>>(S, d_out);
    return 1;
}

// ------------------------------------------------------------------------------------------------------------
// SweepAndPrune3<volume::Sphere>: the same sort / Variance3 / pair search as for Aabb3, on the spheres' extents
// (Bound for Sphere, volume/sphere.rs:19-25: centre + splat(-+ radius)); the pair test is Sphere::intersects
// (volume/sphere.rs:82-91: |c1 - c2|^2 <= (r1 + r2)^2 — touching spheres are reported, unlike boxes).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BT) spheres_to_extents_kernel(const float4* __restrict__ spheres, uint32_t n, float* __restrict__ boxes) {
    uint32_t i = blockIdx.x * BT + threadIdx.x;
    if (i >= n) return;
    const float4 s = spheres[i];
    float* o = boxes + 6 * (size_t)i;
    o[0] = s.x + -s.w; o[1] = s.y + -s.w; o[2] = s.z + -s.w;
    o[3] = s.x + s.w;  o[4] = s.y + s.w;  o[5] = s.z + s.w;
}

int launch_spheres_to_extents(cudaStream_t st, const float4* d_spheres, uint32_t n, float* d_boxes6) {
    if (n == 0) return 0;
    spheres_to_extents_kernel<<<(n + BT - 1) / BT, BT, 0, st>>>(d_spheres, n, d_boxes6);
    return 1;
}

// The pair search works on strict box overlap. Sphere::intersects holds for touching spheres and is evaluated in f32, so
// the search runs on boxes grown by a margin that covers both (relative 2^-16 of the radius and of the coordinate
// magnitude — orders of magnitude above the rounding of the f32 test): a superset, filtered exactly afterwards.
B3_DEV float grow_margin(float lo, float hi) { return (fabsf(lo) + fabsf(hi)) * 1.52587890625e-05f + 1e-30f; }
__global__ void __launch_bounds__(BT) sap_inflate_kernel(float4* __restrict__ lo, float4* __restrict__ hi, float* __restrict__ sorted6, uint32_t n) {
    uint32_t k = blockIdx.x * BT + threadIdx.x;
    if (k >= n) return;
    float4 a = lo[k], b = hi[k];
    const float mx = grow_margin(a.x, b.x), my = grow_margin(a.y, b.y), mz = grow_margin(a.z, b.z);
    a.x -= mx; a.y -= my; a.z -= mz; b.x += mx; b.y += my; b.z += mz;
    lo[k] = a; hi[k] = b;
    if (sorted6) { float* o = sorted6 + 6 * (size_t)k; o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = b.x; o[4] = b.y; o[5] = b.z; }
}
int launch_sap_inflate(cudaStream_t st, const SapScratch& S, float* d_sorted6, uint32_t n) {
    if (n == 0) return 0;
    sap_inflate_kernel<<<(n + BT - 1) / BT, BT, 0, st>>>(S.lo, S.hi, d_sorted6, n);
    return 1;
}

// Boxes that cover the spheres with the same margin: the node bounds of a BVH over volume::Sphere values. A query that the
// sphere's own f32 predicate accepts (touching counts, rounding either way) must reach the leaf, so the cover is a superset.
__global__ void __launch_bounds__(BT) sphere_cover_boxes_kernel(const float4* __restrict__ spheres, uint32_t n, float* __restrict__ boxes6) {
    uint32_t k = blockIdx.x * BT + threadIdx.x;
    if (k >= n) return;
    const float4 s = spheres[k];
    float lo[3] = {s.x - s.w, s.y - s.w, s.z - s.w}, hi[3] = {s.x + s.w, s.y + s.w, s.z + s.w};
    float* o = boxes6 + 6 * (size_t)k;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float m = grow_margin(lo[a], hi[a]);
        o[a] = lo[a] - m; o[3 + a] = hi[a] + m;
    }
}
int launch_sphere_cover_boxes(cudaStream_t st, const float4* d_spheres, uint32_t n, float* d_boxes6) {
    if (n == 0) return 0;
    sphere_cover_boxes_kernel<<<(n + BT - 1) / BT, BT, 0, st>>>(d_spheres, n, d_boxes6);
    return 1;
}

__global__ void __launch_bounds__(BT) gather_spheres_kernel(const float4* __restrict__ spheres, const uint32_t* __restrict__ perm, uint32_t n,
                                                            float4* __restrict__ sorted) {
    uint32_t k = blockIdx.x * BT + threadIdx.x;
    if (k < n) sorted[k] = spheres[perm[k]];
}
int launch_gather_spheres(cudaStream_t st, const float4* d_spheres, const uint32_t* d_perm, uint32_t n, float4* d_sorted) {
    if (n == 0) return 0;
    gather_spheres_kernel<<<(n + BT - 1) / BT, BT, 0, st>>>(d_spheres, d_perm, n, d_sorted);
    return 1;
}

// Keeps (i, j), i < j in sorted order, iff the reference reports it: i is still active when j arrives
// (max_extent_i[axis] >= min_extent_j[axis], sweep_prune.rs:102-105) and the spheres intersect. Order-preserving.
struct SpherePairKeep {
    const float4* s;
    int axis;
    __device__ bool operator()(const uint2& p) const {
        const float4 a = s[p.x], b = s[p.y];
        const float ca = axis == 0 ? a.x : (axis == 1 ? a.y : a.z), cb = axis == 0 ? b.x : (axis == 1 ? b.y : b.z);
        if (!((ca + a.w) >= (cb + -b.w))) return false;
        const V3 d = v3(a.x - b.x, a.y - b.y, a.z - b.z);
        const float rr = a.w + b.w;
        return dot(d, d) <= rr * rr;
    }
};
size_t sphere_filter_tmp_bytes(uint64_t n_pairs) {
    size_t bytes = 0;
    cub::DeviceSelect::If((void*)nullptr, bytes, (const uint2*)nullptr, (uint2*)nullptr, (unsigned long long*)nullptr, (long long)n_pairs, SpherePairKeep{nullptr, 0});
    return bytes;
}
int launch_sphere_pair_filter(cudaStream_t st, const float4* d_sorted_spheres, int axis, const uint2* d_in, uint64_t n_pairs, uint2* d_out,
                              unsigned long long* d_count, void* d_tmp, size_t tmp_bytes) {
    cub::DeviceSelect::If(d_tmp, tmp_bytes, d_in, d_out, d_count, (long long)n_pairs, SpherePairKeep{d_sorted_spheres, axis}, st);
    return 2;
}

// ------------------------------------------------------------------------------------------------------------
// LBVH build (Karras 2012): Morton codes of box centres, radix sort, radix-tree topology, bottom-up bounds
// ------------------------------------------------------------------------------------------------------------
B3_DEV uint32_t fkey(float f) { uint32_t b = __float_as_uint(f); return (b & 0x80000000u) ? ~b : (b | 0x80000000u); }
B3_DEV float funkey(uint32_t k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k); }

__global__ void bvh_scene_init_kernel(uint32_t* scene) {
    if (threadIdx.x < 3) scene[threadIdx.x] = 0xFFFFFFFFu;
    else if (threadIdx.x < 6) scene[threadIdx.x] = 0u;
}

// scene bounds of the box centres (min / max per axis as order-preserving keys): grid-stride accumulation per thread, one
// warp reduction (redux.sync) and six atomics per warp at the end
__global__ void __launch_bounds__(BT) bvh_scene_kernel(const float* __restrict__ aabbs, uint32_t n, uint32_t* __restrict__ scene) {
    uint32_t kmin[3] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu}, kmax[3] = {0u, 0u, 0u};
    for (uint32_t i = blockIdx.x * BT + threadIdx.x; i < n; i += gridDim.x * BT) {
        const float* b = aabbs + 6 * (size_t)i;
        const float c[3] = {(b[0] + b[3]) * 0.5f, (b[1] + b[4]) * 0.5f, (b[2] + b[5]) * 0.5f};
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            if (c[a] == c[a]) {
                const uint32_t k = fkey(c[a]);
                kmin[a] = min(kmin[a], k);
                kmax[a] = max(kmax[a], k);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const uint32_t lo = __reduce_min_sync(0xFFFFFFFFu, kmin[a]), hi = __reduce_max_sync(0xFFFFFFFFu, kmax[a]);
        if ((threadIdx.x & 31) == 0) { atomicMin(scene + a, lo); atomicMax(scene + 3 + a, hi); }
    }
}

B3_DEV uint32_t expand10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void __launch_bounds__(BT) bvh_morton_kernel(const float* __restrict__ aabbs, uint32_t n, const uint32_t* __restrict__ scene,
                                                        uint32_t* __restrict__ morton, uint32_t* __restrict__ idx) {
    uint32_t i = blockIdx.x * BT + threadIdx.x;
    if (i >= n) return;
    const float* b = aabbs + 6 * (size_t)i;
    uint32_t q[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float lo = funkey(scene[a]), hi = funkey(scene[3 + a]);
        float c = (b[a] + b[3 + a]) * 0.5f;
        float ext = hi - lo;
        float t = ext > 0.f ? (c - lo) / ext : 0.f;
        t = fminf(fmaxf(t * 1024.f, 0.f), 1023.f);  // NaN -> 0
        q[a] = (uint32_t)t;
    }
    morton[i] = (expand10(q[0]) << 2) | (expand10(q[1]) << 1) | expand10(q[2]);
    idx[i] = i;
}

B3_DEV int delta(const uint32_t* __restrict__ m, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint32_t a = m[i], b = m[j];
    if (a != b) return __clz(a ^ b);
    return 32 + __clz((uint32_t)i ^ (uint32_t)j);
}

__global__ void __launch_bounds__(BT) bvh_topology_kernel(const uint32_t* __restrict__ morton, uint32_t n, int2* __restrict__ children,
                                                          uint32_t* __restrict__ parent) {
    int i = blockIdx.x * BT + threadIdx.x;
    const int ni = (int)n;
    if (i >= ni - 1) return;
    int d = (delta(morton, ni, i, i + 1) - delta(morton, ni, i, i - 1)) >= 0 ? 1 : -1;
    int dmin = delta(morton, ni, i, i - d);
    int lmax = 2;
    while (delta(morton, ni, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(morton, ni, i, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = delta(morton, ni, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(morton, ni, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + min(d, 0);
    int left = (min(i, j) == gamma) ? (ni - 1) + gamma : gamma;
    int right = (max(i, j) == gamma + 1) ? (ni - 1) + gamma + 1 : gamma + 1;
    children[i] = make_int2(left, right);
    parent[left] = (uint32_t)i;
    parent[right] = (uint32_t)i;
    if (i == 0) parent[0] = 0xFFFFFFFFu;
}

// Ropes: rope[node] = the node a depth-first, left-first traversal continues with once it is done with `node`'s subtree —
// the right sibling of the nearest ancestor-or-self that is a left child, BVH_END above the root. With them the traversal
// needs no stack (dbvt/mod.rs:481-515 keeps a fixed 256-entry one and indexes past it on a deeper tree): hit an internal node
// -> its left child; miss, or a leaf -> the rope.
__global__ void __launch_bounds__(BT) bvh_ropes_kernel(const int2* __restrict__ children, const uint32_t* __restrict__ parent, uint32_t n,
                                                       uint32_t* __restrict__ rope) {
    const uint32_t node = blockIdx.x * BT + threadIdx.x;
    if (node >= 2 * n - 1) return;
    uint32_t cur = node, r = BVH_END;
    for (;;) {
        const uint32_t p = parent[cur];
        if (p == 0xFFFFFFFFu) break;
        const int2 ch = children[p];
        if ((uint32_t)ch.x == cur) { r = (uint32_t)ch.y; break; }
        cur = p;
    }
    rope[node] = r;
}

// Node record: two float4 side by side (one 32-byte sector): (min.xyz, A) and (max.xyz, rope) with A = the left child of an
// internal node or the value index of a leaf, stored as int bits: a traversal step reads one sector and nothing else.
__global__ void __launch_bounds__(BT) bvh_refit_kernel(const float* __restrict__ aabbs, const uint32_t* __restrict__ leaf_value, uint32_t n,
                                                       const int2* __restrict__ children, const uint32_t* __restrict__ parent,
                                                       const uint32_t* __restrict__ rope, uint32_t* __restrict__ flags, float4* __restrict__ nodes) {
    uint32_t k = blockIdx.x * BT + threadIdx.x;
    if (k >= n) return;
    const uint32_t value = leaf_value[k];
    const float* b = aabbs + 6 * (size_t)value;
    uint32_t node = (n - 1) + k;
    nodes[2 * (size_t)node] = make_float4(b[0], b[1], b[2], __int_as_float((int)value));
    nodes[2 * (size_t)node + 1] = make_float4(b[3], b[4], b[5], __int_as_float((int)rope[node]));
    if (n == 1) return;
    __threadfence();
    uint32_t p = parent[node];
    while (p != 0xFFFFFFFFu) {
        if (atomicAdd(flags + p, 1u) == 0u) return;  // the second child to arrive computes the parent
        __threadfence();
        int2 ch = children[p];
        float4 l0 = __ldcg(nodes + 2 * (size_t)ch.x), h0 = __ldcg(nodes + 2 * (size_t)ch.x + 1);
        float4 l1 = __ldcg(nodes + 2 * (size_t)ch.y), h1 = __ldcg(nodes + 2 * (size_t)ch.y + 1);
        // union = grow(min).grow(max) with Rust's NaN-ignoring min/max (aabb3.rs:146-152)
        nodes[2 * (size_t)p] = make_float4(rmin(l0.x, l1.x), rmin(l0.y, l1.y), rmin(l0.z, l1.z), __int_as_float(ch.x));
        nodes[2 * (size_t)p + 1] = make_float4(rmax(h0.x, h1.x), rmax(h0.y, h1.y), rmax(h0.z, h1.z), __int_as_float((int)rope[p]));
        __threadfence();
        p = parent[p];
    }
}

// Refit without rebuilding (the device counterpart of update_node + do_refit, dbvt/mod.rs:561-589): the topology stays,
// leaf bounds are re-read from the new values and the internal bounds recomputed bottom-up. Query results depend only on
// "every branch bound contains its descendants", which holds again afterwards; only the tree's quality degrades as the
// values move away from the Morton order they were built in.
__global__ void __launch_bounds__(BT) bvh_parents_kernel(const int2* __restrict__ children, uint32_t n, uint32_t* __restrict__ parent) {
    uint32_t i = blockIdx.x * BT + threadIdx.x;
    if (i + 1 >= n) return;
    const int2 ch = children[i];
    parent[ch.x] = i;
    parent[ch.y] = i;
    if (i == 0) parent[0] = 0xFFFFFFFFu;
}

int launch_bvh_refit(cudaStream_t st, const float* d_aabbs, uint32_t n, const int2* children, const uint32_t* leaf_value, uint32_t* parent,
                     const uint32_t* rope, uint32_t* flags, float4* nodes) {
    if (n == 0) return 0;
    int launched = 1;
    if (n > 1) {
        cudaMemsetAsync(flags, 0, (size_t)(n - 1) * 4, st);
        bvh_parents_kernel<<<(n + BT - 1) / BT, BT, 0, st>>>(children, n, parent);
        ++launched;
    }
    bvh_refit_kernel<<<(n + BT - 1) / BT, BT, 0, st>>>(d_aabbs, leaf_value, n, children, parent, rope, flags, nodes);
    return launched;
}

size_t bvh_cub_tmp_bytes(uint32_t n) {
    size_t a = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, a, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    return a + 256;
}

int launch_bvh_build(cudaStream_t st, const float* d_aabbs, uint32_t n, const BvhBuildScratch& W, float4* nodes, int2* children,
                     uint32_t* leaf_value, uint32_t* rope) {
    if (n == 0) return 0;
    int launched = 0;
    const int g = (n + BT - 1) / BT;
    uint32_t* scene = reinterpret_cast<uint32_t*>(W.scene);
    bvh_scene_init_kernel<<<1, 32, 0, st>>>(scene); ++launched;
    bvh_scene_kernel<<<min(g, 148 * 16), BT, 0, st>>>(d_aabbs, n, scene); ++launched;
    bvh_morton_kernel<<<g, BT, 0, st>>>(d_aabbs, n, scene, W.morton_in, W.idx_in); ++launched;
    size_t tmp = W.cub_tmp_bytes;
    cub::DeviceRadixSort::SortPairs(W.cub_tmp, tmp, W.morton_in, W.morton_out, W.idx_in, leaf_value, (int)n, 0, 30, st); launched += 4;
    if (n > 1) {
        cudaMemsetAsync(W.flags, 0, (size_t)(n - 1) * 4, st);
        bvh_topology_kernel<<<(n - 1 + BT - 1) / BT, BT, 0, st>>>(W.morton_out, n, children, W.parent); ++launched;
        bvh_ropes_kernel<<<(2 * n - 1 + BT - 1) / BT, BT, 0, st>>>(children, W.parent, n, rope); ++launched;
    } else {
        cudaMemsetAsync(rope, 0xFF, 4, st);  // a single leaf: nothing follows it
    }
    bvh_refit_kernel<<<g, BT, 0, st>>>(d_aabbs, leaf_value, n, children, W.parent, rope, W.flags, nodes); ++launched;
    return launched;
}

// ------------------------------------------------------------------------------------------------------------
// visitors (dbvt/visitor.rs, dbvt/util.rs) and the predicates they call
// ------------------------------------------------------------------------------------------------------------
// Continuous<Aabb3> for Ray (aabb3.rs:165-195)
B3_DEV bool ray_aabb(V3 o, V3 d, float4 mn, float4 mx, V3& out, bool* entry = nullptr) {
    float ix = 1.0f / d.x, iy = 1.0f / d.y, iz = 1.0f / d.z;
    float t1 = (mn.x - o.x) * ix, t2 = (mx.x - o.x) * ix;
    float tmin = rmin(t1, t2), tmax = rmax(t1, t2);
    t1 = (mn.y - o.y) * iy; t2 = (mx.y - o.y) * iy;
    tmin = rmax(tmin, rmin(t1, t2)); tmax = rmin(tmax, rmax(t1, t2));
    t1 = (mn.z - o.z) * iz; t2 = (mx.z - o.z) * iz;
    tmin = rmax(tmin, rmin(t1, t2)); tmax = rmin(tmax, rmax(t1, t2));
    if ((tmin < 0.0f && tmax < 0.0f) || tmax < tmin) return false;
    float t = (tmin >= 0.0f) ? tmin : tmax;
    if (entry) *entry = tmin >= 0.0f;  // the returned point is where the ray ENTERS the box (origin outside)
    out = o + d * t;
    return true;
}

// PlaneBound for Vec3 (bound.rs:48-58): In = 0, Cross = 1, Out = 2
B3_DEV int relate_point(V3 p, float4 pl) {
    float dist = dot(p, v3(pl.x, pl.y, pl.z));
    return dist > pl.w ? 0 : (dist < pl.w ? 2 : 1);
}
// PlaneBound for Aabb3 (aabb3.rs:246-257)
B3_DEV int relate_box(float4 mn, float4 mx, float4 pl) {
    int first = relate_point(v3(mn.x, mn.y, mn.z), pl);
    if (relate_point(v3(mx.x, mn.y, mn.z), pl) != first) return 1;
    if (relate_point(v3(mn.x, mx.y, mn.z), pl) != first) return 1;
    if (relate_point(v3(mx.x, mx.y, mn.z), pl) != first) return 1;
    if (relate_point(v3(mn.x, mn.y, mx.z), pl) != first) return 1;
    if (relate_point(v3(mx.x, mn.y, mx.z), pl) != first) return 1;
    if (relate_point(v3(mn.x, mx.y, mx.z), pl) != first) return 1;
    if (relate_point(v3(mx.x, mx.y, mx.z), pl) != first) return 1;
    return first;
}
// Frustum::contains (frustum.rs:78-95): fold max over [left, right, top, bottom, near, far]; planes are stored
// left,right,bottom,top,near,far (the struct's field order) — max is order independent.
B3_DEV int frustum_relation(float4 mn, float4 mx, const float4* pl) {
    int cur = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) {
        cur = max(cur, relate_box(mn, mx, pl[k]));
        if (cur == 2) break;  // Out stays Out whatever the other planes say (the fold is a max): same value, fewer planes
    }
    return cur;
}

// Stackless depth-first traversal, left child first (the order the reference's explicit stack gives: it pushes left, right
// and pops right first — result SETS are what is compared, see bam3d_gpu.h). `visit(leaf, A, mn, mx)` is called for every
// node reached; for an internal node its return value says whether to descend; A = value index of a leaf.
template <class F>
B3_DEV void bvh_traverse(const BvhDev& T, F&& visit) {
    if (T.n == 0) return;
    const uint32_t first_leaf = T.n - 1;
    uint32_t node = 0;  // the root (a single leaf has id 0 too)
    while (node != BVH_END) {
        const float4 mn = __ldg(T.nodes + 2 * (size_t)node), mx = __ldg(T.nodes + 2 * (size_t)node + 1);
        const bool leaf = node >= first_leaf;
        const bool down = visit(leaf, (uint32_t)__float_as_int(mn.w), mn, mx);
        node = (down && !leaf) ? (uint32_t)__float_as_int(mn.w) : (uint32_t)__float_as_int(mx.w);
    }
}

struct Query {
    float4 qlo, qhi;     // BQ_AABB
    V3 o, d;             // BQ_RAY
    float4 pl[6];        // BQ_FRUSTUM
    float4 sph;          // BQ_SPHERE: centre, radius
};

template <int KIND>
B3_DEV bool accept(const Query& q, float4 mn, float4 mx, V3& point, int& rel) {
    if (KIND == BQ_AABB) {
        // DiscreteVisitor: bound.intersects(query) with self = the tree bound (aabb3.rs:237-244)
        return mx.x > q.qlo.x && mn.x < q.qhi.x && mx.y > q.qlo.y && mn.y < q.qhi.y && mx.z > q.qlo.z && mn.z < q.qhi.z;
    } else if (KIND == BQ_RAY) {
        return ray_aabb(q.o, q.d, mn, mx, point);
    } else if (KIND == BQ_SPHERE) {
        // node test only (the leaf is decided by sphere_leaf): the node box against the query sphere's grown extent, touching included
        const float4 c = q.sph;
        const float mx_ = grow_margin(c.x - c.w, c.x + c.w), my_ = grow_margin(c.y - c.w, c.y + c.w), mz_ = grow_margin(c.z - c.w, c.z + c.w);
        return mx.x >= c.x - c.w - mx_ && mn.x <= c.x + c.w + mx_ && mx.y >= c.y - c.w - my_ && mn.y <= c.y + c.w + my_ &&
               mx.z >= c.z - c.w - mz_ && mn.z <= c.z + c.w + mz_;
    } else {
        rel = frustum_relation(mn, mx, q.pl);
        return rel != 2;
    }
}

// The leaf predicate of a tree over volume::Sphere bounds (volume/sphere.rs), `s` = the value's bound:
//   BQ_SPHERE  DiscreteVisitor: Discrete<Sphere> (:82-91), touching spheres intersect;
//   BQ_RAY     ContinuousVisitor: Continuous<Ray> (:50-67), the first point on the sphere (None when the centre is behind the origin);
//   BQ_FRUSTUM FrustumVisitor: Frustum::contains over PlaneBound for Sphere (:93-104), fold max over the six planes.
template <int KIND>
B3_DEV bool sphere_leaf(const Query& q, float4 s, V3& point, int& rel) {
    const V3 c = v3(s.x, s.y, s.z);
    if (KIND == BQ_SPHERE) {
        const V3 distance = c - v3(q.sph.x, q.sph.y, q.sph.z);
        const float radiuses = s.w + q.sph.w;
        return dot(distance, distance) <= radiuses * radiuses;
    } else if (KIND == BQ_RAY) {
        const V3 l = c - q.o;
        const float tca = dot(l, q.d);
        if (tca < 0.f) return false;
        const float d2 = dot(l, l) - tca * tca;
        if (d2 > s.w * s.w) return false;
        const float thc = sqrtf(s.w * s.w - d2);
        point = q.o + q.d * (tca - thc);
        return true;
    } else if (KIND == BQ_FRUSTUM) {
        int cur = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const float dist = dot(c, v3(q.pl[k].x, q.pl[k].y, q.pl[k].z)) - q.pl[k].w;
            const int r = dist > s.w ? 0 : (dist < -s.w ? 2 : 1);
            cur = max(cur, r);
        }
        rel = cur;
        return cur != 2;
    } else {
        return false;  // an Aabb3 query against sphere bounds is not a reference operation
    }
}

template <int KIND>
B3_DEV void load_query(Query& q, const float* __restrict__ queries, uint32_t qi) {
    if (KIND == BQ_AABB) {
        const float* b = queries + 6 * (size_t)qi;
        q.qlo = make_float4(b[0], b[1], b[2], 0.f); q.qhi = make_float4(b[3], b[4], b[5], 0.f);
    } else if (KIND == BQ_RAY) {
        const float* b = queries + 6 * (size_t)qi;
        q.o = v3(b[0], b[1], b[2]); q.d = v3(b[3], b[4], b[5]);
    } else if (KIND == BQ_SPHERE) {
        q.sph = reinterpret_cast<const float4*>(queries)[qi];
    } else {
        const float4* b = reinterpret_cast<const float4*>(queries) + 6 * (size_t)qi;
#pragma unroll
        for (int k = 0; k < 6; ++k) q.pl[k] = b[k];
    }
}

// One thread per query, stackless. FILL = false counts hits; FILL = true writes them at offsets[qi].
template <int KIND, bool FILL>
__global__ void __launch_bounds__(BT) bvh_query_kernel(BvhDev T, const float* __restrict__ queries, uint32_t nq, uint32_t* __restrict__ counts,
                                                       const unsigned long long* __restrict__ offsets, uint32_t* __restrict__ values,
                                                       void* __restrict__ payload) {
    uint32_t qi = blockIdx.x * BT + threadIdx.x;
    if (qi >= nq) return;
    Query q;
    load_query<KIND>(q, queries, qi);
    unsigned long long w = FILL ? offsets[qi] : 0ull;
    uint32_t cnt = 0;
    bvh_traverse(T, [&](bool leaf, uint32_t value, float4 mn, float4 mx) {
        V3 point; int rel = 0;
        if (!accept<KIND>(q, mn, mx, point, rel)) return false;
        if (leaf && T.spheres && !sphere_leaf<KIND>(q, __ldg(T.spheres + value), point, rel)) return false;
        if (leaf) {
            if (FILL) {
                values[w] = value;
                if (KIND == BQ_RAY) { float* p = reinterpret_cast<float*>(payload) + 3 * w; p[0] = point.x; p[1] = point.y; p[2] = point.z; }
                if (KIND == BQ_FRUSTUM) reinterpret_cast<uint8_t*>(payload)[w] = (uint8_t)rel;
                ++w;
            }
            ++cnt;
        }
        return true;
    });
    if (!FILL) counts[qi] = cnt;
}

// The same in ONE pass for device-resident callers: every hit is appended to (out_query, out_value, out_payload) through a
// warp-aggregated atomic, in no particular order (sort by out_query for the per-query lists). *count keeps counting past
// `capacity`; hits beyond it are dropped.
template <int KIND>
__global__ void __launch_bounds__(BT) bvh_query_append_kernel(BvhDev T, const float* __restrict__ queries, uint32_t nq, uint32_t* __restrict__ out_query,
                                                              uint32_t* __restrict__ out_value, void* __restrict__ out_payload,
                                                              unsigned long long capacity, unsigned long long* __restrict__ count) {
    uint32_t qi = blockIdx.x * BT + threadIdx.x;
    if (qi >= nq) return;
    Query q;
    load_query<KIND>(q, queries, qi);
    const uint32_t lane = threadIdx.x & 31u;
    bvh_traverse(T, [&](bool leaf, uint32_t value, float4 mn, float4 mx) {
        V3 point; int rel = 0;
        if (!accept<KIND>(q, mn, mx, point, rel)) return false;
        if (leaf && T.spheres && !sphere_leaf<KIND>(q, __ldg(T.spheres + value), point, rel)) return false;
        if (leaf) {
            const uint32_t act = __activemask();
            const uint32_t leader = __ffs(act) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(count, (unsigned long long)__popc(act));
            base = __shfl_sync(act, base, leader);
            const unsigned long long w = base + __popc(act & ((1u << lane) - 1u));
            if (w < capacity) {
                out_query[w] = qi;
                out_value[w] = value;
                if (KIND == BQ_RAY) { float* p = reinterpret_cast<float*>(out_payload) + 3 * w; p[0] = point.x; p[1] = point.y; p[2] = point.z; }
                if (KIND == BQ_FRUSTUM) reinterpret_cast<uint8_t*>(out_payload)[w] = (uint8_t)rel;
            }
        }
        return true;
    });
}

int launch_bvh_query_append(cudaStream_t st, const BvhDev& T, int kind, const float* d_queries, uint32_t nq, uint32_t* d_out_query,
                            uint32_t* d_out_value, void* d_out_payload, uint64_t capacity, unsigned long long* d_count) {
    cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st);
    if (nq == 0 || T.n == 0) return 0;
    const int g = (nq + BT - 1) / BT;
    if (kind == BQ_AABB) bvh_query_append_kernel<BQ_AABB><<<g, BT, 0, st>>>(T, d_queries, nq, d_out_query, d_out_value, d_out_payload, capacity, d_count);
    else if (kind == BQ_SPHERE) bvh_query_append_kernel<BQ_SPHERE><<<g, BT, 0, st>>>(T, d_queries, nq, d_out_query, d_out_value, d_out_payload, capacity, d_count);
    else if (kind == BQ_RAY) bvh_query_append_kernel<BQ_RAY><<<g, BT, 0, st>>>(T, d_queries, nq, d_out_query, d_out_value, d_out_payload, capacity, d_count);
    else bvh_query_append_kernel<BQ_FRUSTUM><<<g, BT, 0, st>>>(T, d_queries, nq, d_out_query, d_out_value, d_out_payload, capacity, d_count);
    return 1;
}

// Few large queries, one pass: every (query, value) leaf test, a tile of 256 values per block against the queries of the grid's
// y-slice staged in shared memory; a thread keeps its box in registers. Hits are collected per block (shared counter) and flushed
// with one global atomic per block and query chunk, so the append counter is not hammered by 10^8 warp-level atomics.
constexpr int BRUTE_Q = 16;  // queries per shared-memory chunk
template <int KIND>
__global__ void __launch_bounds__(256) bvh_brute_append_kernel(const float* __restrict__ aabbs, uint32_t n, const float* __restrict__ queries, uint32_t nq,
                                                               uint32_t q_per_block, uint32_t* __restrict__ out_query, uint32_t* __restrict__ out_value,
                                                               void* __restrict__ out_payload, unsigned long long capacity,
                                                               unsigned long long* __restrict__ count) {
    constexpr int QF = KIND == BQ_FRUSTUM ? 24 : 6;
    __shared__ float sq[BRUTE_Q * QF];
    __shared__ uint32_t s_n;
    __shared__ unsigned long long s_base;
    __shared__ uint32_t s_q[256 * 4], s_v[256 * 4];
    __shared__ float s_p[KIND == BQ_RAY ? 256 * 4 * 3 : 1];
    __shared__ uint8_t s_r[KIND == BQ_FRUSTUM ? 256 * 4 : 1];
    const uint32_t v = blockIdx.x * 256 + threadIdx.x;
    float4 mn = make_float4(0.f, 0.f, 0.f, 0.f), mx = mn;
    if (v < n) { const float* b = aabbs + 6 * (size_t)v; mn = make_float4(b[0], b[1], b[2], 0.f); mx = make_float4(b[3], b[4], b[5], 0.f); }
    const uint32_t q_begin = blockIdx.y * q_per_block, q_end = min(nq, q_begin + q_per_block);
    for (uint32_t q0 = q_begin; q0 < q_end; q0 += BRUTE_Q) {
        const uint32_t nqc = min((uint32_t)BRUTE_Q, q_end - q0);
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < nqc * QF; i += 256) sq[i] = queries[(size_t)q0 * QF + i];
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
        // at most 4 hits per thread are staged per flush; a chunk is processed in sub-chunks of 4 queries
        for (uint32_t qs = 0; qs < nqc; qs += 4) {
            for (uint32_t k = qs; k < min(nqc, qs + 4); ++k) {
                Query q;
                if (KIND == BQ_AABB) { const float* b = sq + k * 6; q.qlo = make_float4(b[0], b[1], b[2], 0.f); q.qhi = make_float4(b[3], b[4], b[5], 0.f); }
                else if (KIND == BQ_RAY) { const float* b = sq + k * 6; q.o = v3(b[0], b[1], b[2]); q.d = v3(b[3], b[4], b[5]); }
                else {
#pragma unroll
                    for (int j = 0; j < 6; ++j) q.pl[j] = make_float4(sq[k * 24 + 4 * j], sq[k * 24 + 4 * j + 1], sq[k * 24 + 4 * j + 2], sq[k * 24 + 4 * j + 3]);
                }
                V3 point = v3(0.f, 0.f, 0.f); int rel = 0;
                if (v < n && accept<KIND>(q, mn, mx, point, rel)) {
                    const uint32_t at = atomicAdd(&s_n, 1u);
                    s_q[at] = q0 + k; s_v[at] = v;
                    if (KIND == BQ_RAY) { s_p[3 * at] = point.x; s_p[3 * at + 1] = point.y; s_p[3 * at + 2] = point.z; }
                    if (KIND == BQ_FRUSTUM) s_r[at] = (uint8_t)rel;
                }
            }
            __syncthreads();
            const uint32_t cnt = s_n;
            if (cnt) {
                if (threadIdx.x == 0) s_base = atomicAdd(count, (unsigned long long)cnt);
                __syncthreads();
                const unsigned long long base = s_base;
                for (uint32_t i = threadIdx.x; i < cnt; i += 256) {
                    const unsigned long long w = base + i;
                    if (w < capacity) {
                        out_query[w] = s_q[i]; out_value[w] = s_v[i];
                        if (KIND == BQ_RAY) { float* p = reinterpret_cast<float*>(out_payload) + 3 * w; p[0] = s_p[3 * i]; p[1] = s_p[3 * i + 1]; p[2] = s_p[3 * i + 2]; }
                        if (KIND == BQ_FRUSTUM) reinterpret_cast<uint8_t*>(out_payload)[w] = s_r[i];
                    }
                }
                __syncthreads();
                if (threadIdx.x == 0) s_n = 0;
                __syncthreads();
            }
        }
    }
}

int launch_bvh_brute_append(cudaStream_t st, const float* d_aabbs, uint32_t n, int kind, const float* d_queries, uint32_t nq, uint32_t* d_out_query,
                            uint32_t* d_out_value, void* d_out_payload, uint64_t capacity, unsigned long long* d_count) {
    cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st);
    if (nq == 0 || n == 0) return 0;
    const uint32_t ntiles = (n + 255) / 256;
    // enough blocks to fill the machine a few times over, as few passes over the boxes as that allows
    uint32_t ysplit = (148u * 32u + ntiles - 1) / ntiles;
    if (ysplit < 1) ysplit = 1;
    if (ysplit > nq) ysplit = nq;
    const uint32_t q_per_block = (nq + ysplit - 1) / ysplit;
    const dim3 grid(ntiles, (nq + q_per_block - 1) / q_per_block);
    if (kind == BQ_AABB) bvh_brute_append_kernel<BQ_AABB><<<grid, 256, 0, st>>>(d_aabbs, n, d_queries, nq, q_per_block, d_out_query, d_out_value, d_out_payload, capacity, d_count);
    else if (kind == BQ_RAY) bvh_brute_append_kernel<BQ_RAY><<<grid, 256, 0, st>>>(d_aabbs, n, d_queries, nq, q_per_block, d_out_query, d_out_value, d_out_payload, capacity, d_count);
    else bvh_brute_append_kernel<BQ_FRUSTUM><<<grid, 256, 0, st>>>(d_aabbs, n, d_queries, nq, q_per_block, d_out_query, d_out_value, d_out_payload, capacity, d_count);
    return 1;
}

int launch_bvh_query(cudaStream_t st, const BvhDev& T, int kind, const float* d_queries, uint32_t nq, uint32_t* d_counts,
                     const unsigned long long* d_offsets, uint32_t* d_values, void* d_payload) {
    if (nq == 0) return 0;
    const int g = (nq + BT - 1) / BT;
    const bool fill = d_offsets != nullptr;
#define B3_Q(KIND)                                                                                                        \
    if (fill) bvh_query_kernel<KIND, true><<<g, BT, 0, st>>>(T, d_queries, nq, d_counts, d_offsets, d_values, d_payload); \
    else bvh_query_kernel<KIND, false><<<g, BT, 0, st>>>(T, d_queries, nq, d_counts, d_offsets, d_values, d_payload);
    if (kind == BQ_AABB) { B3_Q(BQ_AABB) }
    else if (kind == BQ_SPHERE) { B3_Q(BQ_SPHERE) }
    else if (kind == BQ_RAY) { B3_Q(BQ_RAY) }
    else { B3_Q(BQ_FRUSTUM) }
#undef B3_Q
    return 1;
}

// Few, large queries (a frustum returns ~10 % of all values) leave a thread-per-query traversal with no parallelism:
// there the leaf predicate is simply evaluated for every (query, value) — exactly the set the tree query returns,
// since the query result is "every value whose real bound the visitor accepts" — one block per (tile of 256 values,
// query), results compacted in value-index order.
template <int KIND, bool FILL>
__global__ void __launch_bounds__(256) bvh_brute_kernel(const float* __restrict__ aabbs, uint32_t n, const float* __restrict__ queries,
                                                        uint32_t ntiles, uint32_t* __restrict__ counts, const unsigned long long* __restrict__ offsets,
                                                        uint32_t* __restrict__ values, void* __restrict__ payload) {
    __shared__ uint32_t warp_cnt[8];
    const uint32_t qi = blockIdx.y, tile = blockIdx.x, v = tile * 256 + threadIdx.x;
    Query q;
    load_query<KIND>(q, queries, qi);
    bool hit = false;
    V3 point = v3(0.f, 0.f, 0.f); int rel = 0;
    if (v < n) {
        const float* b = aabbs + 6 * (size_t)v;
        hit = accept<KIND>(q, make_float4(b[0], b[1], b[2], 0.f), make_float4(b[3], b[4], b[5], 0.f), point, rel);
    }
    const uint32_t mask = __ballot_sync(0xFFFFFFFFu, hit);
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[warp] = __popc(mask);
    __syncthreads();
    if (!FILL) {
        if (threadIdx.x == 0) {
            uint32_t c = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) c += warp_cnt[w];
            counts[(size_t)qi * ntiles + tile] = c;
        }
        return;
    }
    if (hit) {
        uint32_t before = 0;
        for (uint32_t w = 0; w < warp; ++w) before += warp_cnt[w];
        const unsigned long long slot = offsets[(size_t)qi * ntiles + tile] + before + __popc(mask & ((1u << lane) - 1u));
        values[slot] = v;
        if (KIND == BQ_RAY) { float* p = reinterpret_cast<float*>(payload) + 3 * slot; p[0] = point.x; p[1] = point.y; p[2] = point.z; }
        if (KIND == BQ_FRUSTUM) reinterpret_cast<uint8_t*>(payload)[slot] = (uint8_t)rel;
    }
}

int launch_bvh_brute(cudaStream_t st, const float* d_aabbs, uint32_t n, int kind, const float* d_queries, uint32_t nq, uint32_t* d_counts,
                     const unsigned long long* d_offsets, uint32_t* d_values, void* d_payload) {
    if (nq == 0 || n == 0) return 0;
    const uint32_t ntiles = (n + 255) / 256;
    dim3 g(ntiles, nq);
    const bool fill = d_offsets != nullptr;
#define B3_B(KIND)                                                                                                                  \
    if (fill) bvh_brute_kernel<KIND, true><<<g, 256, 0, st>>>(d_aabbs, n, d_queries, ntiles, d_counts, d_offsets, d_values, d_payload); \
    else bvh_brute_kernel<KIND, false><<<g, 256, 0, st>>>(d_aabbs, n, d_queries, ntiles, d_counts, d_offsets, d_values, d_payload);
    if (kind == BQ_AABB) { B3_B(BQ_AABB) }
    else if (kind == BQ_RAY) { B3_B(BQ_RAY) }
    else { B3_B(BQ_FRUSTUM) }
#undef B3_B
    return 1;
}

// query_ray_closest (dbvt/util.rs:77-103): t = (point - origin) . direction over the leaves the ray hits; the smallest
// t wins. Two properties of the reference make its answer depend on its (randomised) tree shape: it keeps the FIRST leaf
// reaching the minimum in traversal order, and RayClosestVisitor (:46-62) prunes a branch when the branch bound's own t
// is not below the current minimum — but for a bound that CONTAINS the ray origin the slab test returns the EXIT point
// (aabb3.rs:190-193), so such a branch can be pruned although it holds closer leaves. The device returns the
// shape-independent answer: the minimum of the reference's per-leaf t over every leaf the ray hits (what the
// reference returns whenever its pruning does not misfire), exact-t ties resolved to the lowest value index. A branch
// is pruned only on a true lower bound (its entry t when the origin is outside it).
__global__ void __launch_bounds__(BT) bvh_ray_closest_kernel(BvhDev T, const float* __restrict__ rays, uint32_t nq, uint32_t* __restrict__ out_value,
                                                             float* __restrict__ out_point_t) {
    uint32_t qi = blockIdx.x * BT + threadIdx.x;
    if (qi >= nq) return;
    const float* b = rays + 6 * (size_t)qi;
    V3 o = v3(b[0], b[1], b[2]), d = v3(b[3], b[4], b[5]);
    float best_t = INFINITY;
    uint32_t best_v = 0xFFFFFFFFu;
    V3 best_p = v3(0.f, 0.f, 0.f);
    bvh_traverse(T, [&](bool leaf, uint32_t v, float4 mn, float4 mx) {
        V3 p;
        bool entered_from_outside;
        if (!ray_aabb(o, d, mn, mx, p, &entered_from_outside)) return false;
        float t = dot(p - o, d);
        if (leaf) {
            if (T.spheres) {  // sphere bounds: the leaf's own Continuous<Ray> point (t < 0 when the origin is inside the sphere, as in the reference)
                Query q;
                q.o = o; q.d = d;
                int rel = 0;
                if (!sphere_leaf<BQ_RAY>(q, __ldg(T.spheres + v), p, rel)) return false;
                t = dot(p - o, d);
            }
            if (t < best_t || (t == best_t && v < best_v)) { best_t = t; best_v = v; best_p = p; }
            return false;
        }
        // (a node of a sphere tree is a box that covers its spheres: its entry t bounds theirs from below when the origin is outside it)
        return !(entered_from_outside && t > best_t);
    });
    out_value[qi] = best_v;
    float* op = out_point_t + 4 * (size_t)qi;
    op[0] = best_p.x; op[1] = best_p.y; op[2] = best_p.z; op[3] = best_t;
}

int launch_bvh_ray_closest(cudaStream_t st, const BvhDev& T, const float* d_rays, uint32_t nq, uint32_t* d_value, float* d_point_t) {
    if (nq == 0) return 0;
    bvh_ray_closest_kernel<<<(nq + BT - 1) / BT, BT, 0, st>>>(T, d_rays, nq, d_value, d_point_t);
    return 1;
}

// DbvtBroadPhase::find_collider_pairs (broad_phase/dbvt.rs:26-63): every dirty value queries the tree with its own
// bound (DiscreteVisitor); pairs are ordered (low, high) and unique. A pair of two dirty values is found from both
// sides — it is emitted from the lower index only; the caller sorts the keys, which gives the reference's sorted list.
// dirty == nullptr: every value is dirty (used by the BVH strategy of sweep-and-prune, where value index = sorted slot
// and keys are packed (high << key_shift | low) to sort into the reference's (j, i) emission order).
template <bool FILL>
__global__ void __launch_bounds__(BT) bvh_collider_pairs_kernel(BvhDev T, const float* __restrict__ aabbs, const uint8_t* __restrict__ dirty,
                                                                uint32_t v_begin, uint32_t v_end, uint32_t* __restrict__ counts,
                                                                const unsigned long long* __restrict__ offsets,
                                                                unsigned long long* __restrict__ keys, int key_shift) {
    uint32_t v = blockIdx.x * BT + threadIdx.x;
    if (v >= T.n) return;
    uint32_t cnt = 0;
    // dirty == nullptr (sweep-and-prune): a pair is emitted from its HIGHER slot, and only the slots of the shard
    // [v_begin, v_end) emit — shards then partition the (j, i)-ordered pair list into consecutive pieces.
    if (dirty ? dirty[v] != 0 : (v >= v_begin && v < v_end)) {
        const float* b = aabbs + 6 * (size_t)v;
        Query q;
        q.qlo = make_float4(b[0], b[1], b[2], 0.f); q.qhi = make_float4(b[3], b[4], b[5], 0.f);
        unsigned long long w = FILL ? offsets[v] : 0ull;
        bvh_traverse(T, [&](bool leaf, uint32_t h, float4 mn, float4 mx) {
            V3 pt; int rel;
            if (!accept<BQ_AABB>(q, mn, mx, pt, rel)) return false;
            if (leaf && T.spheres) {  // sphere bounds: Discrete<Sphere> decides (the cover boxes only route the traversal)
                const float4 a = __ldg(T.spheres + v), c = __ldg(T.spheres + h);
                const V3 dist = v3(c.x, c.y, c.z) - v3(a.x, a.y, a.z);   // bound(h).intersects(bound(v)): self = the tree value
                const float radiuses = c.w + a.w;
                if (!(dot(dist, dist) <= radiuses * radiuses)) return true;
            }
            if (leaf && h != v && (dirty ? (!dirty[h] || v < h) : h < v)) {
                if (FILL) keys[w++] = key_shift ? (((unsigned long long)max(v, h) << key_shift) | min(v, h))
                                                : (((unsigned long long)min(v, h) << 32) | max(v, h));
                ++cnt;
            }
            return true;
        });
    }
    if (!FILL) counts[v] = cnt;
}

// Sweep-and-prune over the sorted slots, BVH strategy, in ONE pass: the slots of the shard [v_begin, v_end) query the tree
// over the sorted boxes and append every overlapping pair (h < v) to the key list through a warp-aggregated atomic. The
// keys are radix-sorted into the reference's (j, i) order afterwards anyway, so their emission order is free — which is
// what lets this skip the count + scan + fill structure the DBVT queries need. *count keeps counting past `capacity`.
__global__ void __launch_bounds__(BT) bvh_overlap_append_kernel(BvhDev T, const float* __restrict__ aabbs, uint32_t v_begin, uint32_t v_end,
                                                                int key_shift, unsigned long long* __restrict__ keys,
                                                                unsigned long long capacity, unsigned long long* __restrict__ count) {
    // threads walk the values in the tree's leaf (Morton) order, so that the queries of a warp are neighbours in space
    // and descend through the same nodes
    const uint32_t slot = blockIdx.x * BT + threadIdx.x;
    if (slot >= T.n) return;
    const uint32_t v = __ldg(T.leaf_value + slot);
    if (v < v_begin || v >= v_end) return;
    const float* b = aabbs + 6 * (size_t)v;
    Query q;
    q.qlo = make_float4(b[0], b[1], b[2], 0.f); q.qhi = make_float4(b[3], b[4], b[5], 0.f);
    const uint32_t lane = threadIdx.x & 31u;
    bvh_traverse(T, [&](bool leaf, uint32_t h, float4 mn, float4 mx) {
        V3 pt; int rel;
        if (!accept<BQ_AABB>(q, mn, mx, pt, rel)) return false;
        if (leaf && h < v) {
            const uint32_t act = __activemask();
            const uint32_t leader = __ffs(act) - 1;
            unsigned long long base = 0;
            if (lane == leader) base = atomicAdd(count, (unsigned long long)__popc(act));
            base = __shfl_sync(act, base, leader);
            const unsigned long long at = base + __popc(act & ((1u << lane) - 1u));
            if (at < capacity) keys[at] = ((unsigned long long)v << key_shift) | h;
        }
        return true;
    });
}

int launch_bvh_overlap_append(cudaStream_t st, const BvhDev& T, const float* d_aabbs, uint32_t v_begin, uint32_t v_end, int key_shift,
                              unsigned long long* d_pair_keys, uint64_t capacity, unsigned long long* d_count) {
    cudaMemsetAsync(d_count, 0, sizeof(unsigned long long), st);
    if (T.n == 0 || v_end <= v_begin) return 0;
    bvh_overlap_append_kernel<<<(T.n + BT - 1) / BT, BT, 0, st>>>(T, d_aabbs, v_begin, v_end, key_shift, d_pair_keys, capacity, d_count);
    return 1;
}

int launch_bvh_collider_pairs(cudaStream_t st, const BvhDev& T, const float* d_aabbs, const uint8_t* d_dirty, uint32_t v_begin,
                              uint32_t v_end, uint32_t* d_counts, const unsigned long long* d_offsets, unsigned long long* d_pair_keys,
                              int key_shift) {
    if (T.n == 0) return 0;
    const int g = (T.n + BT - 1) / BT;
    if (d_offsets) bvh_collider_pairs_kernel<true><<<g, BT, 0, st>>>(T, d_aabbs, d_dirty, v_begin, v_end, d_counts, d_offsets, d_pair_keys, key_shift);
    else bvh_collider_pairs_kernel<false><<<g, BT, 0, st>>>(T, d_aabbs, d_dirty, v_begin, v_end, d_counts, d_offsets, d_pair_keys, key_shift);
    return 1;
}

// SweepAndPrune3 returns indices into the slice it has just re-sorted (sweep_prune.rs:65-69,80): pair (i, j) means the shapes at
// SORTED positions i and j. The narrow phase reads the caller's (unsorted) shape set, so the glue between the two is
// out[k] = (perm[i], perm[j]) with perm[slot] = original index of the box at that sorted position.
__global__ void __launch_bounds__(BT) remap_pairs_kernel(const uint2* __restrict__ sorted_pairs, const uint32_t* __restrict__ perm, uint64_t n,
                                                         uint2* __restrict__ out) {
    for (uint64_t k = (uint64_t)blockIdx.x * BT + threadIdx.x; k < n; k += (uint64_t)gridDim.x * BT) {
        const uint2 p = sorted_pairs[k];
        out[k] = make_uint2(__ldg(perm + p.x), __ldg(perm + p.y));
    }
}
int launch_remap_pairs(cudaStream_t st, const uint2* d_sorted_pairs, const uint32_t* d_perm, uint64_t n, uint2* d_out) {
    if (n == 0) return 0;
    uint64_t g = (n + BT - 1) / BT;
    if (g > 148ull * 32) g = 148ull * 32;
    remap_pairs_kernel<<<(unsigned)g, BT, 0, st>>>(d_sorted_pairs, d_perm, n, d_out);
    return 1;
}

// ------------------------------------------------------------------------------------------------------------
// small utilities over CUB
// ------------------------------------------------------------------------------------------------------------
struct U32ToU64 { __host__ __device__ unsigned long long operator()(uint32_t v) const { return v; } };

size_t scan_tmp_bytes(uint32_t n) {
    size_t a = 0;
    cub::TransformInputIterator<unsigned long long, U32ToU64, const uint32_t*> it(nullptr, U32ToU64());
    cub::DeviceScan::ExclusiveSum(nullptr, a, it, (unsigned long long*)nullptr, (int)(n + 1));
    return a + 256;
}
// d_out gets n + 1 entries: exclusive prefix sums of d_in[0..n) and the total at d_out[n] (d_in must have n + 1 readable
// entries with d_in[n] ignored: the caller zero-pads).
int launch_exclusive_scan_u32_to_u64(cudaStream_t st, const uint32_t* d_in, unsigned long long* d_out, uint32_t n, void* tmp, size_t tmp_bytes) {
    cub::TransformInputIterator<unsigned long long, U32ToU64, const uint32_t*> it(d_in, U32ToU64());
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, it, d_out, (int)(n + 1), st);
    return 2;
}
size_t sort_u64_tmp_bytes(uint64_t n) {
    size_t a = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, a, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (long long)n);
    return a + 256;
}
int launch_sort_u64(cudaStream_t st, unsigned long long* d_keys, unsigned long long* d_alt, uint64_t n, void* tmp, size_t tmp_bytes, bool* in_alt) {
    *in_alt = false;
    if (n == 0) return 0;
    cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, d_keys, d_alt, (long long)n, 0, 64, st);
    *in_alt = true;
    return 8;
}
int launch_unpack_pairs(cudaStream_t st, const unsigned long long* d_keys, uint64_t n, uint2* d_pairs, bool hi_first) {
    if (n == 0) return 0;
    unpack_pairs_kernel<<<(unsigned)((n + BT - 1) / BT), BT, 0, st>>>(d_keys, n, 32, d_pairs, hi_first ? 1 : 0);
    return 1;
}

}  // namespace b3
