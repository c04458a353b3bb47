// Host-callable launchers of the broad-phase kernels (implemented in broad.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace b3 {

struct ShapeSetDev;

// ---- sweep and prune ------------------------------------------------------------------------------------------
struct SapScratch {
    unsigned long long* keys_in;   // n
    unsigned long long* keys_out;  // n
    uint32_t* vals_in;             // n
    uint32_t* vals_out;            // n  (== perm)
    float4* lo;                    // n  sorted boxes: min xyz
    float4* hi;                    // n  sorted boxes: max xyz
    void* cub_tmp;
    size_t cub_tmp_bytes;
    unsigned long long* pair_keys; // capacity (j << 32 | i), unsorted then sorted in place via pair_keys_alt
    unsigned long long* pair_keys_alt;
};
size_t sap_cub_tmp_bytes(uint32_t n, uint64_t pair_capacity);
int sap_bits_for(uint32_t n);
// phase 1: keys, stable radix sort, gather of the sorted boxes (+ optional 6-float rows), candidate-test count
int launch_sap_sort(cudaStream_t st, const float* d_aabbs, uint32_t n, int axis, const SapScratch& S, uint32_t* d_perm, float* d_sorted6,
                    unsigned long long* d_candidates);
// 6-float rows -> the (lo, hi) float4 arrays the variance kernels read, order unchanged
int launch_boxes_to_lohi(cudaStream_t st, const float* d_aabbs, uint32_t n, float4* lo, float4* hi);
// Variance3 in the order of S.lo / S.hi (six chains; bit-identical to the sequential loop); d_sums: 6 floats; d_next_axis: device i32
// d_scratch: sap_variance_scratch_bytes(n) (the six planar value streams)
size_t sap_variance_scratch_bytes(uint32_t n);
int launch_sap_variance(cudaStream_t st, const SapScratch& S, uint32_t n, void* d_scratch, float* d_sums, int* d_next_axis);
// phase 2a: tiled forward sweep; d_count: device u64 (pairs found, may exceed capacity)
// Only pairs whose higher sorted slot j lies in [j_begin, j_end) are emitted (the multi-GPU shard of the sweep).
int launch_sap_sweep(cudaStream_t st, uint32_t n, int axis, uint32_t j_begin, uint32_t j_end, const SapScratch& S, uint64_t capacity,
                     unsigned long long* d_count);
// order the first `count` pair keys by (j, i) and unpack them into d_pairs
int launch_sap_finish(cudaStream_t st, const SapScratch& S, uint64_t count, uint2* d_pairs, uint32_t n);

// ---- world AABBs of packed shapes (ComputeBound<Aabb3> + Aabb::transform) ---------------------------------------
int launch_world_aabbs(cudaStream_t st, const ShapeSetDev& S, float* d_out6);
// ComputeBound<volume::Sphere> + transform_volume per shape: n x (centre.xyz, radius)
int launch_world_spheres(cudaStream_t st, const ShapeSetDev& S, float4* d_out);
// SweepAndPrune3<volume::Sphere>: extents of the spheres as boxes (sort keys, Variance3), growth of the sorted boxes for the
// candidate search, spheres in sorted order, and the exact order-preserving pair filter
int launch_spheres_to_extents(cudaStream_t st, const float4* d_spheres, uint32_t n, float* d_boxes6);
// the same grown by grow_margin: the node boxes of a BVH over sphere bounds
int launch_sphere_cover_boxes(cudaStream_t st, const float4* d_spheres, uint32_t n, float* d_boxes6);
int launch_sap_inflate(cudaStream_t st, const SapScratch& S, float* d_sorted6, uint32_t n);
int launch_gather_spheres(cudaStream_t st, const float4* d_spheres, const uint32_t* d_perm, uint32_t n, float4* d_sorted);
size_t sphere_filter_tmp_bytes(uint64_t n_pairs);
int launch_sphere_pair_filter(cudaStream_t st, const float4* d_sorted_spheres, int axis, const uint2* d_in, uint64_t n_pairs, uint2* d_out,
                              unsigned long long* d_count, void* d_tmp, size_t tmp_bytes);

// ---- static BVH answering the DBVT queries ---------------------------------------------------------------------
constexpr uint32_t BVH_END = 0xFFFFFFFFu;  // rope of the last node in traversal order
struct BvhDev {
    // n leaves (values), n - 1 internal nodes. Node ids: internal 0..n-2 (root = 0), leaf k -> (n - 1) + k.
    uint32_t n;
    const float4* nodes;     // 2 x (2n - 1): (min.xyz, A), (max.xyz, rope) as int bits — one 32-byte sector per node; A = left child
                             // (internal) or value index (leaf); rope = where a left-first depth-first traversal goes after this subtree
    const uint32_t* leaf_value;  // n : value index of leaf k (Morton order)
    const float4* spheres = nullptr;  // n x (centre.xyz, radius) when the values' bounds are volume::Sphere: the node boxes are then a
                                      // conservative cover (grown sphere extents) and every leaf is decided by the sphere's own predicate
};
struct BvhBuildScratch {
    uint32_t* morton_in; uint32_t* morton_out; uint32_t* idx_in; uint32_t* idx_out;
    void* cub_tmp; size_t cub_tmp_bytes;
    uint32_t* parent;   // 2n - 1
    uint32_t* flags;    // n - 1
    float* scene;       // 6 floats: scene bounds
};
size_t bvh_cub_tmp_bytes(uint32_t n);
// rope: 2n - 1 words, kept with the tree (the refit needs them again)
int launch_bvh_build(cudaStream_t st, const float* d_aabbs, uint32_t n, const BvhBuildScratch& W, float4* nodes, int2* children,
                     uint32_t* leaf_value, uint32_t* rope);

// bounds of an existing tree recomputed from new values (same topology); parent: 2n - 1 words, flags: n - 1 words of scratch
int launch_bvh_refit(cudaStream_t st, const float* d_aabbs, uint32_t n, const int2* children, const uint32_t* leaf_value, uint32_t* parent,
                     const uint32_t* rope, uint32_t* flags, float4* nodes);

enum BvhQueryKind { BQ_AABB = 0, BQ_RAY = 1, BQ_FRUSTUM = 2, BQ_SPHERE = 3 };  // BQ_SPHERE: nq x 4 f32 (centre, radius), sphere trees only
// pass 1 (d_offsets == nullptr): per-query hit counts into d_counts. pass 2: fill d_values / d_payload at d_offsets.
//   BQ_AABB: queries nq x 6 f32, no payload;  BQ_RAY: nq x 6 f32 (origin, direction), payload 3 f32 (hit point) per hit;
//   BQ_FRUSTUM: nq x 24 f32 (6 planes n.xyz,d in the order left,right,bottom,top,near,far), payload 1 byte (Relation).
int launch_bvh_query(cudaStream_t st, const BvhDev& T, int kind, const float* d_queries, uint32_t nq, uint32_t* d_counts,
                     const unsigned long long* d_offsets, uint32_t* d_values, void* d_payload);
// all (query, value) tests, for few large queries: d_counts / d_offsets are indexed [query * ntiles + tile], ntiles = ceil(n / 256)
int launch_bvh_brute(cudaStream_t st, const float* d_aabbs, uint32_t n, int kind, const float* d_queries, uint32_t nq, uint32_t* d_counts,
                     const unsigned long long* d_offsets, uint32_t* d_values, void* d_payload);
// one pass, device-resident: hits appended as (query, value, payload) in arbitrary order; *d_count (device u64) keeps counting
// past `capacity`. _append: tree traversal, one thread per query; _brute_append: every (query, value) test, for few large queries.
int launch_bvh_query_append(cudaStream_t st, const BvhDev& T, int kind, const float* d_queries, uint32_t nq, uint32_t* d_out_query,
                            uint32_t* d_out_value, void* d_out_payload, uint64_t capacity, unsigned long long* d_count);
int launch_bvh_brute_append(cudaStream_t st, const float* d_aabbs, uint32_t n, int kind, const float* d_queries, uint32_t nq, uint32_t* d_out_query,
                            uint32_t* d_out_value, void* d_out_payload, uint64_t capacity, unsigned long long* d_count);
int launch_bvh_ray_closest(cudaStream_t st, const BvhDev& T, const float* d_rays, uint32_t nq, uint32_t* d_value, float* d_point_t);
// DbvtBroadPhase::find_collider_pairs: count / fill passes over the dirty values.
// d_dirty == nullptr (sweep-and-prune over sorted slots): pairs are emitted from their higher slot, for the slots in
// [v_begin, v_end) only.
int launch_bvh_collider_pairs(cudaStream_t st, const BvhDev& T, const float* d_aabbs, const uint8_t* d_dirty, uint32_t v_begin,
                              uint32_t v_end, uint32_t* d_counts, const unsigned long long* d_offsets, unsigned long long* d_pair_keys,
                              int key_shift);
// sweep-and-prune, BVH strategy: one pass, pairs (h < v) of the slots v in [v_begin, v_end) appended to d_pair_keys as
// (v << key_shift | h) in arbitrary order; *d_count (device u64) = pairs found (may exceed capacity)
int launch_bvh_overlap_append(cudaStream_t st, const BvhDev& T, const float* d_aabbs, uint32_t v_begin, uint32_t v_end, int key_shift,
                              unsigned long long* d_pair_keys, uint64_t capacity, unsigned long long* d_count);
// sorted-position pairs -> shape-index pairs through perm (the glue between SweepAndPrune3's output and the narrow phase)
int launch_remap_pairs(cudaStream_t st, const uint2* d_sorted_pairs, const uint32_t* d_perm, uint64_t n, uint2* d_out);
int launch_exclusive_scan_u32_to_u64(cudaStream_t st, const uint32_t* d_in, unsigned long long* d_out, uint32_t n, void* tmp, size_t tmp_bytes);
size_t scan_tmp_bytes(uint32_t n);
int launch_sort_u64(cudaStream_t st, unsigned long long* d_keys, unsigned long long* d_alt, uint64_t n, void* tmp, size_t tmp_bytes, bool* in_alt);
size_t sort_u64_tmp_bytes(uint64_t n);
int launch_unpack_pairs(cudaStream_t st, const unsigned long long* d_keys, uint64_t n, uint2* d_pairs, bool hi_first);

}  // namespace b3
