// Host-callable launchers of the narrow-phase kernels (implemented in narrow.cu).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

namespace b3 {

struct ShapeRec;

// Device view of a packed shape set (bam3d_shapes) and its mesh library.
struct ShapeSetDev {
    const ShapeRec* recs;          // n x 96 B
    const uint32_t* meta;          // n : kind | mesh_id << 8
    const float4* mesh_verts;      // polyhedron vertices (xyz, w unused)
    const uint32_t* mesh_offsets;  // n_meshes + 1
    uint32_t n;
    uint32_t* err;                 // device flag word: bit 0 is raised by a pair that names a shape index >= n (the index is clamped, so
                                   // the kernels never read out of bounds; the entry points turn the flag into BAM3D_ERR_ARG)
};

struct SettingsDev {
    float distance_tolerance, continuous_tolerance, epa_tolerance;
    uint32_t max_iterations, toi_iteration_cap;
};

// Scratch the binning pass needs (owned by the ctx, sized for n pairs).
struct BinScratch {
    uint8_t* keys;     // n
    uint32_t* order;   // n : pair indices, thread-path pairs grouped by (kindL,kindR) first, polyhedron pairs last
    uint32_t* hist;    // 64 : per-bin counts
    uint32_t* start;   // 64 : per-bin start offsets
    uint32_t* cursor;  // 64 : per-bin running cursors
    uint32_t* ranges;  // 4  : [0]=thread begin, [1]=thread end, [2]=warp begin, [3]=warp end
};

// Each launcher returns the number of kernels it enqueued (for the launch counter); errors via cudaGetLastError.
int launch_pack_shapes(cudaStream_t st, const uint8_t* d_kind, const float* d_params, const float* d_xform,
                       const uint32_t* d_mesh_id, uint32_t n, uint32_t n_meshes, ShapeRec* recs, uint32_t* meta,
                       uint32_t* d_err, bool affine48 = false);
// planes of the half-edge meshes' faces (2 x uint4 per face: indices, then vertex-slot base -> plane); *d_err != 0: degenerate face
int launch_mesh_face_planes(cudaStream_t st, const float4* d_verts, uint4* d_faces, uint64_t nfaces, uint32_t* d_err);
int launch_bin_pairs(cudaStream_t st, const ShapeSetDev& S, const uint2* d_pairs, uint32_t n, const BinScratch& B);

// B may be null (no binning: the selected kernel covers every pair). run_thread / run_warp pick the execution shapes.
// d_scratch: intersect_scratch_bytes(n) of device memory for the survivor list of the two-launch thread path, or null to
// run the single-kernel thread path. refill_min > 0 (with d_scratch: 4 bytes suffice): the thread path runs as a persistent
// kernel whose lanes take new pairs once `refill_min` lanes of a warp are idle ("intersect_refill").
size_t intersect_scratch_bytes(uint32_t n);
int launch_intersect(cudaStream_t st, const ShapeSetDev& S, const uint2* d_pairs, uint32_t n, const BinScratch* B,
                     bool run_thread, bool run_warp, const SettingsDev& cfg, void* d_scratch, int refill_min, uint8_t* d_status);
// d_overflow_scratch: contact_overflow_scratch_bytes() of device memory for the EPA overflow tier (epa_big.cuh);
// d_queue: contact_queue_bytes(n) for the GJK -> EPA hit queue of the thread-per-pair path.
size_t contact_overflow_scratch_bytes();
size_t contact_queue_bytes(uint32_t n);
// default_key: the (kindL * 7 + kindR) bin of every pair when B is null (a single-kind shape set), else ignored.
// epa_mode: 1 = shared-memory EPA in size tiers (epa_tiers.cuh), 0 = per-thread EPA (epa.cuh).
// epa_pipeline: the thread-per-pair EPA iteration with its memory round trips overlapped (epa_pipe.cuh; needs the slab).
// side (nullable): a second stream and two events; when given, the EPA overflow tier runs on it concurrently with the
// thread-per-pair EPA kernel and is joined back into `st` before the call returns.
struct ContactSide { cudaStream_t stream; cudaEvent_t fork, join; };
int launch_contact(cudaStream_t st, const ShapeSetDev& S, const uint2* d_pairs, uint32_t n, const BinScratch* B,
                   bool run_thread, bool run_warp, const SettingsDev& cfg, int strategy, uint32_t default_key, int epa_mode,
                   bool epa_pipeline, const ContactSide* side, void* d_poly_slab, void* d_overflow_scratch, void* d_huge_slab, void* d_queue, float4* d_contacts,
                   uint8_t* d_status);
// d_huge_slab (nullable): contact_huge_slab_bytes() for the third EPA tier (epa_huge.cuh: polytopes beyond 1024 faces, one block per pair);
// without it such pairs keep BAM3D_STATUS_CAPACITY.
size_t contact_huge_slab_bytes();
// d_poly_slab (nullable): contact_poly_slab_bytes() of device memory; when given, the thread-per-pair EPA keeps its polytopes
// there (one contiguous record per thread) instead of in per-thread local memory.
size_t contact_poly_slab_bytes();
// d_cursor / refill_min: as launch_toi ("distance_refill").
int launch_distance(cudaStream_t st, const ShapeSetDev& S, const uint2* d_pairs, uint32_t n, const BinScratch* B,
                    bool run_thread, bool run_warp, const SettingsDev& cfg, void* d_cursor, int refill_min, float* d_ref, float* d_geo,
                    uint8_t* d_status);
// d_cursor (nullable): 4 bytes of device scratch; when given, the thread path runs as a persistent kernel whose lanes take a new
// pair when theirs is finished ("toi_refill"; `refill_min` = how many lanes of a warp wait before they fetch together), instead of
// one pair per thread.
int launch_toi(cudaStream_t st, const ShapeSetDev& S0, const ShapeSetDev& S1, const uint2* d_pairs, uint32_t n,
               const BinScratch* B, bool run_thread, bool run_warp, const SettingsDev& cfg, void* d_cursor, int refill_min,
               float4* d_contacts, uint8_t* d_status);
// Ray casts against primitives: d_rays = n_rays x 6 f32 (origin, direction; query 2: segment start, end), d_casts = n x
// (ray index, shape index); query 0 = intersects_transformed, 1 = intersection_transformed, 2 = swept particle.
int launch_ray_cast(cudaStream_t st, const ShapeSetDev& S, const float* d_rays, const uint2* d_casts, uint64_t n, int query, float* d_points,
                    uint8_t* d_status);
// compound shapes (gjk/mod.rs:362-523): world transforms of the primitives, the primitive-pair expansion, the reductions
int launch_compound_compose(cudaStream_t st, const float* d_comp_xform, const float* d_local, const uint32_t* d_owner, uint32_t n_prims, float* d_out);
int launch_compound_expand(cudaStream_t st, const uint2* d_comp_pairs, const unsigned long long* d_seg, uint32_t n_pairs, uint64_t total,
                           const uint32_t* d_offsets, uint2* d_flat);
// mode 0 intersection_complex, 1 distance_complex, 2 intersection_complex_time_of_impact
int launch_compound_reduce(cudaStream_t st, int mode, int strategy, const unsigned long long* d_seg, uint32_t n_pairs, const float4* d_contacts,
                           const float* d_ref, const uint8_t* d_status, float4* d_out_contacts, float* d_out_value, uint8_t* d_out_status);
int launch_compact_contacts(cudaStream_t st, const float4* d_contacts, const uint8_t* d_status, uint64_t n,
                            uint64_t base, void* d_out, uint64_t capacity, unsigned long long* d_count);

}  // namespace b3
