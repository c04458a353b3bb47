// Internal definitions of the opaque C-ABI handles (include/bam3d_gpu.h).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

struct bam3d_ctx;

namespace b3 { struct ShapeRec; }

struct DevBuf {  // grow-only device scratch owned by a ctx
    void* ptr = nullptr;
    size_t bytes = 0;
    int32_t reserve(bam3d_ctx* ctx, size_t need);
    void release();
};

struct bam3d_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    bool opt_binning = true;
    bool opt_epa_tiers = false;   // contact path: shared-memory EPA in size tiers (epa_tiers.cuh) instead of per-thread local-memory EPA
    bool opt_intersect_split = false;  // intersect, thread path: first-support test and the GJK loop as two launches (measured slower on C1: 0.176 vs 0.167 ms)
    int opt_intersect_refill = 0;  // intersect, thread path: k >= 1 = persistent kernel whose lanes refill once k of a warp are idle (gjk_intersect_refill_kernel)
    int opt_toi_refill = 8;  // time of impact, thread path: 0 = one pair per thread (1.16 ms per 2^20 mixed pairs); k >= 1 = persistent kernel whose lanes refill once k of a warp are idle (gjk_toi_refill_kernel: 0.56 ms at k = 8)
    int opt_distance_refill = 8;  // distance, thread path: the same lane-refill scheme (gjk_distance_refill_kernel); 0 = one pair per thread
    bool opt_epa_global_polytopes = true;   // thread-per-pair EPA: polytopes in a global slab (contiguous per thread) instead of local memory
    bool opt_epa_pipeline = false;  // thread-per-pair EPA: cp.async face ring + batched walk / moves / new faces (epa_pipe.cuh); 0 = epa.cuh's iteration
    bool opt_epa_huge = true;  // contact path: pairs whose polytope outgrows the 1024-face overflow tier go to the block-per-pair tier (epa_huge.cuh, 2^18 faces) instead of ending as BAM3D_STATUS_CAPACITY
    bool opt_epa_overlap = true;  // contact path: run the EPA overflow tier on the side stream next to the EPA kernel
    int opt_sap_strategy = 0;  // 0 auto, 1 tiled sweep, 2 BVH overlap query
    int opt_bvh_strategy = 0;  // variable-length queries: 0 auto, 1 tree traversal, 2 all (query, value) tests
    cudaStream_t stream2 = nullptr;  // side stream (Variance3 runs beside the pair search)
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    uint32_t* d_flags = nullptr;   // device flag word of this ctx (bit 0: a pair named a shape index out of range)
    uint32_t* h_flags = nullptr;   // pinned copy
    // opt-in timing of the dominant kernel of each narrow-phase call (CUDA events on the ctx stream)
    bool opt_timing = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing_events;
    // scratch for the host-buffer entry points and the binning pass
    DevBuf pairs, status, contacts, f0, f1, keys, order, small, epa_overflow, epa_queue, epa_slab, epa_huge;
    // broad-phase scratch
    DevBuf bp0, bp1, bp2, bp3, bp4, bp5, bp6, bp7, cub_tmp, sph0, sph1, sph2, pipe0, pipe1, pipe2, pipe3;
    void release_scratch() {
        DevBuf* all[] = {&pairs, &status, &contacts, &f0, &f1, &keys, &order, &small, &epa_overflow, &epa_queue, &epa_slab, &epa_huge, &bp0, &bp1, &bp2, &bp3, &bp4, &bp5, &bp6, &bp7, &cub_tmp, &sph0, &sph1, &sph2, &pipe0, &pipe1, &pipe2, &pipe3};
        for (DevBuf* b : all) b->release();
    }
};

struct bam3d_meshes {
    float4* verts = nullptr;      // per mesh: MESH_HEADER_SLOTS header slots (support.cuh: MeshHeader), then xyz + unused w per vertex
    uint32_t* offsets = nullptr;  // n_meshes + 1, in float4 slots of `verts` (headers included)
    uint32_t n_meshes = 0;
    uint64_t nverts = 0;
    // half-edge structures of the meshes uploaded with faces (ConvexPolyhedron::new_with_faces); null when there are none
    uint32_t* vert_edge = nullptr;
    uint4* edges = nullptr;
    uint4* faces = nullptr;
    uint64_t n_edges = 0, n_faces = 0;
};

struct bam3d_shapes {
    b3::ShapeRec* recs = nullptr;
    uint32_t* meta = nullptr;
    // inputs kept on the device so that set_transforms can re-pack without a re-upload of kinds / params
    uint8_t* d_kind = nullptr;
    float* d_params = nullptr;
    float* d_xform = nullptr;
    uint32_t* d_mesh_id = nullptr;
    uint32_t n = 0;
    const bam3d_meshes* meshes = nullptr;
    uint32_t kind_hist[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    uint32_t* err_flag = nullptr;  // where the kernels flag an out-of-range pair index (null: the calling ctx's own word); frame slots set their own
};

int32_t b3_fail(bam3d_ctx* ctx, int32_t code, const char* what, cudaError_t e);
