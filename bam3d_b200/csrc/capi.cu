// C ABI (include/bam3d_gpu.h) over the CUDA kernels: context, device-resident shape sets, and the narrow-phase
// entry points. Host layer duties (north_star subsystem 1): take primitives + glam Mat4 in the caller's host
// arrays, pack them into device buffers, run the kernels on the ctx stream, copy results back.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "../../include/bam3d_gpu.h"
#include "vecmath.cuh"
#include "support.cuh"
#include "narrow.h"
#include "broad.h"
#include "ctx.h"

using namespace b3;

// ------------------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------------------
int32_t b3_fail(bam3d_ctx* ctx, int32_t code, const char* what, cudaError_t e) {
    if (ctx) {
        char buf[512];
        if (e != cudaSuccess) std::snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
        else std::snprintf(buf, sizeof buf, "%s", what);
        ctx->err = buf;
    }
    return code;
}

int32_t DevBuf::reserve(bam3d_ctx* ctx, size_t need) {
    if (need <= bytes) return BAM3D_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; bytes = 0;
    size_t want = need + need / 4 + 256;
    cudaError_t e = cudaMalloc(&ptr, want);
    if (e != cudaSuccess) { ptr = nullptr; return b3_fail(ctx, BAM3D_ERR_OOM, "cudaMalloc(scratch)", e); }
    bytes = want;
    return BAM3D_OK;
}
void DevBuf::release() { if (ptr) cudaFree(ptr); ptr = nullptr; bytes = 0; }

static SettingsDev to_dev(const bam3d_gjk_settings* s) {
    bam3d_gjk_settings d;
    bam3d_default_settings(&d);
    if (s) d = *s;
    SettingsDev r;
    r.distance_tolerance = d.distance_tolerance; r.continuous_tolerance = d.continuous_tolerance;
    r.epa_tolerance = d.epa_tolerance; r.max_iterations = d.max_iterations;
    r.toi_iteration_cap = d.toi_iteration_cap ? d.toi_iteration_cap : 100u;
    return r;
}

static ShapeSetDev view(const bam3d_ctx* ctx, const bam3d_shapes* s) {
    ShapeSetDev v;
    v.recs = s->recs; v.meta = s->meta; v.n = s->n;
    v.err = s->err_flag ? s->err_flag : ctx->d_flags;
    v.mesh_verts = s->meshes ? s->meshes->verts : nullptr;
    v.mesh_offsets = s->meshes ? s->meshes->offsets : nullptr;
    return v;
}

// Decide execution shapes for a shape set: which kernels run and whether pairs are binned first.
struct Plan { bool bin, run_thread, run_warp; };
static Plan plan_for(const bam3d_ctx* ctx, const bam3d_shapes* s) {
    int analytic_kinds = 0;
    for (int k = 0; k < 6; ++k) analytic_kinds += s->kind_hist[k] ? 1 : 0;
    bool has_poly = s->kind_hist[BAM3D_KIND_POLYHEDRON] != 0;
    Plan p;
    if (!has_poly) { p.run_thread = true; p.run_warp = false; p.bin = analytic_kinds > 1; }
    else if (analytic_kinds == 0) { p.run_thread = false; p.run_warp = true; p.bin = false; }
    else { p.run_thread = true; p.run_warp = true; p.bin = true; }
    if (!ctx->opt_binning && !(p.run_thread && p.run_warp)) p.bin = false;
    return p;
}

static int32_t prepare_bins(bam3d_ctx* ctx, const Plan& plan, const ShapeSetDev& S, const uint2* d_pairs, uint32_t n,
                            BinScratch& B, const BinScratch** out) {
    *out = nullptr;
    if (!plan.bin) return BAM3D_OK;
    int32_t rc;
    if ((rc = ctx->keys.reserve(ctx, n)) || (rc = ctx->order.reserve(ctx, (size_t)n * 4)) || (rc = ctx->small.reserve(ctx, 1024))) return rc;
    B.keys = (uint8_t*)ctx->keys.ptr; B.order = (uint32_t*)ctx->order.ptr;
    uint32_t* sm = (uint32_t*)ctx->small.ptr;
    B.hist = sm; B.start = sm + 64; B.cursor = sm + 128; B.ranges = sm + 192;
    ctx->launches += launch_bin_pairs(ctx->stream, S, d_pairs, n, B);
    *out = &B;
    return BAM3D_OK;
}

#define CU_TRY(call, what)                                                         \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) return b3_fail(ctx, BAM3D_ERR_CUDA, what, e_);      \
    } while (0)

static int32_t check_launch(bam3d_ctx* ctx, const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return b3_fail(ctx, BAM3D_ERR_CUDA, what, e);
    return BAM3D_OK;
}

// Bracket the dominant kernel(s) of a call with events when the "timing" option is on.
struct KernelTimer {
    bam3d_ctx* ctx; cudaEvent_t a = nullptr, b = nullptr;
    explicit KernelTimer(bam3d_ctx* c) : ctx(c) {
        if (!ctx->opt_timing) return;
        cudaEventCreate(&a); cudaEventCreate(&b);
        cudaEventRecord(a, ctx->stream);
    }
    ~KernelTimer() {
        if (!a) return;
        cudaEventRecord(b, ctx->stream);
        ctx->timing_events.emplace_back(a, b);
    }
};

// the kernels walk batches with 32-bit grid-stride counters (strides up to 148 x 32 x 128): stay 2^24 below 2^32
static int32_t check_n(bam3d_ctx* ctx, uint64_t n) {
    if (n > 0xFF000000ull) return b3_fail(ctx, BAM3D_ERR_ARG, "batch too large: n must be <= 2^32 - 2^24 items per call", cudaSuccess);
    return BAM3D_OK;
}
static int32_t check_pairs_over(bam3d_ctx* ctx, const bam3d_shapes* shapes, uint64_t n) {
    if (n && shapes->n == 0) return b3_fail(ctx, BAM3D_ERR_ARG, "pairs over an empty shape set", cudaSuccess);
    return BAM3D_OK;
}

// After a synchronisation point: has a kernel of this ctx met a pair whose shape index is out of range? (The index was clamped,
// nothing was read out of bounds, but that pair's result is meaningless.) Reads and clears the flag.
static int32_t take_index_flag(bam3d_ctx* ctx, const char* what) {
    cudaError_t e = cudaMemcpyAsync(ctx->h_flags, ctx->d_flags, 4, cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return b3_fail(ctx, BAM3D_ERR_CUDA, what, e);
    if (*ctx->h_flags & 1u) {
        cudaMemsetAsync(ctx->d_flags, 0, 4, ctx->stream);
        char msg[192];
        std::snprintf(msg, sizeof msg, "%s: a pair names a shape index >= the shape set's size", what);
        return b3_fail(ctx, BAM3D_ERR_ARG, msg, cudaSuccess);
    }
    return BAM3D_OK;
}

extern "C" {

// ------------------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------------------
void bam3d_default_settings(bam3d_gjk_settings* out) {
    out->distance_tolerance = 0.000001f;
    out->continuous_tolerance = 0.000001f;
    out->epa_tolerance = 0.00001f;
    out->max_iterations = 100;
    out->toi_iteration_cap = 100;
}

// The side stream carries the EPA overflow tiers: the longest jobs of a contact call, few blocks, each needing most of an SM.
// Default priority. (BAM3D_SIDE_PRIORITY=1 creates it at the highest priority so that its blocks are placed first whenever a
// multiprocessor drains: measured, no gain on C2 / C5 with two compute lanes — and the overflow consumer SPINS until its
// producers have exited, so favouring it over the producers of other lanes is the wrong way round.)
static cudaError_t create_side_stream(cudaStream_t* s) {
    int least = 0, greatest = 0;
    bool high = false;
    if (const char* e = std::getenv("BAM3D_SIDE_PRIORITY")) high = std::atoi(e) != 0;
    if (!high || cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking);
    return cudaStreamCreateWithPriority(s, cudaStreamNonBlocking, greatest);
}

int32_t bam3d_ctx_create(int32_t device_ordinal, bam3d_ctx** out_ctx) {
    if (!out_ctx) return BAM3D_ERR_ARG;
    *out_ctx = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device_ordinal < 0 || device_ordinal >= count) return BAM3D_ERR_CUDA;  // no CPU fallback
    if (cudaSetDevice(device_ordinal) != cudaSuccess) return BAM3D_ERR_CUDA;
    bam3d_ctx* ctx = new (std::nothrow) bam3d_ctx();
    if (!ctx) return BAM3D_ERR_OOM;
    ctx->device = device_ordinal;
#ifdef B3_EXPERIMENTS
    if (const char* e = std::getenv("BAM3D_EPA_TIERS")) ctx->opt_epa_tiers = std::atoi(e) != 0;      // development A/B switches
    if (const char* e = std::getenv("BAM3D_EPA_PIPELINE")) ctx->opt_epa_pipeline = std::atoi(e) != 0;
    if (const char* e = std::getenv("BAM3D_INTERSECT_SPLIT")) ctx->opt_intersect_split = std::atoi(e) != 0;
    if (const char* e = std::getenv("BAM3D_INTERSECT_REFILL")) ctx->opt_intersect_refill = std::atoi(e);
#endif
    if (const char* e = std::getenv("BAM3D_EPA_OVERLAP")) ctx->opt_epa_overlap = std::atoi(e) != 0;
    if (const char* e = std::getenv("BAM3D_EPA_HUGE")) ctx->opt_epa_huge = std::atoi(e) != 0;
    if (const char* e = std::getenv("BAM3D_EPA_GLOBAL")) ctx->opt_epa_global_polytopes = std::atoi(e) != 0;
    if (const char* e = std::getenv("BAM3D_TOI_REFILL")) ctx->opt_toi_refill = std::atoi(e);
    if (const char* e = std::getenv("BAM3D_DISTANCE_REFILL")) ctx->opt_distance_refill = std::atoi(e);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return BAM3D_ERR_CUDA; }
    if (cudaMalloc(&ctx->d_flags, 64) != cudaSuccess || cudaMemset(ctx->d_flags, 0, 64) != cudaSuccess ||
        cudaHostAlloc(&ctx->h_flags, 64, cudaHostAllocDefault) != cudaSuccess ||
        create_side_stream(&ctx->stream2) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_a, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&ctx->ev_b, cudaEventDisableTiming) != cudaSuccess) { bam3d_ctx_destroy(ctx); return BAM3D_ERR_CUDA; }
    *out_ctx = ctx;
    return BAM3D_OK;
}

void bam3d_ctx_destroy(bam3d_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& ev : ctx->timing_events) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
    ctx->timing_events.clear();
    if (ctx->d_flags) cudaFree(ctx->d_flags);
    if (ctx->h_flags) cudaFreeHost(ctx->h_flags);
    ctx->release_scratch();
    if (ctx->stream2) { cudaStreamSynchronize(ctx->stream2); cudaStreamDestroy(ctx->stream2); }
    if (ctx->ev_a) cudaEventDestroy(ctx->ev_a);
    if (ctx->ev_b) cudaEventDestroy(ctx->ev_b);
    cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* bam3d_last_error(const bam3d_ctx* ctx) { return ctx ? ctx->err.c_str() : "no context (CUDA device unavailable?)"; }
int32_t bam3d_ctx_synchronize(bam3d_ctx* ctx) {
    if (!ctx) return BAM3D_ERR_ARG;
    // also the place where a device-pointer (_dev) call's out-of-range pair index comes to light
    return take_index_flag(ctx, "bam3d_ctx_synchronize");
}
void* bam3d_ctx_stream(bam3d_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
uint64_t bam3d_ctx_launch_count(const bam3d_ctx* ctx) { return ctx ? ctx->launches : 0; }
int32_t bam3d_ctx_set_option(bam3d_ctx* ctx, const char* name, int32_t value) {
    if (!ctx || !name) return BAM3D_ERR_ARG;
    if (!std::strcmp(name, "binning")) { ctx->opt_binning = value != 0; return BAM3D_OK; }
    if (!std::strcmp(name, "timing")) { ctx->opt_timing = value != 0; return BAM3D_OK; }
#ifndef B3_EXPERIMENTS
    if ((!std::strcmp(name, "epa_tiers") || !std::strcmp(name, "intersect_split") || !std::strcmp(name, "intersect_refill") || !std::strcmp(name, "epa_pipeline")) && value != 0)
        return b3_fail(ctx, BAM3D_ERR_ARG, "this switch selects a measured-slower experiment kernel: build the library with -DB3_EXPERIMENTS to have it", cudaSuccess);
#endif
    if (!std::strcmp(name, "epa_tiers")) { ctx->opt_epa_tiers = value != 0; return BAM3D_OK; }
    if (!std::strcmp(name, "epa_pipeline")) { ctx->opt_epa_pipeline = value != 0; return BAM3D_OK; }
    if (!std::strcmp(name, "epa_huge")) { ctx->opt_epa_huge = value != 0; return BAM3D_OK; }
    if (!std::strcmp(name, "epa_overlap")) { ctx->opt_epa_overlap = value != 0; return BAM3D_OK; }
    if (!std::strcmp(name, "epa_global_polytopes")) { ctx->opt_epa_global_polytopes = value != 0; return BAM3D_OK; }
    if (!std::strcmp(name, "intersect_split")) { ctx->opt_intersect_split = value != 0; return BAM3D_OK; }
    if (!std::strcmp(name, "intersect_refill")) { if (value < 0 || value > 32) return b3_fail(ctx, BAM3D_ERR_ARG, "intersect_refill must be 0 (off) or the number of idle lanes (1..32) that triggers a refill", cudaSuccess); ctx->opt_intersect_refill = value; return BAM3D_OK; }
    if (!std::strcmp(name, "toi_refill")) { if (value < 0 || value > 32) return b3_fail(ctx, BAM3D_ERR_ARG, "toi_refill must be 0 (off) or the number of idle lanes (1..32) that triggers a refill", cudaSuccess); ctx->opt_toi_refill = value; return BAM3D_OK; }
    if (!std::strcmp(name, "distance_refill")) { if (value < 0 || value > 32) return b3_fail(ctx, BAM3D_ERR_ARG, "distance_refill must be 0 (off) or the number of idle lanes (1..32) that triggers a refill", cudaSuccess); ctx->opt_distance_refill = value; return BAM3D_OK; }
    if (!std::strcmp(name, "bvh_strategy")) { if (value < 0 || value > 2) return b3_fail(ctx, BAM3D_ERR_ARG, "bvh_strategy must be 0 (auto), 1 (traversal) or 2 (all pairs)", cudaSuccess); ctx->opt_bvh_strategy = value; return BAM3D_OK; }
    if (!std::strcmp(name, "sap_strategy")) { if (value < 0 || value > 2) return b3_fail(ctx, BAM3D_ERR_ARG, "sap_strategy must be 0 (auto), 1 (sweep) or 2 (bvh)", cudaSuccess); ctx->opt_sap_strategy = value; return BAM3D_OK; }
    return b3_fail(ctx, BAM3D_ERR_ARG, "unknown option", cudaSuccess);
}

int32_t bam3d_ctx_read_timing(bam3d_ctx* ctx, double* out_total_ms, uint64_t* out_count) {
    if (!ctx || !out_total_ms || !out_count) return BAM3D_ERR_ARG;
    CU_TRY(cudaStreamSynchronize(ctx->stream), "cudaStreamSynchronize");
    double total = 0.0;
    for (auto& ev : ctx->timing_events) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev.first, ev.second);
        total += ms;
        cudaEventDestroy(ev.first); cudaEventDestroy(ev.second);
    }
    *out_total_ms = total; *out_count = ctx->timing_events.size();
    ctx->timing_events.clear();
    return BAM3D_OK;
}

int32_t bam3d_host_alloc(bam3d_ctx* ctx, uint64_t bytes, void** out_ptr) {
    if (!ctx || !out_ptr) return BAM3D_ERR_ARG;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    CU_TRY(cudaHostAlloc(out_ptr, bytes ? bytes : 1, cudaHostAllocDefault), "cudaHostAlloc");
    return BAM3D_OK;
}
void bam3d_host_free(bam3d_ctx* ctx, void* ptr) { (void)ctx; if (ptr) cudaFreeHost(ptr); }

// ------------------------------------------------------------------------------------------------------------
// meshes / shapes
// ------------------------------------------------------------------------------------------------------------
// build_half_edges (primitive/polyhedron.rs:246-340) for one mesh, in the reference's order: faces one by one, a pair of
// twin half-edges created the first time either direction is seen, and first-writer-wins for vertex.edge, edge.left_face
// and face.edge — hill_climb_support_point and the ray walk follow exactly these links. Indices are local to the mesh.
struct HostHalfEdges {
    std::vector<uint32_t> vert_edge;                       // per vertex
    std::vector<uint32_t> target, twin, next, prev, left;  // per half-edge
    std::vector<uint32_t> fv, fedge;                       // per face: 3 vertices, first edge
};
static bool build_half_edges_host(uint32_t nverts, const uint32_t* faces, uint32_t nfaces, HostHalfEdges& H) {
    H.vert_edge.assign(nverts, 0);
    std::vector<uint8_t> vert_ready(nverts, 0), edge_ready;
    std::map<std::pair<uint32_t, uint32_t>, uint32_t> edge_map;
    for (uint32_t f = 0; f < nfaces; ++f) {
        const uint32_t v[3] = {faces[3 * f], faces[3 * f + 1], faces[3 * f + 2]};
        if (v[0] >= nverts || v[1] >= nverts || v[2] >= nverts) return false;
        uint32_t fe[3] = {0, 0, 0};
        bool face_ready = false;
        uint32_t first_edge = 0;
        for (int j = 0; j < 3; ++j) {
            const int i = (j == 0) ? 2 : j - 1;
            const uint32_t v0 = v[i], v1 = v[j];
            uint32_t e01, e10;
            auto it = edge_map.find({v0, v1});
            if (it != edge_map.end()) { e01 = it->second; e10 = H.twin[e01]; }
            else {
                e01 = (uint32_t)H.target.size(); e10 = e01 + 1;
                H.target.push_back(v1); H.twin.push_back(e10); H.next.push_back(0); H.prev.push_back(0); H.left.push_back(0); edge_ready.push_back(0);
                H.target.push_back(v0); H.twin.push_back(e01); H.next.push_back(0); H.prev.push_back(0); H.left.push_back(0); edge_ready.push_back(0);
            }
            edge_map[{v0, v1}] = e01;
            edge_map[{v1, v0}] = e10;
            if (!edge_ready[e01]) { H.left[e01] = f; edge_ready[e01] = 1; }
            if (!vert_ready[v0]) { H.vert_edge[v0] = e01; vert_ready[v0] = 1; }
            if (!face_ready) { first_edge = e01; face_ready = true; }
            fe[i] = e01;
        }
        H.fv.push_back(v[0]); H.fv.push_back(v[1]); H.fv.push_back(v[2]); H.fedge.push_back(first_edge);
        for (int j = 0; j < 3; ++j) {
            const int i = (j == 0) ? 2 : j - 1;
            H.next[fe[i]] = fe[j];
            H.prev[fe[j]] = fe[i];
        }
    }
    return true;
}

int32_t bam3d_meshes_upload(bam3d_ctx* ctx, const float* verts_xyz, const uint32_t* offsets, uint32_t n_meshes, bam3d_meshes** out) {
    return bam3d_meshes_upload_faces(ctx, verts_xyz, offsets, n_meshes, nullptr, nullptr, out);
}

int32_t bam3d_meshes_upload_faces(bam3d_ctx* ctx, const float* verts_xyz, const uint32_t* offsets, uint32_t n_meshes, const uint32_t* faces,
                                  const uint32_t* face_offsets, bam3d_meshes** out) {
    if (!ctx || !out || !offsets || (n_meshes && !verts_xyz) || ((faces == nullptr) != (face_offsets == nullptr)))
        return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_meshes_upload: null argument", cudaSuccess);
    *out = nullptr;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    for (uint32_t m = 0; m < n_meshes; ++m)
        if (offsets[m + 1] < offsets[m] || (face_offsets && face_offsets[m + 1] < face_offsets[m]))
            return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_meshes_upload: offsets must be non-decreasing", cudaSuccess);
    const uint64_t nverts = offsets[n_meshes];
    // host-side half-edge structures, concatenated over the meshes that have faces
    std::vector<uint32_t> h_vert_edge(nverts ? nverts : 1, 0);
    std::vector<uint4> h_edges, h_faces;
    std::vector<uint64_t> edge_base(n_meshes + 1, 0), face_base(n_meshes + 1, 0);
    for (uint32_t m = 0; m < n_meshes; ++m) {
        edge_base[m] = h_edges.size() / 2; face_base[m] = h_faces.size() / 2;
        const uint32_t nf = face_offsets ? face_offsets[m + 1] - face_offsets[m] : 0;
        if (nf == 0) continue;
        HostHalfEdges H;
        const uint32_t nv = offsets[m + 1] - offsets[m];
        if (!build_half_edges_host(nv, faces + 3 * (size_t)face_offsets[m], nf, H))
            return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_meshes_upload_faces: face vertex index out of range", cudaSuccess);
        for (uint32_t v = 0; v < nv; ++v) h_vert_edge[offsets[m] + v] = H.vert_edge[v];
        for (size_t e = 0; e < H.target.size(); ++e) {
            h_edges.push_back(make_uint4(H.target[e], H.twin[e], H.next[e], H.prev[e]));
            h_edges.push_back(make_uint4(H.left[e], 0, 0, 0));
        }
        for (uint32_t f = 0; f < nf; ++f) {
            h_faces.push_back(make_uint4(H.fv[3 * f], H.fv[3 * f + 1], H.fv[3 * f + 2], H.fedge[f]));
            h_faces.push_back(make_uint4(offsets[m] + MESH_HEADER_SLOTS * (m + 1), 0, 0, 0));  // vertex slot base: replaced by the plane on the device
        }
    }
    edge_base[n_meshes] = h_edges.size() / 2; face_base[n_meshes] = h_faces.size() / 2;
    bam3d_meshes* M = new (std::nothrow) bam3d_meshes();
    if (!M) return BAM3D_ERR_OOM;
    M->n_meshes = n_meshes; M->nverts = nverts; M->n_edges = edge_base[n_meshes]; M->n_faces = face_base[n_meshes];
    const size_t slots = nverts + (size_t)MESH_HEADER_SLOTS * n_meshes;
    cudaError_t e;
    uint32_t* d_err = nullptr;
    if ((e = cudaMalloc(&M->verts, (slots ? slots : 1) * sizeof(float4))) != cudaSuccess ||
        (e = cudaMalloc(&M->offsets, ((size_t)n_meshes + 1) * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&M->vert_edge, h_vert_edge.size() * sizeof(uint32_t))) != cudaSuccess ||
        (e = cudaMalloc(&M->edges, (h_edges.size() ? h_edges.size() : 1) * sizeof(uint4))) != cudaSuccess ||
        (e = cudaMalloc(&M->faces, (h_faces.size() ? h_faces.size() : 1) * sizeof(uint4))) != cudaSuccess ||
        (e = cudaMalloc(&d_err, sizeof(uint32_t))) != cudaSuccess) {
        bam3d_meshes_free(ctx, M); if (d_err) cudaFree(d_err);
        return b3_fail(ctx, BAM3D_ERR_OOM, "cudaMalloc(meshes)", e);
    }
    // vertex slots with a header in front of every mesh; device offsets count the header slots
    std::vector<float4> h_verts(slots ? slots : 1, make_float4(0.f, 0.f, 0.f, 0.f));
    std::vector<uint32_t> h_off(n_meshes + 1);
    for (uint32_t m = 0; m <= n_meshes; ++m) h_off[m] = offsets[m] + MESH_HEADER_SLOTS * m;
    for (uint32_t m = 0; m < n_meshes; ++m) {
        MeshHeader hd;
        hd.vert_edge = M->vert_edge + offsets[m];
        hd.edges = M->edges + 2 * edge_base[m];
        hd.faces = M->faces + 2 * face_base[m];
        hd.n_faces = (uint32_t)(face_base[m + 1] - face_base[m]);
        hd.n_edges = (uint32_t)(edge_base[m + 1] - edge_base[m]);
        std::memcpy(&h_verts[h_off[m]], &hd, sizeof(hd));
        for (uint32_t v = offsets[m]; v < offsets[m + 1]; ++v)
            h_verts[h_off[m] + MESH_HEADER_SLOTS + (v - offsets[m])] = make_float4(verts_xyz[3 * (size_t)v], verts_xyz[3 * (size_t)v + 1], verts_xyz[3 * (size_t)v + 2], 0.f);
    }
    cudaMemcpyAsync(M->verts, h_verts.data(), h_verts.size() * sizeof(float4), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(M->offsets, h_off.data(), h_off.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemcpyAsync(M->vert_edge, h_vert_edge.data(), h_vert_edge.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream);
    if (!h_edges.empty()) cudaMemcpyAsync(M->edges, h_edges.data(), h_edges.size() * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream);
    if (!h_faces.empty()) cudaMemcpyAsync(M->faces, h_faces.data(), h_faces.size() * sizeof(uint4), cudaMemcpyHostToDevice, ctx->stream);
    cudaMemsetAsync(d_err, 0, sizeof(uint32_t), ctx->stream);
    // face planes (Plane::from_points, plane.rs:69-86) on the device: same unfused f32 arithmetic as everything else
    ctx->launches += launch_mesh_face_planes(ctx->stream, M->verts, M->faces, M->n_faces, d_err);
    uint32_t h_err = 0;
    cudaMemcpyAsync(&h_err, d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);  // also keeps the host vectors alive until the copies are done
    cudaFree(d_err);
    if (e != cudaSuccess) { bam3d_meshes_free(ctx, M); return b3_fail(ctx, BAM3D_ERR_CUDA, "bam3d_meshes_upload", e); }
    if (h_err) {
        bam3d_meshes_free(ctx, M);
        return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_meshes_upload_faces: degenerate face (the reference panics in Plane::from_points(..).unwrap())", cudaSuccess);
    }
    *out = M;
    return BAM3D_OK;
}

void bam3d_meshes_free(bam3d_ctx* ctx, bam3d_meshes* M) {
    (void)ctx;
    if (!M) return;
    if (M->verts) cudaFree(M->verts);
    if (M->offsets) cudaFree(M->offsets);
    if (M->vert_edge) cudaFree(M->vert_edge);
    if (M->edges) cudaFree(M->edges);
    if (M->faces) cudaFree(M->faces);
    delete M;
}

static int32_t run_pack(bam3d_ctx* ctx, bam3d_shapes* S, bool affine48 = false) {
    int32_t rc = ctx->small.reserve(ctx, 1024);
    if (rc) return rc;
    uint32_t* d_err = (uint32_t*)ctx->small.ptr + 200;
    cudaMemsetAsync(d_err, 0, sizeof(uint32_t), ctx->stream);
    ctx->launches += launch_pack_shapes(ctx->stream, S->d_kind, S->d_params, S->d_xform, S->d_mesh_id, S->n,
                                        S->meshes ? S->meshes->n_meshes : 0, S->recs, S->meta, d_err, affine48);
    if ((rc = check_launch(ctx, "pack_shapes_kernel"))) return rc;
    uint32_t h_err = 0;
    CU_TRY(cudaMemcpyAsync(&h_err, d_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(err)");
    CU_TRY(cudaStreamSynchronize(ctx->stream), "pack_shapes_kernel");
    if (h_err & 1u) return b3_fail(ctx, BAM3D_ERR_NON_AFFINE, "a transform's 4th row is not (0,0,0,1)", cudaSuccess);
    if (h_err & 2u) return b3_fail(ctx, BAM3D_ERR_ARG, "unknown primitive kind", cudaSuccess);
    if (h_err & 4u) return b3_fail(ctx, BAM3D_ERR_ARG, "polyhedron mesh_id out of range (or no mesh library given)", cudaSuccess);
    return BAM3D_OK;
}

int32_t bam3d_shapes_upload(bam3d_ctx* ctx, const uint8_t* kind, const float* params, const float* xform, const uint32_t* mesh_id,
                            uint32_t n, const bam3d_meshes* meshes, bam3d_shapes** out) {
    if (!ctx || !out || (n && (!kind || !params || !xform))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_shapes_upload: null argument", cudaSuccess);
    *out = nullptr;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    bam3d_shapes* S = new (std::nothrow) bam3d_shapes();
    if (!S) return BAM3D_ERR_OOM;
    S->n = n; S->meshes = meshes;
    for (uint32_t i = 0; i < n; ++i) if (kind[i] < 7) S->kind_hist[kind[i]]++;
    size_t nn = n ? n : 1;
    cudaError_t e;
    if ((e = cudaMalloc(&S->recs, nn * 96)) != cudaSuccess || (e = cudaMalloc(&S->meta, nn * 4)) != cudaSuccess ||
        (e = cudaMalloc(&S->d_kind, nn)) != cudaSuccess || (e = cudaMalloc(&S->d_params, nn * 16)) != cudaSuccess ||
        (e = cudaMalloc(&S->d_xform, nn * 64)) != cudaSuccess || (mesh_id && (e = cudaMalloc(&S->d_mesh_id, nn * 4)) != cudaSuccess)) {
        bam3d_shapes_free(ctx, S);
        return b3_fail(ctx, BAM3D_ERR_OOM, "cudaMalloc(shapes)", e);
    }
    if (n) {
        cudaMemcpyAsync(S->d_kind, kind, n, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(S->d_params, params, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(S->d_xform, xform, (size_t)n * 64, cudaMemcpyHostToDevice, ctx->stream);
        if (mesh_id) cudaMemcpyAsync(S->d_mesh_id, mesh_id, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream);
    }
    int32_t rc = run_pack(ctx, S);
    if (rc) { bam3d_shapes_free(ctx, S); return rc; }
    *out = S;
    return BAM3D_OK;
}

int32_t bam3d_shapes_set_transforms(bam3d_ctx* ctx, bam3d_shapes* S, const float* xform) {
    if (!ctx || !S || !xform) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_shapes_set_transforms: null argument", cudaSuccess);
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    if (S->n) CU_TRY(cudaMemcpyAsync(S->d_xform, xform, (size_t)S->n * 64, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(xform)");
    return run_pack(ctx, S);
}

int32_t bam3d_shapes_set_transforms_affine(bam3d_ctx* ctx, bam3d_shapes* S, const float* xform12) {
    if (!ctx || !S || !xform12) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_shapes_set_transforms_affine: null argument", cudaSuccess);
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    if (S->n) CU_TRY(cudaMemcpyAsync(S->d_xform, xform12, (size_t)S->n * 48, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(xform)");
    return run_pack(ctx, S, true);
}

void bam3d_shapes_free(bam3d_ctx* ctx, bam3d_shapes* S) {
    (void)ctx;
    if (!S) return;
    if (S->recs) cudaFree(S->recs);
    if (S->meta) cudaFree(S->meta);
    if (S->d_kind) cudaFree(S->d_kind);
    if (S->d_params) cudaFree(S->d_params);
    if (S->d_xform) cudaFree(S->d_xform);
    if (S->d_mesh_id) cudaFree(S->d_mesh_id);
    delete S;
}
uint32_t bam3d_shapes_count(const bam3d_shapes* S) { return S ? S->n : 0; }

// ------------------------------------------------------------------------------------------------------------
// narrow phase, device buffers
// ------------------------------------------------------------------------------------------------------------
int32_t bam3d_gjk_intersect_dev(bam3d_ctx* ctx, const bam3d_shapes* shapes, const uint32_t* d_pairs, uint64_t n,
                                const bam3d_gjk_settings* settings, uint8_t* d_status) {
    if (!ctx || !shapes || (n && (!d_pairs || !d_status))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gjk_intersect: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, n)) || (rc = check_pairs_over(ctx, shapes, n))) return rc;
    if (n == 0) return BAM3D_OK;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    ShapeSetDev S = view(ctx, shapes);
    Plan plan = plan_for(ctx, shapes);
    BinScratch B; const BinScratch* pb;
    if ((rc = prepare_bins(ctx, plan, S, (const uint2*)d_pairs, (uint32_t)n, B, &pb))) return rc;
    void* split_scratch = nullptr;
    const int refill_min = ctx->opt_intersect_refill > 0 ? ctx->opt_intersect_refill : 0;
    if (plan.run_thread && (ctx->opt_intersect_split || refill_min)) {
        if ((rc = ctx->epa_queue.reserve(ctx, intersect_scratch_bytes((uint32_t)n)))) return rc;
        split_scratch = ctx->epa_queue.ptr;
    }
    { KernelTimer kt(ctx);
      ctx->launches += launch_intersect(ctx->stream, S, (const uint2*)d_pairs, (uint32_t)n, pb, plan.run_thread, plan.run_warp, to_dev(settings),
                                        split_scratch, refill_min, d_status); }
    return check_launch(ctx, "gjk_intersect_kernel");
}

int32_t bam3d_gjk_contact_dev(bam3d_ctx* ctx, const bam3d_shapes* shapes, const uint32_t* d_pairs, uint64_t n,
                              const bam3d_gjk_settings* settings, int32_t strategy, bam3d_contact* d_contacts, uint8_t* d_status) {
    if (!ctx || !shapes || (n && (!d_pairs || !d_status || !d_contacts))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gjk_contact: null argument", cudaSuccess);
    if (strategy != BAM3D_FULL_RESOLUTION && strategy != BAM3D_COLLISION_ONLY) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gjk_contact: unknown strategy", cudaSuccess);
    SettingsDev cfg = to_dev(settings);
    if (strategy == BAM3D_FULL_RESOLUTION && cfg.max_iterations > 100)
        return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gjk_contact: max_iterations > 100 exceeds the device EPA polytope capacity", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, n)) || (rc = check_pairs_over(ctx, shapes, n))) return rc;
    if (n == 0) return BAM3D_OK;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    ShapeSetDev S = view(ctx, shapes);
    Plan plan = plan_for(ctx, shapes);
    BinScratch B; const BinScratch* pb;
    if ((rc = prepare_bins(ctx, plan, S, (const uint2*)d_pairs, (uint32_t)n, B, &pb))) return rc;
    if ((rc = ctx->epa_overflow.reserve(ctx, contact_overflow_scratch_bytes()))) return rc;
    if (plan.run_thread && (rc = ctx->epa_queue.reserve(ctx, contact_queue_bytes(strategy == BAM3D_FULL_RESOLUTION ? (uint32_t)n : 0)))) return rc;
    { KernelTimer kt(ctx);
      // without binning every thread-path pair has the same kind on both sides: tell the EPA which bin that is
      ContactSide side{ctx->stream2, ctx->ev_a, ctx->ev_b};
      void* poly_slab = nullptr;
      if (ctx->opt_epa_global_polytopes && plan.run_thread && strategy == BAM3D_FULL_RESOLUTION) {
          if ((rc = ctx->epa_slab.reserve(ctx, contact_poly_slab_bytes()))) return rc;
          poly_slab = ctx->epa_slab.ptr;
      }
      void* huge_slab = nullptr;
      if (ctx->opt_epa_huge && strategy == BAM3D_FULL_RESOLUTION) {
          if ((rc = ctx->epa_huge.reserve(ctx, contact_huge_slab_bytes()))) return rc;
          huge_slab = ctx->epa_huge.ptr;
      }
      uint32_t default_key = 0;
      for (uint32_t k = 0; k < 6; ++k) if (shapes->kind_hist[k]) default_key = k * 7 + k;
      ctx->launches += launch_contact(ctx->stream, S, (const uint2*)d_pairs, (uint32_t)n, pb, plan.run_thread, plan.run_warp, cfg, strategy,
                                      default_key, ctx->opt_epa_tiers ? 1 : 0, ctx->opt_epa_pipeline, ctx->opt_epa_overlap ? &side : nullptr, poly_slab, ctx->epa_overflow.ptr, huge_slab, ctx->epa_queue.ptr, (float4*)d_contacts,
                                      d_status); }
    return check_launch(ctx, "gjk_contact_kernel");
}

int32_t bam3d_gjk_distance_dev(bam3d_ctx* ctx, const bam3d_shapes* shapes, const uint32_t* d_pairs, uint64_t n,
                               const bam3d_gjk_settings* settings, float* d_ref_value, float* d_distance, uint8_t* d_status) {
    if (!ctx || !shapes || (n && (!d_pairs || !d_status || !d_ref_value || !d_distance))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gjk_distance: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, n)) || (rc = check_pairs_over(ctx, shapes, n))) return rc;
    if (n == 0) return BAM3D_OK;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    ShapeSetDev S = view(ctx, shapes);
    Plan plan = plan_for(ctx, shapes);
    BinScratch B; const BinScratch* pb;
    if ((rc = prepare_bins(ctx, plan, S, (const uint2*)d_pairs, (uint32_t)n, B, &pb))) return rc;
    void* cursor = nullptr;
    if (plan.run_thread && ctx->opt_distance_refill > 0) {
        if ((rc = ctx->epa_queue.reserve(ctx, 256))) return rc;
        cursor = ctx->epa_queue.ptr;
    }
    { KernelTimer kt(ctx);
      ctx->launches += launch_distance(ctx->stream, S, (const uint2*)d_pairs, (uint32_t)n, pb, plan.run_thread, plan.run_warp, to_dev(settings),
                                       cursor, ctx->opt_distance_refill, d_ref_value, d_distance, d_status); }
    return check_launch(ctx, "gjk_distance_kernel");
}

int32_t bam3d_gjk_time_of_impact_dev(bam3d_ctx* ctx, const bam3d_shapes* s0, const bam3d_shapes* s1, const uint32_t* d_pairs, uint64_t n,
                                     const bam3d_gjk_settings* settings, bam3d_contact* d_contacts, uint8_t* d_status) {
    if (!ctx || !s0 || !s1 || (n && (!d_pairs || !d_status || !d_contacts))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gjk_time_of_impact: null argument", cudaSuccess);
    if (s0->n != s1->n) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gjk_time_of_impact: start/end shape sets differ in size", cudaSuccess);
    // the end set contributes transforms only; its primitives must be the start set's (the same mesh library, or a second
    // upload of the same one — only the sizes can be compared here)
    const bool same_meshes = s0->meshes == s1->meshes ||
                             (s0->meshes && s1->meshes && s0->meshes->n_meshes == s1->meshes->n_meshes && s0->meshes->nverts == s1->meshes->nverts);
    if (!same_meshes || std::memcmp(s0->kind_hist, s1->kind_hist, sizeof s0->kind_hist) != 0)
        return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gjk_time_of_impact: the end shape set must hold the same primitives (kinds, mesh library) as the start set", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, n)) || (rc = check_pairs_over(ctx, s0, n))) return rc;
    if (n == 0) return BAM3D_OK;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    ShapeSetDev S0 = view(ctx, s0), S1 = view(ctx, s1);
    Plan plan = plan_for(ctx, s0);
    BinScratch B; const BinScratch* pb;
    if ((rc = prepare_bins(ctx, plan, S0, (const uint2*)d_pairs, (uint32_t)n, B, &pb))) return rc;
    void* cursor = nullptr;
    if (plan.run_thread && ctx->opt_toi_refill > 0) {
        if ((rc = ctx->epa_queue.reserve(ctx, 256))) return rc;
        cursor = ctx->epa_queue.ptr;
    }
    { KernelTimer kt(ctx);
      ctx->launches += launch_toi(ctx->stream, S0, S1, (const uint2*)d_pairs, (uint32_t)n, pb, plan.run_thread, plan.run_warp, to_dev(settings),
                                  cursor, ctx->opt_toi_refill, (float4*)d_contacts, d_status); }
    return check_launch(ctx, "gjk_toi_kernel");
}

int32_t bam3d_compact_contacts_dev(bam3d_ctx* ctx, const bam3d_contact* d_contacts, const uint8_t* d_status, uint64_t n, uint64_t base,
                                   bam3d_contact40* d_out, uint64_t capacity, uint64_t* d_count) {
    if (!ctx || !d_count || (n && (!d_contacts || !d_status || !d_out))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_compact_contacts: null argument", cudaSuccess);
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    ctx->launches += launch_compact_contacts(ctx->stream, (const float4*)d_contacts, d_status, n, base, d_out, capacity,
                                             (unsigned long long*)d_count);
    return check_launch(ctx, "compact_contacts_kernel");
}

// ------------------------------------------------------------------------------------------------------------
// narrow phase, host buffers: H2D pairs -> kernels -> D2H results, synchronous on return
// ------------------------------------------------------------------------------------------------------------
static int32_t stage_pairs(bam3d_ctx* ctx, const uint32_t* pairs, uint64_t n) {
    int32_t rc = ctx->pairs.reserve(ctx, n * 8);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(ctx->pairs.ptr, pairs, n * 8, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(pairs)");
    return BAM3D_OK;
}

int32_t bam3d_gjk_intersect(bam3d_ctx* ctx, const bam3d_shapes* shapes, const uint32_t* pairs, uint64_t n,
                            const bam3d_gjk_settings* settings, uint8_t* out_status) {
    if (!ctx || !shapes || (n && (!pairs || !out_status))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gjk_intersect: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, n))) return rc;
    if (n == 0) return BAM3D_OK;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    if ((rc = stage_pairs(ctx, pairs, n)) || (rc = ctx->status.reserve(ctx, n))) return rc;
    if ((rc = bam3d_gjk_intersect_dev(ctx, shapes, (const uint32_t*)ctx->pairs.ptr, n, settings, (uint8_t*)ctx->status.ptr))) return rc;
    CU_TRY(cudaMemcpyAsync(out_status, ctx->status.ptr, n, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(status)");
    return take_index_flag(ctx, "bam3d_gjk_intersect");
}

int32_t bam3d_gjk_contact(bam3d_ctx* ctx, const bam3d_shapes* shapes, const uint32_t* pairs, uint64_t n,
                          const bam3d_gjk_settings* settings, int32_t strategy, bam3d_contact* out_contacts, uint8_t* out_status) {
    if (!ctx || !shapes || (n && (!pairs || !out_status || !out_contacts))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gjk_contact: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, n))) return rc;
    if (n == 0) return BAM3D_OK;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    if ((rc = stage_pairs(ctx, pairs, n)) || (rc = ctx->status.reserve(ctx, n)) || (rc = ctx->contacts.reserve(ctx, n * 32))) return rc;
    if ((rc = bam3d_gjk_contact_dev(ctx, shapes, (const uint32_t*)ctx->pairs.ptr, n, settings, strategy, (bam3d_contact*)ctx->contacts.ptr,
                                    (uint8_t*)ctx->status.ptr))) return rc;
    CU_TRY(cudaMemcpyAsync(out_contacts, ctx->contacts.ptr, n * 32, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(contacts)");
    CU_TRY(cudaMemcpyAsync(out_status, ctx->status.ptr, n, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(status)");
    return take_index_flag(ctx, "bam3d_gjk_contact");
}

int32_t bam3d_gjk_distance(bam3d_ctx* ctx, const bam3d_shapes* shapes, const uint32_t* pairs, uint64_t n,
                           const bam3d_gjk_settings* settings, float* out_ref_value, float* out_distance, uint8_t* out_status) {
    if (!ctx || !shapes || (n && (!pairs || !out_status || !out_ref_value || !out_distance))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gjk_distance: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, n))) return rc;
    if (n == 0) return BAM3D_OK;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    if ((rc = stage_pairs(ctx, pairs, n)) || (rc = ctx->status.reserve(ctx, n)) || (rc = ctx->f0.reserve(ctx, n * 4)) || (rc = ctx->f1.reserve(ctx, n * 4))) return rc;
    if ((rc = bam3d_gjk_distance_dev(ctx, shapes, (const uint32_t*)ctx->pairs.ptr, n, settings, (float*)ctx->f0.ptr, (float*)ctx->f1.ptr,
                                     (uint8_t*)ctx->status.ptr))) return rc;
    CU_TRY(cudaMemcpyAsync(out_ref_value, ctx->f0.ptr, n * 4, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(ref)");
    CU_TRY(cudaMemcpyAsync(out_distance, ctx->f1.ptr, n * 4, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(dist)");
    CU_TRY(cudaMemcpyAsync(out_status, ctx->status.ptr, n, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(status)");
    return take_index_flag(ctx, "bam3d_gjk_distance");
}

// ---- ray casts against primitives ---------------------------------------------------------------
int32_t bam3d_ray_cast_dev(bam3d_ctx* ctx, const bam3d_shapes* shapes, const float* d_rays, uint64_t n_rays, const uint32_t* d_casts,
                           uint64_t n, int32_t query, float* d_points, uint8_t* d_status) {
    if (!ctx || !shapes || (n && (!d_rays || !d_casts || !d_points || !d_status))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_ray_cast: null argument", cudaSuccess);
    if (query < 0 || query > 2) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_ray_cast: query must be 0 (intersects), 1 (intersection) or 2 (swept particle)", cudaSuccess);
    (void)n_rays;
    int32_t rc;
    if ((rc = check_n(ctx, n))) return rc;
    if (n == 0) return BAM3D_OK;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    { KernelTimer kt(ctx);
      ctx->launches += launch_ray_cast(ctx->stream, view(ctx, shapes), d_rays, (const uint2*)d_casts, n, query, d_points, d_status); }
    return check_launch(ctx, "ray_cast_kernel");
}

int32_t bam3d_ray_cast(bam3d_ctx* ctx, const bam3d_shapes* shapes, const float* rays, uint64_t n_rays, const uint32_t* casts, uint64_t n,
                       int32_t query, float* out_points, uint8_t* out_status) {
    if (!ctx || !shapes || (n && (!rays || !casts || !out_points || !out_status))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_ray_cast: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, n)) || (rc = check_n(ctx, n_rays))) return rc;
    if (n == 0) return BAM3D_OK;
    for (uint64_t i = 0; i < n; ++i)
        if (casts[2 * i] >= n_rays || casts[2 * i + 1] >= shapes->n) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_ray_cast: cast index out of range", cudaSuccess);
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    if ((rc = stage_pairs(ctx, casts, n)) || (rc = ctx->status.reserve(ctx, n)) || (rc = ctx->f0.reserve(ctx, n * 12)) || (rc = ctx->f1.reserve(ctx, n_rays * 24))) return rc;
    CU_TRY(cudaMemcpyAsync(ctx->f1.ptr, rays, n_rays * 24, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(rays)");
    if ((rc = bam3d_ray_cast_dev(ctx, shapes, (const float*)ctx->f1.ptr, n_rays, (const uint32_t*)ctx->pairs.ptr, n, query, (float*)ctx->f0.ptr,
                                 (uint8_t*)ctx->status.ptr))) return rc;
    CU_TRY(cudaMemcpyAsync(out_points, ctx->f0.ptr, n * 12, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(points)");
    CU_TRY(cudaMemcpyAsync(out_status, ctx->status.ptr, n, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(status)");
    CU_TRY(cudaStreamSynchronize(ctx->stream), "bam3d_ray_cast");
    return BAM3D_OK;
}

int32_t bam3d_gjk_time_of_impact(bam3d_ctx* ctx, const bam3d_shapes* s0, const bam3d_shapes* s1, const uint32_t* pairs, uint64_t n,
                                 const bam3d_gjk_settings* settings, bam3d_contact* out_contacts, uint8_t* out_status) {
    if (!ctx || !s0 || !s1 || (n && (!pairs || !out_status || !out_contacts))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gjk_time_of_impact: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, n))) return rc;
    if (n == 0) return BAM3D_OK;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    if ((rc = stage_pairs(ctx, pairs, n)) || (rc = ctx->status.reserve(ctx, n)) || (rc = ctx->contacts.reserve(ctx, n * 32))) return rc;
    if ((rc = bam3d_gjk_time_of_impact_dev(ctx, s0, s1, (const uint32_t*)ctx->pairs.ptr, n, settings, (bam3d_contact*)ctx->contacts.ptr,
                                           (uint8_t*)ctx->status.ptr))) return rc;
    CU_TRY(cudaMemcpyAsync(out_contacts, ctx->contacts.ptr, n * 32, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(contacts)");
    CU_TRY(cudaMemcpyAsync(out_status, ctx->status.ptr, n, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(status)");
    return take_index_flag(ctx, "bam3d_gjk_time_of_impact");
}

// ---- compound shapes (gjk/mod.rs:362-523) ----------------------------------------------------------------------
}  // extern "C"

struct bam3d_compounds {
    bam3d_shapes* prims = nullptr;      // every primitive of every compound; d_xform is rewritten by each call (compose)
    bam3d_shapes* prims_end = nullptr;  // time of impact: the same primitives at the end transforms (allocated on first use)
    float* d_local = nullptr;           // n_prims x 16: local-to-model transforms
    uint32_t* d_owner = nullptr;        // n_prims: compound of each primitive
    uint32_t* d_offsets = nullptr;      // n_compounds + 1
    std::vector<uint32_t> offsets;      // host copy (segment sizes)
    std::vector<uint8_t> kind; std::vector<float> params; std::vector<uint32_t> mesh_id;  // for prims_end
    const bam3d_meshes* meshes = nullptr;
    uint32_t n_compounds = 0, n_prims = 0;
};

// world transforms of a compound set's primitives for the given compound transforms (host array), then the pack kernel
static int32_t compounds_place(bam3d_ctx* ctx, bam3d_compounds* K, bam3d_shapes* target, const float* comp_xform) {
    int32_t rc;
    const size_t nc = K->n_compounds ? K->n_compounds : 1;
    if ((rc = ctx->f1.reserve(ctx, nc * 64))) return rc;
    if (K->n_compounds) CU_TRY(cudaMemcpyAsync(ctx->f1.ptr, comp_xform, (size_t)K->n_compounds * 64, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(comp_xform)");
    ctx->launches += launch_compound_compose(ctx->stream, (const float*)ctx->f1.ptr, K->d_local, K->d_owner, K->n_prims, target->d_xform);
    if ((rc = check_launch(ctx, "compound_compose_kernel"))) return rc;
    return run_pack(ctx, target);
}

// mode 0 contact, 1 distance, 2 time of impact
static int32_t compounds_run(bam3d_ctx* ctx, bam3d_compounds* K, int mode, const float* comp_xform, const float* comp_xform_end, const uint32_t* pairs,
                             uint64_t n, const bam3d_gjk_settings* settings, int32_t strategy, bam3d_contact* out_contacts, float* out_value,
                             uint8_t* out_status) {
    if (!ctx || !K || (K->n_compounds && !comp_xform) || (n && (!pairs || !out_status)) || (n && mode != 1 && !out_contacts) || (n && mode == 1 && !out_value) ||
        (mode == 2 && K->n_compounds && !comp_xform_end))
        return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gjk_*_complex: null argument", cudaSuccess);
    if (mode != 1 && strategy != BAM3D_FULL_RESOLUTION && strategy != BAM3D_COLLISION_ONLY) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gjk_*_complex: unknown strategy", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, n))) return rc;
    if (n == 0) return BAM3D_OK;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    // segments: compound pair k owns |left| x |right| primitive pairs
    std::vector<unsigned long long> seg(n + 1, 0);
    for (uint64_t k = 0; k < n; ++k) {
        const uint32_t l = pairs[2 * k], r = pairs[2 * k + 1];
        if (l >= K->n_compounds || r >= K->n_compounds) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_gjk_*_complex: compound index out of range", cudaSuccess);
        seg[k + 1] = seg[k] + (unsigned long long)(K->offsets[l + 1] - K->offsets[l]) * (K->offsets[r + 1] - K->offsets[r]);
    }
    const uint64_t total = seg[n];
    if ((rc = check_n(ctx, total))) return rc;
    if ((rc = compounds_place(ctx, K, K->prims, comp_xform))) return rc;
    if (mode == 2) {
        if (!K->prims_end) {  // first use: an identically-typed set whose transforms the compose kernel fills
            std::vector<float> ident((size_t)K->n_prims * 16, 0.f);
            for (uint32_t i = 0; i < K->n_prims; ++i) ident[16 * (size_t)i] = ident[16 * (size_t)i + 5] = ident[16 * (size_t)i + 10] = ident[16 * (size_t)i + 15] = 1.f;
            if ((rc = bam3d_shapes_upload(ctx, K->kind.data(), K->params.data(), ident.data(), K->mesh_id.empty() ? nullptr : K->mesh_id.data(), K->n_prims, K->meshes,
                                          &K->prims_end))) return rc;
        }
        if ((rc = compounds_place(ctx, K, K->prims_end, comp_xform_end))) return rc;
    }
    const size_t tt = total ? total : 1;
    if ((rc = ctx->pipe0.reserve(ctx, (n + 1) * 8)) || (rc = ctx->pipe1.reserve(ctx, n * 8)) || (rc = ctx->pairs.reserve(ctx, tt * 8)) ||
        (rc = ctx->status.reserve(ctx, tt)) || (rc = ctx->contacts.reserve(ctx, tt * 32)) || (rc = ctx->f0.reserve(ctx, tt * 4)) || (rc = ctx->f1.reserve(ctx, tt * 4)) ||
        (rc = ctx->pipe2.reserve(ctx, n * 32)) || (rc = ctx->pipe3.reserve(ctx, n + 16))) return rc;
    CU_TRY(cudaMemcpyAsync(ctx->pipe0.ptr, seg.data(), (n + 1) * 8, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(segments)");
    CU_TRY(cudaMemcpyAsync(ctx->pipe1.ptr, pairs, n * 8, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(pairs)");
    ctx->launches += launch_compound_expand(ctx->stream, (const uint2*)ctx->pipe1.ptr, (const unsigned long long*)ctx->pipe0.ptr, (uint32_t)n, total, K->d_offsets,
                                            (uint2*)ctx->pairs.ptr);
    if ((rc = check_launch(ctx, "compound_expand_kernel"))) return rc;
    if (total) {
        if (mode == 0) rc = bam3d_gjk_contact_dev(ctx, K->prims, (const uint32_t*)ctx->pairs.ptr, total, settings, strategy, (bam3d_contact*)ctx->contacts.ptr, (uint8_t*)ctx->status.ptr);
        else if (mode == 1) rc = bam3d_gjk_distance_dev(ctx, K->prims, (const uint32_t*)ctx->pairs.ptr, total, settings, (float*)ctx->f0.ptr, (float*)ctx->f1.ptr, (uint8_t*)ctx->status.ptr);
        else rc = bam3d_gjk_time_of_impact_dev(ctx, K->prims, K->prims_end, (const uint32_t*)ctx->pairs.ptr, total, settings, (bam3d_contact*)ctx->contacts.ptr, (uint8_t*)ctx->status.ptr);
        if (rc) return rc;
    }
    ctx->launches += launch_compound_reduce(ctx->stream, mode, strategy, (const unsigned long long*)ctx->pipe0.ptr, (uint32_t)n, (const float4*)ctx->contacts.ptr,
                                            (const float*)ctx->f0.ptr, (const uint8_t*)ctx->status.ptr, (float4*)ctx->pipe2.ptr, (float*)ctx->pipe2.ptr, (uint8_t*)ctx->pipe3.ptr);
    if ((rc = check_launch(ctx, "compound_reduce_kernel"))) return rc;
    if (mode == 1) CU_TRY(cudaMemcpyAsync(out_value, ctx->pipe2.ptr, n * 4, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(value)");
    else CU_TRY(cudaMemcpyAsync(out_contacts, ctx->pipe2.ptr, n * 32, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(contacts)");
    CU_TRY(cudaMemcpyAsync(out_status, ctx->pipe3.ptr, n, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(status)");
    CU_TRY(cudaStreamSynchronize(ctx->stream), "bam3d_gjk_*_complex");
    return BAM3D_OK;
}

extern "C" {

void bam3d_compounds_free(bam3d_ctx* ctx, bam3d_compounds* K) {
    if (!K) return;
    bam3d_shapes_free(ctx, K->prims);
    bam3d_shapes_free(ctx, K->prims_end);
    if (K->d_local) cudaFree(K->d_local);
    if (K->d_owner) cudaFree(K->d_owner);
    if (K->d_offsets) cudaFree(K->d_offsets);
    delete K;
}

int32_t bam3d_compounds_upload(bam3d_ctx* ctx, const uint8_t* kind, const float* params, const float* local_xform, const uint32_t* mesh_id,
                               uint32_t n_prims, const uint32_t* comp_offsets, uint32_t n_compounds, const bam3d_meshes* meshes, bam3d_compounds** out) {
    if (!ctx || !out || !comp_offsets || (n_prims && (!kind || !params || !local_xform))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_compounds_upload: null argument", cudaSuccess);
    *out = nullptr;
    if (comp_offsets[0] != 0 || comp_offsets[n_compounds] != n_prims) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_compounds_upload: comp_offsets must run from 0 to n_prims", cudaSuccess);
    for (uint32_t c = 0; c < n_compounds; ++c)
        if (comp_offsets[c + 1] < comp_offsets[c]) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_compounds_upload: comp_offsets must be non-decreasing", cudaSuccess);
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    bam3d_compounds* K = new (std::nothrow) bam3d_compounds();
    if (!K) return BAM3D_ERR_OOM;
    K->n_compounds = n_compounds; K->n_prims = n_prims; K->meshes = meshes;
    K->offsets.assign(comp_offsets, comp_offsets + n_compounds + 1);
    K->kind.assign(kind, kind + n_prims); K->params.assign(params, params + 4 * (size_t)n_prims);
    if (mesh_id) K->mesh_id.assign(mesh_id, mesh_id + n_prims);
    std::vector<uint32_t> owner(n_prims ? n_prims : 1, 0);
    for (uint32_t c = 0; c < n_compounds; ++c) for (uint32_t i = comp_offsets[c]; i < comp_offsets[c + 1]; ++i) owner[i] = c;
    // the local transforms double as the initial world transforms: validates kinds, mesh ids and affinity once
    int32_t rc = bam3d_shapes_upload(ctx, kind, params, local_xform, mesh_id, n_prims, meshes, &K->prims);
    if (rc) { bam3d_compounds_free(ctx, K); return rc; }
    const size_t np = n_prims ? n_prims : 1;
    cudaError_t e;
    if ((e = cudaMalloc(&K->d_local, np * 64)) != cudaSuccess || (e = cudaMalloc(&K->d_owner, np * 4)) != cudaSuccess ||
        (e = cudaMalloc(&K->d_offsets, ((size_t)n_compounds + 1) * 4)) != cudaSuccess) {
        bam3d_compounds_free(ctx, K);
        return b3_fail(ctx, BAM3D_ERR_OOM, "cudaMalloc(compounds)", e);
    }
    if (n_prims) {
        cudaMemcpyAsync(K->d_local, local_xform, (size_t)n_prims * 64, cudaMemcpyHostToDevice, ctx->stream);
        cudaMemcpyAsync(K->d_owner, owner.data(), (size_t)n_prims * 4, cudaMemcpyHostToDevice, ctx->stream);
    }
    cudaMemcpyAsync(K->d_offsets, comp_offsets, ((size_t)n_compounds + 1) * 4, cudaMemcpyHostToDevice, ctx->stream);
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { bam3d_compounds_free(ctx, K); return b3_fail(ctx, BAM3D_ERR_CUDA, "bam3d_compounds_upload", e); }
    *out = K;
    return BAM3D_OK;
}

int32_t bam3d_gjk_contact_complex(bam3d_ctx* ctx, bam3d_compounds* K, const float* comp_xform, const uint32_t* pairs, uint64_t n,
                                  const bam3d_gjk_settings* settings, int32_t strategy, bam3d_contact* out_contacts, uint8_t* out_status) {
    return compounds_run(ctx, K, 0, comp_xform, nullptr, pairs, n, settings, strategy, out_contacts, nullptr, out_status);
}
int32_t bam3d_gjk_distance_complex(bam3d_ctx* ctx, bam3d_compounds* K, const float* comp_xform, const uint32_t* pairs, uint64_t n,
                                   const bam3d_gjk_settings* settings, float* out_value, uint8_t* out_status) {
    return compounds_run(ctx, K, 1, comp_xform, nullptr, pairs, n, settings, 0, nullptr, out_value, out_status);
}
int32_t bam3d_gjk_time_of_impact_complex(bam3d_ctx* ctx, bam3d_compounds* K, const float* comp_xform_start, const float* comp_xform_end,
                                         const uint32_t* pairs, uint64_t n, const bam3d_gjk_settings* settings, int32_t strategy,
                                         bam3d_contact* out_contacts, uint8_t* out_status) {
    return compounds_run(ctx, K, 2, comp_xform_start, comp_xform_end, pairs, n, settings, strategy, out_contacts, nullptr, out_status);
}

}  // extern "C"

// ============================================================================================================
// broad phase
// ============================================================================================================
struct bam3d_bvh {
    uint32_t n = 0;
    float* d_aabbs = nullptr;      // n x 6 (kept for DbvtBroadPhase queries)
    float4* nodes = nullptr;       // 2 x (2n - 1): node records (broad.h: BvhDev)
    int2* children = nullptr;      // n - 1 (build / refit only)
    uint32_t* leaf_value = nullptr;
    uint32_t* rope = nullptr;      // 2n - 1 (written into the node records; kept for the refit)
    float4* d_spheres = nullptr;   // n x (centre, radius) for a tree over volume::Sphere bounds (d_aabbs then holds the cover boxes)
};

static BvhDev bvh_view(const bam3d_bvh* t) {
    BvhDev v; v.n = t->n; v.nodes = t->nodes; v.leaf_value = t->leaf_value; v.spheres = t->d_spheres;
    return v;
}

extern "C" {

// Frustum::from_mat4 (frustum.rs:46-75) on the host: rows of the matrix; NB near / far are built from the same rows as
// left / right (w + x, w - x) exactly as the reference does. out24 = 6 x (n.xyz, d) in the order left, right, bottom,
// top, near, far. Returns BAM3D_ERR_ARG when a plane normal is zero (the reference returns None).
int32_t bam3d_frustum_from_mat4(const float* m, float* out24) {
    if (!m || !out24) return BAM3D_ERR_ARG;
    // transpose: row r = (m[r], m[4 + r], m[8 + r], m[12 + r])
    auto row = [&](int r, float* v) { v[0] = m[r]; v[1] = m[4 + r]; v[2] = m[8 + r]; v[3] = m[12 + r]; };
    float x[4], y[4], w[4];
    row(0, x); row(1, y); row(3, w);
    const float* src[6] = {x, x, y, y, x, x};
    const float sgn[6] = {1.f, -1.f, 1.f, -1.f, 1.f, -1.f};
    for (int k = 0; k < 6; ++k) {
        float v[4];
        for (int c = 0; c < 4; ++c) v[c] = sgn[k] > 0.f ? w[c] + src[k][c] : w[c] - src[k][c];
        // Plane::from_vector4_alt: n = v.xyz, d = -v.w; Plane::normalize: * 1 / sqrt(n.n)
        float nx = v[0], ny = v[1], nz = v[2], d = -v[3];
        if (nx == 0.f && ny == 0.f && nz == 0.f) return BAM3D_ERR_ARG;
        float denom = 1.0f / sqrtf((nx * nx + ny * ny) + nz * nz);
        out24[4 * k] = nx * denom; out24[4 * k + 1] = ny * denom; out24[4 * k + 2] = nz * denom; out24[4 * k + 3] = d * denom;
    }
    return BAM3D_OK;
}

// ---- SweepAndPrune3 --------------------------------------------------------------------------------------------
// Pair search strategies (identical output): the tiled forward sweep does one test per (i, j) whose sweep-axis
// intervals overlap — n x ~24k tests for the C4 scene (a 1-axis sweep prunes little in a 3-D uniform cloud) — while a
// BVH overlap query over the SORTED boxes (value index = sorted slot) does O(log n + hits) per box. "auto" counts the
// sweep's candidate tests exactly (binary search per slot) and takes the BVH when they exceed 64 per box.
// shard / n_shards: only the pairs whose higher sorted slot j lies in [n * shard / n_shards, n * (shard + 1) / n_shards)
// are searched and returned (the sort is replicated on every GPU, the sweep is sharded by sorted-index range;
// the shards' outputs, concatenated in shard order, are the unsharded output).
// d_spheres != nullptr: SweepAndPrune3<volume::Sphere> — d_aabbs then holds the spheres' extents (sort keys and Variance3 are
// exactly the Aabb3 path's on those), the candidate search runs on the grown boxes with the BVH strategy (the tiled sweep relies on
// sorted mins, which growing does not keep), and the candidates are filtered by the reference's two tests (still-active on the sweep
// axis, Sphere::intersects) in order; d_perm must then be given.
static int32_t sap_run(bam3d_ctx* ctx, const float* d_aabbs, uint64_t n64, int32_t* inout_axis, uint32_t* d_perm, uint32_t* d_pairs,
                       uint64_t capacity, uint64_t* out_count, uint32_t shard = 0, uint32_t n_shards = 1, const float4* d_spheres = nullptr) {
    const uint32_t n = (uint32_t)n64;
    if (n_shards == 0 || shard >= n_shards) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_sap_find_pairs: shard index out of range", cudaSuccess);
    const uint32_t j_begin = (uint32_t)(n64 * shard / n_shards), j_end = (uint32_t)(n64 * (shard + 1) / n_shards);
    int axis = *inout_axis;
    if (axis < 0 || axis > 2) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_sap_find_pairs: sweep axis must be 0, 1 or 2", cudaSuccess);
    *out_count = 0;
    if (n <= 1) {
        if (n == 1 && d_perm) cudaMemsetAsync(d_perm, 0, 4, ctx->stream);
        return BAM3D_OK;
    }
    int32_t rc;
    SapScratch S;
    const size_t cap1 = capacity ? capacity : 1;
    size_t tmp_bytes = sap_cub_tmp_bytes(n, cap1);
    const size_t bvh_tmp = bvh_cub_tmp_bytes(n), scan_tmp = scan_tmp_bytes(n);
    if (bvh_tmp > tmp_bytes) tmp_bytes = bvh_tmp;
    if (scan_tmp > tmp_bytes) tmp_bytes = scan_tmp;
    if ((rc = ctx->bp0.reserve(ctx, (size_t)n * 8 * 2)) || (rc = ctx->bp1.reserve(ctx, (size_t)n * 4 * 2)) || (rc = ctx->bp2.reserve(ctx, (size_t)n * 16 * 2)) ||
        (rc = ctx->bp3.reserve(ctx, cap1 * 8 * 2)) || (rc = ctx->cub_tmp.reserve(ctx, tmp_bytes)) || (rc = ctx->small.reserve(ctx, 1024)) ||
        (rc = ctx->bp7.reserve(ctx, sap_variance_scratch_bytes(n))))
        return rc;
    S.keys_in = (unsigned long long*)ctx->bp0.ptr; S.keys_out = S.keys_in + n;
    S.vals_in = (uint32_t*)ctx->bp1.ptr; S.vals_out = S.vals_in + n;
    S.lo = (float4*)ctx->bp2.ptr; S.hi = S.lo + n;
    S.pair_keys = (unsigned long long*)ctx->bp3.ptr; S.pair_keys_alt = S.pair_keys + cap1;
    S.cub_tmp = ctx->cub_tmp.ptr; S.cub_tmp_bytes = tmp_bytes;
    unsigned long long* d_count = (unsigned long long*)((char*)ctx->small.ptr + 832);
    int* d_axis = (int*)((char*)ctx->small.ptr + 840);
    unsigned long long* d_cand = (unsigned long long*)((char*)ctx->small.ptr + 848);
    float* d_sums = (float*)((char*)ctx->small.ptr + 864);
    const int strategy = d_spheres ? 2 : ctx->opt_sap_strategy;
    // BVH strategy scratch (one slab): sorted 6-float rows, node bounds, topology, per-slot counts / offsets
    float* d_sorted6 = nullptr;
    char* slab = nullptr;
    const size_t nn = n;
    if (strategy != 1) {
        const size_t bytes = nn * 24 + 2 * nn * 16 * 2 + nn * 8 + nn * 4 + 2 * nn * 4 + nn * 4 * 4 + 2 * nn * 4 + nn * 4 + (nn + 1) * 4 + (nn + 1) * 8 + 17 * 256;
        if ((rc = ctx->bp6.reserve(ctx, bytes))) return rc;
        slab = (char*)ctx->bp6.ptr;
        d_sorted6 = (float*)slab;
    }
    { KernelTimer kt(ctx);
      ctx->launches += launch_sap_sort(ctx->stream, d_aabbs, n, axis, S, d_perm, d_sorted6, d_cand); }
    // Variance3 on the side stream, beside the pair search
    cudaEventRecord(ctx->ev_a, ctx->stream);
    cudaStreamWaitEvent(ctx->stream2, ctx->ev_a, 0);
    ctx->launches += launch_sap_variance(ctx->stream2, S, n, ctx->bp7.ptr, d_sums, d_axis);
    cudaEventRecord(ctx->ev_b, ctx->stream2);
    if (d_spheres) {  // Variance3 reads the exact extents: grow the sorted boxes only after it has
        cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0);
        ctx->launches += launch_sap_inflate(ctx->stream, S, d_sorted6, n);
    }
    if ((rc = check_launch(ctx, "sap sort"))) return rc;
    bool use_bvh = strategy == 2;
    if (strategy == 0) {
        unsigned long long h_cand = 0;
        CU_TRY(cudaMemcpyAsync(&h_cand, d_cand, 8, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(candidates)");
        CU_TRY(cudaStreamSynchronize(ctx->stream), "sap sort");
        use_bvh = h_cand > 64ull * n;
    }
    unsigned long long h_count = 0;
    if (!use_bvh) {
        { KernelTimer kt(ctx);
          ctx->launches += launch_sap_sweep(ctx->stream, n, axis, j_begin, j_end, S, capacity, d_count); }
        CU_TRY(cudaMemcpyAsync(&h_count, d_count, 8, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(count)");
        CU_TRY(cudaStreamSynchronize(ctx->stream), "sap sweep");
    } else {
        char* q = slab;
        auto carve = [&](size_t bytes) { char* r = q; q += (bytes + 255) & ~(size_t)255; return r; };
        carve(nn * 24);  // d_sorted6
        float4* nnodes = (float4*)carve(2 * 2 * nn * 16);
        int2* children = (int2*)carve(nn * 8);
        uint32_t* leaf_value = (uint32_t*)carve(nn * 4);
        uint32_t* rope = (uint32_t*)carve(2 * nn * 4);
        BvhBuildScratch W;
        W.morton_in = (uint32_t*)carve(nn * 4); W.morton_out = (uint32_t*)carve(nn * 4);
        W.idx_in = (uint32_t*)carve(nn * 4); W.idx_out = (uint32_t*)carve(nn * 4);
        W.parent = (uint32_t*)carve(2 * nn * 4);
        W.flags = (uint32_t*)carve(nn * 4);
        uint32_t* d_counts = (uint32_t*)carve((nn + 1) * 4);
        unsigned long long* d_offsets = (unsigned long long*)carve((nn + 1) * 8);
        W.cub_tmp = ctx->cub_tmp.ptr; W.cub_tmp_bytes = tmp_bytes;
        W.scene = (float*)((char*)ctx->small.ptr + 896);
        BvhDev T; T.n = n; T.nodes = nnodes; T.leaf_value = leaf_value;
        const int shift = sap_bits_for(n);
        { KernelTimer kt(ctx);
          ctx->launches += launch_bvh_build(ctx->stream, d_sorted6, n, W, nnodes, children, leaf_value, rope);
          ctx->launches += launch_bvh_overlap_append(ctx->stream, T, d_sorted6, j_begin, j_end, shift, S.pair_keys, capacity, d_count); }
        CU_TRY(cudaMemcpyAsync(&h_count, d_count, 8, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(count)");
        CU_TRY(cudaStreamSynchronize(ctx->stream), "sap bvh pair search");
        (void)d_counts; (void)d_offsets;
    }
    if ((rc = check_launch(ctx, "sap pair search"))) return rc;
    *out_count = h_count;
    if (h_count > capacity) {
        cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0);  // keep the side stream ordered before the scratch is reused
        char msg[160];
        std::snprintf(msg, sizeof msg, "bam3d_sap_find_pairs: %llu pairs found, capacity %llu", h_count, (unsigned long long)capacity);
        return b3_fail(ctx, BAM3D_ERR_CAPACITY, msg, cudaSuccess);
    }
    if (!d_spheres) {
        KernelTimer kt(ctx);
        ctx->launches += launch_sap_finish(ctx->stream, S, h_count, (uint2*)d_pairs, n);
    } else {
        const size_t filter_tmp = sphere_filter_tmp_bytes(h_count);
        const size_t cand_bytes = (((size_t)(h_count ? h_count : 1) * 8) + 255) & ~(size_t)255;  // keeps the float4 array behind it aligned
        if ((rc = ctx->sph1.reserve(ctx, cand_bytes + (size_t)n * 16)) || (rc = ctx->sph2.reserve(ctx, filter_tmp + 256))) return rc;
        uint2* cand = (uint2*)ctx->sph1.ptr;
        float4* sorted_spheres = (float4*)((char*)ctx->sph1.ptr + cand_bytes);
        KernelTimer kt(ctx);
        ctx->launches += launch_sap_finish(ctx->stream, S, h_count, cand, n);
        ctx->launches += launch_gather_spheres(ctx->stream, d_spheres, d_perm, n, sorted_spheres);
        ctx->launches += launch_sphere_pair_filter(ctx->stream, sorted_spheres, axis, cand, h_count, (uint2*)d_pairs, d_count, ctx->sph2.ptr, filter_tmp);
        CU_TRY(cudaMemcpyAsync(&h_count, d_count, 8, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(count)");
    }
    int h_axis = axis;
    cudaStreamWaitEvent(ctx->stream, ctx->ev_b, 0);
    CU_TRY(cudaMemcpyAsync(&h_axis, d_axis, 4, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(axis)");
    CU_TRY(cudaStreamSynchronize(ctx->stream), "sap finish");
    *inout_axis = h_axis;
    if (d_spheres) *out_count = h_count;
    return check_launch(ctx, "sap finish");
}

// SweepAndPrune3<volume::Sphere>::find_collider_pairs; d_spheres = n x (centre.xyz, radius)
static int32_t sap_run_spheres(bam3d_ctx* ctx, const float4* d_spheres, uint64_t n, int32_t* inout_axis, uint32_t* d_perm, uint32_t* d_pairs,
                               uint64_t capacity, uint64_t* out_count) {
    int32_t rc;
    const size_t nn = n ? n : 1;
    if ((rc = ctx->sph0.reserve(ctx, nn * 24 + nn * 4))) return rc;
    float* d_extents = (float*)ctx->sph0.ptr;
    if (!d_perm) d_perm = (uint32_t*)((char*)ctx->sph0.ptr + nn * 24);
    ctx->launches += launch_spheres_to_extents(ctx->stream, d_spheres, (uint32_t)n, d_extents);
    return sap_run(ctx, d_extents, n, inout_axis, d_perm, d_pairs, capacity, out_count, 0, 1, d_spheres);
}

int32_t bam3d_sap_find_pairs_shard_dev(bam3d_ctx* ctx, const float* d_aabbs, uint64_t n, int32_t* inout_sweep_axis, uint32_t* d_perm,
                                       uint32_t* d_pairs, uint64_t capacity, uint64_t* out_count, uint32_t shard, uint32_t n_shards) {
    if (!ctx || !inout_sweep_axis || !out_count || (n && !d_aabbs) || (capacity && !d_pairs)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_sap_find_pairs: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, n)) || (rc = check_n(ctx, capacity))) return rc;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    return sap_run(ctx, d_aabbs, n, inout_sweep_axis, d_perm, d_pairs, capacity, out_count, shard, n_shards);
}
int32_t bam3d_sap_find_pairs_dev(bam3d_ctx* ctx, const float* d_aabbs, uint64_t n, int32_t* inout_sweep_axis, uint32_t* d_perm,
                                 uint32_t* d_pairs, uint64_t capacity, uint64_t* out_count) {
    return bam3d_sap_find_pairs_shard_dev(ctx, d_aabbs, n, inout_sweep_axis, d_perm, d_pairs, capacity, out_count, 0, 1);
}

int32_t bam3d_sap_find_pairs(bam3d_ctx* ctx, const float* aabbs, uint64_t n, int32_t* inout_sweep_axis, uint32_t* out_perm,
                             uint32_t* out_pairs, uint64_t capacity, uint64_t* out_count) {
    return bam3d_sap_find_pairs_shard(ctx, aabbs, n, inout_sweep_axis, out_perm, out_pairs, capacity, out_count, 0, 1);
}
int32_t bam3d_sap_find_pairs_shard(bam3d_ctx* ctx, const float* aabbs, uint64_t n, int32_t* inout_sweep_axis, uint32_t* out_perm,
                                   uint32_t* out_pairs, uint64_t capacity, uint64_t* out_count, uint32_t shard, uint32_t n_shards) {
    if (!ctx || !inout_sweep_axis || !out_count || (n && !aabbs) || (capacity && !out_pairs)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_sap_find_pairs: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, n)) || (rc = check_n(ctx, capacity))) return rc;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    if ((rc = ctx->bp4.reserve(ctx, (size_t)(n ? n : 1) * 24 + (size_t)(n ? n : 1) * 4)) || (rc = ctx->bp5.reserve(ctx, (capacity ? capacity : 1) * 8))) return rc;
    float* d_aabbs = (float*)ctx->bp4.ptr;
    uint32_t* d_perm = (uint32_t*)((char*)ctx->bp4.ptr + (size_t)(n ? n : 1) * 24);
    uint32_t* d_pairs = (uint32_t*)ctx->bp5.ptr;
    if (n) CU_TRY(cudaMemcpyAsync(d_aabbs, aabbs, n * 24, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(aabbs)");
    rc = sap_run(ctx, d_aabbs, n, inout_sweep_axis, d_perm, d_pairs, capacity, out_count, shard, n_shards);
    if (rc != BAM3D_OK) return rc;
    if (out_perm && n) CU_TRY(cudaMemcpyAsync(out_perm, d_perm, n * 4, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(perm)");
    if (*out_count) CU_TRY(cudaMemcpyAsync(out_pairs, d_pairs, *out_count * 8, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(pairs)");
    CU_TRY(cudaStreamSynchronize(ctx->stream), "bam3d_sap_find_pairs");
    return BAM3D_OK;
}

int32_t bam3d_sap_find_pairs_spheres_dev(bam3d_ctx* ctx, const float* d_spheres, uint64_t n, int32_t* inout_sweep_axis, uint32_t* d_perm,
                                         uint32_t* d_pairs, uint64_t capacity, uint64_t* out_count) {
    if (!ctx || !inout_sweep_axis || !out_count || (n && !d_spheres) || (capacity && !d_pairs)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_sap_find_pairs_spheres: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, n)) || (rc = check_n(ctx, capacity))) return rc;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    return sap_run_spheres(ctx, (const float4*)d_spheres, n, inout_sweep_axis, d_perm, d_pairs, capacity, out_count);
}
int32_t bam3d_sap_find_pairs_spheres(bam3d_ctx* ctx, const float* spheres, uint64_t n, int32_t* inout_sweep_axis, uint32_t* out_perm,
                                     uint32_t* out_pairs, uint64_t capacity, uint64_t* out_count) {
    if (!ctx || !inout_sweep_axis || !out_count || (n && !spheres) || (capacity && !out_pairs)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_sap_find_pairs_spheres: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, n)) || (rc = check_n(ctx, capacity))) return rc;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t nn = n ? n : 1;
    if ((rc = ctx->bp4.reserve(ctx, nn * 16 + nn * 4)) || (rc = ctx->bp5.reserve(ctx, (capacity ? capacity : 1) * 8))) return rc;
    float4* d_spheres = (float4*)ctx->bp4.ptr;
    uint32_t* d_perm = (uint32_t*)((char*)ctx->bp4.ptr + nn * 16);
    uint32_t* d_pairs = (uint32_t*)ctx->bp5.ptr;
    if (n) CU_TRY(cudaMemcpyAsync(d_spheres, spheres, n * 16, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(spheres)");
    rc = sap_run_spheres(ctx, d_spheres, n, inout_sweep_axis, d_perm, d_pairs, capacity, out_count);
    if (rc != BAM3D_OK) return rc;
    if (out_perm && n) CU_TRY(cudaMemcpyAsync(out_perm, d_perm, n * 4, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(perm)");
    if (*out_count) CU_TRY(cudaMemcpyAsync(out_pairs, d_pairs, *out_count * 8, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(pairs)");
    CU_TRY(cudaStreamSynchronize(ctx->stream), "bam3d_sap_find_pairs_spheres");
    return BAM3D_OK;
}

// Variance3 (sweep_prune.rs:166-216) over the boxes in the order given: add_bound for each, then compute_axis.
int32_t bam3d_sap_variance(bam3d_ctx* ctx, const float* aabbs, uint64_t n64, float* out_sums6, int32_t* out_axis) {
    if (!ctx || !out_sums6 || !out_axis || (n64 && !aabbs)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_sap_variance: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, n64))) return rc;
    const uint32_t n = (uint32_t)n64;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t nn = n ? n : 1;
    if ((rc = ctx->bp4.reserve(ctx, nn * 24)) || (rc = ctx->bp2.reserve(ctx, nn * 16 * 2)) || (rc = ctx->small.reserve(ctx, 1024)) ||
        (rc = ctx->bp7.reserve(ctx, sap_variance_scratch_bytes(n) + 64))) return rc;
    SapScratch S{};
    S.lo = (float4*)ctx->bp2.ptr; S.hi = S.lo + nn;
    float* d_sums = (float*)((char*)ctx->small.ptr + 864);
    int* d_axis = (int*)((char*)ctx->small.ptr + 840);
    if (n) CU_TRY(cudaMemcpyAsync(ctx->bp4.ptr, aabbs, (size_t)n * 24, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(aabbs)");
    ctx->launches += launch_boxes_to_lohi(ctx->stream, (const float*)ctx->bp4.ptr, n, S.lo, S.hi);
    { KernelTimer kt(ctx);
      ctx->launches += launch_sap_variance(ctx->stream, S, n, ctx->bp7.ptr, d_sums, d_axis); }
    if ((rc = check_launch(ctx, "sap_variance_chain_kernel"))) return rc;
#ifdef B3_VAR_TRACE
    { int h[6]; cudaMemcpyAsync(h, d_sums + 16, 24, cudaMemcpyDeviceToHost, ctx->stream); cudaStreamSynchronize(ctx->stream);
      fprintf(stderr, "[variance trace] sequential tiles per chain: %d %d %d %d %d %d of %u\n", h[0], h[1], h[2], h[3], h[4], h[5], (n + 4095) / 4096);
      cudaMemsetAsync(d_sums + 16, 0, 24, ctx->stream); }
#endif
    CU_TRY(cudaMemcpyAsync(out_sums6, d_sums, 24, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(sums)");
    CU_TRY(cudaMemcpyAsync(out_axis, d_axis, 4, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(axis)");
    CU_TRY(cudaStreamSynchronize(ctx->stream), "bam3d_sap_variance");
    return BAM3D_OK;
}

// ---- world AABBs -----------------------------------------------------------------------------------------------
int32_t bam3d_shapes_world_aabbs_dev(bam3d_ctx* ctx, const bam3d_shapes* shapes, float* d_out6) {
    if (!ctx || !shapes || (shapes->n && !d_out6)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_shapes_world_aabbs: null argument", cudaSuccess);
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    ctx->launches += launch_world_aabbs(ctx->stream, view(ctx, shapes), d_out6);
    return check_launch(ctx, "world_aabb_kernel");
}
int32_t bam3d_shapes_world_aabbs(bam3d_ctx* ctx, const bam3d_shapes* shapes, float* out6) {
    if (!ctx || !shapes || (shapes->n && !out6)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_shapes_world_aabbs: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = ctx->bp4.reserve(ctx, (size_t)(shapes->n ? shapes->n : 1) * 24))) return rc;
    if ((rc = bam3d_shapes_world_aabbs_dev(ctx, shapes, (float*)ctx->bp4.ptr))) return rc;
    if (shapes->n) CU_TRY(cudaMemcpyAsync(out6, ctx->bp4.ptr, (size_t)shapes->n * 24, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(aabbs)");
    CU_TRY(cudaStreamSynchronize(ctx->stream), "bam3d_shapes_world_aabbs");
    return BAM3D_OK;
}

int32_t bam3d_shapes_world_spheres_dev(bam3d_ctx* ctx, const bam3d_shapes* shapes, float* d_out4) {
    if (!ctx || !shapes || (shapes->n && !d_out4)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_shapes_world_spheres: null argument", cudaSuccess);
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    ctx->launches += launch_world_spheres(ctx->stream, view(ctx, shapes), (float4*)d_out4);
    return check_launch(ctx, "world_sphere_kernel");
}
int32_t bam3d_shapes_world_spheres(bam3d_ctx* ctx, const bam3d_shapes* shapes, float* out4) {
    if (!ctx || !shapes || (shapes->n && !out4)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_shapes_world_spheres: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = ctx->bp4.reserve(ctx, (size_t)(shapes->n ? shapes->n : 1) * 16))) return rc;
    if ((rc = bam3d_shapes_world_spheres_dev(ctx, shapes, (float*)ctx->bp4.ptr))) return rc;
    if (shapes->n) CU_TRY(cudaMemcpyAsync(out4, ctx->bp4.ptr, (size_t)shapes->n * 16, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(spheres)");
    CU_TRY(cudaStreamSynchronize(ctx->stream), "bam3d_shapes_world_spheres");
    return BAM3D_OK;
}

// ---- broad -> narrow pipeline (BASELINE configs[3]: "SweepAndPrune3 ... then narrow phase") ---------------------------------
// world AABBs of the shape set (ComputeBound<Aabb3> + Aabb::transform) -> SweepAndPrune3::find_collider_pairs over them ->
// sorted-position pairs mapped back to shape indices through perm -> GJK::intersection on every candidate pair. Everything
// stays on the device; the one host synchronisation is sweep-and-prune's own (the pair count sizes its final sort and the
// narrow-phase launch).
int32_t bam3d_pipeline_contacts_dev(bam3d_ctx* ctx, const bam3d_shapes* shapes, int32_t* inout_sweep_axis, const bam3d_gjk_settings* settings,
                                    int32_t strategy, uint32_t shard, uint32_t n_shards, uint32_t* d_pairs, uint32_t* d_sorted_pairs,
                                    uint32_t* d_perm, bam3d_contact* d_contacts, uint8_t* d_status, uint64_t capacity, uint64_t* out_count) {
    if (!ctx || !shapes || !inout_sweep_axis || !out_count || (capacity && (!d_pairs || !d_contacts || !d_status)))
        return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_pipeline_contacts: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, capacity))) return rc;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    const uint64_t n = shapes->n;
    *out_count = 0;
    if ((rc = ctx->pipe0.reserve(ctx, (size_t)(n ? n : 1) * 24))) return rc;
    if (!d_perm) { if ((rc = ctx->pipe1.reserve(ctx, (size_t)(n ? n : 1) * 4))) return rc; d_perm = (uint32_t*)ctx->pipe1.ptr; }
    if (!d_sorted_pairs) { if ((rc = ctx->pipe2.reserve(ctx, (size_t)(capacity ? capacity : 1) * 8))) return rc; d_sorted_pairs = (uint32_t*)ctx->pipe2.ptr; }
    if ((rc = bam3d_shapes_world_aabbs_dev(ctx, shapes, (float*)ctx->pipe0.ptr))) return rc;
    if ((rc = sap_run(ctx, (const float*)ctx->pipe0.ptr, n, inout_sweep_axis, d_perm, d_sorted_pairs, capacity, out_count, shard, n_shards))) return rc;
    const uint64_t m = *out_count;
    if (m == 0) return BAM3D_OK;
    ctx->launches += launch_remap_pairs(ctx->stream, (const uint2*)d_sorted_pairs, d_perm, m, (uint2*)d_pairs);
    if ((rc = check_launch(ctx, "remap_pairs_kernel"))) return rc;
    return bam3d_gjk_contact_dev(ctx, shapes, d_pairs, m, settings, strategy, d_contacts, d_status);
}

int32_t bam3d_pipeline_contacts(bam3d_ctx* ctx, const bam3d_shapes* shapes, int32_t* inout_sweep_axis, const bam3d_gjk_settings* settings,
                                int32_t strategy, uint32_t* out_pairs, uint32_t* out_sorted_pairs, uint32_t* out_perm, bam3d_contact* out_contacts,
                                uint8_t* out_status, uint64_t capacity, uint64_t* out_count) {
    if (!ctx || !shapes || !inout_sweep_axis || !out_count || (capacity && (!out_pairs || !out_contacts || !out_status)))
        return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_pipeline_contacts: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, capacity))) return rc;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    const size_t cap = capacity ? capacity : 1, nn = shapes->n ? shapes->n : 1;
    if ((rc = ctx->pipe3.reserve(ctx, cap * 8 + 256)) || (rc = ctx->contacts.reserve(ctx, cap * 32)) || (rc = ctx->status.reserve(ctx, cap)) ||
        (rc = ctx->pipe1.reserve(ctx, nn * 4)) || (rc = ctx->pipe2.reserve(ctx, cap * 8))) return rc;
    rc = bam3d_pipeline_contacts_dev(ctx, shapes, inout_sweep_axis, settings, strategy, 0, 1, (uint32_t*)ctx->pipe3.ptr, (uint32_t*)ctx->pipe2.ptr,
                                     (uint32_t*)ctx->pipe1.ptr, (bam3d_contact*)ctx->contacts.ptr, (uint8_t*)ctx->status.ptr, capacity, out_count);
    if (rc) return rc;
    const uint64_t m = *out_count;
    if (out_perm && shapes->n) CU_TRY(cudaMemcpyAsync(out_perm, ctx->pipe1.ptr, (size_t)shapes->n * 4, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(perm)");
    if (m) {
        CU_TRY(cudaMemcpyAsync(out_pairs, ctx->pipe3.ptr, m * 8, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(pairs)");
        if (out_sorted_pairs) CU_TRY(cudaMemcpyAsync(out_sorted_pairs, ctx->pipe2.ptr, m * 8, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(sorted pairs)");
        CU_TRY(cudaMemcpyAsync(out_contacts, ctx->contacts.ptr, m * 32, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(contacts)");
        CU_TRY(cudaMemcpyAsync(out_status, ctx->status.ptr, m, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(status)");
    }
    CU_TRY(cudaStreamSynchronize(ctx->stream), "bam3d_pipeline_contacts");
    return BAM3D_OK;
}

// ---- BVH (DBVT queries) ----------------------------------------------------------------------------------------
void bam3d_bvh_free(bam3d_ctx* ctx, bam3d_bvh* t) {
    (void)ctx;
    if (!t) return;
    if (t->d_aabbs) cudaFree(t->d_aabbs);
    if (t->nodes) cudaFree(t->nodes);
    if (t->children) cudaFree(t->children);
    if (t->leaf_value) cudaFree(t->leaf_value);
    if (t->rope) cudaFree(t->rope);
    if (t->d_spheres) cudaFree(t->d_spheres);
    delete t;
}

// `spheres`: the values' bounds are volume::Sphere (n x 4 floats in `aabbs`' place); the tree is then built over their cover boxes
static int32_t bvh_build_impl(bam3d_ctx* ctx, const float* aabbs, bool on_device, uint64_t n64, bam3d_bvh** out, bool spheres = false) {
    *out = nullptr;
    int32_t rc;
    if ((rc = check_n(ctx, n64))) return rc;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    const uint32_t n = (uint32_t)n64;
    bam3d_bvh* t = new (std::nothrow) bam3d_bvh();
    if (!t) return BAM3D_ERR_OOM;
    t->n = n;
    const size_t nn = n ? n : 1;
    cudaError_t e;
    if ((e = cudaMalloc(&t->d_aabbs, nn * 24)) != cudaSuccess || (e = cudaMalloc(&t->nodes, (2 * 2 * nn) * 16)) != cudaSuccess ||
        (e = cudaMalloc(&t->children, nn * 8)) != cudaSuccess ||
        (e = cudaMalloc(&t->leaf_value, nn * 4)) != cudaSuccess || (e = cudaMalloc(&t->rope, 2 * nn * 4)) != cudaSuccess) {
        bam3d_bvh_free(ctx, t);
        return b3_fail(ctx, BAM3D_ERR_OOM, "cudaMalloc(bvh)", e);
    }
    if (spheres) {
        if ((e = cudaMalloc(&t->d_spheres, nn * 16)) != cudaSuccess) { bam3d_bvh_free(ctx, t); return b3_fail(ctx, BAM3D_ERR_OOM, "cudaMalloc(bvh spheres)", e); }
        if (n) {
            cudaMemcpyAsync(t->d_spheres, aabbs, (size_t)n * 16, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream);
            ctx->launches += launch_sphere_cover_boxes(ctx->stream, t->d_spheres, n, t->d_aabbs);
        }
    } else if (n) {
        cudaMemcpyAsync(t->d_aabbs, aabbs, (size_t)n * 24, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream);
    }
    BvhBuildScratch W;
    const size_t tmp_bytes = bvh_cub_tmp_bytes(n ? n : 1);
    if ((rc = ctx->bp0.reserve(ctx, nn * 4 * 4)) || (rc = ctx->bp1.reserve(ctx, (2 * nn) * 4 + nn * 4)) || (rc = ctx->cub_tmp.reserve(ctx, tmp_bytes)) ||
        (rc = ctx->small.reserve(ctx, 1024))) { bam3d_bvh_free(ctx, t); return rc; }
    W.morton_in = (uint32_t*)ctx->bp0.ptr; W.morton_out = W.morton_in + nn; W.idx_in = W.morton_out + nn; W.idx_out = W.idx_in + nn;
    W.parent = (uint32_t*)ctx->bp1.ptr; W.flags = W.parent + 2 * nn;
    W.cub_tmp = ctx->cub_tmp.ptr; W.cub_tmp_bytes = tmp_bytes;
    W.scene = (float*)((char*)ctx->small.ptr + 896);
    { KernelTimer kt(ctx);
      ctx->launches += launch_bvh_build(ctx->stream, t->d_aabbs, n, W, t->nodes, t->children, t->leaf_value, t->rope); }
    e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) { bam3d_bvh_free(ctx, t); return b3_fail(ctx, BAM3D_ERR_CUDA, "bam3d_bvh_build", e); }
    *out = t;
    return BAM3D_OK;
}

static int32_t bvh_refit_impl(bam3d_ctx* ctx, bam3d_bvh* t, const float* aabbs, bool on_device) {
    if (!ctx || !t || (t->n && !aabbs)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_refit: null argument", cudaSuccess);
    if (t->d_spheres) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_refit: a tree over sphere bounds is rebuilt, not refitted", cudaSuccess);
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    const uint32_t n = t->n;
    if (n == 0) return BAM3D_OK;
    int32_t rc;
    const size_t nn = n;
    if ((rc = ctx->bp1.reserve(ctx, (2 * nn) * 4 + nn * 4))) return rc;
    uint32_t* parent = (uint32_t*)ctx->bp1.ptr;
    uint32_t* flags = parent + 2 * nn;
    CU_TRY(cudaMemcpyAsync(t->d_aabbs, aabbs, nn * 24, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(aabbs)");
    { KernelTimer kt(ctx);
      ctx->launches += launch_bvh_refit(ctx->stream, t->d_aabbs, n, t->children, t->leaf_value, parent, t->rope, flags, t->nodes); }
    if ((rc = check_launch(ctx, "bvh_refit_kernel"))) return rc;
    if (!on_device) CU_TRY(cudaStreamSynchronize(ctx->stream), "bam3d_bvh_refit");
    return BAM3D_OK;
}
int32_t bam3d_bvh_refit(bam3d_ctx* ctx, bam3d_bvh* tree, const float* aabbs) { return bvh_refit_impl(ctx, tree, aabbs, false); }
int32_t bam3d_bvh_refit_dev(bam3d_ctx* ctx, bam3d_bvh* tree, const float* d_aabbs) { return bvh_refit_impl(ctx, tree, d_aabbs, true); }

int32_t bam3d_bvh_build(bam3d_ctx* ctx, const float* aabbs, uint64_t n, bam3d_bvh** out) {
    if (!ctx || !out || (n && !aabbs)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_build: null argument", cudaSuccess);
    return bvh_build_impl(ctx, aabbs, false, n, out);
}
int32_t bam3d_bvh_build_dev(bam3d_ctx* ctx, const float* d_aabbs, uint64_t n, bam3d_bvh** out) {
    if (!ctx || !out || (n && !d_aabbs)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_build: null argument", cudaSuccess);
    return bvh_build_impl(ctx, d_aabbs, true, n, out);
}
int32_t bam3d_bvh_build_spheres(bam3d_ctx* ctx, const float* spheres, uint64_t n, bam3d_bvh** out) {
    if (!ctx || !out || (n && !spheres)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_build_spheres: null argument", cudaSuccess);
    return bvh_build_impl(ctx, spheres, false, n, out, true);
}
int32_t bam3d_bvh_build_spheres_dev(bam3d_ctx* ctx, const float* d_spheres, uint64_t n, bam3d_bvh** out) {
    if (!ctx || !out || (n && !d_spheres)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_build_spheres: null argument", cudaSuccess);
    return bvh_build_impl(ctx, d_spheres, true, n, out, true);
}
uint32_t bam3d_bvh_count(const bam3d_bvh* t) { return t ? t->n : 0; }

// which queries a tree answers: Aabb3 queries need box bounds, Sphere queries sphere bounds; rays and frusta work on both
static int32_t bvh_kind_check(bam3d_ctx* ctx, const bam3d_bvh* t, int kind) {
    if (kind == BQ_AABB && t->d_spheres) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_query_aabb: the tree's bounds are spheres (use bam3d_bvh_query_sphere)", cudaSuccess);
    if (kind == BQ_SPHERE && !t->d_spheres) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_query_sphere: the tree's bounds are boxes (use bam3d_bvh_query_aabb)", cudaSuccess);
    return BAM3D_OK;
}

// Shared implementation of the variable-length queries: count pass, exclusive scan, fill pass.
static int32_t bvh_query_impl(bam3d_ctx* ctx, const bam3d_bvh* t, int kind, const float* queries, uint64_t nq64, uint64_t* out_offsets,
                              uint32_t* out_values, void* out_payload, size_t payload_stride, uint64_t capacity, uint64_t* out_total) {
    int32_t rc;
    if ((rc = check_n(ctx, nq64)) || (rc = bvh_kind_check(ctx, t, kind))) return rc;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    const uint32_t nq = (uint32_t)nq64;
    *out_total = 0;
    if (nq == 0) { if (out_offsets) out_offsets[0] = 0; return BAM3D_OK; }
    const size_t qfloats = kind == BQ_FRUSTUM ? 24 : (kind == BQ_SPHERE ? 4 : 6);
    // few large queries -> evaluate the leaf predicate for every (query, value); many small ones -> tree traversal
    const uint32_t ntiles = (t->n + 255) / 256;
    bool brute = ctx->opt_bvh_strategy == 2 || (ctx->opt_bvh_strategy == 0 && nq < 2048 && (uint64_t)nq * t->n <= (1ull << 35));
    if (nq > 65535 || t->n == 0 || t->d_spheres) brute = false;  // (the flat scan tests boxes only)
    const size_t ncount = brute ? (size_t)nq * ntiles : (size_t)nq;
    const size_t scan_bytes = scan_tmp_bytes((uint32_t)ncount);
    if (ncount > 0xFFFFFF00ull) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_query: too many (query, tile) cells", cudaSuccess);
    if ((rc = ctx->bp0.reserve(ctx, nq * qfloats * 4)) || (rc = ctx->bp1.reserve(ctx, (ncount + 1) * 4)) ||
        (rc = ctx->bp2.reserve(ctx, (ncount + 1) * 8)) || (rc = ctx->cub_tmp.reserve(ctx, scan_bytes))) return rc;
    float* d_q = (float*)ctx->bp0.ptr;
    uint32_t* d_counts = (uint32_t*)ctx->bp1.ptr;
    unsigned long long* d_offsets = (unsigned long long*)ctx->bp2.ptr;
    CU_TRY(cudaMemcpyAsync(d_q, queries, nq * qfloats * 4, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(queries)");
    cudaMemsetAsync(d_counts + ncount, 0, 4, ctx->stream);
    BvhDev T = bvh_view(t);
    { KernelTimer kt(ctx);
      if (brute) ctx->launches += launch_bvh_brute(ctx->stream, t->d_aabbs, t->n, kind, d_q, nq, d_counts, nullptr, nullptr, nullptr);
      else ctx->launches += launch_bvh_query(ctx->stream, T, kind, d_q, nq, d_counts, nullptr, nullptr, nullptr); }
    ctx->launches += launch_exclusive_scan_u32_to_u64(ctx->stream, d_counts, d_offsets, (uint32_t)ncount, ctx->cub_tmp.ptr, scan_bytes);
    if ((rc = check_launch(ctx, "bvh query (count)"))) return rc;
    unsigned long long total = 0;
    CU_TRY(cudaMemcpyAsync(&total, d_offsets + ncount, 8, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(total)");
    if (out_offsets) {
        // per-query offsets: every ntiles-th entry of the cell offsets when brute
        const size_t stride = brute ? (size_t)ntiles * 8 : 8;
        CU_TRY(cudaMemcpy2DAsync(out_offsets, 8, d_offsets, stride, 8, (size_t)nq + 1, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpy2DAsync(offsets)");
    }
    CU_TRY(cudaStreamSynchronize(ctx->stream), "bvh query (count)");
    *out_total = total;
    if (total > capacity) {
        char msg[160];
        std::snprintf(msg, sizeof msg, "bam3d_bvh_query: %llu hits, capacity %llu", total, (unsigned long long)capacity);
        return b3_fail(ctx, BAM3D_ERR_CAPACITY, msg, cudaSuccess);
    }
    if (total == 0) return BAM3D_OK;
    if ((rc = ctx->bp3.reserve(ctx, total * 4)) || (payload_stride && (rc = ctx->bp4.reserve(ctx, total * payload_stride)))) return rc;
    { KernelTimer kt(ctx);
      if (brute) ctx->launches += launch_bvh_brute(ctx->stream, t->d_aabbs, t->n, kind, d_q, nq, d_counts, d_offsets, (uint32_t*)ctx->bp3.ptr, payload_stride ? ctx->bp4.ptr : nullptr);
      else ctx->launches += launch_bvh_query(ctx->stream, T, kind, d_q, nq, d_counts, d_offsets, (uint32_t*)ctx->bp3.ptr, payload_stride ? ctx->bp4.ptr : nullptr); }
    if ((rc = check_launch(ctx, "bvh query (fill)"))) return rc;
    CU_TRY(cudaMemcpyAsync(out_values, ctx->bp3.ptr, total * 4, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(values)");
    if (payload_stride) CU_TRY(cudaMemcpyAsync(out_payload, ctx->bp4.ptr, total * payload_stride, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(payload)");
    CU_TRY(cudaStreamSynchronize(ctx->stream), "bvh query (fill)");
    return BAM3D_OK;
}

int32_t bam3d_bvh_query_aabb(bam3d_ctx* ctx, const bam3d_bvh* t, const float* queries, uint64_t nq, uint64_t* out_offsets, uint32_t* out_values,
                             uint64_t capacity, uint64_t* out_total) {
    if (!ctx || !t || !out_total || (nq && !queries) || (capacity && !out_values)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_query_aabb: null argument", cudaSuccess);
    return bvh_query_impl(ctx, t, BQ_AABB, queries, nq, out_offsets, out_values, nullptr, 0, capacity, out_total);
}
int32_t bam3d_bvh_query_sphere(bam3d_ctx* ctx, const bam3d_bvh* t, const float* spheres, uint64_t nq, uint64_t* out_offsets, uint32_t* out_values,
                               uint64_t capacity, uint64_t* out_total) {
    if (!ctx || !t || !out_total || (nq && !spheres) || (capacity && !out_values)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_query_sphere: null argument", cudaSuccess);
    return bvh_query_impl(ctx, t, BQ_SPHERE, spheres, nq, out_offsets, out_values, nullptr, 0, capacity, out_total);
}
int32_t bam3d_bvh_query_ray(bam3d_ctx* ctx, const bam3d_bvh* t, const float* rays, uint64_t nq, uint64_t* out_offsets, uint32_t* out_values,
                            float* out_points, uint64_t capacity, uint64_t* out_total) {
    if (!ctx || !t || !out_total || (nq && !rays) || (capacity && (!out_values || !out_points))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_query_ray: null argument", cudaSuccess);
    return bvh_query_impl(ctx, t, BQ_RAY, rays, nq, out_offsets, out_values, out_points, 12, capacity, out_total);
}
int32_t bam3d_bvh_query_frustum(bam3d_ctx* ctx, const bam3d_bvh* t, const float* planes, uint64_t nq, uint64_t* out_offsets, uint32_t* out_values,
                                uint8_t* out_relation, uint64_t capacity, uint64_t* out_total) {
    if (!ctx || !t || !out_total || (nq && !planes) || (capacity && (!out_values || !out_relation))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_query_frustum: null argument", cudaSuccess);
    return bvh_query_impl(ctx, t, BQ_FRUSTUM, planes, nq, out_offsets, out_values, out_relation, 1, capacity, out_total);
}

int32_t bam3d_bvh_query_ray_closest(bam3d_ctx* ctx, const bam3d_bvh* t, const float* rays, uint64_t nq64, uint32_t* out_value, float* out_point_t) {
    if (!ctx || !t || (nq64 && (!rays || !out_value || !out_point_t))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_query_ray_closest: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, nq64))) return rc;
    if (nq64 == 0) return BAM3D_OK;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    const uint32_t nq = (uint32_t)nq64;
    if ((rc = ctx->bp0.reserve(ctx, (size_t)nq * 24)) || (rc = ctx->bp1.reserve(ctx, (size_t)nq * 4)) || (rc = ctx->bp2.reserve(ctx, (size_t)nq * 16))) return rc;
    CU_TRY(cudaMemcpyAsync(ctx->bp0.ptr, rays, (size_t)nq * 24, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(rays)");
    { KernelTimer kt(ctx);
      ctx->launches += launch_bvh_ray_closest(ctx->stream, bvh_view(t), (const float*)ctx->bp0.ptr, nq, (uint32_t*)ctx->bp1.ptr, (float*)ctx->bp2.ptr); }
    if ((rc = check_launch(ctx, "bvh_ray_closest_kernel"))) return rc;
    CU_TRY(cudaMemcpyAsync(out_value, ctx->bp1.ptr, (size_t)nq * 4, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(value)");
    CU_TRY(cudaMemcpyAsync(out_point_t, ctx->bp2.ptr, (size_t)nq * 16, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(point)");
    CU_TRY(cudaStreamSynchronize(ctx->stream), "bam3d_bvh_query_ray_closest");
    return BAM3D_OK;
}

// ---- device-resident queries: one pass, hits appended as (query, value, payload) in arbitrary order, no synchronisation ----
static int32_t bvh_query_dev_impl(bam3d_ctx* ctx, const bam3d_bvh* t, int kind, const float* d_queries, uint64_t nq64, uint32_t* d_out_query,
                                  uint32_t* d_out_values, void* d_out_payload, uint64_t capacity, uint64_t* d_count, const char* what) {
    if (!ctx || !t || !d_count || (nq64 && !d_queries) || (capacity && (!d_out_query || !d_out_values))) return b3_fail(ctx, BAM3D_ERR_ARG, what, cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, nq64)) || (rc = bvh_kind_check(ctx, t, kind))) return rc;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    const uint32_t nq = (uint32_t)nq64;
    // few large queries -> every (query, value) test; many small ones -> tree traversal (as the host-buffer forms decide)
    bool brute = ctx->opt_bvh_strategy == 2 || (ctx->opt_bvh_strategy == 0 && nq < 2048 && (uint64_t)nq * t->n <= (1ull << 35));
    if (t->n == 0 || t->d_spheres) brute = false;
    { KernelTimer kt(ctx);
      if (brute) ctx->launches += launch_bvh_brute_append(ctx->stream, t->d_aabbs, t->n, kind, d_queries, nq, d_out_query, d_out_values, d_out_payload, capacity,
                                                          (unsigned long long*)d_count);
      else ctx->launches += launch_bvh_query_append(ctx->stream, bvh_view(t), kind, d_queries, nq, d_out_query, d_out_values, d_out_payload, capacity,
                                                    (unsigned long long*)d_count); }
    return check_launch(ctx, "bvh query (append)");
}
int32_t bam3d_bvh_query_aabb_dev(bam3d_ctx* ctx, const bam3d_bvh* t, const float* d_queries, uint64_t nq, uint32_t* d_out_query, uint32_t* d_out_values,
                                 uint64_t capacity, uint64_t* d_count) {
    return bvh_query_dev_impl(ctx, t, BQ_AABB, d_queries, nq, d_out_query, d_out_values, nullptr, capacity, d_count, "bam3d_bvh_query_aabb_dev: null argument");
}
int32_t bam3d_bvh_query_sphere_dev(bam3d_ctx* ctx, const bam3d_bvh* t, const float* d_spheres, uint64_t nq, uint32_t* d_out_query, uint32_t* d_out_values,
                                   uint64_t capacity, uint64_t* d_count) {
    return bvh_query_dev_impl(ctx, t, BQ_SPHERE, d_spheres, nq, d_out_query, d_out_values, nullptr, capacity, d_count, "bam3d_bvh_query_sphere_dev: null argument");
}
int32_t bam3d_bvh_query_ray_dev(bam3d_ctx* ctx, const bam3d_bvh* t, const float* d_rays, uint64_t nq, uint32_t* d_out_query, uint32_t* d_out_values,
                                float* d_out_points, uint64_t capacity, uint64_t* d_count) {
    if (capacity && !d_out_points) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_query_ray_dev: null argument", cudaSuccess);
    return bvh_query_dev_impl(ctx, t, BQ_RAY, d_rays, nq, d_out_query, d_out_values, d_out_points, capacity, d_count, "bam3d_bvh_query_ray_dev: null argument");
}
int32_t bam3d_bvh_query_frustum_dev(bam3d_ctx* ctx, const bam3d_bvh* t, const float* d_planes, uint64_t nq, uint32_t* d_out_query, uint32_t* d_out_values,
                                    uint8_t* d_out_relation, uint64_t capacity, uint64_t* d_count) {
    if (capacity && !d_out_relation) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_query_frustum_dev: null argument", cudaSuccess);
    return bvh_query_dev_impl(ctx, t, BQ_FRUSTUM, d_planes, nq, d_out_query, d_out_values, d_out_relation, capacity, d_count, "bam3d_bvh_query_frustum_dev: null argument");
}
int32_t bam3d_bvh_query_ray_closest_dev(bam3d_ctx* ctx, const bam3d_bvh* t, const float* d_rays, uint64_t nq64, uint32_t* d_out_value, float* d_out_point_t) {
    if (!ctx || !t || (nq64 && (!d_rays || !d_out_value || !d_out_point_t))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_query_ray_closest_dev: null argument", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, nq64))) return rc;
    if (nq64 == 0) return BAM3D_OK;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    { KernelTimer kt(ctx);
      ctx->launches += launch_bvh_ray_closest(ctx->stream, bvh_view(t), d_rays, (uint32_t)nq64, d_out_value, d_out_point_t); }
    return check_launch(ctx, "bvh_ray_closest_kernel");
}

int32_t bam3d_bvh_find_collider_pairs(bam3d_ctx* ctx, const bam3d_bvh* t, const uint8_t* dirty, uint32_t* out_pairs, uint64_t capacity, uint64_t* out_count) {
    if (!ctx || !t || !out_count || (t->n && !dirty) || (capacity && !out_pairs)) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_bvh_find_collider_pairs: null argument", cudaSuccess);
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    *out_count = 0;
    const uint32_t n = t->n;
    if (n == 0) return BAM3D_OK;
    int32_t rc;
    const size_t scan_bytes = scan_tmp_bytes(n);
    if ((rc = ctx->bp0.reserve(ctx, n)) || (rc = ctx->bp1.reserve(ctx, ((size_t)n + 1) * 4)) || (rc = ctx->bp2.reserve(ctx, ((size_t)n + 1) * 8)) ||
        (rc = ctx->cub_tmp.reserve(ctx, scan_bytes))) return rc;
    uint8_t* d_dirty = (uint8_t*)ctx->bp0.ptr;
    uint32_t* d_counts = (uint32_t*)ctx->bp1.ptr;
    unsigned long long* d_offsets = (unsigned long long*)ctx->bp2.ptr;
    CU_TRY(cudaMemcpyAsync(d_dirty, dirty, n, cudaMemcpyHostToDevice, ctx->stream), "cudaMemcpyAsync(dirty)");
    cudaMemsetAsync(d_counts + n, 0, 4, ctx->stream);
    BvhDev T = bvh_view(t);
    { KernelTimer kt(ctx);
      ctx->launches += launch_bvh_collider_pairs(ctx->stream, T, t->d_aabbs, d_dirty, 0, 0, d_counts, nullptr, nullptr, 0); }
    ctx->launches += launch_exclusive_scan_u32_to_u64(ctx->stream, d_counts, d_offsets, n, ctx->cub_tmp.ptr, scan_bytes);
    if ((rc = check_launch(ctx, "bvh collider pairs (count)"))) return rc;
    unsigned long long total = 0;
    CU_TRY(cudaMemcpyAsync(&total, d_offsets + n, 8, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(total)");
    CU_TRY(cudaStreamSynchronize(ctx->stream), "bvh collider pairs (count)");
    *out_count = total;
    if (total > capacity) return b3_fail(ctx, BAM3D_ERR_CAPACITY, "bam3d_bvh_find_collider_pairs: output capacity too small", cudaSuccess);
    if (total == 0) return BAM3D_OK;
    const size_t sort_bytes = sort_u64_tmp_bytes(total);
    if ((rc = ctx->bp3.reserve(ctx, total * 8 * 2)) || (rc = ctx->bp4.reserve(ctx, total * 8)) || (rc = ctx->cub_tmp.reserve(ctx, sort_bytes > scan_bytes ? sort_bytes : scan_bytes))) return rc;
    unsigned long long* keys = (unsigned long long*)ctx->bp3.ptr;
    { KernelTimer kt(ctx);
      ctx->launches += launch_bvh_collider_pairs(ctx->stream, T, t->d_aabbs, d_dirty, 0, 0, d_counts, d_offsets, keys, 0); }
    bool in_alt = false;
    ctx->launches += launch_sort_u64(ctx->stream, keys, keys + total, total, ctx->cub_tmp.ptr, sort_bytes, &in_alt);
    ctx->launches += launch_unpack_pairs(ctx->stream, in_alt ? keys + total : keys, total, (uint2*)ctx->bp4.ptr, true);
    if ((rc = check_launch(ctx, "bvh collider pairs (fill)"))) return rc;
    CU_TRY(cudaMemcpyAsync(out_pairs, ctx->bp4.ptr, total * 8, cudaMemcpyDeviceToHost, ctx->stream), "cudaMemcpyAsync(pairs)");
    CU_TRY(cudaStreamSynchronize(ctx->stream), "bam3d_bvh_find_collider_pairs");
    return BAM3D_OK;
}


// ------------------------------------------------------------------------------------------------------------
// frame queue: set_transforms + narrow phase + results, `depth` frames in flight
// ------------------------------------------------------------------------------------------------------------
// Streams: host->device copies and the pack kernel on `s_in`, device->host copies on `s_out`, and the narrow-phase kernels
// on TWO compute lanes (depth >= 2): lane = frame % 2, each lane a private sub-context with its own stream and its own
// scratch (bins, GJK->EPA queue, polytope slab, overflow list). Frame k+1's upload and frame k-1's download run beside
// frame k's kernels, and frame k+1's kernels start while frame k's persistent EPA kernel drains — its last ~20 % runs
// with most lanes idle (no queued pairs left), which one frame alone cannot fill. A slot is reused once its download is done.
struct bam3d_frames {
    struct Slot {
        bam3d_shapes* shapes = nullptr;
        bam3d_shapes* shapes_end = nullptr;  // time of impact only
        uint32_t* d_pairs = nullptr;
        uint8_t* d_status = nullptr;
        bam3d_contact* d_contacts = nullptr;
        uint32_t* d_err = nullptr;   // pack_shapes_kernel's flags for this frame
        uint32_t* h_err = nullptr;   // pinned copy, read by bam3d_frames_wait
        cudaEvent_t uploaded = nullptr, computed = nullptr, downloaded = nullptr;
        uint64_t frame = ~0ull;      // frame number in flight / last completed in this slot
        bool in_flight = false;
        int32_t launch_rc = BAM3D_OK;
        // gather mode (bam3d_frames_submit_gather): hits compacted on the device, exchanged over NCCL, and the list of ALL
        // ranks' hits copied to the host instead of this rank's dense contacts
        bam3d_contact40* d_compact = nullptr;
        bam3d_contact40* d_gathered = nullptr;
        uint64_t* d_count = nullptr;
        uint64_t gathered_capacity = 0;
        bam3d_comm* comm = nullptr;   // non-null while a gather is pending for this slot
        bam3d_ctx* lane = nullptr;
        uint64_t ticket = 0, total = 0, out_capacity = 0;
        bam3d_contact40* out_gathered = nullptr;
        std::vector<uint64_t> counts;
        cudaEvent_t exchanged = nullptr;
    };
    std::vector<Slot> slots;
    uint64_t max_pairs = 0, next_frame = 0;
    int xform_layout = BAM3D_XFORM_MAT4;  // what submit's xform / xform_end hold
    cudaStream_t s_in = nullptr, s_out = nullptr;
    bam3d_ctx* lanes[2] = {nullptr, nullptr};  // compute lanes (null: the parent ctx computes)
};

// the compute lane of a frame, with the parent's switches
static bam3d_ctx* frames_lane(bam3d_ctx* ctx, bam3d_frames* F, uint64_t frame) {
    bam3d_ctx* c = F->lanes[frame & 1];
    if (!c) return ctx;
    c->opt_binning = ctx->opt_binning; c->opt_epa_tiers = ctx->opt_epa_tiers; c->opt_intersect_split = ctx->opt_intersect_split;
    c->opt_intersect_refill = ctx->opt_intersect_refill; c->opt_toi_refill = ctx->opt_toi_refill;
    c->opt_epa_global_polytopes = ctx->opt_epa_global_polytopes; c->opt_epa_overlap = ctx->opt_epa_overlap;
    c->opt_distance_refill = ctx->opt_distance_refill; c->opt_epa_pipeline = ctx->opt_epa_pipeline; c->opt_epa_huge = ctx->opt_epa_huge;
    c->opt_timing = false;
    return c;
}

static int32_t frames_finish_slot(bam3d_ctx* ctx, bam3d_frames* F, bam3d_frames::Slot& sl) {
    if (!sl.in_flight) return BAM3D_OK;
    if (sl.comm) {
        // second half of the exchange: the counts are on the host by now (the GPU is busy with the next frame); the payload
        // goes over the comm stream, then to the host on the copy-out stream
        bam3d_comm* comm = sl.comm;
        sl.comm = nullptr;
        int32_t rc = bam3d_gather_contacts_finish(sl.lane, comm, sl.ticket, sl.d_compact, sl.d_gathered,
                                                  sl.out_capacity < sl.gathered_capacity ? sl.out_capacity : sl.gathered_capacity,
                                                  sl.counts.data(), &sl.total);
        if (rc) { if (sl.lane != ctx) ctx->err = sl.lane->err; cudaEventSynchronize(sl.downloaded); sl.in_flight = false; return rc; }
        cudaStream_t cs = (cudaStream_t)bam3d_comm_stream(comm);
        cudaEventRecord(sl.exchanged, cs);
        cudaStreamWaitEvent(F->s_out, sl.exchanged, 0);
        if (sl.total) cudaMemcpyAsync(sl.out_gathered, sl.d_gathered, sl.total * sizeof(bam3d_contact40), cudaMemcpyDeviceToHost, F->s_out);
        cudaEventRecord(sl.downloaded, F->s_out);
    }
    cudaError_t e = cudaEventSynchronize(sl.downloaded);
    sl.in_flight = false;
    if (e != cudaSuccess) return b3_fail(ctx, BAM3D_ERR_CUDA, "bam3d_frames_wait", e);
    if (sl.launch_rc) return sl.launch_rc;
    if (*sl.h_err & 1u) return b3_fail(ctx, BAM3D_ERR_NON_AFFINE, "a transform's 4th row is not (0,0,0,1)", cudaSuccess);
    if (*sl.h_err) return b3_fail(ctx, BAM3D_ERR_ARG, "pack_shapes_kernel rejected the frame's shapes", cudaSuccess);
    if (sl.h_err[1] & 1u) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_frames: a pair of this frame names a shape index >= the shape set's size", cudaSuccess);
    return BAM3D_OK;
}

void bam3d_frames_free(bam3d_ctx* ctx, bam3d_frames* F) {
    if (!F) return;
    if (ctx) cudaSetDevice(ctx->device);
    for (auto& sl : F->slots) {
        if (sl.in_flight && sl.downloaded) cudaEventSynchronize(sl.downloaded);
        bam3d_shapes_free(ctx, sl.shapes);
        bam3d_shapes_free(ctx, sl.shapes_end);
        if (sl.d_pairs) cudaFree(sl.d_pairs);
        if (sl.d_status) cudaFree(sl.d_status);
        if (sl.d_contacts) cudaFree(sl.d_contacts);
        if (sl.d_err) cudaFree(sl.d_err);
        if (sl.d_compact) cudaFree(sl.d_compact);
        if (sl.d_gathered) cudaFree(sl.d_gathered);
        if (sl.d_count) cudaFree(sl.d_count);
        if (sl.exchanged) cudaEventDestroy(sl.exchanged);
        if (sl.h_err) cudaFreeHost(sl.h_err);
        if (sl.uploaded) cudaEventDestroy(sl.uploaded);
        if (sl.computed) cudaEventDestroy(sl.computed);
        if (sl.downloaded) cudaEventDestroy(sl.downloaded);
    }
    if (F->s_in) { cudaStreamSynchronize(F->s_in); cudaStreamDestroy(F->s_in); }
    if (F->s_out) { cudaStreamSynchronize(F->s_out); cudaStreamDestroy(F->s_out); }
    for (bam3d_ctx* c : F->lanes) if (c) bam3d_ctx_destroy(c);
    delete F;
}

int32_t bam3d_frames_create(bam3d_ctx* ctx, const uint8_t* kind, const float* params, const float* xform, const uint32_t* mesh_id,
                            uint32_t n_shapes, const bam3d_meshes* meshes, uint64_t max_pairs, uint32_t depth, int32_t with_end,
                            bam3d_frames** out) {
    if (!ctx || !out) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_frames_create: null argument", cudaSuccess);
    *out = nullptr;
    if (depth < 1 || depth > 8) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_frames_create: depth must be 1..8", cudaSuccess);
    int32_t rc;
    if ((rc = check_n(ctx, max_pairs))) return rc;
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    bam3d_frames* F = new (std::nothrow) bam3d_frames();
    if (!F) return BAM3D_ERR_OOM;
    F->max_pairs = max_pairs;
    F->slots.resize(depth);
    cudaError_t e = cudaSuccess;
    if ((e = cudaStreamCreateWithFlags(&F->s_in, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&F->s_out, cudaStreamNonBlocking)) != cudaSuccess) {
        bam3d_frames_free(ctx, F);
        return b3_fail(ctx, BAM3D_ERR_CUDA, "cudaStreamCreate(frames)", e);
    }
    if (depth >= 2 && !std::getenv("BAM3D_FRAMES_ONE_LANE")) {  // (development knob: the single-lane behaviour for A/B timing)
        for (auto& lane : F->lanes)
            if ((rc = bam3d_ctx_create(ctx->device, &lane))) {
                bam3d_frames_free(ctx, F);
                return b3_fail(ctx, rc, "bam3d_frames_create: compute lane", cudaSuccess);
            }
    }
    const size_t np = max_pairs ? max_pairs : 1;
    for (auto& sl : F->slots) {
        // the synchronous upload validates kinds, mesh ids and the initial transforms once; a frame can then only fail on
        // a non-affine transform, which is reported by bam3d_frames_wait for that frame
        if ((rc = bam3d_shapes_upload(ctx, kind, params, xform, mesh_id, n_shapes, meshes, &sl.shapes)) ||
            (with_end && (rc = bam3d_shapes_upload(ctx, kind, params, xform, mesh_id, n_shapes, meshes, &sl.shapes_end)))) {
            bam3d_frames_free(ctx, F);
            return rc;
        }
        if ((e = cudaMalloc(&sl.d_pairs, np * 8)) != cudaSuccess || (e = cudaMalloc(&sl.d_status, np)) != cudaSuccess ||
            (e = cudaMalloc(&sl.d_contacts, np * 32)) != cudaSuccess || (e = cudaMalloc(&sl.d_err, 8)) != cudaSuccess ||
            (e = cudaMallocHost(&sl.h_err, 8)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&sl.uploaded, cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&sl.computed, cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&sl.downloaded, cudaEventDisableTiming)) != cudaSuccess) {
            bam3d_frames_free(ctx, F);
            return b3_fail(ctx, BAM3D_ERR_OOM, "bam3d_frames_create: device / pinned allocation", e);
        }
        sl.h_err[0] = sl.h_err[1] = 0;
    }
    *out = F;
    return BAM3D_OK;
}

}  // extern "C"

struct FrameGather { bam3d_comm* comm; uint64_t base; bam3d_contact40* out; uint64_t capacity; };

static int32_t frames_submit_impl(bam3d_ctx* ctx, bam3d_frames* F, int32_t query, const float* xform, const float* xform_end,
                                  const uint32_t* pairs, uint64_t n, const bam3d_gjk_settings* settings, int32_t strategy,
                                  bam3d_contact* out_contacts, uint8_t* out_status, uint64_t* out_frame, const FrameGather* G) {
    if (!ctx || !F || !xform || (n && (!pairs || !out_status))) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_frames_submit: null argument", cudaSuccess);
    if (query < BAM3D_FRAME_INTERSECT || query > BAM3D_FRAME_TIME_OF_IMPACT) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_frames_submit: unknown query", cudaSuccess);
    if (!G && query != BAM3D_FRAME_INTERSECT && n && !out_contacts) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_frames_submit: out_contacts is null", cudaSuccess);
    if (G && (query == BAM3D_FRAME_INTERSECT || !G->comm || (G->capacity && !G->out)))
        return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_frames_submit_gather: needs a contact-producing query, a communicator and an output buffer", cudaSuccess);
    if (n > F->max_pairs) return b3_fail(ctx, BAM3D_ERR_CAPACITY, "bam3d_frames_submit: more pairs than max_pairs", cudaSuccess);
    bam3d_frames::Slot& sl = F->slots[F->next_frame % F->slots.size()];
    if (query == BAM3D_FRAME_TIME_OF_IMPACT && (!xform_end || !sl.shapes_end))
        return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_frames_submit: time of impact needs xform_end and a queue created with with_end", cudaSuccess);
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    // back-pressure: the frame `depth` submissions ago must have left this slot (its status is what bam3d_frames_wait on
    // it returns; a caller that skipped the wait loses that frame's error code, not its data)
    if (sl.in_flight) (void)frames_finish_slot(ctx, F, sl);
    if (G && !sl.d_compact) {  // first gather through this slot: its exchange buffers
        const size_t np = F->max_pairs ? F->max_pairs : 1;
        const size_t cap = (size_t)np * (size_t)bam3d_comm_size(G->comm);
        cudaError_t ea;
        if ((ea = cudaMalloc(&sl.d_compact, np * sizeof(bam3d_contact40))) != cudaSuccess || (ea = cudaMalloc(&sl.d_gathered, cap * sizeof(bam3d_contact40))) != cudaSuccess ||
            (ea = cudaMalloc(&sl.d_count, 8)) != cudaSuccess || (ea = cudaEventCreateWithFlags(&sl.exchanged, cudaEventDisableTiming)) != cudaSuccess)
            return b3_fail(ctx, BAM3D_ERR_OOM, "bam3d_frames_submit_gather: exchange buffers", ea);
        sl.gathered_capacity = cap;
        sl.counts.assign((size_t)bam3d_comm_size(G->comm), 0);
    }
    const uint32_t ns = sl.shapes->n;
    const uint32_t n_meshes = sl.shapes->meshes ? sl.shapes->meshes->n_meshes : 0;
    sl.frame = F->next_frame++;
    sl.launch_rc = BAM3D_OK;
    if (out_frame) *out_frame = sl.frame;
    // upload + pack on the copy-in stream
    CU_TRY(cudaMemsetAsync(sl.d_err, 0, 8, F->s_in), "cudaMemsetAsync(err)");  // [0] pack_shapes_kernel's flags, [1] pair-index flag
    sl.shapes->err_flag = sl.d_err + 1;
    if (sl.shapes_end) sl.shapes_end->err_flag = sl.d_err + 1;
    if (ns) {
        const bool affine48 = F->xform_layout == BAM3D_XFORM_AFFINE3X4;
        const size_t xbytes = (size_t)ns * (affine48 ? 48 : 64);
        CU_TRY(cudaMemcpyAsync(sl.shapes->d_xform, xform, xbytes, cudaMemcpyHostToDevice, F->s_in), "cudaMemcpyAsync(xform)");
        ctx->launches += launch_pack_shapes(F->s_in, sl.shapes->d_kind, sl.shapes->d_params, sl.shapes->d_xform, sl.shapes->d_mesh_id, ns,
                                            n_meshes, sl.shapes->recs, sl.shapes->meta, sl.d_err, affine48);
        if (query == BAM3D_FRAME_TIME_OF_IMPACT) {
            CU_TRY(cudaMemcpyAsync(sl.shapes_end->d_xform, xform_end, xbytes, cudaMemcpyHostToDevice, F->s_in), "cudaMemcpyAsync(xform_end)");
            ctx->launches += launch_pack_shapes(F->s_in, sl.shapes_end->d_kind, sl.shapes_end->d_params, sl.shapes_end->d_xform,
                                                sl.shapes_end->d_mesh_id, ns, n_meshes, sl.shapes_end->recs, sl.shapes_end->meta, sl.d_err, affine48);
        }
    }
    if (n) CU_TRY(cudaMemcpyAsync(sl.d_pairs, pairs, n * 8, cudaMemcpyHostToDevice, F->s_in), "cudaMemcpyAsync(pairs)");
    CU_TRY(cudaEventRecord(sl.uploaded, F->s_in), "cudaEventRecord");
    // kernels on the frame's compute lane
    bam3d_ctx* cc = frames_lane(ctx, F, sl.frame);
    CU_TRY(cudaStreamWaitEvent(cc->stream, sl.uploaded, 0), "cudaStreamWaitEvent");
    int32_t rc = BAM3D_OK;
    const uint64_t launches0 = cc->launches;
    if (n) {
        if (query == BAM3D_FRAME_INTERSECT) rc = bam3d_gjk_intersect_dev(cc, sl.shapes, sl.d_pairs, n, settings, sl.d_status);
        else if (query == BAM3D_FRAME_CONTACT) rc = bam3d_gjk_contact_dev(cc, sl.shapes, sl.d_pairs, n, settings, strategy, sl.d_contacts, sl.d_status);
        else rc = bam3d_gjk_time_of_impact_dev(cc, sl.shapes, sl.shapes_end, sl.d_pairs, n, settings, sl.d_contacts, sl.d_status);
    }
    if (!rc && G) {  // hits only, tagged with the global pair index; then the first half of the exchange (record counts)
        rc = bam3d_compact_contacts_dev(cc, sl.d_contacts, sl.d_status, n, G->base, sl.d_compact, n, sl.d_count);
        if (!rc) rc = bam3d_gather_contacts_begin(cc, G->comm, sl.d_count, &sl.ticket);
    }
    if (cc != ctx) { ctx->launches += cc->launches - launches0; if (rc) ctx->err = cc->err; }
    CU_TRY(cudaEventRecord(sl.computed, cc->stream), "cudaEventRecord");
    if (rc) {  // nothing of this frame is copied back; the slot is free again once the upload has drained
        cudaEventSynchronize(sl.computed);
        return rc;
    }
    // results on the copy-out stream
    CU_TRY(cudaStreamWaitEvent(F->s_out, sl.computed, 0), "cudaStreamWaitEvent");
    if (G) { sl.comm = G->comm; sl.lane = cc; sl.out_gathered = G->out; sl.out_capacity = G->capacity; sl.total = 0; }
    if (n && query != BAM3D_FRAME_INTERSECT && out_contacts) CU_TRY(cudaMemcpyAsync(out_contacts, sl.d_contacts, n * 32, cudaMemcpyDeviceToHost, F->s_out), "cudaMemcpyAsync(contacts)");
    if (n) CU_TRY(cudaMemcpyAsync(out_status, sl.d_status, n, cudaMemcpyDeviceToHost, F->s_out), "cudaMemcpyAsync(status)");
    CU_TRY(cudaMemcpyAsync(sl.h_err, sl.d_err, 8, cudaMemcpyDeviceToHost, F->s_out), "cudaMemcpyAsync(err)");
    CU_TRY(cudaEventRecord(sl.downloaded, F->s_out), "cudaEventRecord");
    sl.in_flight = true;
    return BAM3D_OK;
}

extern "C" {

int32_t bam3d_frames_set_transform_layout(bam3d_ctx* ctx, bam3d_frames* F, int32_t layout) {
    if (!ctx || !F || (layout != BAM3D_XFORM_MAT4 && layout != BAM3D_XFORM_AFFINE3X4))
        return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_frames_set_transform_layout: unknown layout", cudaSuccess);
    F->xform_layout = layout;
    return BAM3D_OK;
}

int32_t bam3d_frames_submit(bam3d_ctx* ctx, bam3d_frames* F, int32_t query, const float* xform, const float* xform_end,
                            const uint32_t* pairs, uint64_t n, const bam3d_gjk_settings* settings, int32_t strategy,
                            bam3d_contact* out_contacts, uint8_t* out_status, uint64_t* out_frame) {
    return frames_submit_impl(ctx, F, query, xform, xform_end, pairs, n, settings, strategy, out_contacts, out_status, out_frame, nullptr);
}

int32_t bam3d_frames_submit_gather(bam3d_ctx* ctx, bam3d_frames* F, bam3d_comm* comm, int32_t query, const float* xform, const float* xform_end,
                                   const uint32_t* pairs, uint64_t n, const bam3d_gjk_settings* settings, int32_t strategy,
                                   uint64_t pair_index_base, bam3d_contact40* out_gathered, uint64_t capacity, uint8_t* out_status,
                                   uint64_t* out_frame) {
    FrameGather G{comm, pair_index_base, out_gathered, capacity};
    return frames_submit_impl(ctx, F, query, xform, xform_end, pairs, n, settings, strategy, nullptr, out_status, out_frame, &G);
}

int32_t bam3d_frames_wait(bam3d_ctx* ctx, bam3d_frames* F, uint64_t frame) {
    if (!ctx || !F) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_frames_wait: null argument", cudaSuccess);
    if (frame >= F->next_frame) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_frames_wait: frame was never submitted", cudaSuccess);
    bam3d_frames::Slot& sl = F->slots[frame % F->slots.size()];
    if (sl.frame != frame) return BAM3D_OK;  // an older frame: its slot was reused, so it completed long ago
    CU_TRY(cudaSetDevice(ctx->device), "cudaSetDevice");
    return frames_finish_slot(ctx, F, sl);
}

int32_t bam3d_frames_gathered(bam3d_ctx* ctx, bam3d_frames* F, uint64_t frame, uint64_t* out_counts, uint64_t* out_total) {
    if (!ctx || !F || !out_total) return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_frames_gathered: null argument", cudaSuccess);
    bam3d_frames::Slot& sl = F->slots[frame % F->slots.size()];
    if (frame >= F->next_frame || sl.frame != frame || sl.in_flight || sl.counts.empty())
        return b3_fail(ctx, BAM3D_ERR_ARG, "bam3d_frames_gathered: call it after bam3d_frames_wait for that frame, before its slot is reused", cudaSuccess);
    if (out_counts) for (size_t r = 0; r < sl.counts.size(); ++r) out_counts[r] = sl.counts[r];
    *out_total = sl.total;
    return BAM3D_OK;
}

}  // extern "C"
