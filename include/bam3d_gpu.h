/* bam3d_gpu.h — C ABI of the B200-native batched collision pipeline that sits behind bam3d's public API.
 *
 * The reference crate (DallasC/bam3d) is a pure-Rust library with no FFI of its own; this ABI is what a
 * `bam3d` build with a `gpu` feature would bind through `extern "C"` (INTEGRATION.md shows the Rust block and
 * the build.rs that compiles the CUDA sources with nvcc for sm_100a). Each entry point names the reference
 * interface it stands in for (paths relative to the reference's src/).
 *
 * Conventions
 *  - every function returns int32_t: 0 = BAM3D_OK, negative = error (bam3d_last_error gives the text);
 *  - the caller owns every host buffer; device memory is owned by the ctx / opaque handles;
 *  - per-item outcomes (hit, miss, not converged, "the reference would have panicked") are returned in
 *    `out_status[]`, never as a function error;  hit  <=>  status == BAM3D_STATUS_OK;
 *  - one ctx per host thread (a ctx is not thread-safe); any number of ctxs may exist;
 *  - host-buffer entry points are synchronous on return; `_dev` entry points take device pointers, enqueue on
 *    the ctx stream (bam3d_ctx_stream) and return without synchronising;
 *  - matrices are glam `Mat4` column-major arrays of 16 f32 and must be affine (4th row exactly 0,0,0,1);
 *  - there is no CPU fallback: without a CUDA device every call fails with BAM3D_ERR_CUDA.
 */
#ifndef BAM3D_GPU_H
#define BAM3D_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BAM3D_OK 0
#define BAM3D_ERR_CUDA (-1)        /* a CUDA runtime call failed */
#define BAM3D_ERR_ARG (-2)         /* bad argument (null pointer, index out of range, unsupported setting) */
#define BAM3D_ERR_NON_AFFINE (-3)  /* a transform's 4th row is not (0,0,0,1) */
#define BAM3D_ERR_CAPACITY (-4)    /* output capacity too small; the required count is reported */
#define BAM3D_ERR_OOM (-5)
#define BAM3D_ERR_COMM (-6)        /* multi-GPU exchange: libnccl could not be loaded, or an NCCL call failed */

/* primitive kinds, in the order of `enum Primitive3` (primitive/primitive3.rs:16-33); Cube == Cuboid */
#define BAM3D_KIND_PARTICLE 0   /* primitive/particle.rs   params: -                                  */
#define BAM3D_KIND_QUAD 1       /* primitive/quad.rs       params: half_dim.x, half_dim.y              */
#define BAM3D_KIND_SPHERE 2     /* primitive/sphere.rs     params: radius                              */
#define BAM3D_KIND_CUBOID 3     /* primitive/cuboid.rs     params: half_dim.x, half_dim.y, half_dim.z  */
#define BAM3D_KIND_CYLINDER 4   /* primitive/cylinder.rs   params: half_height, radius                 */
#define BAM3D_KIND_CAPSULE 5    /* primitive/capsule.rs    params: half_height, radius                 */
#define BAM3D_KIND_POLYHEDRON 6 /* primitive/polyhedron.rs ConvexPolyhedron::new (VertexOnly); mesh_id  */

/* per-item status */
#define BAM3D_STATUS_OK 0         /* Some(..): hit / converged                                          */
#define BAM3D_STATUS_MISS 1       /* None: no intersection (distance: the shapes collide)               */
#define BAM3D_STATUS_MAX_ITER 2   /* None: max_iterations reached (TOI: the added iteration cap)        */
#define BAM3D_STATUS_REF_PANIC 3  /* the reference would panic here (simplex3d.rs:134 unwrap, :138 assert!) */
#define BAM3D_STATUS_NAN 4        /* None through a NaN comparison                                      */
#define BAM3D_STATUS_CAPACITY 5   /* EPA polytope outgrew 2^18 faces (never seen). The reference has no face limit: on near-degenerate
                                   * input (a Cylinder cap against a flat face) its polytope pinches and grows multiplicatively — 44 of 2^20
                                   * mixed pairs pass 1024 faces, the largest reaches 121 962 — and it still returns a Contact when its
                                   * iteration cap ends the loop; three EPA tiers (thread / warp / block per pair) return that Contact */
#define BAM3D_STATUS_UNSUPPORTED 6 /* ray cast against a ConvexPolyhedron uploaded without faces (or > 4096 faces)   */

/* CollisionStrategy (contact.rs:12-18) */
#define BAM3D_FULL_RESOLUTION 0
#define BAM3D_COLLISION_ONLY 1

typedef struct bam3d_ctx bam3d_ctx;
typedef struct bam3d_meshes bam3d_meshes; /* convex-polyhedron vertex library resident on the device */
typedef struct bam3d_shapes bam3d_shapes; /* packed primitives + transforms resident on the device   */
typedef struct bam3d_bvh bam3d_bvh;       /* static BVH answering the DBVT queries                     */

/* GJK::new_with_settings (algorithm/minkowski/gjk/mod.rs:45-58). max_iterations is shared by GJK and EPA, as in
 * the reference, and must be <= 100 for contact generation (device polytope capacity). toi_iteration_cap bounds
 * the reference's uncapped time-of-impact loop (gjk/mod.rs:166); 0 = default (100). */
typedef struct bam3d_gjk_settings {
    float distance_tolerance;   /* GJK_DISTANCE_TOLERANCE   = 1e-6 (gjk/mod.rs:19) */
    float continuous_tolerance; /* GJK_CONTINUOUS_TOLERANCE = 1e-6 (gjk/mod.rs:20) */
    float epa_tolerance;        /* EPA_TOLERANCE            = 1e-5 (epa/mod.rs:12) */
    uint32_t max_iterations;    /* MAX_ITERATIONS           = 100  (gjk/mod.rs:18, epa/mod.rs:13) */
    uint32_t toi_iteration_cap;
} bam3d_gjk_settings;

/* Contact (contact.rs:22-38) without the strategy tag: 32 bytes, one per pair. Zero for non-hits. */
typedef struct bam3d_contact {
    float normal[3];
    float penetration_depth;
    float contact_point[3];
    float time_of_impact;
} bam3d_contact;

/* Compacted contact for the multi-GPU gather: only hits, tagged with the pair's index in the batch. */
typedef struct bam3d_contact40 {
    uint32_t pair_index;
    uint32_t status;
    bam3d_contact contact;
} bam3d_contact40;

/* ---- context ------------------------------------------------------------------------------------------ */
int32_t bam3d_ctx_create(int32_t device_ordinal, bam3d_ctx** out_ctx);
void bam3d_ctx_destroy(bam3d_ctx* ctx);
const char* bam3d_last_error(const bam3d_ctx* ctx);
int32_t bam3d_ctx_synchronize(bam3d_ctx* ctx);
void* bam3d_ctx_stream(bam3d_ctx* ctx); /* the ctx's cudaStream_t */
void bam3d_default_settings(bam3d_gjk_settings* out); /* GJK::new() (gjk/mod.rs:34-42) */
/* switches: "binning" (default 1) sorts pairs by (kindL, kindR) before the thread-per-pair kernels; "timing"
 * (default 0) brackets the main kernel(s) of every narrow-phase / broad-phase call with CUDA events; "sap_strategy"
 * picks how sweep-and-prune finds its pairs — 0 auto, 1 tiled forward sweep, 2 BVH overlap query (same output);
 * "bvh_strategy" picks how the variable-length BVH queries run — 0 auto, 1 tree traversal (thread per query), 2 every
 * (query, value) test (for few large queries); "epa_overlap" (default 1) runs the EPA overflow tiers on a side stream next
 * to the EPA kernel; "epa_huge" (default 1) enables the third EPA tier (0: pairs beyond 1024 faces keep
 * BAM3D_STATUS_CAPACITY, as in the first release); "epa_global_polytopes" (default 1) keeps the thread-per-pair EPA's
 * polytopes in a global-memory slab, one contiguous record per thread (582 MB of scratch), instead of per-thread local
 * memory; "distance_refill" (default 8; 0..32) and "toi_refill" (default 8; 0..32) run bam3d_gjk_distance's /
 * bam3d_gjk_time_of_impact's thread path as a persistent kernel in which every lane owns one pair and idle lanes fetch new
 * pairs together once that many lanes of a warp are idle (0 = one pair per thread: same results, less than half the speed on
 * mixed scenes). Every switch changes scheduling only, never a result bit. "epa_tiers", "epa_pipeline", "intersect_split" and
 * "intersect_refill" select kernels that were measured slower than the defaults; they exist only in a library built with
 * -DB3_EXPERIMENTS and are refused (BAM3D_ERR_ARG) otherwise. */
int32_t bam3d_ctx_set_option(bam3d_ctx* ctx, const char* name, int32_t value);
/* synchronise, then return and reset the summed device time (ms) and number of bracketed calls ("timing" = 1) */
int32_t bam3d_ctx_read_timing(bam3d_ctx* ctx, double* out_total_ms, uint64_t* out_count);
/* page-locked host memory so that the host-buffer entry points copy at full PCIe / C2C rate */
int32_t bam3d_host_alloc(bam3d_ctx* ctx, uint64_t bytes, void** out_ptr);
void bam3d_host_free(bam3d_ctx* ctx, void* ptr);

/* ---- host layer: pack primitives + Mat4 into device buffers ------------------------------------------- */
/* ConvexPolyhedron::new(vertices) for a library of meshes (primitive/polyhedron.rs:69-95). verts_xyz holds
 * offsets[n_meshes] vertices (3 f32 each); mesh m owns vertices offsets[m] .. offsets[m+1]. */
int32_t bam3d_meshes_upload(bam3d_ctx* ctx, const float* verts_xyz, const uint32_t* offsets, uint32_t n_meshes,
                            bam3d_meshes** out);
/* ConvexPolyhedron::new_with_faces (primitive/polyhedron.rs:98-115, build_half_edges :246-340): the same library, where
 * mesh m additionally has the triangles faces[3 * face_offsets[m]] .. faces[3 * face_offsets[m + 1]) (vertex indices local
 * to the mesh, counter-clockwise seen from outside). A mesh with faces is in HalfEdge mode: its support point is the
 * reference's hill climb over the half-edge structure (:153-183) and ray casts walk its faces (:373-522); a mesh without
 * faces stays VertexOnly (brute-force support, no ray casts). A degenerate face (the reference panics) or an index out of
 * range fails the call with BAM3D_ERR_ARG. */
int32_t bam3d_meshes_upload_faces(bam3d_ctx* ctx, const float* verts_xyz, const uint32_t* offsets, uint32_t n_meshes,
                                  const uint32_t* faces, const uint32_t* face_offsets, bam3d_meshes** out);
void bam3d_meshes_free(bam3d_ctx* ctx, bam3d_meshes* meshes);

/* Primitive3 values and their model-to-world Mat4 (primitive/primitive3.rs, traits.rs:78-100). kind[n],
 * params[n*4] (see BAM3D_KIND_*), xform[n*16] column-major; mesh_id[n] only read for polyhedra (may be NULL
 * when there are none); meshes may be NULL likewise. Computes Mat4::inverse() once per shape on the device
 * (the reference recomputes it inside every support_point call: primitive/util.rs:11, sphere.rs:22, ...). */
int32_t bam3d_shapes_upload(bam3d_ctx* ctx, const uint8_t* kind, const float* params, const float* xform,
                            const uint32_t* mesh_id, uint32_t n, const bam3d_meshes* meshes, bam3d_shapes** out);
/* replace every transform (next frame) without re-sending kinds / params */
int32_t bam3d_shapes_set_transforms(bam3d_ctx* ctx, bam3d_shapes* shapes, const float* xform);
/* The same from 12 floats per shape: the xyz of the x, y, z and w columns — the column-major Mat4 without its fourth row. A
 * model transform has (0, 0, 0, 1) there (the Mat4 form fails with BAM3D_ERR_NON_AFFINE otherwise), so nothing is lost and a
 * quarter less crosses PCIe per frame; the records built on the device are bit-identical. */
int32_t bam3d_shapes_set_transforms_affine(bam3d_ctx* ctx, bam3d_shapes* shapes, const float* xform12);
#define BAM3D_XFORM_MAT4 0       /* 16 floats per shape, column-major (glam Mat4) */
#define BAM3D_XFORM_AFFINE3X4 1  /* 12 floats per shape: the four columns' xyz */
void bam3d_shapes_free(bam3d_ctx* ctx, bam3d_shapes* shapes);
uint32_t bam3d_shapes_count(const bam3d_shapes* shapes);

/* ---- narrow phase, host buffers ------------------------------------------------------------------------ */
/* pairs: n x (left shape index, right shape index). */

/* GJK::intersect (gjk/mod.rs:74-113) per pair: status OK <=> Some(simplex). */
int32_t bam3d_gjk_intersect(bam3d_ctx* ctx, const bam3d_shapes* shapes, const uint32_t* pairs, uint64_t n,
                            const bam3d_gjk_settings* settings, uint8_t* out_status);
/* GJK::intersection (gjk/mod.rs:317-341) per pair = intersect + EPA3::process (epa/epa3d.rs:16-52) for
 * FullResolution, or Contact::new(CollisionOnly) (all-zero payload). */
int32_t bam3d_gjk_contact(bam3d_ctx* ctx, const bam3d_shapes* shapes, const uint32_t* pairs, uint64_t n,
                          const bam3d_gjk_settings* settings, int32_t strategy, bam3d_contact* out_contacts,
                          uint8_t* out_status);
/* GJK::distance (gjk/mod.rs:232-280) per pair. out_ref_value = what the reference returns for Some (dp - d0, the
 * convergence residual, :273-275); out_distance = sqrt(d.d), the geometric separation (:274, commented out in the
 * reference). Both 0 unless status == OK. */
int32_t bam3d_gjk_distance(bam3d_ctx* ctx, const bam3d_shapes* shapes, const uint32_t* pairs, uint64_t n,
                           const bam3d_gjk_settings* settings, float* out_ref_value, float* out_distance,
                           uint8_t* out_status);
/* GJK::intersection_time_of_impact (gjk/mod.rs:130-217) per pair; shapes_start / shapes_end hold the same
 * primitives with their transforms at the start / end of the step (Range<&Mat4>). */
int32_t bam3d_gjk_time_of_impact(bam3d_ctx* ctx, const bam3d_shapes* shapes_start, const bam3d_shapes* shapes_end,
                                 const uint32_t* pairs, uint64_t n, const bam3d_gjk_settings* settings,
                                 bam3d_contact* out_contacts, uint8_t* out_status);

/* ---- compound shapes --------------------------------------------------------------------------------------- */
/* The reference's `&[(Primitive, Mat4)]` slices (gjk/mod.rs:362-371): compound c owns the primitives comp_offsets[c] ..
 * comp_offsets[c + 1] of (kind, params, local_xform = local-to-model Mat4, mesh_id); comp_offsets has n_compounds + 1 entries. */
typedef struct bam3d_compounds bam3d_compounds;
int32_t bam3d_compounds_upload(bam3d_ctx* ctx, const uint8_t* kind, const float* params, const float* local_xform, const uint32_t* mesh_id,
                               uint32_t n_prims, const uint32_t* comp_offsets, uint32_t n_compounds, const bam3d_meshes* meshes,
                               bam3d_compounds** out);
void bam3d_compounds_free(bam3d_ctx* ctx, bam3d_compounds* compounds);
/* pairs: n x (left compound, right compound); comp_xform: n_compounds x 16, the model-to-world Mat4 of every compound. On the
 * device: world transform of every primitive = comp_xform.mul_mat4(local_xform) (gjk/mod.rs:377-379, glam's operation order),
 * every primitive x primitive combination in the reference's loop order through the batched kernels, then the reference's
 * reduction per compound pair:
 *  - GJK::intersection_complex (gjk/mod.rs:362-407): CollisionOnly = the first hit; FullResolution = the LAST contact of maximum
 *    penetration depth (Iterator::max_by); BAM3D_STATUS_REF_PANIC where partial_cmp(..).unwrap() would meet a NaN depth;
 *  - GJK::distance_complex (:422-456): None (MISS) at the first primitive pair that is not Some, else f32::min over all;
 *  - GJK::intersection_complex_time_of_impact (:478-523): CollisionOnly = the first hit; FullResolution = the FIRST contact of
 *    minimum time of impact (min_by, unwrap_or(Equal)); REF_PANIC where a primitive pair the reference evaluates would panic.
 * A primitive pair the device could not decide (BAM3D_STATUS_CAPACITY; the time-of-impact iteration cap BAM3D_STATUS_MAX_ITER,
 * which the reference does not have) makes its compound pair undecided too: that status is returned instead of a contact. */
int32_t bam3d_gjk_contact_complex(bam3d_ctx* ctx, bam3d_compounds* compounds, const float* comp_xform, const uint32_t* pairs, uint64_t n,
                                  const bam3d_gjk_settings* settings, int32_t strategy, bam3d_contact* out_contacts, uint8_t* out_status);
int32_t bam3d_gjk_distance_complex(bam3d_ctx* ctx, bam3d_compounds* compounds, const float* comp_xform, const uint32_t* pairs, uint64_t n,
                                   const bam3d_gjk_settings* settings, float* out_value, uint8_t* out_status);
int32_t bam3d_gjk_time_of_impact_complex(bam3d_ctx* ctx, bam3d_compounds* compounds, const float* comp_xform_start,
                                         const float* comp_xform_end, const uint32_t* pairs, uint64_t n, const bam3d_gjk_settings* settings,
                                         int32_t strategy, bam3d_contact* out_contacts, uint8_t* out_status);

/* ---- narrow phase, device buffers (asynchronous on the ctx stream) -------------------------------------- */
int32_t bam3d_gjk_intersect_dev(bam3d_ctx* ctx, const bam3d_shapes* shapes, const uint32_t* d_pairs, uint64_t n,
                                const bam3d_gjk_settings* settings, uint8_t* d_status);
int32_t bam3d_gjk_contact_dev(bam3d_ctx* ctx, const bam3d_shapes* shapes, const uint32_t* d_pairs, uint64_t n,
                              const bam3d_gjk_settings* settings, int32_t strategy, bam3d_contact* d_contacts,
                              uint8_t* d_status);
int32_t bam3d_gjk_distance_dev(bam3d_ctx* ctx, const bam3d_shapes* shapes, const uint32_t* d_pairs, uint64_t n,
                               const bam3d_gjk_settings* settings, float* d_ref_value, float* d_distance,
                               uint8_t* d_status);
int32_t bam3d_gjk_time_of_impact_dev(bam3d_ctx* ctx, const bam3d_shapes* shapes_start,
                                     const bam3d_shapes* shapes_end, const uint32_t* d_pairs, uint64_t n,
                                     const bam3d_gjk_settings* settings, bam3d_contact* d_contacts,
                                     uint8_t* d_status);
/* Ray casts against primitives: Discrete/ContinuousTransformed<Ray> for Primitive3 (primitive3.rs:133-166, ray.rs:58-78;
 * per primitive: ray.rs:35-57 Vec3, sphere.rs:45-76, cuboid.rs:85-103 and quad.rs:80-96 through aabb3.rs:165-235,
 * cylinder.rs:82-204, capsule.rs:69-178, particle.rs:53-69) and the swept particle (Particle, Range<Vec3>)
 * (particle.rs:39-128). rays: n_rays x 6 f32 = origin.xyz, direction.xyz as given (the reference does not normalise) —
 * for query 2 the 6 floats are the particle's start and end point. casts: n x (ray index, shape index).
 * query 0 = intersects_transformed (bool), 1 = intersection_transformed (Option<Vec3>), 2 = swept particle (both the
 * Discrete and the Continuous form: the hit, kept only within the travelled distance).
 * out_status: BAM3D_STATUS_OK = true / Some, BAM3D_STATUS_MISS = false / None, BAM3D_STATUS_UNSUPPORTED for a
 * ConvexPolyhedron without faces (VertexOnly: the reference's walk panics on it) or with more than 4096; out_points: n x 3 f32 world-space hit points (zero when there is none, and for query 0). */
int32_t bam3d_ray_cast(bam3d_ctx* ctx, const bam3d_shapes* shapes, const float* rays, uint64_t n_rays, const uint32_t* casts,
                       uint64_t n, int32_t query, float* out_points, uint8_t* out_status);
int32_t bam3d_ray_cast_dev(bam3d_ctx* ctx, const bam3d_shapes* shapes, const float* d_rays, uint64_t n_rays,
                           const uint32_t* d_casts, uint64_t n, int32_t query, float* d_points, uint8_t* d_status);

/* Compact the hits of a contact batch into d_out (capacity records), pair_index = index + pair_index_base.
 * *d_count (device u32/u64 pair: one uint64) receives the number of hits. Feeds the multi-GPU gather. */
int32_t bam3d_compact_contacts_dev(bam3d_ctx* ctx, const bam3d_contact* d_contacts, const uint8_t* d_status,
                                   uint64_t n, uint64_t pair_index_base, bam3d_contact40* d_out, uint64_t capacity,
                                   uint64_t* d_count);


/* ---- broad phase ---------------------------------------------------------------------------------------- */
/* Aabb3 values are 6 f32: min.xyz, max.xyz (volume/aabb/aabb3.rs:15-21). NaN bounds are not supported (the
 * reference's sort comparator treats NaN as Equal, which is not an ordering). */

/* SweepAndPrune3::find_collider_pairs (algorithm/broad_phase/sweep_prune.rs:69-128). *inout_sweep_axis is the
 * SweepAndPrune's sweep_axis state (:28,:42-52): read for this call, replaced by the axis Variance3 picks for the next
 * one (:123-125). The reference re-sorts the caller's slice in place (stable, by (min, max) on the sweep axis) and
 * returns indices into the SORTED slice; here out_perm[k] = index in `aabbs` of the box at sorted position k (may be
 * NULL), and out_pairs holds (i, j), i < j, sorted positions, in the reference's emission order (j ascending, then i).
 * The pair set is every strictly overlapping pair (aabb3.rs:237-244). If more than `capacity` pairs exist the call
 * returns BAM3D_ERR_CAPACITY with *out_count = the number required (the axis is left unchanged). */
int32_t bam3d_sap_find_pairs(bam3d_ctx* ctx, const float* aabbs, uint64_t n, int32_t* inout_sweep_axis, uint32_t* out_perm,
                             uint32_t* out_pairs, uint64_t capacity, uint64_t* out_count);
/* same with device buffers (d_perm may be NULL); synchronises once internally to size the final pair sort */
int32_t bam3d_sap_find_pairs_dev(bam3d_ctx* ctx, const float* d_aabbs, uint64_t n, int32_t* inout_sweep_axis,
                                 uint32_t* d_perm, uint32_t* d_pairs, uint64_t capacity, uint64_t* out_count);

/* Variance3 (sweep_prune.rs:166-216) on its own: add_bound for every box in the order given, then compute_axis(n).
 * out_sums6 = csum.xyz, csumsq.xyz — the same bits as the reference's sequential f32 accumulation. */
int32_t bam3d_sap_variance(bam3d_ctx* ctx, const float* aabbs, uint64_t n, float* out_sums6, int32_t* out_axis);

/* Multi-GPU form: every GPU sorts the full box set (identical perm and next axis everywhere) and searches
 * only the pairs whose higher sorted position j lies in [n * shard / n_shards, n * (shard + 1) / n_shards). The shards'
 * outputs concatenated in shard order are exactly the output of bam3d_sap_find_pairs, so each GPU can feed its slice
 * straight into its narrow phase with no pair exchange. capacity / *out_count refer to this shard's pairs. */
int32_t bam3d_sap_find_pairs_shard(bam3d_ctx* ctx, const float* aabbs, uint64_t n, int32_t* inout_sweep_axis, uint32_t* out_perm,
                                   uint32_t* out_pairs, uint64_t capacity, uint64_t* out_count, uint32_t shard, uint32_t n_shards);
int32_t bam3d_sap_find_pairs_shard_dev(bam3d_ctx* ctx, const float* d_aabbs, uint64_t n, int32_t* inout_sweep_axis,
                                       uint32_t* d_perm, uint32_t* d_pairs, uint64_t capacity, uint64_t* out_count,
                                       uint32_t shard, uint32_t n_shards);

/* ComputeBound<Aabb3> of every packed primitive (sphere.rs:27-34, cuboid.rs:66-73, cylinder.rs:64-71,
 * capsule.rs:51-58, quad.rs:62-69, polyhedron.rs:84-86, primitive3.rs:83-96) moved to world space with
 * Aabb::transform (aabb3.rs:89-96): n x 6 f32. */
int32_t bam3d_shapes_world_aabbs(bam3d_ctx* ctx, const bam3d_shapes* shapes, float* out_aabbs);
int32_t bam3d_shapes_world_aabbs_dev(bam3d_ctx* ctx, const bam3d_shapes* shapes, float* d_out_aabbs);

/* volume::Sphere as the bound type (volume/sphere.rs). A bounding sphere is 4 f32: centre.xyz, radius.
 * ComputeBound<volume::Sphere> of every packed primitive (primitive3.rs:98-114 — particle: radius 0; quad.rs:71-78; sphere.rs:36-43;
 * cuboid.rs:75-83: the largest half extent, as the reference writes it; cylinder.rs:73-80; capsule.rs:60-67; polyhedron.rs:87-92,
 * 363-370: the largest vertex norm) moved to world space with Bound::transform_volume (volume/sphere.rs:34-39: the centre is
 * transformed, the radius kept): n x 4 f32. */
int32_t bam3d_shapes_world_spheres(bam3d_ctx* ctx, const bam3d_shapes* shapes, float* out_spheres);
int32_t bam3d_shapes_world_spheres_dev(bam3d_ctx* ctx, const bam3d_shapes* shapes, float* d_out_spheres);
/* SweepAndPrune3<volume::Sphere>::find_collider_pairs (sweep_prune.rs:69-128 with Bound for Sphere, volume/sphere.rs:17-48): sorted
 * (stable) by (min_extent, max_extent) = centre -+ radius on the sweep axis; (i, j), i < j in sorted order, is reported iff i is
 * still active when j arrives (max_extent_i >= min_extent_j on the sweep axis) and Sphere::intersects holds (|ci - cj|^2 <=
 * (ri + rj)^2, volume/sphere.rs:82-91 — touching spheres count); Variance3 over the extents picks the next axis. Arguments as for
 * bam3d_sap_find_pairs; `capacity` must also cover the candidate pairs of the box search (a few per cent more than the result). */
int32_t bam3d_sap_find_pairs_spheres(bam3d_ctx* ctx, const float* spheres, uint64_t n, int32_t* inout_sweep_axis, uint32_t* out_perm,
                                     uint32_t* out_pairs, uint64_t capacity, uint64_t* out_count);
int32_t bam3d_sap_find_pairs_spheres_dev(bam3d_ctx* ctx, const float* d_spheres, uint64_t n, int32_t* inout_sweep_axis, uint32_t* d_perm,
                                         uint32_t* d_pairs, uint64_t capacity, uint64_t* out_count);

/* Broad phase -> narrow phase in one call, everything resident (BASELINE configs[3]): world Aabb3 of every shape
 * (bam3d_shapes_world_aabbs) -> SweepAndPrune3::find_collider_pairs over them (bam3d_sap_find_pairs[_shard]) -> GJK::intersection
 * (bam3d_gjk_contact) on every candidate pair. The reference leaves this glue to its caller, and find_collider_pairs returns
 * indices into the slice it has just re-sorted (sweep_prune.rs:65-69,80): out_pairs[k] = (perm[i], perm[j]) are the SHAPE
 * indices of candidate k, in the reference's emission order; out_sorted_pairs (may be NULL) the (i, j) sorted positions as the
 * reference returns them, out_perm (may be NULL) the permutation. contacts / status are per candidate pair. If more than
 * `capacity` candidates exist: BAM3D_ERR_CAPACITY with *out_count = the number required. The _dev form takes a shard
 * (multi-GPU: the sort is replicated, each GPU searches and resolves its own slice of the pair list). */
int32_t bam3d_pipeline_contacts(bam3d_ctx* ctx, const bam3d_shapes* shapes, int32_t* inout_sweep_axis, const bam3d_gjk_settings* settings,
                                int32_t strategy, uint32_t* out_pairs, uint32_t* out_sorted_pairs, uint32_t* out_perm,
                                bam3d_contact* out_contacts, uint8_t* out_status, uint64_t capacity, uint64_t* out_count);
int32_t bam3d_pipeline_contacts_dev(bam3d_ctx* ctx, const bam3d_shapes* shapes, int32_t* inout_sweep_axis, const bam3d_gjk_settings* settings,
                                    int32_t strategy, uint32_t shard, uint32_t n_shards, uint32_t* d_pairs, uint32_t* d_sorted_pairs,
                                    uint32_t* d_perm, bam3d_contact* d_contacts, uint8_t* d_status, uint64_t capacity, uint64_t* out_count);

/* Static BVH over n values' bounds, answering DynamicBoundingVolumeTree::query_for_indices (dbvt/mod.rs:481-515).
 * Value index k = position in `aabbs` (the tree's `values` order). The reference's insert / remove / rotate
 * maintenance (dbvt/mod.rs:531-1235) is not reproduced: its tree shape is randomised (rand::thread_rng, :901) and
 * query result SETS do not depend on the shape. Results of one query are returned in the device tree's traversal
 * order (deterministic for given boxes); compare with the reference as sets. The traversal is stackless (every node
 * record carries the node to continue with when its subtree is done), so no tree depth can overflow anything — the
 * reference indexes past its fixed 256-entry stack on a deep enough tree. */
int32_t bam3d_bvh_build(bam3d_ctx* ctx, const float* aabbs, uint64_t n, bam3d_bvh** out);
int32_t bam3d_bvh_build_dev(bam3d_ctx* ctx, const float* d_aabbs, uint64_t n, bam3d_bvh** out);
/* New bounds for the values of an existing tree, topology unchanged: the device form of update_node + do_refit
 * (dbvt/mod.rs:561-589) for every value at once. `aabbs` holds bam3d_bvh_count(tree) boxes in value order. Query results
 * are those of a tree built over the new boxes (result sets do not depend on the tree's shape); rebuild when the values
 * have moved far from where they were at build time. */
int32_t bam3d_bvh_refit(bam3d_ctx* ctx, bam3d_bvh* tree, const float* aabbs);
int32_t bam3d_bvh_refit_dev(bam3d_ctx* ctx, bam3d_bvh* tree, const float* d_aabbs);
void bam3d_bvh_free(bam3d_ctx* ctx, bam3d_bvh* bvh);
uint32_t bam3d_bvh_count(const bam3d_bvh* bvh);
/* Variable-length results: hits of query q are out_values[out_offsets[q] .. out_offsets[q + 1]) (out_offsets has nq + 1
 * entries, may be NULL). If the total exceeds `capacity` the call returns BAM3D_ERR_CAPACITY with *out_total (and
 * out_offsets) filled, so the caller can size its buffers and call again. */
/* DiscreteVisitor (dbvt/visitor.rs:57-90): values whose bound strictly overlaps the query Aabb3 (nq x 6 f32). */
int32_t bam3d_bvh_query_aabb(bam3d_ctx* ctx, const bam3d_bvh* bvh, const float* queries, uint64_t nq, uint64_t* out_offsets,
                             uint32_t* out_values, uint64_t capacity, uint64_t* out_total);
/* ContinuousVisitor<Ray> == query_ray (dbvt/util.rs:114-129): rays are nq x (origin.xyz, direction.xyz); out_points holds
 * the slab-test entry point per hit (aabb3.rs:165-195), 3 f32 each. */
int32_t bam3d_bvh_query_ray(bam3d_ctx* ctx, const bam3d_bvh* bvh, const float* rays, uint64_t nq, uint64_t* out_offsets,
                            uint32_t* out_values, float* out_points, uint64_t capacity, uint64_t* out_total);
/* FrustumVisitor (dbvt/visitor.rs:99-134): planes are nq x 6 x (n.xyz, d) in the order left, right, bottom, top, near,
 * far (bam3d_frustum_from_mat4); out_relation is Relation as u8: 0 In, 1 Cross (Out is never reported). */
int32_t bam3d_bvh_query_frustum(bam3d_ctx* ctx, const bam3d_bvh* bvh, const float* planes, uint64_t nq, uint64_t* out_offsets,
                                uint32_t* out_values, uint8_t* out_relation, uint64_t capacity, uint64_t* out_total);
/* query_ray_closest (dbvt/util.rs:77-103): out_value[q] = value index or 0xFFFFFFFF (None); out_point_t[q] = point.xyz, t.
 * Exact ties in t resolve to the lowest value index (the reference keeps the first in its randomised traversal order). */
int32_t bam3d_bvh_query_ray_closest(bam3d_ctx* ctx, const bam3d_bvh* bvh, const float* rays, uint64_t nq, uint32_t* out_value,
                                    float* out_point_t);
/* The same queries for a resident frame loop (dbvt/mod.rs:481-515 is what such a loop calls every frame): device pointers, ONE
 * traversal, no synchronisation. Hits are appended in arbitrary order as parallel arrays — d_out_query[i] = query index,
 * d_out_values[i] = value index, plus the visitor's result (hit point / Relation) — up to `capacity`; *d_count (device u64)
 * receives the number of hits found, which may exceed capacity (the excess is dropped). Sort by d_out_query for per-query lists. */
int32_t bam3d_bvh_query_aabb_dev(bam3d_ctx* ctx, const bam3d_bvh* bvh, const float* d_queries, uint64_t nq, uint32_t* d_out_query,
                                 uint32_t* d_out_values, uint64_t capacity, uint64_t* d_count);
int32_t bam3d_bvh_query_ray_dev(bam3d_ctx* ctx, const bam3d_bvh* bvh, const float* d_rays, uint64_t nq, uint32_t* d_out_query,
                                uint32_t* d_out_values, float* d_out_points, uint64_t capacity, uint64_t* d_count);
int32_t bam3d_bvh_query_frustum_dev(bam3d_ctx* ctx, const bam3d_bvh* bvh, const float* d_planes, uint64_t nq, uint32_t* d_out_query,
                                    uint32_t* d_out_values, uint8_t* d_out_relation, uint64_t capacity, uint64_t* d_count);
int32_t bam3d_bvh_query_ray_closest_dev(bam3d_ctx* ctx, const bam3d_bvh* bvh, const float* d_rays, uint64_t nq, uint32_t* d_out_value,
                                        float* d_out_point_t);
/* The same tree over values whose bound is volume::Sphere (volume/sphere.rs:17-196; DynamicBoundingVolumeTree is generic over the
 * bound): spheres = n x (centre.xyz, radius). The node boxes are a conservative cover of the spheres; every leaf is decided by the
 * sphere's own predicate, so result sets are those of the leaf predicate over all values (as for boxes, independent of tree shape):
 *   bam3d_bvh_query_sphere[_dev]   DiscreteVisitor with a Sphere: Discrete<Sphere> for Sphere (:82-91; touching spheres intersect)
 *   bam3d_bvh_query_ray[_dev]      ContinuousVisitor: Continuous<Ray> for Sphere (:50-67) — the first point on the sphere, no hit when
 *                                  the centre is behind the origin; directions are unit vectors, as Ray::new's callers provide
 *   bam3d_bvh_query_frustum[_dev]  FrustumVisitor: Frustum::contains over PlaneBound for Sphere (:93-104)
 *   bam3d_bvh_query_ray_closest[_dev]  the smallest t = (point - origin) . direction over the spheres the ray meets (negative when the
 *                                  origin is inside that sphere: Continuous<Ray> returns the point behind it), ties to the lowest value
 *   bam3d_bvh_find_collider_pairs  DbvtBroadPhase with Discrete<Sphere> as the leaf test
 * bam3d_bvh_query_aabb and bam3d_bvh_refit return BAM3D_ERR_ARG on such a tree (rebuild instead of refit), bam3d_bvh_query_sphere on
 * a tree over boxes likewise. */
int32_t bam3d_bvh_build_spheres(bam3d_ctx* ctx, const float* spheres, uint64_t n, bam3d_bvh** out);
int32_t bam3d_bvh_build_spheres_dev(bam3d_ctx* ctx, const float* d_spheres, uint64_t n, bam3d_bvh** out);
int32_t bam3d_bvh_query_sphere(bam3d_ctx* ctx, const bam3d_bvh* bvh, const float* spheres, uint64_t nq, uint64_t* out_offsets, uint32_t* out_values,
                               uint64_t capacity, uint64_t* out_total);
int32_t bam3d_bvh_query_sphere_dev(bam3d_ctx* ctx, const bam3d_bvh* bvh, const float* d_spheres, uint64_t nq, uint32_t* d_out_query,
                                   uint32_t* d_out_values, uint64_t capacity, uint64_t* d_count);
/* DbvtBroadPhase::find_collider_pairs (algorithm/broad_phase/dbvt.rs:26-63): every value with dirty[k] != 0 is queried
 * against the tree; out_pairs = sorted unique (low, high) value-index pairs. */
int32_t bam3d_bvh_find_collider_pairs(bam3d_ctx* ctx, const bam3d_bvh* bvh, const uint8_t* dirty, uint32_t* out_pairs,
                                      uint64_t capacity, uint64_t* out_count);
/* Frustum::from_mat4 (frustum.rs:46-75) on the host; out_planes = 6 x (n.xyz, d). Returns BAM3D_ERR_ARG for None. */
int32_t bam3d_frustum_from_mat4(const float* mat4_cols, float* out_planes);

/* ---- frame queue: the per-frame sequence with `depth` frames in flight -------------------------------------- */
/* A simulation loop calls, every frame: new Mat4 for every shape (bam3d_shapes_set_transforms), the narrow phase over the
 * frame's candidate pairs (GJK::intersect / intersection / intersection_time_of_impact, gjk/mod.rs:74-113,:317-341,:130-217),
 * results back in host memory. Done one call after the other, the PCIe copies (64 B per shape in, 33 B per pair out) and the
 * kernels take turns. A frame queue keeps `depth` frames in flight instead: host->device copies + the pack kernel on one
 * stream, the narrow-phase kernels on the ctx stream, device->host copies on a third, so that frame k+1's upload and frame
 * k-1's download run beside frame k's kernels. Results are bit-identical to the one-call-after-the-other sequence.
 * Host buffers should be page-locked (bam3d_host_alloc) — pageable memory still works but copies synchronously — and
 * must stay untouched from submit until bam3d_frames_wait for that frame returns.
 * create: kind / params / mesh_id / meshes as for bam3d_shapes_upload, `xform` = initial transforms (validated
 * synchronously together with kinds and mesh ids); with_end != 0 also allocates the end-of-step sets time of impact needs. */
typedef struct bam3d_frames bam3d_frames;
#define BAM3D_FRAME_INTERSECT 0        /* bam3d_gjk_intersect: out_status only (out_contacts may be NULL) */
#define BAM3D_FRAME_CONTACT 1          /* bam3d_gjk_contact with `strategy` */
#define BAM3D_FRAME_TIME_OF_IMPACT 2   /* bam3d_gjk_time_of_impact between xform and xform_end */
int32_t bam3d_frames_create(bam3d_ctx* ctx, const uint8_t* kind, const float* params, const float* xform, const uint32_t* mesh_id,
                            uint32_t n_shapes, const bam3d_meshes* meshes, uint64_t max_pairs, uint32_t depth, int32_t with_end,
                            bam3d_frames** out);
/* Enqueue one frame and return at once; *out_frame (may be NULL) = its number (0, 1, 2, ... in submission order). Blocks only
 * while the frame `depth` submissions ago is still in flight. xform_end is read for BAM3D_FRAME_TIME_OF_IMPACT only. */
/* Layout of the xform / xform_end arrays later submits pass (BAM3D_XFORM_*; default BAM3D_XFORM_MAT4). */
int32_t bam3d_frames_set_transform_layout(bam3d_ctx* ctx, bam3d_frames* frames, int32_t layout);
int32_t bam3d_frames_submit(bam3d_ctx* ctx, bam3d_frames* frames, int32_t query, const float* xform, const float* xform_end,
                            const uint32_t* pairs, uint64_t n, const bam3d_gjk_settings* settings, int32_t strategy,
                            bam3d_contact* out_contacts, uint8_t* out_status, uint64_t* out_frame);
/* Block until that frame's results are in the host buffers given to its submit. Returns the frame's own status:
 * BAM3D_ERR_NON_AFFINE if one of its transforms had a 4th row other than (0,0,0,1) (its results are then meaningless). */
int32_t bam3d_frames_wait(bam3d_ctx* ctx, bam3d_frames* frames, uint64_t frame);
void bam3d_frames_free(bam3d_ctx* ctx, bam3d_frames* frames);

/* ---- multi-GPU: NCCL gather of the ranks' compacted contacts ------------------------------------------ */
/* One process per GPU; pair lists shard across ranks with no data-path collective (each rank runs the narrow phase on its
 * slice), and the one exchange step is the gather of the ranks' compacted contact lists so that every rank ends with the
 * contacts of all pairs. The reference has no distributed layer; this is the `gpu` feature's addition (north_star). NCCL is
 * bound at run time (dlopen of libnccl.so.2): every call returns BAM3D_ERR_COMM when it is not installed. */
typedef struct bam3d_comm bam3d_comm;
#define BAM3D_UNIQUE_ID_BYTES 256   /* two ncclUniqueId: counts and payloads travel on separate communicators */
/* rank 0: a fresh rendezvous id (ncclGetUniqueId); hand the bytes to the other ranks by any host-side means */
int32_t bam3d_comm_unique_id(uint8_t* out_id);
/* every rank, collectively: join the communicator on `device` */
int32_t bam3d_comm_create(int32_t device, const uint8_t* id, int32_t rank, int32_t n_ranks, bam3d_comm** out);
void bam3d_comm_destroy(bam3d_comm* comm);
int32_t bam3d_comm_rank(const bam3d_comm* comm);
int32_t bam3d_comm_size(const bam3d_comm* comm);
void* bam3d_comm_stream(bam3d_comm* comm);          /* the cudaStream_t the exchange runs on */
const char* bam3d_comm_unavailable_reason(void);    /* "" when libnccl was loaded */
/* The exchange, in two halves so that it overlaps the next step's kernels. `begin` (asynchronous): after everything
 * enqueued on the ctx stream so far, all-gather the ranks' record counts (*d_count as written by
 * bam3d_compact_contacts_dev) and copy them to the host. `finish`: waits on the host for that ticket's counts, then
 * enqueues — on the comm stream — one broadcast per rank with the exact sizes, so that d_gathered holds the ranks' records
 * concatenated in rank order (out_counts[n_ranks] on the host, may be NULL; *out_total = their sum; BAM3D_ERR_CAPACITY
 * if it exceeds `capacity` records). bam3d_comm_join makes the ctx stream wait for the exchange before it reads
 * d_gathered. Every rank must call begin / finish in the same order; at most 16 tickets may be open. Counts and payloads use
 * separate communicators and streams, so a payload never queues behind the count exchange of a later step. */
int32_t bam3d_gather_contacts_begin(bam3d_ctx* ctx, bam3d_comm* comm, const uint64_t* d_count, uint64_t* out_ticket);
int32_t bam3d_gather_contacts_finish(bam3d_ctx* ctx, bam3d_comm* comm, uint64_t ticket, const bam3d_contact40* d_compact,
                                     bam3d_contact40* d_gathered, uint64_t capacity, uint64_t* out_counts, uint64_t* out_total);
int32_t bam3d_comm_join(bam3d_ctx* ctx, bam3d_comm* comm);
/* finer than join: the ctx stream waits only for that ticket's payload (before it overwrites d_compact / reads d_gathered) */
int32_t bam3d_gather_contacts_wait(bam3d_ctx* ctx, bam3d_comm* comm, uint64_t ticket);
/* begin + finish + join in one call (no overlap) */
int32_t bam3d_gather_contacts_dev(bam3d_ctx* ctx, bam3d_comm* comm, const bam3d_contact40* d_compact, const uint64_t* d_count,
                                  bam3d_contact40* d_gathered, uint64_t capacity, uint64_t* out_counts, uint64_t* out_total);
/* host buffers: this rank's dense results (as bam3d_gjk_contact / _time_of_impact returned them for its n pairs, whose global
 * indices start at pair_index_base) in, the hits of ALL ranks out (rank order), synchronous on return */
int32_t bam3d_gather_contacts(bam3d_ctx* ctx, bam3d_comm* comm, const bam3d_contact* contacts, const uint8_t* status, uint64_t n,
                              uint64_t pair_index_base, bam3d_contact40* out, uint64_t capacity, uint64_t* out_counts,
                              uint64_t* out_total);

/* Frame queue with the exchange inside: like bam3d_frames_submit for a contact-producing query, but the frame's hits are
 * compacted on the device (pair_index = index in `pairs` + pair_index_base), exchanged over `comm`, and bam3d_frames_wait
 * leaves the hits of ALL ranks (rank order) in out_gathered (host, `capacity` records) instead of this rank's dense contacts;
 * out_status still receives this rank's per-pair status. The record counts travel while the next frame computes; the
 * payload is exchanged inside bam3d_frames_wait. Every rank must submit and wait in the same order.
 * bam3d_frames_gathered (after the wait): per-rank record counts (n_ranks, may be NULL) and their sum. */
int32_t bam3d_frames_submit_gather(bam3d_ctx* ctx, bam3d_frames* frames, bam3d_comm* comm, int32_t query, const float* xform,
                                   const float* xform_end, const uint32_t* pairs, uint64_t n, const bam3d_gjk_settings* settings,
                                   int32_t strategy, uint64_t pair_index_base, bam3d_contact40* out_gathered, uint64_t capacity,
                                   uint8_t* out_status, uint64_t* out_frame);
int32_t bam3d_frames_gathered(bam3d_ctx* ctx, bam3d_frames* frames, uint64_t frame, uint64_t* out_counts, uint64_t* out_total);

/* number of kernels this ctx has launched so far (bench.py's gpu_launches) */
uint64_t bam3d_ctx_launch_count(const bam3d_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* BAM3D_GPU_H */
