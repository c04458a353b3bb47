// ORACLE — TEST INFRASTRUCTURE ONLY (see glam_mini.hpp header). C ABI over the CPU restatement so that tests/
// (ctypes), __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs can call it. Batched
// entry points take the same SoA layout the GPU C-ABI (include/bam3d_gpu.h) takes and hand the batch out in chunks
// to `nthreads` std::threads (the reference itself is single-threaded; threads only replicate it).
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>
#include "bam3d_oracle.hpp"
#include "broad_oracle.hpp"
#include "ray_oracle.hpp"

using namespace orc;

namespace {

// Half-edge structures of the mesh library's polyhedra that were built with faces (orc_set_mesh_faces); mesh m is in
// HalfEdge mode iff g_half_edges[m] has faces. Process-wide, set by the tests before the batch calls that use it.
std::vector<HalfEdgeMesh> g_half_edges;

struct ShapeTable {
    const uint8_t* kind; const float* params; const float* xform; const uint32_t* mesh_id;
    const float* mesh_verts; const uint32_t* mesh_offsets;
    Shape shape(uint32_t i) const {
        Shape s; s.kind = kind[i]; s.p0 = params[4 * i]; s.p1 = params[4 * i + 1]; s.p2 = params[4 * i + 2];
        if (s.kind == POLYHEDRON) {
            uint32_t m = mesh_id[i];
            s.verts = mesh_verts + 3 * (size_t)mesh_offsets[m];
            s.nverts = mesh_offsets[m + 1] - mesh_offsets[m];
            if (m < g_half_edges.size() && !g_half_edges[m].faces.empty()) s.he = &g_half_edges[m];
        }
        return s;
    }
    Mat4 mat(uint32_t i) const { return Mat4::from_cols_array(xform + 16 * (size_t)i); }
};

// f(begin, end, thread) over [0, n) in chunks of 256 handed out by an atomic counter: a few pairs cost 10^4 times the median
// (pinched EPA polytopes, time-of-impact sweeps that run to the cap), so a static split leaves most threads idle behind one.
template <class F>
void parallel_for(uint64_t n, int nthreads, F f) {
    if (nthreads <= 1 || n < 64) { f(0, n, 0); return; }
    const uint64_t chunk = 256;
    std::atomic<uint64_t> next{0};
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t) {
        th.emplace_back([&, t] {
            for (;;) {
                const uint64_t b = next.fetch_add(chunk);
                if (b >= n) break;
                f(b, std::min<uint64_t>(n, b + chunk), t);
            }
        });
    }
    for (auto& x : th) x.join();
}

void put_contact(float* out, const Contact& c) {
    out[0] = c.normal.x; out[1] = c.normal.y; out[2] = c.normal.z; out[3] = c.penetration_depth;
    out[4] = c.contact_point.x; out[5] = c.contact_point.y; out[6] = c.contact_point.z; out[7] = c.time_of_impact;
}

template <class R> void sort_hits(std::vector<std::pair<uint32_t, R>>& h) {
    std::stable_sort(h.begin(), h.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
}

}  // namespace

extern "C" {

struct orc_settings { float distance_tolerance, continuous_tolerance, epa_tolerance; uint32_t max_iterations; };
struct orc_counters { uint64_t support_calls, gjk_iterations, epa_iterations; uint32_t max_faces, max_verts, max_edges, pad; };

static void merge(orc_counters* dst, const std::vector<Counters>& cs) {
    if (!dst) return;
    std::memset(dst, 0, sizeof(*dst));
    for (auto& c : cs) {
        dst->support_calls += c.support_calls; dst->gjk_iterations += c.gjk_iterations; dst->epa_iterations += c.epa_iterations;
        if (c.max_faces > dst->max_faces) dst->max_faces = c.max_faces;
        if (c.max_verts > dst->max_verts) dst->max_verts = c.max_verts;
        if (c.max_edges > dst->max_edges) dst->max_edges = c.max_edges;
    }
}

static GJK make_gjk(const orc_settings* s) {
    return s ? GJK(s->distance_tolerance, s->continuous_tolerance, s->epa_tolerance, s->max_iterations) : GJK();
}

// GJK::intersect over a batch (gjk/mod.rs:74-113). out_hit[i] in {0,1}; out_status[i] = orc::Status.
void orc_gjk_intersect_batch(const uint8_t* kind, const float* params, const float* xform, const uint32_t* mesh_id,
                             const float* mesh_verts, const uint32_t* mesh_offsets, const uint32_t* pairs, uint64_t n,
                             const orc_settings* settings, uint8_t* out_hit, uint8_t* out_status, int nthreads,
                             orc_counters* counters) {
    ShapeTable T{kind, params, xform, mesh_id, mesh_verts, mesh_offsets};
    GJK gjk = make_gjk(settings);
    std::vector<Counters> cs(nthreads > 0 ? nthreads : 1);
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e, int t) {
        for (uint64_t i = b; i < e; ++i) {
            uint32_t l = pairs[2 * i], r = pairs[2 * i + 1];
            Simplex s; Status st;
            bool hit = gjk.intersect(T.shape(l), T.mat(l), T.shape(r), T.mat(r), s, &st, counters ? &cs[t] : nullptr);
            out_hit[i] = hit ? 1 : 0;
            if (out_status) out_status[i] = (uint8_t)st;
        }
    });
    merge(counters, cs);
}

// GJK::intersection over a batch (gjk/mod.rs:317-341). strategy: 0 FullResolution, 1 CollisionOnly.
// out_contact: 8 floats per pair {normal xyz, depth, point xyz, toi}; zero-filled on miss.
void orc_gjk_contact_batch(const uint8_t* kind, const float* params, const float* xform, const uint32_t* mesh_id,
                           const float* mesh_verts, const uint32_t* mesh_offsets, const uint32_t* pairs, uint64_t n,
                           const orc_settings* settings, int strategy, float* out_contact, uint8_t* out_hit,
                           uint8_t* out_status, int nthreads, orc_counters* counters) {
    ShapeTable T{kind, params, xform, mesh_id, mesh_verts, mesh_offsets};
    GJK gjk = make_gjk(settings);
    std::vector<Counters> cs(nthreads > 0 ? nthreads : 1);
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e, int t) {
        for (uint64_t i = b; i < e; ++i) {
            uint32_t l = pairs[2 * i], r = pairs[2 * i + 1];
            Contact c; Status st;
            bool hit = gjk.intersection((CollisionStrategy)strategy, T.shape(l), T.mat(l), T.shape(r), T.mat(r), c, &st,
                                        counters ? &cs[t] : nullptr);
            out_hit[i] = hit ? 1 : 0;
            if (out_status) out_status[i] = (uint8_t)st;
            if (hit) put_contact(out_contact + 8 * i, c); else std::memset(out_contact + 8 * i, 0, 32);
        }
    });
    merge(counters, cs);
}

// Per-pair EPA statistics (polytope face high-water, EPA iterations) for sizing the device tiers; FullResolution.
void orc_gjk_contact_stats_batch(const uint8_t* kind, const float* params, const float* xform, const uint32_t* mesh_id,
                                 const float* mesh_verts, const uint32_t* mesh_offsets, const uint32_t* pairs, uint64_t n,
                                 uint32_t* out_max_faces, uint32_t* out_epa_iterations, uint8_t* out_status, int nthreads,
                                 uint32_t* out_max_edges) {
    ShapeTable T{kind, params, xform, mesh_id, mesh_verts, mesh_offsets};
    GJK gjk;
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t i = b; i < e; ++i) {
            uint32_t l = pairs[2 * i], r = pairs[2 * i + 1];
            Contact c; Status st; Counters cnt;
            gjk.intersection(CollisionStrategy::FullResolution, T.shape(l), T.mat(l), T.shape(r), T.mat(r), c, &st, &cnt);
            out_max_faces[i] = cnt.max_faces; out_epa_iterations[i] = (uint32_t)cnt.epa_iterations; out_status[i] = (uint8_t)st;
            if (out_max_edges) out_max_edges[i] = cnt.max_edges;
        }
    });
}

// The same with the oracle's EPA face limit replaced by `face_limit` (the reference's Vec is unbounded): what the reference
// itself returns for the pairs the device reports as BAM3D_STATUS_CAPACITY. Single-threaded per pair; a pair whose polytope
// passes face_limit still ends as Status::Capacity (the reference would keep growing it until the iteration cap or OOM).
void orc_gjk_contact_face_limit(const uint8_t* kind, const float* params, const float* xform, const uint32_t* mesh_id,
                                const float* mesh_verts, const uint32_t* mesh_offsets, const uint32_t* pairs, uint64_t n,
                                uint64_t face_limit, float* out_contact, uint8_t* out_status, uint32_t* out_max_faces,
                                uint32_t* out_epa_iterations, uint64_t* out_removed_faces, int nthreads, int in_phases) {
    ShapeTable T{kind, params, xform, mesh_id, mesh_verts, mesh_offsets};
    GJK gjk;
    gjk.epa.face_limit = (size_t)face_limit;
    gjk.epa.in_phases = in_phases != 0;
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t i = b; i < e; ++i) {
            uint32_t l = pairs[2 * i], r = pairs[2 * i + 1];
            Contact c; Status st; Counters cnt;
            bool hit = gjk.intersection(CollisionStrategy::FullResolution, T.shape(l), T.mat(l), T.shape(r), T.mat(r), c, &st, &cnt);
            out_status[i] = (uint8_t)st;
            out_max_faces[i] = cnt.max_faces; out_epa_iterations[i] = (uint32_t)cnt.epa_iterations;
            if (out_removed_faces) out_removed_faces[i] = cnt.removed_faces;
            if (hit) put_contact(out_contact + 8 * i, c); else std::memset(out_contact + 8 * i, 0, 32);
        }
    });
}

// GJK::distance over a batch (gjk/mod.rs:232-280). out_some[i]=1 for Some; out_ref = dp-d0 (what the reference
// returns), out_geo = sqrt(d.d).
void orc_gjk_distance_batch(const uint8_t* kind, const float* params, const float* xform, const uint32_t* mesh_id,
                            const float* mesh_verts, const uint32_t* mesh_offsets, const uint32_t* pairs, uint64_t n,
                            const orc_settings* settings, float* out_ref, float* out_geo, uint8_t* out_some,
                            uint8_t* out_status, int nthreads, orc_counters* counters) {
    ShapeTable T{kind, params, xform, mesh_id, mesh_verts, mesh_offsets};
    GJK gjk = make_gjk(settings);
    std::vector<Counters> cs(nthreads > 0 ? nthreads : 1);
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e, int t) {
        for (uint64_t i = b; i < e; ++i) {
            uint32_t l = pairs[2 * i], r = pairs[2 * i + 1];
            float rv = 0.f, geo = 0.f; Status st;
            bool some = gjk.distance(T.shape(l), T.mat(l), T.shape(r), T.mat(r), rv, geo, st, counters ? &cs[t] : nullptr);
            out_some[i] = some ? 1 : 0; out_ref[i] = some ? rv : 0.f; out_geo[i] = some ? geo : 0.f;
            if (out_status) out_status[i] = (uint8_t)st;
        }
    });
    merge(counters, cs);
}

// GJK::intersection_time_of_impact over a batch (gjk/mod.rs:130-217). xform0 = start transforms, xform1 = end.
void orc_gjk_toi_batch(const uint8_t* kind, const float* params, const float* xform0, const float* xform1,
                       const uint32_t* mesh_id, const float* mesh_verts, const uint32_t* mesh_offsets,
                       const uint32_t* pairs, uint64_t n, const orc_settings* settings, uint32_t toi_cap,
                       float* out_contact, uint8_t* out_hit, uint8_t* out_status, int nthreads, orc_counters* counters) {
    ShapeTable T0{kind, params, xform0, mesh_id, mesh_verts, mesh_offsets};
    ShapeTable T1{kind, params, xform1, mesh_id, mesh_verts, mesh_offsets};
    GJK gjk = make_gjk(settings);
    if (toi_cap) gjk.toi_iteration_cap = toi_cap;
    std::vector<Counters> cs(nthreads > 0 ? nthreads : 1);
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e, int t) {
        for (uint64_t i = b; i < e; ++i) {
            uint32_t l = pairs[2 * i], r = pairs[2 * i + 1];
            Contact c; Status st;
            bool hit = gjk.intersection_time_of_impact(T0.shape(l), T0.mat(l), T1.mat(l), T0.shape(r), T0.mat(r), T1.mat(r),
                                                       c, st, counters ? &cs[t] : nullptr);
            out_hit[i] = hit ? 1 : 0;
            if (out_status) out_status[i] = (uint8_t)st;
            if (hit) put_contact(out_contact + 8 * i, c); else std::memset(out_contact + 8 * i, 0, 32);
        }
    });
    merge(counters, cs);
}

// Compound shapes (gjk/mod.rs:362-523). A compound = primitives [begin, end) of the shape table, whose `xform` entries
// are the LOCAL transforms; comp_xform0/1 are the compounds' model-to-world transforms at the start / end of the step.
// mode: 0 intersection_complex, 1 distance_complex, 2 intersection_complex_time_of_impact. out: 8 floats per compound
// pair (contact layout; distance in [3]).
void orc_gjk_complex_batch(const uint8_t* kind, const float* params, const float* xform, const uint32_t* mesh_id, const float* mesh_verts,
                           const uint32_t* mesh_offsets, const uint32_t* comp_offsets, const float* comp_xform0, const float* comp_xform1,
                           const uint32_t* pairs, uint64_t n, int mode, int strategy, float* out, uint8_t* out_status) {
    ShapeTable T{kind, params, xform, mesh_id, mesh_verts, mesh_offsets};
    GJK gjk;
    auto compound = [&](uint32_t c) {
        std::vector<std::pair<Shape, Mat4>> v;
        for (uint32_t i = comp_offsets[c]; i < comp_offsets[c + 1]; ++i) v.emplace_back(T.shape(i), T.mat(i));
        return v;
    };
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t l = pairs[2 * i], r = pairs[2 * i + 1];
        auto L = compound(l), R = compound(r);
        Mat4 l0 = Mat4::from_cols_array(comp_xform0 + 16 * (size_t)l), r0 = Mat4::from_cols_array(comp_xform0 + 16 * (size_t)r);
        Contact c; Status st = Status::Miss; bool ok = false;
        std::memset(out + 8 * i, 0, 32);
        if (mode == 0) ok = gjk.intersection_complex((CollisionStrategy)strategy, L, l0, R, r0, c, st);
        else if (mode == 1) { float d = 0.f; ok = gjk.distance_complex(L, l0, R, r0, d, st); if (ok) out[8 * i + 3] = d; }
        else {
            Mat4 l1 = Mat4::from_cols_array(comp_xform1 + 16 * (size_t)l), r1 = Mat4::from_cols_array(comp_xform1 + 16 * (size_t)r);
            ok = gjk.intersection_complex_time_of_impact((CollisionStrategy)strategy, L, l0, l1, R, r0, r1, c, st);
        }
        if (ok && mode != 1) put_contact(out + 8 * i, c);
        out_status[i] = ok ? 0 : (uint8_t)st;
    }
}

// ---- single-call helpers used by the parity / golden tests -------------------------------------------------
// ConvexPolyhedron::new_with_faces for the meshes of a library: faces = (a, b, c) vertex indices local to the mesh,
// face_offsets[m] .. face_offsets[m + 1] the faces of mesh m (none = VertexOnly). Returns 0, or 1 + the index of the first
// mesh with a degenerate / out-of-range face (the reference panics there). n_meshes = 0 clears the registry.
int32_t orc_set_mesh_faces(const float* mesh_verts, const uint32_t* mesh_offsets, uint32_t n_meshes, const uint32_t* faces,
                           const uint32_t* face_offsets) {
    g_half_edges.clear();
    g_half_edges.resize(n_meshes);
    for (uint32_t m = 0; m < n_meshes; ++m) {
        const uint32_t nf = face_offsets[m + 1] - face_offsets[m];
        if (nf == 0) continue;
        if (!g_half_edges[m].build(mesh_verts + 3 * (size_t)mesh_offsets[m], mesh_offsets[m + 1] - mesh_offsets[m],
                                   faces + 3 * (size_t)face_offsets[m], nf)) { g_half_edges.clear(); return 1 + (int32_t)m; }
    }
    return 0;
}

// Ray casts against primitives (ray_oracle.hpp). rays: n_rays x 6 f32 (origin, direction) — or, for query 2, segments
// (start, end) of a swept Particle; casts: n x (ray index, shape index). query: 0 = intersects_transformed,
// 1 = intersection_transformed, 2 = (Particle, Range<Vec3>). out_status: 0 hit, 1 miss, 6 unsupported (a ConvexPolyhedron
// without faces);
// out_points: n x 3 f32, the world-space hit point for queries 1 and 2 (zero otherwise).
void orc_ray_cast_batch(const uint8_t* kind, const float* params, const float* xform, const uint32_t* mesh_id, const float* mesh_verts,
                        const uint32_t* mesh_offsets, const float* rays, const uint32_t* casts, uint64_t n, int32_t query,
                        float* out_points, uint8_t* out_status, int nthreads) {
    ShapeTable T{kind, params, xform, mesh_id, mesh_verts, mesh_offsets};
    parallel_for(n, nthreads, [&](uint64_t b, uint64_t e, int) {
        for (uint64_t i = b; i < e; ++i) {
            const float* r = rays + 6 * (size_t)casts[2 * i];
            const uint32_t si = casts[2 * i + 1];
            Vec3 p = Vec3::zero();
            uint8_t st = 1;
            Shape s0 = (kind[si] == POLYHEDRON && !mesh_id) ? Shape() : T.shape(si);
            if (kind[si] == POLYHEDRON && !s0.he) {
                st = 6;  // VertexOnly polyhedron: no faces to walk
            } else {
                Shape s = s0;
                Mat4 m = T.mat(si);
                bool hit;
                if (query == 0) hit = ray_intersects_transformed(s, Ray(Vec3(r[0], r[1], r[2]), Vec3(r[3], r[4], r[5])), m);
                else if (query == 1) hit = ray_intersection_transformed(s, Ray(Vec3(r[0], r[1], r[2]), Vec3(r[3], r[4], r[5])), m, p);
                else hit = particle_sweep_transformed(s, Vec3(r[0], r[1], r[2]), Vec3(r[3], r[4], r[5]), m, p);
                st = hit ? 0 : 1;
                if (!hit) p = Vec3::zero();
            }
            out_status[i] = st;
            out_points[3 * i] = p.x; out_points[3 * i + 1] = p.y; out_points[3 * i + 2] = p.z;
        }
    });
}

void orc_support_point(uint8_t kind, const float* params4, const float* verts, uint32_t nverts, const float* dir3,
                       const float* xform16, float* out3) {
    Shape s; s.kind = kind; s.p0 = params4[0]; s.p1 = params4[1]; s.p2 = params4[2]; s.verts = verts; s.nverts = nverts;
    Vec3 p = support_point(s, Vec3(dir3[0], dir3[1], dir3[2]), Mat4::from_cols_array(xform16));
    out3[0] = p.x; out3[1] = p.y; out3[2] = p.z;
}
void orc_mat4_from_srt(const float* scale3, const float* quat4, const float* trans3, float* out16) {
    Mat4::from_scale_rotation_translation(Vec3(scale3[0], scale3[1], scale3[2]), Quat(quat4[0], quat4[1], quat4[2], quat4[3]),
                                          Vec3(trans3[0], trans3[1], trans3[2])).to_cols_array(out16);
}
void orc_mat4_inverse(const float* in16, float* out16) { Mat4::from_cols_array(in16).inverse().to_cols_array(out16); }
void orc_quat_from_rotation_z(float angle, float* out4) { Quat q = Quat::from_rotation_z(angle); out4[0] = q.x; out4[1] = q.y; out4[2] = q.z; out4[3] = q.w; }
void orc_translation_interpolate(const float* a16, const float* b16, float amount, float* out16) {
    translation_interpolate(Mat4::from_cols_array(a16), Mat4::from_cols_array(b16), amount).to_cols_array(out16);
}

// ComputeBound<Aabb3> for each primitive (sphere.rs:27-34, cuboid.rs:66-73, cylinder.rs:64-71, capsule.rs:51-58,
// quad.rs:62-69, polyhedron.rs:84-86 + :357-361, primitive3.rs:83-96) followed by Aabb::transform (aabb3.rs:89-96).
void orc_world_aabb_batch(const uint8_t* kind, const float* params, const float* xform, const uint32_t* mesh_id,
                          const float* mesh_verts, const uint32_t* mesh_offsets, uint64_t n, float* out6) {
    ShapeTable T{kind, params, xform, mesh_id, mesh_verts, mesh_offsets};
    for (uint64_t i = 0; i < n; ++i) {
        Shape s = T.shape((uint32_t)i);
        Aabb3 b;
        switch (s.kind) {
            case PARTICLE: b = Aabb3::zero(); break;
            case QUAD: b = Aabb3(Vec3(-s.p0, -s.p1, 0.f), Vec3(s.p0, s.p1, 0.f)); break;
            case SPHERE: b = Aabb3(Vec3::splat(-s.p0), Vec3::splat(s.p0)); break;
            case CUBOID: b = Aabb3(-Vec3(s.p0, s.p1, s.p2), Vec3(s.p0, s.p1, s.p2)); break;
            case CYLINDER: b = Aabb3(Vec3(-s.p1, -s.p0, -s.p1), Vec3(s.p1, s.p0, s.p1)); break;
            case CAPSULE: b = Aabb3(Vec3(-s.p1, -s.p0 - s.p1, -s.p1), Vec3(s.p1, s.p0 + s.p1, s.p1)); break;
            case POLYHEDRON:
                b = Aabb3::zero();
                for (uint32_t k = 0; k < s.nverts; ++k) b = b.grow(Vec3(s.verts[3 * k], s.verts[3 * k + 1], s.verts[3 * k + 2]));
                break;
        }
        Aabb3 w = b.transform(T.mat((uint32_t)i));
        float* o = out6 + 6 * i;
        o[0] = w.min.x; o[1] = w.min.y; o[2] = w.min.z; o[3] = w.max.x; o[4] = w.max.y; o[5] = w.max.z;
    }
}

// ComputeBound<volume::Sphere> for each primitive (primitive3.rs:98-114; quad.rs:71-78, sphere.rs:36-43, cuboid.rs:75-83,
// cylinder.rs:73-80, capsule.rs:60-67, polyhedron.rs:87-92 + :363-370) followed by Bound::transform_volume (volume/sphere.rs:34-39).
void orc_world_sphere_batch(const uint8_t* kind, const float* params, const float* xform, const uint32_t* mesh_id,
                            const float* mesh_verts, const uint32_t* mesh_offsets, uint64_t n, float* out4) {
    ShapeTable T{kind, params, xform, mesh_id, mesh_verts, mesh_offsets};
    for (uint64_t i = 0; i < n; ++i) {
        Shape s = T.shape((uint32_t)i);
        BSphere b = BSphere::empty();
        switch (s.kind) {
            case PARTICLE: break;
            case QUAD: b.radius = rmax(s.p0, s.p1); break;
            case SPHERE: b.radius = s.p0; break;
            case CUBOID: b.radius = rmax(rmax(s.p0, s.p1), s.p2); break;
            case CYLINDER: b.radius = std::sqrt((s.p1 * s.p1) + (s.p0 * s.p0)); break;
            case CAPSULE: b.radius = s.p0 + s.p1; break;
            case POLYHEDRON: {
                // .map(|p| p.dot(p).sqrt()).max_by(partial_cmp.unwrap_or(Equal)).unwrap_or(0.): max_by keeps the accumulated
                // element only when it compares Greater than the next one
                bool any = false;
                float acc = 0.f;
                for (uint32_t k = 0; k < s.nverts; ++k) {
                    Vec3 p(s.verts[3 * k], s.verts[3 * k + 1], s.verts[3 * k + 2]);
                    float len = std::sqrt(dot(p, p));
                    if (!any || !(acc > len)) acc = len;
                    any = true;
                }
                b.radius = acc;
                break;
            }
        }
        BSphere w = b.transform_volume(T.mat((uint32_t)i));
        float* o = out4 + 4 * i;
        o[0] = w.center.x; o[1] = w.center.y; o[2] = w.center.z; o[3] = w.radius;
    }
}

// ---- broad phase ---------------------------------------------------------------------------------------------
static Aabb3 box6(const float* a) { Aabb3 b; b.min = Vec3(a[0], a[1], a[2]); b.max = Vec3(a[3], a[4], a[5]); return b; }

// SweepAndPrune3::find_collider_pairs (sweep_prune.rs:69-128). aabbs: n x {min xyz, max xyz}, NOT modified;
// out_perm[k] = original index at sorted position k; pairs are (i, j) indices into the sorted order.
// Returns the number of pairs found (may exceed pair_capacity; only the first pair_capacity are written).
uint64_t orc_sap_find_pairs(const float* aabbs, uint64_t n, int32_t* inout_axis, uint32_t* out_perm, uint32_t* out_pairs,
                            uint64_t pair_capacity) {
    std::vector<Aabb3> b(n);
    for (uint64_t i = 0; i < n; ++i) b[i] = box6(aabbs + 6 * i);
    SweepAndPrune3 sap; sap.sweep_axis = *inout_axis;
    std::vector<uint32_t> perm;
    auto pairs = sap.find_collider_pairs(b, &perm);
    *inout_axis = sap.sweep_axis;
    if (out_perm) std::memcpy(out_perm, perm.data(), n * sizeof(uint32_t));
    uint64_t w = std::min<uint64_t>(pairs.size(), pair_capacity);
    for (uint64_t i = 0; i < w; ++i) { out_pairs[2 * i] = pairs[i].first; out_pairs[2 * i + 1] = pairs[i].second; }
    return pairs.size();
}

// SweepAndPrune3<volume::Sphere>::find_collider_pairs; spheres: n x {centre xyz, radius}. Same conventions as orc_sap_find_pairs.
uint64_t orc_sap_find_pairs_spheres(const float* spheres, uint64_t n, int32_t* inout_axis, uint32_t* out_perm, uint32_t* out_pairs,
                                    uint64_t pair_capacity) {
    std::vector<BSphere> b(n);
    for (uint64_t i = 0; i < n; ++i) b[i] = BSphere(Vec3(spheres[4 * i], spheres[4 * i + 1], spheres[4 * i + 2]), spheres[4 * i + 3]);
    SweepAndPrune3Sphere sap; sap.sweep_axis = *inout_axis;
    std::vector<uint32_t> perm;
    auto pairs = sap.find_collider_pairs(b, &perm);
    *inout_axis = sap.sweep_axis;
    if (out_perm) std::memcpy(out_perm, perm.data(), n * sizeof(uint32_t));
    uint64_t w = std::min<uint64_t>(pairs.size(), pair_capacity);
    for (uint64_t i = 0; i < w; ++i) { out_pairs[2 * i] = pairs[i].first; out_pairs[2 * i + 1] = pairs[i].second; }
    return pairs.size();
}

// Variance3 alone (sweep_prune.rs:192-215) over boxes in the given order.
int32_t orc_sap_variance_axis(const float* aabbs, uint64_t n) {
    Variance3 v; v.clear();
    for (uint64_t i = 0; i < n; ++i) v.add_bound(box6(aabbs + 6 * i));
    return v.compute_axis((float)n);
}

// Variance3 over the boxes in the given order: out6 = csum.xyz, csumsq.xyz; returns compute_axis(n).
int32_t orc_sap_variance_sums(const float* aabbs, uint64_t n, float* out6) {
    Variance3 v; v.clear();
    for (uint64_t i = 0; i < n; ++i) v.add_bound(box6(aabbs + 6 * i));
    out6[0] = v.csum.x; out6[1] = v.csum.y; out6[2] = v.csum.z; out6[3] = v.csumsq.x; out6[4] = v.csumsq.y; out6[5] = v.csumsq.z;
    return v.compute_axis((float)n);
}

struct orc_dbvt { Dbvt tree; };
orc_dbvt* orc_dbvt_build(const float* aabbs, const float* fat_aabbs, uint64_t n) {
    orc_dbvt* t = new orc_dbvt();
    for (uint64_t i = 0; i < n; ++i) {
        TreeValue v; v.bound = box6(aabbs + 6 * i); v.fat = box6((fat_aabbs ? fat_aabbs : aabbs) + 6 * i);
        t->tree.insert(v);
    }
    t->tree.refit();
    return t;
}
void orc_dbvt_free(orc_dbvt* t) { delete t; }
int32_t orc_dbvt_stack_overflow(orc_dbvt* t) { return t->tree.stack_overflow ? 1 : 0; }


// DiscreteVisitor query (visitor.rs:57-90). Results sorted by value index. Returns hit count.
uint64_t orc_dbvt_query_aabb(orc_dbvt* t, const float* q6, uint32_t* out_values, uint64_t cap) {
    DiscreteVisitor vis; vis.bound = box6(q6);
    std::vector<std::pair<uint32_t, uint8_t>> h;
    t->tree.query_for_indices<DiscreteVisitor, uint8_t>(vis, h);
    sort_hits(h);
    for (uint64_t i = 0; i < h.size() && i < cap; ++i) out_values[i] = h[i].first;
    return h.size();
}
// ContinuousVisitor<Ray> query == query_ray (util.rs:114-129). out_points: 3 floats per hit.
uint64_t orc_dbvt_query_ray(orc_dbvt* t, const float* ray6, uint32_t* out_values, float* out_points, uint64_t cap) {
    ContinuousRayVisitor vis; vis.ray = Ray(Vec3(ray6[0], ray6[1], ray6[2]), Vec3(ray6[3], ray6[4], ray6[5]));
    std::vector<std::pair<uint32_t, Vec3>> h;
    t->tree.query_for_indices<ContinuousRayVisitor, Vec3>(vis, h);
    sort_hits(h);
    for (uint64_t i = 0; i < h.size() && i < cap; ++i) {
        out_values[i] = h[i].first; out_points[3 * i] = h[i].second.x; out_points[3 * i + 1] = h[i].second.y; out_points[3 * i + 2] = h[i].second.z;
    }
    return h.size();
}
// query_ray_closest (util.rs:77-103). Returns 1 if a value was hit.
int32_t orc_dbvt_query_ray_closest(orc_dbvt* t, const float* ray6, uint32_t* out_value, float* out_point3, float* out_t) {
    uint32_t v = 0; Vec3 p; float tt = 0.f;
    bool ok = query_ray_closest(t->tree, Ray(Vec3(ray6[0], ray6[1], ray6[2]), Vec3(ray6[3], ray6[4], ray6[5])), v, p, tt);
    if (ok) { *out_value = v; out_point3[0] = p.x; out_point3[1] = p.y; out_point3[2] = p.z; *out_t = tt; }
    return ok ? 1 : 0;
}
// Frustum::from_mat4 (frustum.rs:46-75): out 6 planes x {n xyz, d} in the order left,right,bottom,top,near,far.
int32_t orc_frustum_from_mat4(const float* m16, float* out24) {
    Frustum f;
    if (!Frustum::from_mat4(Mat4::from_cols_array(m16), f)) return 0;
    const Plane* ps[6] = {&f.left, &f.right, &f.bottom, &f.top, &f.near_, &f.far_};
    for (int i = 0; i < 6; ++i) { out24[4 * i] = ps[i]->n.x; out24[4 * i + 1] = ps[i]->n.y; out24[4 * i + 2] = ps[i]->n.z; out24[4 * i + 3] = ps[i]->d; }
    return 1;
}
void orc_perspective_rh_gl(float fov, float aspect, float zn, float zf, float* out16) { Mat4::perspective_rh_gl(fov, aspect, zn, zf).to_cols_array(out16); }
// FrustumVisitor query (visitor.rs:99-134). planes24 in the order left,right,bottom,top,near,far.
// Does any BRANCH bound on the path root -> leaf of `value` contain `origin` (min <= o <= max on every axis)? Then that
// branch's visitor distance is the ray's EXIT distance (aabb3.rs:165-195 returns the far slab point when the origin is
// inside), and RayClosestVisitor's `t < self.min` test (dbvt/util.rs:46-62) may prune it although it holds a closer leaf.
int32_t orc_dbvt_origin_inside_ancestor(orc_dbvt* t, uint32_t value, const float* origin3) {
    const Dbvt& T = t->tree;
    uint32_t leaf = 0;
    for (uint32_t i = 1; i < T.nodes.size(); ++i)
        if (T.nodes[i].kind == Dbvt::Leaf && T.nodes[i].value == value) { leaf = i; break; }
    if (!leaf) return -1;
    for (uint32_t i = T.nodes[leaf].parent; i != 0; i = T.nodes[i].parent) {
        const Aabb3& b = T.nodes[i].bound;
        if (origin3[0] >= b.min.x && origin3[0] <= b.max.x && origin3[1] >= b.min.y && origin3[1] <= b.max.y &&
            origin3[2] >= b.min.z && origin3[2] <= b.max.z) return 1;
    }
    return 0;
}

uint64_t orc_dbvt_query_frustum(orc_dbvt* t, const float* planes24, uint32_t* out_values, uint8_t* out_relation, uint64_t cap) {
    FrustumVisitor vis;
    Plane* ps[6] = {&vis.frustum.left, &vis.frustum.right, &vis.frustum.bottom, &vis.frustum.top, &vis.frustum.near_, &vis.frustum.far_};
    for (int i = 0; i < 6; ++i) *ps[i] = Plane(Vec3(planes24[4 * i], planes24[4 * i + 1], planes24[4 * i + 2]), planes24[4 * i + 3]);
    std::vector<std::pair<uint32_t, uint8_t>> h;
    t->tree.query_for_indices<FrustumVisitor, uint8_t>(vis, h);
    sort_hits(h);
    for (uint64_t i = 0; i < h.size() && i < cap; ++i) { out_values[i] = h[i].first; out_relation[i] = h[i].second; }
    return h.size();
}
// DbvtBroadPhase::find_collider_pairs (broad_phase/dbvt.rs:26-63).
uint64_t orc_dbvt_find_collider_pairs(orc_dbvt* t, const uint8_t* dirty, uint32_t* out_pairs, uint64_t cap) {
    auto pr = dbvt_find_collider_pairs(t->tree, dirty);
    for (uint64_t i = 0; i < pr.size() && i < cap; ++i) { out_pairs[2 * i] = pr[i].first; out_pairs[2 * i + 1] = pr[i].second; }
    return pr.size();
}
// A DynamicBoundingVolumeTree whose values are bounded by volume::Sphere (the tree is generic over the bound): the leaf predicate of
// the matching visitor applied to every value — the result set of query_for_indices, which does not depend on the tree's shape
// (branch bounds contain their leaves, the visitors are monotone). kind 0: DiscreteVisitor with a Sphere (query = centre, radius;
// sphere.rs:82-91), 1: ContinuousVisitor (query = ray origin, direction; :50-67; out_points), 2: FrustumVisitor (query = 6 planes
// n.xyz, d in the order left, right, bottom, top, near, far; :93-104 under frustum.rs:78-95; out_relation). Values ascending.
uint64_t orc_sphere_values_query(const float* spheres, uint64_t n, int32_t kind, const float* query, uint32_t* out_values, float* out_points,
                                 uint8_t* out_relation, uint64_t cap) {
    uint64_t hits = 0;
    Frustum fr;
    if (kind == 2) {
        Plane* ps[6] = {&fr.left, &fr.right, &fr.bottom, &fr.top, &fr.near_, &fr.far_};
        for (int i = 0; i < 6; ++i) *ps[i] = Plane(Vec3(query[4 * i], query[4 * i + 1], query[4 * i + 2]), query[4 * i + 3]);
    }
    for (uint64_t v = 0; v < n; ++v) {
        const BSphere s(Vec3(spheres[4 * v], spheres[4 * v + 1], spheres[4 * v + 2]), spheres[4 * v + 3]);
        bool hit = false;
        Vec3 p = Vec3::zero();
        Relation rel = In;
        if (kind == 0) {
            hit = s.intersects(BSphere(Vec3(query[0], query[1], query[2]), query[3]));
        } else if (kind == 1) {
            hit = s.intersection(Ray(Vec3(query[0], query[1], query[2]), Vec3(query[3], query[4], query[5])), p);
        } else {
            rel = fr.contains(s);
            hit = rel != Out;
        }
        if (!hit) continue;
        if (hits < cap) {
            out_values[hits] = (uint32_t)v;
            if (kind == 1 && out_points) { out_points[3 * hits] = p.x; out_points[3 * hits + 1] = p.y; out_points[3 * hits + 2] = p.z; }
            if (kind == 2 && out_relation) out_relation[hits] = (uint8_t)rel;
        }
        ++hits;
    }
    return hits;
}
// Ray / frustum vs one Aabb3 (aabb3.rs:165-257, frustum.rs:78-95) for direct leaf-level parity checks.
int32_t orc_ray_aabb(const float* ray6, const float* box, float* out_point3) {
    Vec3 p; bool ok = ray_aabb_intersection(Ray(Vec3(ray6[0], ray6[1], ray6[2]), Vec3(ray6[3], ray6[4], ray6[5])), box6(box), p);
    if (ok) { out_point3[0] = p.x; out_point3[1] = p.y; out_point3[2] = p.z; }
    return ok ? 1 : 0;
}

}  // extern "C"
