// ORACLE — TEST INFRASTRUCTURE ONLY (see glam_mini.hpp header). CPU restatement of the bam3d narrow phase,
// following the reference line by line (same operation order, f32 everywhere, no FMA, the reference's
// per-support-call Mat4::inverse(), Vec/SmallVec containers with the reference's remove/swap_remove order).
// Every function cites the reference file:line it restates (paths relative to /root/reference/src).
//
// Pinning: the reference cannot be compiled here (no cargo/rustc), so this restatement is pinned by the
// reference's own known-answer tests, ported bit-exactly in oracle/kat_main.cpp. Paths with NO reference test
// (rotated / mixed-shape GJK+EPA pairs, Particle/Primitive3 dispatch, intersection_complex*) are
// "parity unpinned": no reference value stands behind them. For GJK::intersection / GJK::distance on Sphere, Cuboid, Cylinder and
// Capsule pairs (rotated ones and the pinched EPA polytopes included) this restatement is additionally checked bit for bit against a
// second one written independently from the Rust source (tests/replay_gjk_epa.py, tests/test_host_cpu.py).
//
// Rust panics on this path (simplex3d.rs:134 unwrap, :138-142 assert!) are mapped to Status::WouldPanic; the
// uncapped TOI loop (gjk/mod.rs:166) is capped at `toi_iteration_cap` and reported as Status::MaxIter.
#pragma once
#include <cstdint>
#include <utility>
#include <map>
#include <vector>
#include "glam_mini.hpp"

namespace orc {

// ---------------------------------------------------------------------------------------------------------
// Primitives (primitive/*.rs). Kind numbering follows the order of `enum Primitive3` (primitive3.rs:16-33)
// with Cube folded into Cuboid (Cube forwards to its inner Cuboid, cuboid.rs:133-137).
// ---------------------------------------------------------------------------------------------------------
enum Kind : uint8_t { PARTICLE = 0, QUAD = 1, SPHERE = 2, CUBOID = 3, CYLINDER = 4, CAPSULE = 5, POLYHEDRON = 6 };

// ConvexPolyhedron::new_with_faces: the half-edge structure of build_half_edges (primitive/polyhedron.rs:246-340), built
// in the same order (face by face, edge pairs created on first sight, first-writer-wins for vertex.edge / edge.left_face /
// face.edge) because hill_climb_support_point and the ray walk depend on those links.
struct HalfEdgeMesh {
    struct Edge { uint32_t target_vertex, left_face, next_edge, previous_edge, twin_edge; bool ready; };
    struct Face { uint32_t edge; uint32_t v[3]; Vec3 n; float d; bool ready; };
    std::vector<uint32_t> vertex_edge;
    std::vector<Edge> edges;
    std::vector<Face> faces;
    // false if a face is degenerate (Plane::from_points(..).unwrap() would panic, polyhedron.rs:270-274)
    bool build(const float* verts, uint32_t nverts, const uint32_t* in_faces, uint32_t nfaces) {
        vertex_edge.assign(nverts, 0);
        std::vector<bool> vertex_ready(nverts, false);
        edges.clear(); faces.clear();
        std::map<std::pair<uint32_t, uint32_t>, uint32_t> edge_map;
        auto pos = [&](uint32_t i) { return Vec3(verts[3 * i], verts[3 * i + 1], verts[3 * i + 2]); };
        for (uint32_t f = 0; f < nfaces; ++f) {
            const uint32_t fv[3] = {in_faces[3 * f], in_faces[3 * f + 1], in_faces[3 * f + 2]};
            if (fv[0] >= nverts || fv[1] >= nverts || fv[2] >= nverts) return false;
            Face face; face.edge = 0; face.v[0] = fv[0]; face.v[1] = fv[1]; face.v[2] = fv[2]; face.ready = false;
            {   // Plane::from_points (plane.rs:69-86)
                Vec3 p1 = pos(fv[0]), v0 = pos(fv[1]) - p1, v1 = pos(fv[2]) - p1;
                Vec3 normal = cross(v0, v1);
                if (all_eq(normal, Vec3::zero())) return false;
                face.n = normalize(normal);
                face.d = -dot(p1, face.n);
            }
            const uint32_t face_index = (uint32_t)faces.size();
            uint32_t face_edge_indices[3] = {0, 0, 0};
            for (int j = 0; j < 3; ++j) {
                const int i = (j == 0) ? 2 : j - 1;
                const uint32_t v0 = fv[i], v1 = fv[j];
                uint32_t e01, e10;
                auto it = edge_map.find({v0, v1});
                if (it != edge_map.end()) { e01 = it->second; e10 = edges[e01].twin_edge; }
                else {
                    e01 = (uint32_t)edges.size(); e10 = e01 + 1;
                    edges.push_back(Edge{v1, 0, 0, 0, e10, false});
                    edges.push_back(Edge{v0, 0, 0, 0, e01, false});
                }
                edge_map[{v0, v1}] = e01;
                edge_map[{v1, v0}] = e10;
                if (!edges[e01].ready) { edges[e01].left_face = face_index; edges[e01].ready = true; }
                if (!vertex_ready[v0]) { vertex_edge[v0] = e01; vertex_ready[v0] = true; }
                if (!face.ready) { face.edge = e01; face.ready = true; }
                face_edge_indices[i] = e01;
            }
            faces.push_back(face);
            for (int j = 0; j < 3; ++j) {
                const int i = (j == 0) ? 2 : j - 1;
                edges[face_edge_indices[i]].next_edge = face_edge_indices[j];
                edges[face_edge_indices[j]].previous_edge = face_edge_indices[i];
            }
        }
        return true;
    }
};

struct Shape {
    uint8_t kind = PARTICLE;
    // SPHERE: p0 = radius. CUBOID: (p0,p1,p2) = half_dim. QUAD: (p0,p1) = half_dim.
    // CYLINDER / CAPSULE: p0 = half_height, p1 = radius.
    float p0 = 0.f, p1 = 0.f, p2 = 0.f;
    const float* verts = nullptr;  // POLYHEDRON: nverts x 3 floats (ConvexPolyhedron::new = VertexOnly mode)
    uint32_t nverts = 0;
    const HalfEdgeMesh* he = nullptr;  // POLYHEDRON built with new_with_faces (HalfEdge mode), else null

    static Shape sphere(float r) { Shape s; s.kind = SPHERE; s.p0 = r; return s; }
    // Cuboid::new(dim) -> half_dim = dim / 2 (cuboid.rs:27-34)
    static Shape cuboid(float dx, float dy, float dz) {
        Shape s; s.kind = CUBOID; Vec3 h = Vec3(dx, dy, dz) / 2.0f; s.p0 = h.x; s.p1 = h.y; s.p2 = h.z; return s;
    }
    static Shape cuboid_half(float hx, float hy, float hz) { Shape s; s.kind = CUBOID; s.p0 = hx; s.p1 = hy; s.p2 = hz; return s; }
    static Shape quad(float dx, float dy) { Shape s; s.kind = QUAD; s.p0 = dx / 2.0f; s.p1 = dy / 2.0f; return s; }
    static Shape cylinder(float hh, float r) { Shape s; s.kind = CYLINDER; s.p0 = hh; s.p1 = r; return s; }
    static Shape capsule(float hh, float r) { Shape s; s.kind = CAPSULE; s.p0 = hh; s.p1 = r; return s; }
    static Shape particle() { Shape s; s.kind = PARTICLE; return s; }
    static Shape polyhedron(const float* v, uint32_t n) { Shape s; s.kind = POLYHEDRON; s.verts = v; s.nverts = n; return s; }
};

struct Counters {
    uint64_t support_calls = 0;   // SupportPoint::from_minkowski calls
    uint64_t gjk_iterations = 0;  // simplex reductions
    uint64_t epa_iterations = 0;
    uint32_t max_faces = 0, max_verts = 0, max_edges = 0;
    uint64_t removed_faces = 0;   // faces Polytope::add removed (the serial part of an EPA iteration)
};

// primitive/util.rs:7-23 get_max_point — inverse() recomputed on every call, fold from (zero, -inf), strict >.
inline Vec3 get_max_point(const Vec3* verts, int n, Vec3 direction, const Mat4& transform) {
    Vec3 d = transform.inverse().transform_vector3(direction);
    Vec3 max_p = Vec3::zero();
    float max_dot = -std::numeric_limits<float>::infinity();
    for (int i = 0; i < n; ++i) {
        float dt = dot(verts[i], d);
        if (dt > max_dot) { max_p = verts[i]; max_dot = dt; }
    }
    return transform.transform_point3(max_p);
}

// Primitive::support_point for every primitive; dispatch as primitive3.rs:116-131.
inline Vec3 support_point(const Shape& s, Vec3 direction, const Mat4& transform) {
    switch (s.kind) {
        case PARTICLE:  // primitive3.rs:119
            return transform.transform_point3(Vec3::zero());
        case QUAD: {  // quad.rs:46-60, corners in the reference's order
            Vec3 c[4] = {Vec3(s.p0, s.p1, 0.f), Vec3(-s.p0, s.p1, 0.f), Vec3(-s.p0, -s.p1, 0.f), Vec3(s.p0, -s.p1, 0.f)};
            return get_max_point(c, 4, direction, transform);
        }
        case SPHERE: {  // sphere.rs:20-25
            Vec3 d = transform.inverse().transform_vector3(direction);
            return transform.transform_point3(d * (s.p0 / std::sqrt(dot(d, d))));
        }
        case CUBOID: {  // cuboid.rs:46-64, corners in the reference's order
            float hx = s.p0, hy = s.p1, hz = s.p2;
            Vec3 c[8] = {Vec3(hx, hy, hz),  Vec3(-hx, hy, hz),  Vec3(-hx, -hy, hz),  Vec3(hx, -hy, hz),
                         Vec3(hx, hy, -hz), Vec3(-hx, hy, -hz), Vec3(-hx, -hy, -hz), Vec3(hx, -hy, -hz)};
            return get_max_point(c, 8, direction, transform);
        }
        case CYLINDER: {  // cylinder.rs:37-62
            Vec3 d = transform.inverse().transform_vector3(direction);
            Vec3 result = d;
            bool negative = std::signbit(result.y);
            result.y = 0.f;
            if (std::fabs(dot(result, result) - 0.f) < F32_EPSILON) {
                result = Vec3::splat(0.f);
            } else {
                result = normalize(result);
                if (std::fabs(result.y - 0.f) < F32_EPSILON && std::fabs(result.x - 0.f) < F32_EPSILON &&
                    std::fabs(result.z - 0.f) < F32_EPSILON) {
                    result = Vec3::splat(0.f);
                } else {
                    result = result * s.p1;
                }
            }
            result.y = negative ? -s.p0 : s.p0;
            return transform.transform_point3(result);
        }
        case CAPSULE: {  // capsule.rs:40-49
            Vec3 d = transform.inverse().transform_vector3(direction);
            Vec3 result = Vec3::zero();
            result.y = rsignum(d.y) * s.p0;
            return transform.transform_point3(result + d * (s.p1 / std::sqrt(dot(d, d))));
        }
        case POLYHEDRON: {  // polyhedron.rs:135-150 (VertexOnly) / :153-183 (HalfEdge) + :342-355
            Vec3 d = transform.inverse().transform_vector3(direction);
            if (s.he) {  // hill_climb_support_point
                auto pos = [&](uint32_t i) { return Vec3(s.verts[3 * i], s.verts[3 * i + 1], s.verts[3 * i + 2]); };
                uint32_t best_index = 0;
                float best_dot = dot(pos(best_index), d);
                for (;;) {
                    const float previous_best_dot = best_dot;
                    uint32_t edge_index = s.he->vertex_edge[best_index];
                    const uint32_t start_edge_index = edge_index;
                    for (;;) {
                        const uint32_t vertex_index = s.he->edges[edge_index].target_vertex;
                        const float dt = dot(pos(vertex_index), d);
                        if (dt > best_dot) { best_index = vertex_index; best_dot = dt; }
                        edge_index = s.he->edges[s.he->edges[edge_index].twin_edge].next_edge;
                        if (start_edge_index == edge_index) break;
                    }
                    if (std::fabs(best_dot - previous_best_dot) < F32_EPSILON) break;
                }
                return transform.transform_point3(pos(best_index));
            }
            Vec3 max_p = Vec3::zero();
            float max_dot = -std::numeric_limits<float>::infinity();
            for (uint32_t i = 0; i < s.nverts; ++i) {
                Vec3 v(s.verts[3 * i], s.verts[3 * i + 1], s.verts[3 * i + 2]);
                float dt = dot(v, d);
                if (dt > max_dot) { max_p = v; max_dot = dt; }
            }
            return transform.transform_point3(max_p);
        }
    }
    return Vec3::zero();
}

// ---------------------------------------------------------------------------------------------------------
// algorithm/minkowski/mod.rs:16-70
// ---------------------------------------------------------------------------------------------------------
struct SupportPoint {
    Vec3 v, sup_a, sup_b;
    // mod.rs:33-51
    static SupportPoint from_minkowski(const Shape& left, const Mat4& lt, const Shape& right, const Mat4& rt,
                                       Vec3 direction, Counters* c = nullptr) {
        if (c) c->support_calls++;
        Vec3 l = support_point(left, direction, lt);
        Vec3 r = support_point(right, -direction, rt);
        SupportPoint p; p.v = l - r; p.sup_a = l; p.sup_b = r;
        return p;
    }
    // mod.rs:60-70  (Sub<Vec3>: shifts only v)
    SupportPoint sub(Vec3 rhs) const { SupportPoint p = *this; p.v = v - rhs; return p; }
};
inline SupportPoint sup_from_v(float x, float y, float z) { SupportPoint s; s.v = Vec3(x, y, z); return s; }

// Simplex = SmallVec<[SupportPoint; 5]> (gjk/simplex/mod.rs:10). Rust Vec::remove shifts the tail down.
struct Simplex {
    SupportPoint p[5];
    int n = 0;
    int len() const { return n; }
    void push(const SupportPoint& s) { p[n++] = s; }
    void remove(int i) { for (int k = i; k + 1 < n; ++k) p[k] = p[k + 1]; --n; }
    void swap(int i, int j) { SupportPoint t = p[i]; p[i] = p[j]; p[j] = t; }
    void clear() { n = 0; }
    SupportPoint& operator[](int i) { return p[i]; }
    const SupportPoint& operator[](int i) const { return p[i]; }
};

enum class Status : uint8_t { Ok = 0, Miss = 1, MaxIter = 2, WouldPanic = 3, Nan = 4, Capacity = 5 };

// ---------------------------------------------------------------------------------------------------------
// primitive/util.rs
// ---------------------------------------------------------------------------------------------------------
// util.rs:44-59 barycentric_vector
inline void barycentric_vector(Vec3 p, Vec3 a, Vec3 b, Vec3 c, float& u, float& v, float& w) {
    Vec3 v0 = b - a, v1 = c - a, v2 = p - a;
    float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    float inv_denom = 1.0f / (d00 * d11 - d01 * d01);
    v = (d11 * d20 - d01 * d21) * inv_denom;
    w = (d00 * d21 - d01 * d20) * inv_denom;
    u = 1.0f - v - w;
}
// util.rs:81-93 get_closest_point_on_edge
inline Vec3 get_closest_point_on_edge(Vec3 start, Vec3 end, Vec3 point) {
    Vec3 line = end - start;
    Vec3 line_dir = normalize(line);
    Vec3 v = point - start;
    float d = dot(v, line_dir);
    if (d < 0.f) return start;
    else if ((d * d) > dot(line, line)) return end;
    else return start + line_dir * d;
}

// ---------------------------------------------------------------------------------------------------------
// gjk/simplex/simplex3d.rs
// ---------------------------------------------------------------------------------------------------------
inline Vec3 cross_aba(Vec3 a, Vec3 b) { return cross(cross(a, b), a); }  // simplex3d.rs:157-159

// simplex3d.rs:163-201
inline void check_side(Vec3 abc, Vec3 ab, Vec3 ac, Vec3 ao, Simplex& simplex, Vec3& v, bool above, bool ignore_ab) {
    Vec3 ab_perp = cross(ab, abc);
    if (!ignore_ab && dot(ab_perp, ao) > 0.f) {
        simplex.remove(0);
        v = cross_aba(ab, ao);
        return;
    }
    Vec3 ac_perp = cross(abc, ac);
    if (dot(ac_perp, ao) > 0.f) {
        simplex.remove(1);
        v = cross_aba(ac, ao);
        return;
    }
    if (above || dot(abc, ao) > 0.f) {
        v = abc;
    } else {
        simplex.swap(0, 1);
        v = -abc;
    }
}

// simplex3d.rs:16-90
inline bool reduce_to_closest_feature(Simplex& simplex, Vec3& v) {
    if (simplex.len() == 4) {
        Vec3 a = simplex[3].v, b = simplex[2].v, c = simplex[1].v, d = simplex[0].v;
        Vec3 ao = -a, ab = b - a, ac = c - a;
        Vec3 abc = cross(ab, ac);
        if (dot(abc, ao) > 0.f) {
            simplex.remove(0);
            check_side(abc, ab, ac, ao, simplex, v, true, false);
        } else {
            Vec3 ad = d - a;
            Vec3 acd = cross(ac, ad);
            if (dot(acd, ao) > 0.f) {
                simplex.remove(2);
                check_side(acd, ac, ad, ao, simplex, v, true, true);
            } else {
                Vec3 adb = cross(ad, ab);
                if (dot(adb, ao) > 0.f) {
                    simplex.remove(1);
                    simplex.swap(0, 1);
                    v = adb;
                } else {
                    return true;
                }
            }
        }
    } else if (simplex.len() == 3) {
        Vec3 a = simplex[2].v, b = simplex[1].v, c = simplex[0].v;
        Vec3 ao = -a, ab = b - a, ac = c - a;
        check_side(cross(ab, ac), ab, ac, ao, simplex, v, false, false);
    } else if (simplex.len() == 2) {
        Vec3 a = simplex[1].v, b = simplex[0].v;
        Vec3 ao = -a, ab = b - a;
        v = cross_aba(ab, ao);
        if (all_eq(v, Vec3::zero())) v.x = 0.1f;
    }
    return false;
}

inline bool in_range(float v) { return v >= 0.f && v <= 1.f; }  // simplex3d.rs:152-154

// simplex3d.rs:122-149 with plane.rs:69-86 (Plane::from_points) and plane.rs:117-129 (Continuous<Ray> for Plane)
inline Vec3 get_closest_point_on_face(Vec3 a, Vec3 b, Vec3 c, Vec3 normal, Status& st) {
    Vec3 ray_origin = Vec3::zero();
    Vec3 ray_dir = -normal;
    // Plane::from_points(ap, bp, cp)
    Vec3 v0 = b - a, v1 = c - a;
    Vec3 pn = cross(v0, v1);
    if (all_eq(pn, Vec3::zero())) { st = Status::WouldPanic; return Vec3::zero(); }  // .unwrap() on None
    Vec3 n = normalize(pn);
    float d = -dot(a, n);
    // plane.intersection(&ray)
    float t = -(d + dot(ray_origin, n)) / dot(ray_dir, n);
    if (t < 0.0f) return Vec3::zero();  // None => Vec3::zero()
    Vec3 point = ray_origin + ray_dir * t;
    float u, v, w;
    barycentric_vector(point, a, b, c, u, v, w);
    if (!(in_range(u) && in_range(v) && in_range(w))) { st = Status::WouldPanic; return point; }  // assert!
    return point;
}

// simplex3d.rs:95-113
inline Vec3 get_closest_point_to_origin(Simplex& simplex, Status& st) {
    Vec3 d = Vec3::zero();
    if (reduce_to_closest_feature(simplex, d)) return d;
    if (simplex.len() == 1) return simplex[0].v;
    else if (simplex.len() == 2) return get_closest_point_on_edge(simplex[1].v, simplex[0].v, Vec3::zero());
    else return get_closest_point_on_face(simplex[2].v, simplex[1].v, simplex[0].v, d, st);
}

// ---------------------------------------------------------------------------------------------------------
// contact.rs
// ---------------------------------------------------------------------------------------------------------
enum class CollisionStrategy : uint8_t { FullResolution = 0, CollisionOnly = 1 };  // contact.rs:12-18
struct Contact {  // contact.rs:22-38
    CollisionStrategy strategy = CollisionStrategy::FullResolution;
    Vec3 normal, contact_point;
    float penetration_depth = 0.f, time_of_impact = 0.f;
};

// ---------------------------------------------------------------------------------------------------------
// epa/epa3d.rs
// ---------------------------------------------------------------------------------------------------------
struct Face {  // epa3d.rs:152-157
    uint32_t vertices[3];
    Vec3 normal;
    float distance;
    // epa3d.rs:160-170
    static Face new_impl(const std::vector<SupportPoint>& s, uint32_t a, uint32_t b, uint32_t c) {
        Vec3 ab = s[b].v - s[a].v;
        Vec3 ac = s[c].v - s[a].v;
        Face f;
        f.normal = normalize(cross(ab, ac));
        f.distance = dot(f.normal, s[a].v);
        f.vertices[0] = a; f.vertices[1] = b; f.vertices[2] = c;
        return f;
    }
};

// epa3d.rs:183-190
inline void remove_or_add_edge(std::vector<std::pair<uint32_t, uint32_t>>& edges, std::pair<uint32_t, uint32_t> edge) {
    for (size_t i = 0; i < edges.size(); ++i) {
        if (edge.first == edges[i].second && edge.second == edges[i].first) {
            edges.erase(edges.begin() + i);  // Vec::remove (order preserving)
            return;
        }
    }
    edges.push_back(edge);
}

// The reference's Vecs are unbounded. On near-degenerate input (a Cylinder cap against a flat face: the cap's rim supports are
// coplanar, visibility becomes rounding noise) its polytope pinches and the face list grows multiplicatively per iteration: 44 of
// 2^20 mixed pairs pass 1024 faces, the largest reaches 121 962 — and the reference still returns a Contact after at most
// max_iterations trips. The limit below is the capacity of the device's largest tier (HUGE_FCAP, bam3d_b200/csrc/epa_huge.cuh);
// a pair that would pass it ends as Status::Capacity on both sides (none is known). EPA3::face_limit can be set per run.
constexpr size_t EPA_FACE_LIMIT = (size_t)1 << 18;

struct Polytope {  // epa3d.rs:97-150
    std::vector<SupportPoint>& vertices;
    std::vector<Face> faces;
    uint32_t max_edges = 0;
    uint64_t removed = 0;
    bool overflow = false;
    size_t face_limit = EPA_FACE_LIMIT;  // the device's capacity by default; tests lift it to see what the reference itself returns
    explicit Polytope(std::vector<SupportPoint>& simplex) : vertices(simplex) {
        // Face::new, epa3d.rs:172-179
        faces.push_back(Face::new_impl(simplex, 3, 2, 1));
        faces.push_back(Face::new_impl(simplex, 3, 1, 0));
        faces.push_back(Face::new_impl(simplex, 3, 0, 2));
        faces.push_back(Face::new_impl(simplex, 2, 0, 1));
    }
    // epa3d.rs:111-119 — strict <, first minimum wins
    const Face& closest_face_to_origin() const {
        const Face* face = &faces[0];
        for (size_t i = 1; i < faces.size(); ++i)
            if (faces[i].distance < face->distance) face = &faces[i];
        return *face;
    }
    // epa3d.rs:121-149
    void add(const SupportPoint& sup) {
        std::vector<std::pair<uint32_t, uint32_t>> edges;
        size_t i = 0;
        while (i < faces.size()) {
            float dt = dot(faces[i].normal, sup.v - vertices[faces[i].vertices[0]].v);
            if (dt > 0.f) {
                Face face = faces[i];
                faces[i] = faces.back();  // swap_remove
                faces.pop_back();
                ++removed;
                for (int e = 0; e < 3; ++e) {
                    remove_or_add_edge(edges, {face.vertices[e], face.vertices[(e + 1) % 3]});
                }
                if (edges.size() > max_edges) max_edges = (uint32_t)edges.size();
            } else {
                ++i;
            }
        }
        if (faces.size() + edges.size() > face_limit) { overflow = true; return; }
        uint32_t n = (uint32_t)vertices.size();
        vertices.push_back(sup);
        for (auto& e : edges) faces.push_back(Face::new_impl(vertices, n, e.first, e.second));
    }

    // The same Polytope::add written the way the device's large-polytope tier evaluates it (bam3d_b200/csrc/epa_huge.cuh): no
    // sequential walk, only prefix sums, scatters and short per-key scans — each phase is data parallel. Executed sequentially
    // here; tests/test_host_cpu.py checks that it builds the same face list, in the same order, as add() above.
    //   visibility flags -> exclusive prefix counts -> S = survivors. The walk `while i < len { if visible(i) swap_remove(i) else
    //   i++ }` is a two-pointer process: holes (visible positions below S, ascending: h_1 < h_2 < ...) are filled by the tail
    //   survivors (non-visible positions >= S) taken from the back (t_1 > t_2 > ...), and the visible faces are REMOVED in the
    //   order h_1, [visible positions in (t_1, F) descending], h_2, [visible in (t_2, t_1) descending], ..., and finally, if
    //   positions remain between S and the last tail survivor taken, S itself followed by the rest of them descending.
    //   The horizon list is the sequence of directed edges of the removed faces in that order under remove_or_add_edge; per
    //   undirected edge that is a first-in-first-out matching of the two directions, whose survivors have a closed form (below),
    //   compacted in sequence order.
    void add_in_phases(const SupportPoint& sup) {
        const size_t F = faces.size();
        std::vector<uint8_t> vis(F);
        for (size_t i = 0; i < F; ++i) vis[i] = dot(faces[i].normal, sup.v - vertices[faces[i].vertices[0]].v) > 0.f ? 1 : 0;
        std::vector<uint32_t> before(F + 1, 0);  // visible positions < p
        for (size_t i = 0; i < F; ++i) before[i + 1] = before[i] + vis[i];
        const uint32_t V = before[F];
        const size_t S = F - V;
        auto vis_after = [&](size_t p) { return V - before[p] - vis[p]; };           // visible positions > p
        auto surv_after = [&](size_t p) { return (uint32_t)(F - 1 - p) - vis_after(p); };  // survivors at positions > p
        const uint32_t Hn = before[S];  // holes below S == tail survivors
        std::vector<uint32_t> tsrc(Hn + 2, (uint32_t)F);  // tsrc[k] = position of the k-th tail survivor from the back; tsrc[0] = F
        for (size_t p = S; p < F; ++p) if (!vis[p]) tsrc[surv_after(p) + 1] = (uint32_t)p;
        const uint32_t t_last = Hn ? tsrc[Hn] : (uint32_t)F;
        // removal order
        std::vector<uint32_t> order(V);
        for (size_t p = 0; p < F; ++p) {
            if (!vis[p]) continue;
            uint32_t idx;
            if (p < S) {  // hole h_k
                const uint32_t k = before[p] + 1;
                idx = (k - 1) + (k >= 2 ? vis_after(tsrc[k - 1]) : 0u);
            } else {
                const uint32_t s = surv_after(p);
                if (s < Hn) idx = (s + 1) + vis_after(p);                  // pulled in while hole h_(s+1) was being filled
                else if (p == S) idx = Hn + vis_after(t_last == F ? F - 1 : t_last) + (t_last == F ? 0u : 0u);
                else idx = Hn + 1 + vis_after(p);
            }
            order[idx] = (uint32_t)p;
        }
        // directed edge events in removal order, then the first-in-first-out matching per undirected edge
        struct Ev { uint32_t a, b; };
        std::vector<Ev> ev(3 * (size_t)V);
        for (uint32_t i = 0; i < V; ++i) {
            const Face& f = faces[order[i]];
            for (int e = 0; e < 3; ++e) ev[3 * (size_t)i + e] = {f.vertices[e], f.vertices[(e + 1) % 3]};
        }
        // remove_or_add_edge over that sequence: per undirected edge the list only ever holds copies of one direction (an arriving
        // edge removes the OLDEST copy of its reverse if there is one, else it is pushed), so what is left are the last
        // |#forward - #backward| events of the majority direction: event i survives iff its rank among the events with the same
        // directed edge, in sequence order, is >= the number of events of the reverse direction.
        std::vector<uint8_t> keep(ev.size(), 0);
        std::map<std::pair<uint32_t, uint32_t>, uint32_t> count, seen;
        for (auto& e : ev) ++count[{e.a, e.b}];
        for (uint32_t i = 0; i < ev.size(); ++i) {
            const uint32_t n_rev = count.count({ev[i].b, ev[i].a}) ? count[{ev[i].b, ev[i].a}] : 0u;
            const uint32_t rank = seen[{ev[i].a, ev[i].b}]++;
            keep[i] = (count[{ev[i].a, ev[i].b}] > n_rev && rank >= n_rev) ? 1 : 0;
        }
#ifdef ORC_RANK_STATS
        {   // directed edges whose events need a rank (both directions present, this one the majority), and the events they hold
            uint32_t keys = 0, events = 0;
            for (auto& kv : count) {
                auto r = count.find({kv.first.second, kv.first.first});
                if (r != count.end() && kv.second > r->second) { ++keys; events += kv.second; }
            }
            fprintf(stderr, "[rank] F=%zu V=%u keys=%u events=%u\n", F, V, keys, events);
        }
#endif
        std::vector<std::pair<uint32_t, uint32_t>> edges;
        for (uint32_t i = 0; i < ev.size(); ++i) if (keep[i]) edges.emplace_back(ev[i].a, ev[i].b);
        if (edges.size() > max_edges) max_edges = (uint32_t)edges.size();
        removed += V;
        if (S + edges.size() > face_limit) { overflow = true; return; }
        // survivors: positions below S stay, holes take the tail survivors from the back
        std::vector<Face> next(S);
        for (size_t p = 0; p < S; ++p) next[p] = vis[p] ? faces[tsrc[before[p] + 1]] : faces[p];
        faces.swap(next);
        uint32_t n = (uint32_t)vertices.size();
        vertices.push_back(sup);
        for (auto& e : edges) faces.push_back(Face::new_impl(vertices, n, e.first, e.second));
    }
};

// epa3d.rs:83-94 point(); :70-77 contact()
inline Contact epa_contact(const Polytope& polytope, const Face& face) {
    float u, v, w;
    barycentric_vector(face.normal * face.distance, polytope.vertices[face.vertices[0]].v,
                       polytope.vertices[face.vertices[1]].v, polytope.vertices[face.vertices[2]].v, u, v, w);
    Contact c;
    c.strategy = CollisionStrategy::FullResolution;
    c.normal = face.normal;
    c.penetration_depth = face.distance;
    c.contact_point = polytope.vertices[face.vertices[0]].sup_a * u + polytope.vertices[face.vertices[1]].sup_a * v +
                      polytope.vertices[face.vertices[2]].sup_a * w;
    c.time_of_impact = 0.f;
    return c;
}

struct EPA3 {  // epa3d.rs:9-67; constants epa/mod.rs:12-13
    float tolerance = 0.00001f;
    uint32_t max_iterations = 100;
    size_t face_limit = EPA_FACE_LIMIT;  // not in the reference (its Vec is unbounded): see the comment at EPA_FACE_LIMIT
    bool in_phases = false;              // Polytope::add_in_phases instead of add (the device's data-parallel formulation)
    // epa3d.rs:16-52
    bool process(std::vector<SupportPoint>& simplex, const Shape& left, const Mat4& lt, const Shape& right,
                 const Mat4& rt, Contact& out, Counters* cnt = nullptr, Status* st = nullptr) const {
        if (simplex.size() < 4) return false;
        Polytope polytope(simplex);
        polytope.face_limit = face_limit;
        uint32_t i = 1;
        for (;;) {
            const Face& face = polytope.closest_face_to_origin();
            SupportPoint p = SupportPoint::from_minkowski(left, lt, right, rt, face.normal, cnt);
            float d = dot(p.v, face.normal);
            if (cnt) cnt->epa_iterations++;
            if (d - face.distance < tolerance || i >= max_iterations) {
                out = epa_contact(polytope, face);
                if (cnt) {
                    if (polytope.faces.size() > cnt->max_faces) cnt->max_faces = (uint32_t)polytope.faces.size();
                    if (polytope.vertices.size() > cnt->max_verts) cnt->max_verts = (uint32_t)polytope.vertices.size();
                    if (polytope.max_edges > cnt->max_edges) cnt->max_edges = polytope.max_edges;
                    cnt->removed_faces += polytope.removed;
                }
                return true;
            }
            if (in_phases) polytope.add_in_phases(p); else polytope.add(p);
            if (polytope.overflow) {
                if (st) *st = Status::Capacity;
                if (cnt) cnt->max_faces = (uint32_t)face_limit + 1;
                return false;
            }
            if (cnt && polytope.faces.size() > cnt->max_faces) cnt->max_faces = (uint32_t)polytope.faces.size();
            i += 1;
        }
    }
};

// ---------------------------------------------------------------------------------------------------------
// traits.rs:180-190 TranslationInterpolate for Mat4
// ---------------------------------------------------------------------------------------------------------
inline Mat4 translation_interpolate(const Mat4& self, const Mat4& other, float amount) {
    Vec3 ss, st, os, ot; Quat sr, orot;
    self.to_scale_rotation_translation(ss, sr, st);
    other.to_scale_rotation_translation(os, orot, ot);
    return Mat4::from_scale_rotation_translation(os, orot, lerp(st, ot, amount));
}

// ---------------------------------------------------------------------------------------------------------
// gjk/mod.rs
// ---------------------------------------------------------------------------------------------------------
struct GJK {
    EPA3 epa;
    float distance_tolerance = 0.000001f;    // gjk/mod.rs:19
    float continuous_tolerance = 0.000001f;  // gjk/mod.rs:20
    uint32_t max_iterations = 100;           // gjk/mod.rs:18
    uint32_t toi_iteration_cap = 100;        // NOT in the reference (its TOI loop is uncapped, gjk/mod.rs:166)

    GJK() {}
    // gjk/mod.rs:45-58 (max_iterations shared with EPA)
    GJK(float dist_tol, float cont_tol, float epa_tol, uint32_t max_it)
        : distance_tolerance(dist_tol), continuous_tolerance(cont_tol), max_iterations(max_it) {
        epa.tolerance = epa_tol; epa.max_iterations = max_it;
    }

    // gjk/mod.rs:74-113. Returns true + the enclosing simplex, or false (miss / not converged).
    bool intersect(const Shape& left, const Mat4& lt, const Shape& right, const Mat4& rt, Simplex& simplex,
                   Status* st = nullptr, Counters* cnt = nullptr) const {
        Vec3 left_pos = lt.transform_point3(Vec3::zero());
        Vec3 right_pos = rt.transform_point3(Vec3::zero());
        Vec3 d = right_pos - left_pos;
        if (all_eq(d, Vec3::zero())) d = Vec3::splat(1.f);
        SupportPoint a = SupportPoint::from_minkowski(left, lt, right, rt, d, cnt);
        if (st) *st = Status::Miss;
        if (dot(a.v, d) <= 0.f) return false;
        simplex.clear();
        simplex.push(a);
        d = -d;
        for (uint32_t it = 0; it < max_iterations; ++it) {
            SupportPoint a2 = SupportPoint::from_minkowski(left, lt, right, rt, d, cnt);
            if (dot(a2.v, d) <= 0.f) {
                return false;
            } else {
                simplex.push(a2);
                if (cnt) cnt->gjk_iterations++;
                if (reduce_to_closest_feature(simplex, d)) {
                    if (st) *st = Status::Ok;
                    return true;
                }
            }
        }
        if (st) *st = Status::MaxIter;
        return false;
    }

    // gjk/mod.rs:317-341
    bool intersection(CollisionStrategy strategy, const Shape& left, const Mat4& lt, const Shape& right,
                      const Mat4& rt, Contact& out, Status* st = nullptr, Counters* cnt = nullptr) const {
        Simplex simplex;
        if (!intersect(left, lt, right, rt, simplex, st, cnt)) return false;
        if (strategy == CollisionStrategy::CollisionOnly) {
            out = Contact();  // Contact::new(CollisionOnly): zero normal/depth/point/toi
            out.strategy = CollisionStrategy::CollisionOnly;
            return true;
        }
        std::vector<SupportPoint> sv(simplex.p, simplex.p + simplex.n);  // simplex.into_vec()
        return epa.process(sv, left, lt, right, rt, out, cnt, st);        // get_contact_manifold, :285-299
    }

    // gjk/mod.rs:362-407 intersection_complex. Shapes are lists of (primitive, local transform). Returns true + contact.
    // `st` = WouldPanic where `partial_cmp(..).unwrap()` meets a NaN depth (:403-405).
    bool intersection_complex(CollisionStrategy strategy, const std::vector<std::pair<Shape, Mat4>>& left, const Mat4& lt,
                              const std::vector<std::pair<Shape, Mat4>>& right, const Mat4& rt, Contact& out, Status& st) const {
        st = Status::Miss;
        std::vector<Contact> contacts;
        bool undecided = false;  // a primitive pair stopped by the face limit (not in the reference): its contact is unknown, so is the maximum
        for (auto& l : left) {
            Mat4 ltf = lt.mul_mat4(l.second);
            for (auto& r : right) {
                Mat4 rtf = rt.mul_mat4(r.second);
                Contact c; Status s1;
                if (intersection(strategy, l.first, ltf, r.first, rtf, c, &s1)) {
                    if (strategy == CollisionStrategy::CollisionOnly) { out = c; st = Status::Ok; return true; }
                    contacts.push_back(c);
                } else if (s1 == Status::Capacity) undecided = true;
            }
        }
        if (undecided) { st = Status::Capacity; return false; }
        if (contacts.empty()) return false;
        // Iterator::max_by: the LAST of several equal maxima is returned; unwrap() panics on NaN
        size_t best = 0;
        for (size_t i = 1; i < contacts.size(); ++i) {
            float a = contacts[best].penetration_depth, b = contacts[i].penetration_depth;
            if (a != a || b != b) { st = Status::WouldPanic; return false; }
            if (!(a > b)) best = i;  // max_by keeps the later element unless the earlier compares Greater
        }
        if (contacts.size() == 1 && contacts[0].penetration_depth != contacts[0].penetration_depth) { /* single element: no comparison */ }
        out = contacts[best]; st = Status::Ok;
        return true;
    }

    // gjk/mod.rs:422-456 distance_complex: None as soon as one primitive pair returns None; else the f32::min of all.
    bool distance_complex(const std::vector<std::pair<Shape, Mat4>>& left, const Mat4& lt, const std::vector<std::pair<Shape, Mat4>>& right,
                          const Mat4& rt, float& out, Status& st) const {
        bool have = false; float mn = 0.f;
        for (auto& l : left) {
            Mat4 ltf = lt.mul_mat4(l.second);
            for (auto& r : right) {
                Mat4 rtf = rt.mul_mat4(r.second);
                float rv, geo; Status s1;
                if (!distance(l.first, ltf, r.first, rtf, rv, geo, s1)) { st = s1 == Status::WouldPanic ? s1 : Status::Miss; return false; }
                mn = have ? rmin(rv, mn) : rv;
                have = true;
            }
        }
        st = have ? Status::Ok : Status::Miss;
        out = mn;
        return have;
    }

    // gjk/mod.rs:478-523 intersection_complex_time_of_impact: CollisionOnly -> first hit; else min_by time_of_impact
    // (the FIRST of several equal minima; partial_cmp(..).unwrap_or(Equal)).
    bool intersection_complex_time_of_impact(CollisionStrategy strategy, const std::vector<std::pair<Shape, Mat4>>& left, const Mat4& lt0,
                                             const Mat4& lt1, const std::vector<std::pair<Shape, Mat4>>& right, const Mat4& rt0, const Mat4& rt1,
                                             Contact& out, Status& st) const {
        st = Status::Miss;
        std::vector<Contact> contacts;
        bool undecided = false;  // a primitive pair stopped by the added iteration cap: the reference's loop has none, its outcome is unknown
        for (auto& l : left) {
            Mat4 l0 = lt0.mul_mat4(l.second), l1 = lt1.mul_mat4(l.second);
            for (auto& r : right) {
                Mat4 r0 = rt0.mul_mat4(r.second), r1 = rt1.mul_mat4(r.second);
                Contact c; Status s1;
                if (intersection_time_of_impact(l.first, l0, l1, r.first, r0, r1, c, s1)) {
                    if (strategy == CollisionStrategy::CollisionOnly) {
                        if (undecided) { st = Status::MaxIter; return false; }
                        c.strategy = CollisionStrategy::CollisionOnly; out = c; st = Status::Ok; return true;
                    }
                    contacts.push_back(c);
                } else if (s1 == Status::WouldPanic) { st = s1; return false; }
                else if (s1 == Status::MaxIter) undecided = true;
            }
        }
        if (undecided) { st = Status::MaxIter; return false; }
        if (contacts.empty()) return false;
        size_t best = 0;
        for (size_t i = 1; i < contacts.size(); ++i)
            if (contacts[i].time_of_impact < contacts[best].time_of_impact) best = i;  // min_by keeps the first unless a later one is Less
        out = contacts[best]; st = Status::Ok;
        return true;
    }

    // gjk/mod.rs:232-280. Returns Some/None; `ref_value` is what the reference returns (dp - d0, the
    // convergence residual); `geometric` is sqrt(d.d) (the commented-out return at :274).
    bool distance(const Shape& left, const Mat4& lt, const Shape& right, const Mat4& rt, float& ref_value,
                  float& geometric, Status& st, Counters* cnt = nullptr) const {
        st = Status::Ok;
        Vec3 right_pos = rt.transform_point3(Vec3::zero());
        Vec3 left_pos = lt.transform_point3(Vec3::zero());
        Simplex simplex;
        Vec3 d = right_pos - left_pos;
        if (std::fabs(d.x - 0.f) < F32_EPSILON && std::fabs(d.y - 0.f) < F32_EPSILON && std::fabs(d.z - 0.f) < F32_EPSILON)
            d = Vec3::splat(1.f);
        simplex.push(SupportPoint::from_minkowski(left, lt, right, rt, d, cnt));
        simplex.push(SupportPoint::from_minkowski(left, lt, right, rt, -d, cnt));
        for (uint32_t it = 0; it < max_iterations; ++it) {
            Vec3 dd = get_closest_point_to_origin(simplex, st);
            if (cnt) cnt->gjk_iterations++;
            if (st == Status::WouldPanic) return false;
            if (std::fabs(dd.x - 0.f) < F32_EPSILON && std::fabs(dd.y - 0.f) < F32_EPSILON && std::fabs(dd.z - 0.f) < F32_EPSILON) {
                st = Status::Miss;  // "colliding" -> None
                return false;
            }
            Vec3 nd = -dd;
            SupportPoint p = SupportPoint::from_minkowski(left, lt, right, rt, nd, cnt);
            float dp = dot(p.v, nd);
            float d0 = dot(simplex[0].v, nd);
            if (dp - d0 < distance_tolerance) {
                ref_value = dp - d0;
                geometric = std::sqrt(dot(nd, nd));
                st = Status::Ok;
                return true;
            }
            simplex.push(p);
        }
        st = Status::MaxIter;
        return false;
    }

    // gjk/mod.rs:130-217. (start, end) transforms per side.
    bool intersection_time_of_impact(const Shape& left, const Mat4& lt0, const Mat4& lt1, const Shape& right,
                                     const Mat4& rt0, const Mat4& rt1, Contact& out, Status& st,
                                     Counters* cnt = nullptr) const {
        st = Status::Ok;
        Vec3 left_lin_vel = lt1.transform_point3(Vec3::zero()) - lt0.transform_point3(Vec3::zero());
        Vec3 right_lin_vel = rt1.transform_point3(Vec3::zero()) - rt0.transform_point3(Vec3::zero());
        Vec3 ray = right_lin_vel - left_lin_vel;
        float lambda = 0.f;
        Vec3 normal = Vec3::zero();
        Vec3 ray_origin = Vec3::zero();
        Simplex simplex;
        SupportPoint p = SupportPoint::from_minkowski(left, lt0, right, rt0, -ray, cnt);
        Vec3 v = p.v;
        uint32_t iters = 0;
        while (dot(v, v) > continuous_tolerance) {
            if (iters++ >= toi_iteration_cap) { st = Status::MaxIter; return false; }  // cap: not in reference
            SupportPoint p2 = SupportPoint::from_minkowski(left, lt0, right, rt0, -v, cnt);
            float vp = dot(v, p2.v);
            float vr = dot(v, ray);
            if (vp > lambda * vr) {
                if (vr > 0.f) {
                    lambda = vp / vr;
                    if (lambda > 1.f) { st = Status::Miss; return false; }
                    ray_origin = ray * lambda;
                    simplex.clear();
                    normal = -v;
                } else {
                    st = Status::Miss;
                    return false;
                }
            }
            if (simplex.len() >= 5) { st = Status::WouldPanic; return false; }  // cannot happen; SmallVec would spill
            simplex.push(p2.sub(ray_origin));
            v = get_closest_point_to_origin(simplex, st);
            if (cnt) cnt->gjk_iterations++;
            if (st == Status::WouldPanic) return false;
        }
        if (dot(v, v) <= continuous_tolerance) {
            Mat4 transform = translation_interpolate(rt0, rt1, lambda);
            out.strategy = CollisionStrategy::FullResolution;
            out.normal = -normalize(normal);
            out.penetration_depth = std::sqrt(dot(v, v));
            out.contact_point = transform.transform_point3(ray_origin);
            out.time_of_impact = lambda;
            return true;
        }
        st = Status::Nan;  // v.v is NaN: loop exits and the <= test fails -> None
        return false;
    }
};

}  // namespace orc
