// The reference's GJK / EPA / SweepAndPrune unit tests (gjk/mod.rs:552-644, quad.rs:113-124), rewritten against the C++
// host layer (include/bam3d.hpp) so that they read like the crate's own tests. Needs a GPU. Exit code 0 = all passed.
#include <cmath>
#include <cstdio>
#include "../../include/bam3d.hpp"

using namespace bam3d;
static int fails = 0;
#define CHECK(c) do { if (!(c)) { ++fails; std::printf("FAIL line %d: %s\n", __LINE__, #c); } } while (0)

static Mat4 transform_3d(float x, float y, float z) { return mat4_from_translation(x, y, z); }  // angle_z = 0 in every KAT used

int main() {
    Context ctx(0);
    GJK gjk(ctx);
    {   // test_gjk_exact_3d
        auto shape = Primitive3::Cuboid(1, 1, 1);
        auto t = transform_3d(0, 0, 0);
        CHECK(gjk.intersection(CollisionStrategy::FullResolution, shape, t, shape, t).has_value());
        CHECK(!gjk.distance(shape, t, shape, t).has_value());
    }
    {   // test_gjk_sphere
        auto shape = Primitive3::Sphere(1);
        auto t = transform_3d(0, 0, 0);
        CHECK(gjk.intersection(CollisionStrategy::FullResolution, shape, t, shape, t).has_value());
        CHECK(gjk.distance(shape, t, shape, t).has_value());
    }
    {   // test_gjk_3d_hit
        auto left = Primitive3::Cuboid(10, 10, 10), right = Primitive3::Cuboid(10, 10, 10);
        auto lt = transform_3d(15, 0, 0), rt = transform_3d(7, 2, 0);
        CHECK(gjk.intersect(left, lt, right, rt));
        auto c = gjk.intersection(CollisionStrategy::FullResolution, left, lt, right, rt);
        CHECK(c.has_value());
        if (c) {
            CHECK(c->normal[0] == -1.f && c->normal[1] == 0.f && c->normal[2] == 0.f);
            CHECK(c->penetration_depth == 2.f);
            CHECK(c->contact_point[0] == 10.f && c->contact_point[1] == 1.0000002f && c->contact_point[2] == 4.999999f);
        }
    }
    {   // test_gjk_distance_3d (the reference returns the convergence residual; the geometric distance is 4)
        auto left = Primitive3::Cuboid(10, 10, 10), right = Primitive3::Cuboid(10, 10, 10);
        CHECK(!gjk.distance(left, transform_3d(15, 0, 0), right, transform_3d(7, 2, 0)).has_value());
        ShapeSet s(ctx, {left, right}, {transform_3d(15, 0, 0), transform_3d(1, 0, 0)});
        auto d = gjk.distance_batch(s, {{0, 1}})[0];
        CHECK(d.has_value() && d->second == 4.f);
    }
    {   // test_gjk_time_of_impact_3d
        auto left = Primitive3::Cuboid(10, 11, 10), right = Primitive3::Cuboid(10, 15, 10);
        auto c = gjk.intersection_time_of_impact(left, transform_3d(0, 0, 0), transform_3d(30, 0, 0), right, transform_3d(15, 0, 0), transform_3d(15, 0, 0));
        CHECK(c.has_value());
        if (c) {
            CHECK(c->time_of_impact == 0.16666667f);
            CHECK(c->normal[0] == -1.f && c->normal[1] == 0.f && c->normal[2] == 0.f);
            CHECK(c->penetration_depth == 0.f);
            CHECK(c->contact_point[0] == 10.f && c->contact_point[1] == 0.f && c->contact_point[2] == 0.f);
        }
        CHECK(!gjk.intersection_time_of_impact(left, transform_3d(0, 0, 0), transform_3d(0, 0, 0), right, transform_3d(15, 0, 0), transform_3d(15, 0, 0)).has_value());
    }
    {   // SweepAndPrune3: strict overlap only; the slice is re-sorted; indices refer to the sorted slice
        std::vector<Aabb3> shapes = {{{5, 0, 0}, {6, 1, 1}}, {{0, 0, 0}, {1, 1, 1}}, {{0.5f, 0.5f, 0.5f}, {1.5f, 1.5f, 1.5f}}, {{1, 0, 0}, {2, 1, 1}}};
        SweepAndPrune3 sap(ctx);
        auto pairs = sap.find_collider_pairs(shapes);
        CHECK(shapes[0].min[0] == 0.f && shapes[1].min[0] == 0.5f && shapes[2].min[0] == 1.f && shapes[3].min[0] == 5.f);
        CHECK(pairs.size() == 2);
        if (pairs.size() == 2) { CHECK(pairs[0].first == 0 && pairs[0].second == 1); CHECK(pairs[1].first == 1 && pairs[1].second == 2); }
    }
    {   // SweepAndPrune3<volume::Sphere>: touching spheres are pairs (volume/sphere.rs:82-91), separated ones are not
        std::vector<Sphere> shapes = {{{5, 0, 0}, 0.5f}, {{0, 0, 0}, 1.f}, {{2, 0, 0}, 1.f}, {{2, 2.1f, 0}, 1.f}};
        SweepAndPrune3Sphere sap(ctx);
        auto pairs = sap.find_collider_pairs(shapes);
        CHECK(shapes[0].center[0] == 0.f && shapes[1].center[1] == 0.f && shapes[2].center[1] == 2.1f && shapes[3].center[0] == 5.f);
        CHECK(pairs.size() == 1);
        if (pairs.size() == 1) CHECK(pairs[0].first == 0 && pairs[0].second == 1);
        ShapeSet set(ctx, {Primitive3::Cuboid(2, 4, 6), Primitive3::Capsule(1, 0.5f)}, {transform_3d(1, 2, 3), transform_3d(-1, 0, 0)});
        auto ws = world_spheres(ctx, set);
        CHECK(ws.size() == 2 && ws[0].radius == 3.f && ws[0].center[0] == 1.f && ws[0].center[2] == 3.f);  // cuboid.rs:75-83: largest half extent
        CHECK(ws[1].radius == 1.5f && ws[1].center[0] == -1.f);                                            // capsule.rs:60-67
    }
    {   // FrameQueue: three frames of test_gjk_3d_hit's cuboids moving apart, two in flight, against the one-shot calls
        std::vector<Primitive3> prims = {Primitive3::Cuboid(10, 10, 10), Primitive3::Cuboid(10, 10, 10)};
        std::vector<std::vector<Mat4>> frames = {{transform_3d(15, 0, 0), transform_3d(7, 2, 0)}, {transform_3d(15, 0, 0), transform_3d(6, 2, 0)},
                                                 {transform_3d(15, 0, 0), transform_3d(4, 2, 0)}};
        std::vector<Pair> pairs = {{0, 1}};
        FrameQueue q(ctx, prims, frames[0], 1, 2);
        bam3d_contact c[3];
        uint8_t st[3] = {0xEE, 0xEE, 0xEE};
        uint64_t ticket[3];
        for (int k = 0; k < 3; ++k) ticket[k] = q.submit_intersection(gjk, CollisionStrategy::FullResolution, frames[k], pairs, &c[k], &st[k]);
        for (int k = 0; k < 3; ++k) {
            q.wait(ticket[k]);
            auto want = gjk.intersection(CollisionStrategy::FullResolution, prims[0], frames[k][0], prims[1], frames[k][1]);
            CHECK((st[k] == BAM3D_STATUS_OK) == want.has_value());
            if (want && st[k] == BAM3D_STATUS_OK) {
                CHECK(c[k].penetration_depth == want->penetration_depth && c[k].normal[0] == want->normal[0]);
                CHECK(c[k].contact_point[0] == want->contact_point[0] && c[k].contact_point[1] == want->contact_point[1] && c[k].contact_point[2] == want->contact_point[2]);
            }
        }
        CHECK(st[0] == BAM3D_STATUS_OK && c[0].penetration_depth == 2.f);  // frame 0 is test_gjk_3d_hit
        CHECK(st[2] != BAM3D_STATUS_OK);                                   // 15 - 4 = 11 > 10: separated
    }
    {   // compound shapes (gjk/mod.rs:362-456): two cuboids per side, local transforms composed with the model transforms
        std::vector<CompoundSet::Part> left = {{Primitive3::Cuboid(10, 10, 10), transform_3d(0, 0, 0)}, {Primitive3::Cuboid(2, 2, 2), transform_3d(0, 20, 0)}};
        std::vector<CompoundSet::Part> right = {{Primitive3::Cuboid(10, 10, 10), transform_3d(0, 0, 0)}, {Primitive3::Sphere(1), transform_3d(0, -20, 0)}};
        auto c = intersection_complex(ctx, gjk, CollisionStrategy::FullResolution, left, transform_3d(15, 0, 0), right, transform_3d(7, 2, 0));
        CHECK(c.has_value());
        if (c) { CHECK(c->penetration_depth == 2.f && c->normal[0] == -1.f); CHECK(c->contact_point[0] == 10.f && c->contact_point[1] == 1.0000002f); }  // == test_gjk_3d_hit
        CHECK(!intersection_complex(ctx, gjk, CollisionStrategy::FullResolution, left, transform_3d(15, 0, 0), right, transform_3d(-40, 2, 0)).has_value());
        CHECK(!distance_complex(ctx, gjk, left, transform_3d(15, 0, 0), right, transform_3d(7, 2, 0)).has_value());  // colliding -> None
        // GJK::distance of a Cuboid against the Sphere part runs to the iteration cap in the reference (None), and one None makes
        // distance_complex None (gjk/mod.rs:444): separated compounds report a distance only when every primitive pair does
        CHECK(!distance_complex(ctx, gjk, left, transform_3d(15, 0, 0), right, transform_3d(-40, 2, 0)).has_value());
        std::vector<CompoundSet::Part> boxes = {{Primitive3::Cuboid(10, 10, 10), transform_3d(0, 0, 0)}, {Primitive3::Cuboid(2, 2, 2), transform_3d(0, -20, 0)}};
        CHECK(distance_complex(ctx, gjk, left, transform_3d(15, 0, 0), boxes, transform_3d(-40, 2, 0)).has_value());
    }
    {   // broad -> narrow in one call: three cuboids in a row, the outer two do not touch; SHAPE indices come back
        ShapeSet set(ctx, {Primitive3::Cuboid(10, 10, 10), Primitive3::Cuboid(10, 10, 10), Primitive3::Cuboid(10, 10, 10)},
                     {transform_3d(30, 0, 0), transform_3d(7, 2, 0), transform_3d(15, 0, 0)});
        int axis = 0;
        auto cand = collide_all(ctx, gjk, CollisionStrategy::FullResolution, set, axis);
        CHECK(cand.size() == 1);
        if (cand.size() == 1) {
            CHECK(cand[0].left == 1 && cand[0].right == 2);   // sorted by min x: shape 1 (x 2..12), shape 2 (10..20), shape 0 (25..35)
            CHECK(cand[0].contact.has_value() && cand[0].contact->penetration_depth == 2.f);
        }
    }
    {   // DBVT queries through the host layer: tests/dbvt.rs:39-53 test_frustum, ray-vs-Aabb3 KATs of tests/aabb.rs:62-157, DbvtBroadPhase
        auto perspective_rh_gl = [](float fov_y, float aspect, float z_near, float z_far) {   // glam Mat4::perspective_rh_gl
            const float inv_length = 1.0f / (z_near - z_far), f = 1.0f / std::tan(0.5f * fov_y), a = f / aspect;
            return Mat4{a, 0, 0, 0, 0, f, 0, 0, 0, 0, (z_near + z_far) * inv_length, -1, 0, 0, (2.0f * z_near * z_far) * inv_length, 0};
        };
        DynamicBoundingVolumeTree tree(ctx, std::vector<Aabb3>{{{0.2f, 0.2f, 0.2f}, {10, 10, 10}}, {{0.2f, 0.2f, -10}, {10, 10, -0.2f}}});
        auto fr = Frustum::from_mat4(perspective_rh_gl(60.0f * (3.14159265358979323846f / 180.0f), 16.f / 9.f, 0.1f, 4.0f));
        CHECK(fr.has_value());
        if (fr) {
            auto hits = tree.query_frustum(*fr);
            CHECK(hits.size() == 1);
            if (hits.size() == 1) CHECK(hits[0].first == 1 && hits[0].second == Relation::Cross);
        }
        DynamicBoundingVolumeTree one(ctx, std::vector<Aabb3>{{{1, 1, 1}, {5, 5, 5}}});
        CHECK(one.query_ray(Ray{{0, 0, 0}, {1, 0, 0}}).empty());
        CHECK(one.query_ray(Ray{{6, 2, 2}, {1, 0, 0}}).empty());
        auto h = one.query_ray(Ray{{0, 6, 0}, {1, -1, 1}});
        CHECK(h.size() == 1);
        if (h.size() == 1) CHECK(h[0].first == 0 && h[0].second[0] == 1.f && h[0].second[1] == 5.f && h[0].second[2] == 1.f);
        auto c = one.query_ray_closest(Ray{{6, 2, 2}, {-1, 0, 0}});
        CHECK(c.has_value());
        if (c) CHECK(c->first == 0 && c->second[0] == 5.f && c->second[1] == 2.f && c->second[2] == 2.f);
        CHECK(!one.query_ray_closest(Ray{{6, 2, 2}, {1, 0, 0}}).has_value());
        CHECK(one.query_discrete(Aabb3{{4, 4, 4}, {6, 6, 6}}).size() == 1 && one.query_discrete(Aabb3{{5, 5, 5}, {6, 6, 6}}).empty());  // strict overlap
        DynamicBoundingVolumeTree three(ctx, std::vector<Aabb3>{{{0, 0, 0}, {2, 2, 2}}, {{1, 1, 1}, {3, 3, 3}}, {{5, 5, 5}, {6, 6, 6}}});
        auto pairs = three.find_collider_pairs({1, 1, 1});
        CHECK(pairs.size() == 1);
        if (pairs.size() == 1) CHECK(pairs[0].first == 0 && pairs[0].second == 1);
        // the same tree type over volume::Sphere bounds: tests/sphere.rs:7-47 test_intersection, tests/frustum.rs test_contains (with the
        // 1-radian field of view for which its expectations hold: DESIGN.md, documented discrepancies)
        DynamicBoundingVolumeTree ball(ctx, std::vector<Sphere>{{{0, 0, 0}, 1.f}});
        auto b0 = ball.query_ray(Ray{{0, 0, 5}, {0, 0, -1}});
        CHECK(b0.size() == 1);
        if (b0.size() == 1) CHECK(b0[0].second[0] == 0.f && b0[0].second[1] == 0.f && b0[0].second[2] == 1.f);
        auto b2 = ball.query_ray(Ray{{1, 0, 5}, {0, 0, -1}});
        CHECK(b2.size() == 1);
        if (b2.size() == 1) CHECK(b2[0].second[0] == 1.f && b2[0].second[1] == 0.f && b2[0].second[2] == 0.f);
        CHECK(ball.query_ray(Ray{{2, 0, 5}, {0, 0, -1}}).empty());
        CHECK(ball.query_discrete(Sphere{{2, 0, 0}, 1.f}).size() == 1 && ball.query_discrete(Sphere{{2.5f, 0, 0}, 1.f}).empty());   // touching spheres intersect
        DynamicBoundingVolumeTree balls(ctx, std::vector<Sphere>{{{0, 0, -5}, 1.f}, {{0, 3, -5}, 1.f}, {{0, 0, 5}, 1.f}});
        auto fr2 = Frustum::from_mat4(perspective_rh_gl(1.0f, 1.f, 1.f, 10.f));
        CHECK(fr2.has_value());
        if (fr2) {
            auto hits = balls.query_frustum(*fr2);
            CHECK(hits.size() == 2);
            for (auto& kv : hits) CHECK((kv.first == 0 && kv.second == Relation::In) || (kv.first == 1 && kv.second == Relation::Cross));
        }
        bool threw = false;
        try { ball.query_discrete(Aabb3{{0, 0, 0}, {1, 1, 1}}); } catch (const Error&) { threw = true; }
        CHECK(threw);   // an Aabb3 query against sphere bounds is not a reference operation
    }
    std::printf("kat_gpu: %s\n", fails ? "FAILED" : "all passed");
    return fails ? 1 : 0;
}
