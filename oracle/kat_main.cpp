// ORACLE — TEST INFRASTRUCTURE ONLY. Known-answer tests: every unit / integration test the reference holds for
// the hot path, ported with the reference's exact expected values (f32 ==, as Rust's assert_eq! does).
// Each test names the reference test it ports (paths relative to /root/reference).
// Exit code 0 iff every pinned KAT passes. Lines starting with "DISCREPANCY" report the two reference tests
// that the reference's own code cannot satisfy; they are printed, not asserted.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "bam3d_oracle.hpp"
#include "broad_oracle.hpp"
#include "ray_oracle.hpp"

using namespace orc;

static int g_fail = 0, g_pass = 0;
static const char* g_test = "";
#define CHECK(cond)                                                                      \
    do {                                                                                 \
        if (cond) { ++g_pass; }                                                          \
        else { ++g_fail; std::printf("FAIL %s: %s (line %d)\n", g_test, #cond, __LINE__); } \
    } while (0)
#define CHECK_EQF(expect, got)                                                                                \
    do {                                                                                                      \
        float e_ = (expect), g_ = (got);                                                                      \
        if (e_ == g_) { ++g_pass; }                                                                           \
        else { ++g_fail; std::printf("FAIL %s: expected %.9g got %.9g (line %d)\n", g_test, e_, g_, __LINE__); } \
    } while (0)
#define CHECK_V3(ex, ey, ez, v) do { Vec3 v_ = (v); CHECK_EQF(ex, v_.x); CHECK_EQF(ey, v_.y); CHECK_EQF(ez, v_.z); } while (0)
#define TEST(name) g_test = name

static Mat4 transform_z(float x, float y, float z, float angle_deg) {  // the tests' transform()/transform_3d()
    return Mat4::from_scale_rotation_translation(Vec3::splat(1.f), Quat::from_rotation_z(to_radians(angle_deg)), Vec3(x, y, z));
}
static SupportPoint sup(float x, float y, float z) { return sup_from_v(x, y, z); }

static Vec3 run_check_side(Simplex& s, bool above, bool ignore) {  // simplex3d.rs:360-372
    Vec3 ab = s[1].v - s[2].v, ac = s[0].v - s[2].v, ao = -s[2].v;
    Vec3 abc = cross(ab, ac);
    Vec3 v = Vec3::zero();
    check_side(abc, ab, ac, ao, s, v, above, ignore);
    return v;
}

static void simplex_tests() {
    {   TEST("simplex3d.rs:212 test_check_side_outside_ab");
        Simplex s; s.push(sup(8, -10, 0)); s.push(sup(-1, -10, 0)); s.push(sup(3, 5, 0));
        Vec3 v = run_check_side(s, false, false);
        CHECK(s.len() == 2); CHECK_V3(-1.f, -10.f, 0.f, s[0].v); CHECK_V3(-375.f, 100.f, 0.f, v); }
    {   TEST("simplex3d.rs:223 test_check_side_outside_ac");
        Simplex s; s.push(sup(2, -10, 0)); s.push(sup(-7, -10, 0)); s.push(sup(-3, 5, 0));
        Vec3 v = run_check_side(s, false, false);
        CHECK(s.len() == 2); CHECK_V3(2.f, -10.f, 0.f, s[0].v); CHECK_V3(300.f, 100.f, 0.f, v); }
    {   TEST("simplex3d.rs:234 test_check_side_above");
        Simplex s; s.push(sup(5, -10, -1)); s.push(sup(-4, -10, -1)); s.push(sup(0, 5, -1));
        Vec3 v = run_check_side(s, false, false);
        CHECK(s.len() == 3); CHECK_V3(5.f, -10.f, -1.f, s[0].v); CHECK_V3(0.f, 0.f, 135.f, v); }
    {   TEST("simplex3d.rs:245 test_check_side_below");
        Simplex s; s.push(sup(5, -10, 1)); s.push(sup(-4, -10, 1)); s.push(sup(0, 5, 1));
        Vec3 v = run_check_side(s, false, false);
        CHECK(s.len() == 3); CHECK_V3(-4.f, -10.f, 1.f, s[0].v); CHECK_V3(0.f, 0.f, -135.f, v); }
    {   TEST("simplex3d.rs:256 test_check_origin_empty");
        Simplex s; Vec3 v = Vec3::zero(); bool hit = reduce_to_closest_feature(s, v);
        CHECK(!hit); CHECK(s.len() == 0); CHECK(all_eq(v, Vec3::zero())); }
    {   TEST("simplex3d.rs:265 test_check_origin_point");
        Simplex s; s.push(sup(8, -10, 0)); Vec3 v = Vec3::zero(); bool hit = reduce_to_closest_feature(s, v);
        CHECK(!hit); CHECK(s.len() == 1); CHECK(all_eq(v, Vec3::zero())); }
    {   TEST("simplex3d.rs:274 test_check_origin_line");
        Simplex s; s.push(sup(8, -10, 0)); s.push(sup(-1, -10, 0)); Vec3 v = Vec3::zero();
        bool hit = reduce_to_closest_feature(s, v);
        CHECK(!hit); CHECK(s.len() == 2); CHECK_V3(0.f, 810.f, 0.f, v); }
    {   TEST("simplex3d.rs:283 test_check_origin_triangle");
        Simplex s; s.push(sup(5, -10, -1)); s.push(sup(-4, -10, -1)); s.push(sup(0, 5, -1)); Vec3 v = Vec3::zero();
        bool hit = reduce_to_closest_feature(s, v);
        CHECK(!hit); CHECK(s.len() == 3); CHECK_V3(0.f, 0.f, 135.f, v); }
    {   TEST("simplex3d.rs:292 test_check_origin_tetrahedron_outside_abc");
        Simplex s; s.push(sup(8, -10, -1)); s.push(sup(-1, -10, -1)); s.push(sup(3, 5, -1)); s.push(sup(3, -3, 5));
        Vec3 v = Vec3::zero(); bool hit = reduce_to_closest_feature(s, v);
        CHECK(!hit); CHECK(s.len() == 3); CHECK_V3(-1.f, -10.f, -1.f, s[0].v); CHECK_V3(-90.f, 24.f, 32.f, v); }
    {   TEST("simplex3d.rs:307 test_check_origin_tetrahedron_outside_acd");
        Simplex s; s.push(sup(8, 0.1f, -1)); s.push(sup(-1, 0.1f, -1)); s.push(sup(3, 15, -1)); s.push(sup(3, 7, 5));
        Vec3 v = Vec3::zero(); bool hit = reduce_to_closest_feature(s, v);
        CHECK(!hit); CHECK(s.len() == 3);
        CHECK_V3(3.f, 7.f, 5.f, s[2].v); CHECK_V3(-1.f, 0.1f, -1.f, s[1].v); CHECK_V3(8.f, 0.1f, -1.f, s[0].v);
        CHECK_V3(0.f, -54.f, 62.1f, v); }
    {   TEST("simplex3d.rs:324 test_check_origin_tetrahedron_outside_adb");
        Simplex s; s.push(sup(2, -10, -1)); s.push(sup(-7, -10, -1)); s.push(sup(-3, 5, -1)); s.push(sup(-3, -3, 5));
        Vec3 v = Vec3::zero(); bool hit = reduce_to_closest_feature(s, v);
        CHECK(!hit); CHECK(s.len() == 3);
        CHECK_V3(-3.f, -3.f, 5.f, s[2].v); CHECK_V3(2.f, -10.f, -1.f, s[1].v); CHECK_V3(-3.f, 5.f, -1.f, s[0].v);
        CHECK_V3(90.f, 30.f, 40.f, v); }
    {   TEST("simplex3d.rs:341 test_check_origin_tetrahedron_inside");
        Simplex s; s.push(sup(3, -3, -1)); s.push(sup(-3, -3, -1)); s.push(sup(0, 3, -1)); s.push(sup(0, 0, 5));
        Vec3 v = Vec3::zero(); bool hit = reduce_to_closest_feature(s, v);
        CHECK(hit); CHECK(s.len() == 4); }
}

static void check_face(const Face& f, uint32_t a, uint32_t b, uint32_t c, float nx, float ny, float nz, float d) {
    CHECK(f.vertices[0] == a && f.vertices[1] == b && f.vertices[2] == c);
    CHECK_EQF(nx, f.normal.x); CHECK_EQF(ny, f.normal.y); CHECK_EQF(nz, f.normal.z); CHECK_EQF(d, f.distance);
}

static void epa_tests() {
    {   TEST("epa3d.rs:199 test_remove_or_add_edge_added");
        std::vector<std::pair<uint32_t, uint32_t>> e = {{1, 2}, {6, 5}};
        remove_or_add_edge(e, {4, 3});
        CHECK(e.size() == 3); CHECK(e[2].first == 4 && e[2].second == 3); }
    {   TEST("epa3d.rs:207 test_remove_or_add_edge_removed");
        std::vector<std::pair<uint32_t, uint32_t>> e = {{1, 2}, {6, 5}};
        remove_or_add_edge(e, {2, 1});
        CHECK(e.size() == 1); CHECK(e[0].first == 6 && e[0].second == 5); }
    std::vector<SupportPoint> base = {sup(3, -3, -1), sup(-3, -3, -1), sup(0, 3, -1), sup(0, 0, 5)};
    {   TEST("epa3d.rs:215 test_face_impl");
        std::vector<SupportPoint> s = base; Polytope p(s);
        CHECK(p.faces.size() == 4);
        check_face(p.faces[0], 3, 2, 1, -0.8728715f, 0.43643576f, 0.21821788f, 1.0910894f);
        check_face(p.faces[1], 3, 1, 0, 0.f, -0.89442724f, 0.44721362f, 2.236068f);
        check_face(p.faces[2], 3, 0, 2, 0.8728715f, 0.43643576f, 0.21821788f, 1.0910894f);
        check_face(p.faces[3], 2, 0, 1, 0.f, 0.f, -1.f, 1.0f); }
    {   TEST("epa3d.rs:249 test_polytope_closest_to_origin");
        std::vector<SupportPoint> s = base; Polytope p(s);
        check_face(p.closest_face_to_origin(), 2, 0, 1, 0.f, 0.f, -1.f, 1.0f); }
    {   TEST("epa3d.rs:262 test_polytope_add");
        std::vector<SupportPoint> s = base; Polytope p(s);
        p.add(sup(0, 0, -2));
        CHECK(p.vertices.size() == 5); CHECK(p.faces.size() == 6);
        CHECK(p.faces[3].vertices[0] == 4 && p.faces[3].vertices[1] == 2 && p.faces[3].vertices[2] == 0);
        CHECK(p.faces[4].vertices[0] == 4 && p.faces[4].vertices[1] == 0 && p.faces[4].vertices[2] == 1);
        CHECK(p.faces[5].vertices[0] == 4 && p.faces[5].vertices[1] == 1 && p.faces[5].vertices[2] == 2); }
    {   TEST("epa3d.rs:279 test_epa_3d");
        Shape left = Shape::cuboid(10, 10, 10), right = Shape::cuboid(10, 10, 10);
        Mat4 lt = transform_z(15, 0, 0, 0), rt = transform_z(7, 2, 0, 0);
        std::vector<SupportPoint> s = {sup(18, -12, 0), sup(-2, 8, 0), sup(-2, -12, 0), sup(8, -2, -10)};
        Contact c; bool ok = EPA3().process(s, left, lt, right, rt, c);
        CHECK(ok); CHECK_V3(-1.f, 0.f, 0.f, c.normal); CHECK_EQF(2.f, c.penetration_depth); }
}

static void gjk_tests() {
    GJK gjk;
    {   TEST("gjk/mod.rs:552 test_gjk_exact_3d");
        Shape shape = Shape::cuboid(1, 1, 1); Mat4 t = transform_z(0, 0, 0, 0);
        Contact c; CHECK(gjk.intersection(CollisionStrategy::FullResolution, shape, t, shape, t, c));
        float rv, geo; Status st; CHECK(!gjk.distance(shape, t, shape, t, rv, geo, st)); }
    {   TEST("gjk/mod.rs:563 test_gjk_sphere");
        Shape shape = Shape::sphere(1); Mat4 t = transform_z(0, 0, 0, 0);
        Contact c; CHECK(gjk.intersection(CollisionStrategy::FullResolution, shape, t, shape, t, c));
        float rv = -1, geo = -1; Status st; bool some = gjk.distance(shape, t, shape, t, rv, geo, st);
        CHECK(some);
        // reference asserts d == 0. exactly; libm/rounding sensitive -> documented discrepancy
        std::printf("DISCREPANCY gjk/mod.rs:572 test_gjk_sphere expects Some(0.), oracle gives %s(%.9g)\n",
                    some ? "Some" : "None", rv); }
    {   TEST("gjk/mod.rs:575 test_gjk_3d_hit");
        Shape left = Shape::cuboid(10, 10, 10), right = Shape::cuboid(10, 10, 10);
        Mat4 lt = transform_z(15, 0, 0, 0), rt = transform_z(7, 2, 0, 0);
        Simplex s; CHECK(gjk.intersect(left, lt, right, rt, s));
        Contact c; Counters cnt;
        CHECK(gjk.intersection(CollisionStrategy::FullResolution, left, lt, right, rt, c, nullptr, &cnt));
        CHECK_V3(-1.f, 0.f, 0.f, c.normal); CHECK_EQF(2.f, c.penetration_depth);
        CHECK_V3(10.f, 1.0000002f, 4.999999f, c.contact_point); }
    {   TEST("gjk/mod.rs:598 test_gjk_distance_3d");
        Shape left = Shape::cuboid(10, 10, 10), right = Shape::cuboid(10, 10, 10);
        Mat4 lt = transform_z(15, 0, 0, 0), rt = transform_z(7, 2, 0, 0);
        float rv = -1, geo = -1; Status st;
        CHECK(!gjk.distance(left, lt, right, rt, rv, geo, st));
        Mat4 rt2 = transform_z(1, 0, 0, 0);
        bool some = gjk.distance(left, lt, right, rt2, rv, geo, st);
        CHECK(some);
        CHECK_EQF(4.f, geo);  // the geometric distance sqrt(d.d) (commented out at gjk/mod.rs:274) IS 4
        // the reference's own assert (Some(4.)) cannot be met by its code, which returns the residual dp - d0
        std::printf("DISCREPANCY gjk/mod.rs:611 test_gjk_distance_3d expects Some(4.), reference code returns the residual: "
                    "oracle gives %s(%.9g), geometric %.9g\n", some ? "Some" : "None", rv, geo); }
    {   TEST("gjk/mod.rs:617 test_gjk_time_of_impact_3d");
        Shape left = Shape::cuboid(10, 11, 10), right = Shape::cuboid(10, 15, 10);
        Mat4 l0 = transform_z(0, 0, 0, 0), l1 = transform_z(30, 0, 0, 0), rt = transform_z(15, 0, 0, 0);
        Contact c; Status st;
        CHECK(gjk.intersection_time_of_impact(left, l0, l1, right, rt, rt, c, st));
        CHECK_EQF(0.16666667f, c.time_of_impact); CHECK_V3(-1.f, 0.f, 0.f, c.normal);
        CHECK_EQF(0.f, c.penetration_depth); CHECK_V3(10.f, 0.f, 0.f, c.contact_point);
        CHECK(!gjk.intersection_time_of_impact(left, l0, l0, right, rt, rt, c, st)); }
    {   TEST("quad.rs:113 test_plane_cuboid_intersect");
        Shape quad = Shape::quad(2, 2), cuboid = Shape::cuboid(1, 1, 1);
        Quat rx = Quat::from_rotation_x(to_radians(1.0f));
        Mat4 t1 = Mat4::from_scale_rotation_translation(Vec3::splat(1.f), rx, Vec3(0, 0, 1));
        Mat4 t2 = Mat4::from_scale_rotation_translation(Vec3::splat(1.f), rx, Vec3(0, 0, 1.1f));
        Simplex s; CHECK(gjk.intersect(quad, t1, cuboid, t2, s)); }
}

static void support_tests() {
    const float PI = 3.14159265358979323846f;
    auto sphere_support = [&](float dx, float dy, float dz, float px, float py, float pz, float rot) {
        Shape s = Shape::sphere(10); Vec3 p = support_point(s, Vec3(dx, dy, dz), transform_z(0, 0, 0, rot));
        CHECK_V3(px, py, pz, p);
    };
    TEST("sphere.rs:87 test_sphere_support_1"); sphere_support(1, 0, 0, 10, 0, 0, 0);
    TEST("sphere.rs:92 test_sphere_support_2"); sphere_support(1, 1, 1, 5.773503f, 5.773503f, 5.773503f, 0);
    TEST("sphere.rs:105 test_sphere_support_3"); sphere_support(1, 0, 0, 10, 0, 0, -PI / 4.f);
    {   TEST("sphere.rs:118 test_sphere_support_4");
        CHECK_V3(10.f, 10.f, 0.f, support_point(Shape::sphere(10), Vec3(1, 0, 0), transform_z(0, 10, 0, 0))); }
    Shape cyl = Shape::cylinder(2, 1);
    {   TEST("cylinder.rs:221 test_cylinder_support_1"); CHECK_V3(1.f, 2.f, 0.f, support_point(cyl, Vec3(1, 0, 0), transform_z(0, 0, 0, 0))); }
    {   TEST("cylinder.rs:230 test_cylinder_support_2"); CHECK_V3(1.f, -2.f, 0.f, support_point(cyl, normalize(Vec3(0.5f, -1, 0)), transform_z(0, 0, 0, 0))); }
    {   TEST("cylinder.rs:239 test_cylinder_support_3"); CHECK_V3(0.f, 2.f, 0.f, support_point(cyl, Vec3(0, 1, 0), transform_z(0, 0, 0, 0))); }
    {   TEST("cylinder.rs:248 test_cylinder_support_4"); CHECK_V3(11.f, 2.f, 0.f, support_point(cyl, Vec3(1, 0, 0), transform_z(10, 0, 0, 0))); }
    {   TEST("cylinder.rs:257 test_cylinder_support_5"); CHECK_V3(1.1081045f, -1.9421906f, 0.0f, support_point(cyl, Vec3(1, 0, 0), transform_z(0, 0, 0, PI))); }
    Shape cap = Shape::capsule(2, 1);
    {   TEST("capsule.rs:195 test_capsule_support_1"); CHECK_V3(1.f, 2.f, 0.f, support_point(cap, Vec3(1, 0, 0), transform_z(0, 0, 0, 0))); }
    {   TEST("capsule.rs:204 test_capsule_support_2"); CHECK_V3(0.44721365f, -2.8944273f, 0.f, support_point(cap, normalize(Vec3(0.5f, -1, 0)), transform_z(0, 0, 0, 0))); }
    {   TEST("capsule.rs:213 test_capsule_support_3"); CHECK_V3(0.f, 3.f, 0.f, support_point(cap, Vec3(0, 1, 0), transform_z(0, 0, 0, 0))); }
    {   TEST("capsule.rs:222 test_capsule_support_4"); CHECK_V3(11.f, 2.f, 0.f, support_point(cap, Vec3(1, 0, 0), transform_z(10, 0, 0, 0))); }
    {   TEST("capsule.rs:231 test_capsule_support_5"); CHECK_V3(1.1096073f, -1.9969943f, 0.f, support_point(cap, Vec3(1, 0, 0), transform_z(0, 0, 0, PI))); }
    {   TEST("polyhedron.rs:532 test_polytope_half_edge (VertexOnly half)");
        float v[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
        Shape p = Shape::polyhedron(v, 4); Mat4 t = transform_z(0, 0, 0, 0);
        CHECK_V3(1.f, 0.f, 0.f, support_point(p, Vec3(1, 0, 0), t));
        CHECK_V3(0.f, 1.f, 0.f, support_point(p, Vec3(0, 1, 0), t)); }
}

static void aabb_tests() {
    {   TEST("tests/aabb.rs:10 test_general");
        Aabb3 a(Vec3(-20, 30, 5), Vec3(10, -10, -5));
        CHECK_V3(-20.f, -10.f, -5.f, a.min); CHECK_V3(10.f, 30.f, 5.f, a.max); CHECK_V3(30.f, 40.f, 10.f, a.dim());
        CHECK_EQF(30.f * 40.f * 10.f, a.volume()); CHECK_V3(-5.f, 10.f, 0.f, a.center());
        CHECK(a.contains(Vec3(0, 0, 0))); CHECK(!a.contains(Vec3(-100, 0, 0))); CHECK(!a.contains(Vec3(100, 0, 0)));
        CHECK(a.contains(Vec3(9, 29, -1))); CHECK(!a.contains(Vec3(10, 30, 5))); CHECK(a.contains(Vec3(-20, -10, -5)));
        CHECK(!a.contains(Vec3(-21, -11, -6)));
        Aabb3 b = a.add_v(Vec3(1, 2, 3)); CHECK_V3(-19.f, -8.f, -2.f, b.min); CHECK_V3(11.f, 32.f, 8.f, b.max);
        b = a.mul_s(2); CHECK_V3(-40.f, -20.f, -10.f, b.min); CHECK_V3(20.f, 60.f, 10.f, b.max);
        b = a.mul_v(Vec3(1, 2, 3)); CHECK_V3(-20.f, -20.f, -15.f, b.min); CHECK_V3(10.f, 60.f, 15.f, b.max);
        b = a.add_margin(Vec3(2, 2, 2)); CHECK_V3(-22.f, -12.f, -7.f, b.min); CHECK_V3(12.f, 32.f, 7.f, b.max); }
    Aabb3 box(Vec3(1, 1, 1), Vec3(5, 5, 5));
    auto miss = [&](Ray r) { Vec3 p; CHECK(!ray_aabb_intersection(r, box, p)); CHECK(!ray_aabb_intersects(r, box)); };
    auto hit = [&](Ray r, float x, float y, float z) { Vec3 p; CHECK(ray_aabb_intersection(r, box, p)); CHECK_V3(x, y, z, p); CHECK(ray_aabb_intersects(r, box)); };
    {   TEST("tests/aabb.rs:62 test_parallel_ray_should_not_intersect");
        miss(Ray(Vec3(0, 0, 0), Vec3(1, 0, 0))); miss(Ray(Vec3(0, 0, 0), Vec3(0, 1, 0))); miss(Ray(Vec3(0, 0, 0), Vec3(0, 0, 1)));
        miss(Ray(Vec3(0, 0, 0), Vec3(0.0001f, 0.0001f, 1.0f))); }
    {   TEST("tests/aabb.rs:84 test_oblique_ray_should_intersect");
        hit(Ray(Vec3(0, 0, 0), normalize(Vec3(1, 1, 1))), 1, 1, 1); hit(Ray(Vec3(0, 6, 0), Vec3(1, -1, 1)), 1, 5, 1); }
    {   TEST("tests/aabb.rs:99 test_pointing_to_other_dir_ray_should_not_intersect");
        miss(Ray(Vec3(0, 2, 2), Vec3(-1, 0.01f, 0.01f))); miss(Ray(Vec3(2, 0, 2), Vec3(0.01f, -1, 0.01f)));
        miss(Ray(Vec3(2, 2, 0), Vec3(0.01f, 0.01f, -1))); miss(Ray(Vec3(6, 2, 2), Vec3(1, 0, 0)));
        miss(Ray(Vec3(2, 6, 2), Vec3(0, 1, 0))); miss(Ray(Vec3(2, 2, 6), Vec3(0, 0, 1))); }
    {   TEST("tests/aabb.rs:134 test_pointing_to_box_dir_ray_should_intersect");
        hit(Ray(Vec3(0, 2, 2), Vec3(1, 0, 0)), 1, 2, 2); hit(Ray(Vec3(2, 0, 2), Vec3(0, 1, 0)), 2, 1, 2);
        hit(Ray(Vec3(2, 2, 0), Vec3(0, 0, 1)), 2, 2, 1); hit(Ray(Vec3(6, 2, 2), Vec3(-1, 0, 0)), 5, 2, 2);
        hit(Ray(Vec3(2, 6, 2), Vec3(0, -1, 0)), 2, 5, 2); hit(Ray(Vec3(2, 2, 6), Vec3(0, 0, -1)), 2, 2, 5); }
    {   TEST("tests/aabb.rs:159 test_corners");
        Vec3 c[8]; Aabb3(Vec3(-20, 30, 5), Vec3(10, -10, -5)).to_corners(c);
        auto has = [&](Vec3 p) { for (auto& q : c) if (all_eq(p, q)) return true; return false; };
        CHECK(has(Vec3(-20, 30, -5))); CHECK(has(Vec3(10, 30, 5))); CHECK(has(Vec3(10, -10, 5))); }
    {   TEST("tests/aabb.rs:169 test_bound");
        Aabb3 a(Vec3(-5, 5, 0), Vec3(5, 10, 1));
        CHECK(a.relate_plane(Plane::from_point_normal(Vec3(0, 0, 0), Vec3(0, 0, 1))) == Cross);
        CHECK(a.relate_plane(Plane::from_point_normal(Vec3(-5, 4, 0), Vec3(0, 1, 0))) == In);
        CHECK(a.relate_plane(Plane::from_point_normal(Vec3(6, 0, 0), Vec3(1, 0, 0))) == Out); }
    {   TEST("tests/aabb.rs:184-222 aabb3 intersects");
        Aabb3 a(Vec3(0, 0, 0), Vec3(10, 10, 10));
        CHECK(!a.intersects(Aabb3(Vec3(15, 15, 15), Vec3(25, 25, 25))));
        CHECK(a.intersects(Aabb3(Vec3(5, 5, 5), Vec3(15, 15, 15))));
        CHECK(!a.intersects(Aabb3(Vec3(5, 11, 11), Vec3(15, 13, 13))));
        CHECK(!a.intersects(Aabb3(Vec3(11, 5, 11), Vec3(15, 13, 13))));
        CHECK(!a.intersects(Aabb3(Vec3(11, 11, 5), Vec3(15, 13, 13)))); }
    {   TEST("tests/aabb.rs:224-268 contains vec3/aabb3");
        Aabb3 a(Vec3(0, 0, 0), Vec3(10, 10, 10));
        CHECK(a.contains(Vec3(3, 3, 3))); CHECK(!a.contains(Vec3(11, 11, 11)));
        CHECK(a.contains(Aabb3(Vec3(1, 1, 1), Vec3(2, 2, 2)))); CHECK(!a.contains(Aabb3(Vec3(11, 11, 11), Vec3(12, 12, 12))));
        CHECK(!a.contains(Aabb3(Vec3(1, 1, 1), Vec3(12, 12, 12)))); }
    {   TEST("tests/aabb.rs:292 test_ray_aabb3_parallel");
        Aabb3 a(Vec3(5, 5, 5), Vec3(10, 10, 10)); Vec3 p;
        Ray r1(Vec3(2, 0, 2), Vec3(0, 1, 0)), r2(Vec3(0, 2, 2), Vec3(1, 0, 0)), r3(Vec3(2, 2, 0), Vec3(0, 0, 1));
        CHECK(!ray_aabb_intersects(r1, a)); CHECK(!ray_aabb_intersection(r1, a, p));
        CHECK(!ray_aabb_intersects(r2, a)); CHECK(!ray_aabb_intersection(r2, a, p));
        CHECK(!ray_aabb_intersects(r3, a)); CHECK(!ray_aabb_intersection(r3, a, p)); }
    {   TEST("tests/aabb.rs:309 test_aabb3_union_aabb3");
        Aabb3 base(Vec3(0, 0, 0), Vec3(10, 10, 10));
        Aabb3 u = base.union_(Aabb3(Vec3(2, 2, 2), Vec3(5, 5, 5))); CHECK_V3(0.f, 0.f, 0.f, u.min); CHECK_V3(10.f, 10.f, 10.f, u.max);
        u = base.union_(Aabb3(Vec3(12, 12, 12), Vec3(15, 15, 15))); CHECK_V3(0.f, 0.f, 0.f, u.min); CHECK_V3(15.f, 15.f, 15.f, u.max);
        u = base.union_(Aabb3(Vec3(2, 2, 2), Vec3(12, 12, 12))); CHECK_V3(0.f, 0.f, 0.f, u.min); CHECK_V3(12.f, 12.f, 12.f, u.max); }
    {   TEST("tests/aabb.rs:354 test_aabb3_surface_area");
        CHECK_EQF(2200.f, Aabb3(Vec3(0, 0, 0), Vec3(10, 20, 30)).surface_area()); }
    {   TEST("tests/aabb.rs:361 test_aabb3_transform");
        Aabb3 t = Aabb3(Vec3(0, 0, 0), Vec3(10, 20, 30)).transform(transform_z(5, 3, 1, 0));
        CHECK_V3(5.f, 3.f, 1.f, t.min); CHECK_V3(15.f, 23.f, 31.f, t.max); }
}

static void plane_point_tests() {
    {   TEST("tests/plane.rs:8 test_from_points");
        Plane p; CHECK(Plane::from_points(Vec3(5, 0, 5), Vec3(5, 5, 5), Vec3(5, 0, -1), p));
        CHECK_V3(-1.f, 0.f, 0.f, p.n); CHECK_EQF(5.f, p.d);
        CHECK(!Plane::from_points(Vec3(0, 5, -5), Vec3(0, 5, 0), Vec3(0, 5, 5), p)); }
    {   TEST("tests/plane.rs:29 test_ray_intersection");
        Plane p0(Vec3(1, 0, 0), -7.f); Ray r0(Vec3(2, 3, 4), normalize(Vec3(1, 1, 1)));
        Vec3 q; CHECK(p0.intersects(r0)); CHECK(p0.intersection(r0, q)); CHECK_V3(7.f, 8.f, 9.f, q);
        Plane p1; Plane::from_points(Vec3(5, 0, 5), Vec3(5, 5, 5), Vec3(5, 0, -1), p1);
        Ray r1(Vec3(0, 0, 0), normalize(Vec3(-1, 0, 0)));
        CHECK(!p1.intersection(r1, q)); CHECK(!p1.intersects(r1)); }
    {   TEST("tests/point.rs:8 test_bound");
        Vec3 point(1, 2, 3), normal(0, -0.8f, -0.36f);
        Plane plane = Plane::from_point_normal(point, normal);
        CHECK(relate_plane(point, plane) == Cross); CHECK(relate_plane(point + normal, plane) == In);
        CHECK(relate_plane(point + normal * -1.0f, plane) == Out); }
}

static void dbvt_tests() {
    TEST("tests/dbvt.rs:39 test_frustum");
    Dbvt tree;
    auto val = [](Aabb3 a) { TreeValue v; v.bound = a; v.fat = a.add_margin(Vec3(3, 3, 3)); return v; };
    tree.insert(val(Aabb3(Vec3(0.2f, 0.2f, 0.2f), Vec3(10, 10, 10))));
    tree.insert(val(Aabb3(Vec3(0.2f, 0.2f, -0.2f), Vec3(10, 10, -10))));
    tree.refit();
    Mat4 proj = Mat4::perspective_rh_gl(to_radians(60.0f), 16.f / 9.f, 0.1f, 4.0f);
    FrustumVisitor vis; CHECK(Frustum::from_mat4(proj, vis.frustum));
    std::vector<std::pair<uint32_t, uint8_t>> res;
    tree.query_for_indices<FrustumVisitor, uint8_t>(vis, res);
    CHECK(res.size() == 1);
    if (res.size() == 1) { CHECK(res[0].second == (uint8_t)Cross); CHECK(res[0].first == 1); /* id 11 = second value */ }
}

static void bound_tests() {
    // ComputeBound<Aabb3> KATs (a16) are asserted through oracle_capi's orc_compute_bound in tests/; the local
    // bounds themselves are closed-form constructor calls (sphere.rs:27-34, cuboid.rs:66-73, cylinder.rs:64-71,
    // capsule.rs:51-58), checked here against the reference's expected boxes.
    TEST("cuboid.rs:173 / sphere.rs:126 / cylinder.rs:212 / capsule.rs:186 bounds");
    Shape c = Shape::cuboid(10, 10, 10); Aabb3 b(Vec3(-c.p0, -c.p1, -c.p2), Vec3(c.p0, c.p1, c.p2));
    CHECK_V3(-5.f, -5.f, -5.f, b.min); CHECK_V3(5.f, 5.f, 5.f, b.max);
}

// Ray casts against primitives: every ray test of primitive/{sphere,cuboid,cylinder,capsule}.rs.
static void ray_tests() {
    Vec3 p;
    const Mat4 I = transform_z(0, 0, 0, 0);
    {   Shape sph = Shape::sphere(10.f);
        TEST("sphere.rs:134 test_ray_discrete");
        CHECK(ray_intersects_local(sph, Ray(Vec3(20, 0, 0), Vec3(-1, 0, 0))));
        CHECK(!ray_intersects_local(sph, Ray(Vec3(20, 0, 0), Vec3(1, 0, 0))));
        CHECK(!ray_intersects_local(sph, Ray(Vec3(20, -15, 0), Vec3(-1, 0, 0))));
        TEST("sphere.rs:146 test_ray_discrete_transformed");
        CHECK(ray_intersects_transformed(sph, Ray(Vec3(20, 0, 0), Vec3(-1, 0, 0)), I));
        CHECK(!ray_intersects_transformed(sph, Ray(Vec3(20, 0, 0), Vec3(1, 0, 0)), I));
        CHECK(!ray_intersects_transformed(sph, Ray(Vec3(20, 0, 0), Vec3(-1, 0, 0)), transform_z(0, 15, 0, 0)));
        TEST("sphere.rs:159 test_ray_continuous");
        CHECK(ray_intersection_local(sph, Ray(Vec3(20, 0, 0), Vec3(-1, 0, 0)), p)); CHECK_V3(10.f, 0.f, 0.f, p);
        CHECK(!ray_intersection_local(sph, Ray(Vec3(20, 0, 0), Vec3(1, 0, 0)), p));
        CHECK(!ray_intersection_local(sph, Ray(Vec3(20, -15, 0), Vec3(-1, 0, 0)), p));
        TEST("sphere.rs:170 test_ray_continuous_transformed");
        CHECK(ray_intersection_transformed(sph, Ray(Vec3(20, 0, 0), Vec3(-1, 0, 0)), I, p)); CHECK_V3(10.f, 0.f, 0.f, p);
        CHECK(!ray_intersection_transformed(sph, Ray(Vec3(20, 0, 0), Vec3(1, 0, 0)), I, p));
        CHECK(!ray_intersection_transformed(sph, Ray(Vec3(20, 0, 0), Vec3(-1, 0, 0)), transform_z(0, 15, 0, 0), p)); }
    {   Shape cub = Shape::cuboid(10, 10, 10);
        TEST("cuboid.rs test_ray_discrete");
        CHECK(ray_intersects_local(cub, Ray(Vec3(10, 0, 0), Vec3(-1, 0, 0))));
        CHECK(!ray_intersects_local(cub, Ray(Vec3(10, 0, 0), Vec3(1, 0, 0))));
        TEST("cuboid.rs test_ray_discrete_transformed");
        CHECK(ray_intersects_transformed(cub, Ray(Vec3(10, 0, 0), Vec3(-1, 0, 0)), transform_z(0, 1, 0, 0)));
        CHECK(!ray_intersects_transformed(cub, Ray(Vec3(10, 0, 0), Vec3(1, 0, 0)), transform_z(0, 1, 0, 0)));
        CHECK(ray_intersects_transformed(cub, Ray(Vec3(10, 0, 0), Vec3(-1, 0, 0)), transform_z(0, 1, 0, 0.3f)));
        TEST("cuboid.rs test_ray_continuous");
        CHECK(ray_intersection_local(cub, Ray(Vec3(10, 0, 0), Vec3(-1, 0, 0)), p)); CHECK_V3(5.f, 0.f, 0.f, p);
        CHECK(!ray_intersection_local(cub, Ray(Vec3(10, 0, 0), Vec3(1, 0, 0)), p));
        TEST("cuboid.rs test_ray_continuous_transformed");
        CHECK(ray_intersection_transformed(cub, Ray(Vec3(10, 0, 0), Vec3(-1, 0, 0)), transform_z(0, 1, 0, 0), p)); CHECK_V3(5.f, 0.f, 0.f, p);
        CHECK(!ray_intersection_transformed(cub, Ray(Vec3(10, 0, 0), Vec3(1, 0, 0)), transform_z(0, 1, 0, 0), p));
        // rotated by 0.3 degrees: the reference asserts |p - (5.000068, 0, 0)| < f32::EPSILON per component
        CHECK(ray_intersection_transformed(cub, Ray(Vec3(10, 0, 0), Vec3(-1, 0, 0)), transform_z(0, 0, 0, 0.3f), p));
        CHECK(std::fabs(p.x - 5.000068f) < F32_EPSILON); CHECK(std::fabs(p.y - 0.f) < F32_EPSILON); CHECK(std::fabs(p.z - 0.f) < F32_EPSILON); }
    {   Shape cyl = Shape::cylinder(2.f, 1.f);
        TEST("cylinder.rs test_discrete_1..4");
        CHECK(ray_intersects_local(cyl, Ray(Vec3(-3, 0, 0), Vec3(1, 0, 0))));
        CHECK(!ray_intersects_local(cyl, Ray(Vec3(-3, 0, 0), Vec3(-1, 0, 0))));
        CHECK(ray_intersects_local(cyl, Ray(Vec3(0, 3, 0), normalize(Vec3(0.1f, -1.f, 0.1f)))));
        CHECK(!ray_intersects_local(cyl, Ray(Vec3(0, 3, 0), normalize(Vec3(0.1f, 1.f, 0.1f)))));
        TEST("cylinder.rs test_continuous_1..5");
        CHECK(ray_intersection_local(cyl, Ray(Vec3(-3, 0, 0), Vec3(1, 0, 0)), p)); CHECK_V3(-1.f, 0.f, 0.f, p);
        CHECK(!ray_intersection_local(cyl, Ray(Vec3(-3, 0, 0), Vec3(-1, 0, 0)), p));
        CHECK(ray_intersection_local(cyl, Ray(Vec3(0, 3, 0), normalize(Vec3(0.1f, -1.f, 0.1f))), p)); CHECK_V3(0.1f, 2.f, 0.1f, p);
        CHECK(!ray_intersection_local(cyl, Ray(Vec3(0, 3, 0), normalize(Vec3(0.1f, 1.f, 0.1f))), p));
        CHECK(ray_intersection_local(cyl, Ray(Vec3(0, 3, 0), normalize(Vec3(0.f, -1.f, 0.f))), p)); CHECK_V3(0.f, 2.f, 0.f, p); }
    {   Shape cap = Shape::capsule(2.f, 1.f);
        TEST("capsule.rs test_discrete_1..5");
        CHECK(ray_intersects_local(cap, Ray(Vec3(-3, 0, 0), Vec3(1, 0, 0))));
        CHECK(!ray_intersects_local(cap, Ray(Vec3(-3, 0, 0), Vec3(-1, 0, 0))));
        CHECK(ray_intersects_local(cap, Ray(Vec3(0, 3, 0), normalize(Vec3(0.1f, -1.f, 0.1f)))));
        CHECK(!ray_intersects_local(cap, Ray(Vec3(0, 3, 0), normalize(Vec3(0.1f, 1.f, 0.1f)))));
        CHECK(ray_intersects_local(cap, Ray(Vec3(0, 3, 0), normalize(Vec3(0.f, -1.f, 0.f)))));
        TEST("capsule.rs test_continuous_1..5");
        CHECK(ray_intersection_local(cap, Ray(Vec3(-3, 0, 0), Vec3(1, 0, 0)), p)); CHECK_V3(-1.f, 0.f, 0.f, p);
        CHECK(!ray_intersection_local(cap, Ray(Vec3(-3, 0, 0), Vec3(-1, 0, 0)), p));
        CHECK(ray_intersection_local(cap, Ray(Vec3(0, 4, 0), normalize(Vec3(0.1f, -1.f, 0.1f))), p)); CHECK_V3(0.10102587f, 2.9897413f, 0.10102587f, p);
        CHECK(!ray_intersection_local(cap, Ray(Vec3(0, 3, 0), normalize(Vec3(0.1f, 1.f, 0.1f))), p));
        CHECK(ray_intersection_local(cap, Ray(Vec3(0, 4, 0), normalize(Vec3(0.f, -1.f, 0.f))), p)); CHECK_V3(0.f, 3.f, 0.f, p); }
    {   // polyhedron.rs tests: the unit tetrahedron built with new_with_faces
        const float verts[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
        const uint32_t faces[12] = {1, 3, 2, 3, 1, 0, 2, 0, 1, 0, 2, 3};
        HalfEdgeMesh he; CHECK(he.build(verts, 4, faces, 4));
        Shape hp = Shape::polyhedron(verts, 4); hp.he = &he;
        Shape vp = Shape::polyhedron(verts, 4);
        TEST("polyhedron.rs:533 test_polytope_half_edge");
        CHECK_V3(1.f, 0.f, 0.f, support_point(vp, Vec3(1, 0, 0), I)); CHECK_V3(1.f, 0.f, 0.f, support_point(hp, Vec3(1, 0, 0), I));
        CHECK_V3(0.f, 1.f, 0.f, support_point(vp, Vec3(0, 1, 0), I)); CHECK_V3(0.f, 1.f, 0.f, support_point(hp, Vec3(0, 1, 0), I));
        const Ray down(Vec3(0.25f, 5.f, 0.25f), Vec3(0, -1, 0)), up(Vec3(0.5f, 5.f, 0.5f), Vec3(0, 1, 0));
        TEST("polyhedron.rs:590 test_ray_discrete / :606 test_ray_discrete_transformed");
        CHECK(ray_intersects_local(hp, down)); CHECK(!ray_intersects_local(hp, up));
        CHECK(ray_intersects_transformed(hp, down, I)); CHECK(!ray_intersects_transformed(hp, up, I));
        CHECK(ray_intersects_transformed(hp, down, transform_z(0, 1, 0, 0))); CHECK(ray_intersects_transformed(hp, down, transform_z(0, 0, 0, 0.3f)));
        TEST("polyhedron.rs:629 test_ray_continuous");
        CHECK(ray_intersection_local(hp, down, p));
        CHECK(std::fabs(p.x - 0.25000018f) < F32_EPSILON); CHECK(std::fabs(p.y - 0.4999997f) < F32_EPSILON); CHECK(std::fabs(p.z - 0.25000018f) < F32_EPSILON);
        CHECK(!ray_intersection_local(hp, up, p));
        CHECK(ray_intersection_local(hp, Ray(Vec3(0, 5, 0), Vec3(0, -1, 0)), p)); CHECK_V3(0.f, 1.f, 0.f, p);
        TEST("polyhedron.rs:651 test_ray_continuous_transformed");
        CHECK(ray_intersection_transformed(hp, down, I, p));
        CHECK(std::fabs(p.x - 0.25000018f) < F32_EPSILON); CHECK(std::fabs(p.y - 0.4999997f) < F32_EPSILON); CHECK(std::fabs(p.z - 0.25000018f) < F32_EPSILON);
        CHECK(!ray_intersection_transformed(hp, up, I, p));
        CHECK(ray_intersection_transformed(hp, down, transform_z(0, 1, 0, 0), p));
        CHECK(std::fabs(p.x - 0.25000018f) < F32_EPSILON); CHECK(std::fabs(p.y - 1.4999998f) < F32_EPSILON); CHECK(std::fabs(p.z - 0.25000018f) < F32_EPSILON);
        CHECK(ray_intersection_transformed(hp, down, transform_z(0, 0, 0, 0.3f), p));
        CHECK(std::fabs(p.x - 0.2500002f) < F32_EPSILON); CHECK(std::fabs(p.y - 0.4987077f) < F32_EPSILON); CHECK(std::fabs(p.z - 0.25000012f) < F32_EPSILON);
        TEST("polyhedron.rs:683 test_intersect_face (must not panic)");
        ray_intersection_local(hp, Ray(Vec3(1, -1, 1), Vec3(0, 1, 0)), p); CHECK(true); }
    {   TEST("tests/point.rs:19 test_ray_intersect");
        Vec3 point(1.f, 2.f, 3.f);
        CHECK(point_ray_intersects(point, Ray(Vec3(1, 2, 0), Vec3(0, 0, 1))));
        CHECK(!point_ray_intersects(point, Ray(Vec3(1, 2, 0), Vec3(0, 0, -1))));
        CHECK(!point_ray_intersects(point, Ray(Vec3(1, 0, 0), Vec3(0, 0, 1)))); }
}

// volume::Sphere (the bounding sphere): every test of tests/sphere.rs, the Sphere cases of tests/aabb.rs and tests/frustum.rs.
static void bounding_sphere_tests() {
    {   TEST("tests/sphere.rs:8 test_intersection");
        BSphere sphere(Vec3(0, 0, 0), 1.f);
        Vec3 dir = normalize(Vec3(0, 0, -5));
        Ray r0(Vec3(0, 0, 5), dir), r1(Vec3(std::cos(1.f), 0, 5), dir), r2(Vec3(1, 0, 5), dir), r3(Vec3(2, 0, 5), dir);
        Vec3 p;
        CHECK(sphere.intersection(r0, p)); CHECK_V3(0.f, 0.f, 1.f, p); CHECK(sphere.intersects(r0));
        CHECK(sphere.intersects(r1));
        CHECK(sphere.intersection(r2, p)); CHECK_V3(1.f, 0.f, 0.f, p); CHECK(sphere.intersects(r2));
        CHECK(!sphere.intersection(r3, p)); CHECK(!sphere.intersects(r3)); }
    {   TEST("tests/sphere.rs:50 test_bound");
        Vec3 point(1, 2, 3), normal(0, 0, 1);
        BSphere sphere(point, 1.f);
        CHECK(sphere.relate_plane(Plane::from_point_normal(point, normal)) == Cross);
        CHECK(sphere.relate_plane(Plane::from_point_normal(point + normal * -3.f, normal)) == In);
        CHECK(sphere.relate_plane(Plane::from_point_normal(point + normal * 3.f, normal)) == Out); }
    {   TEST("tests/sphere.rs:73 test_sphere_contains_point");
        BSphere sphere(Vec3(1, 2, 3), 1.f);
        CHECK(sphere.contains(Vec3(1, 2.2f, 3))); CHECK(!sphere.contains(Vec3(11, 2.2f, 3))); }
    {   TEST("tests/sphere.rs:87 test_sphere_contains_line");
        BSphere sphere(Vec3(1, 2, 3), 1.f);
        CHECK(sphere.contains_line(Vec3(1, 2, 3), Vec3(1, 2.2f, 3)));
        CHECK(!sphere.contains_line(Vec3(11, 2, 3), Vec3(11, 2.2f, 3)));
        CHECK(!sphere.contains_line(Vec3(1, 2, 3), Vec3(11, 2.2f, 3))); }
    {   TEST("tests/sphere.rs:103 test_sphere_contains_sphere");
        BSphere sphere(Vec3(1, 2, 3), 1.f);
        CHECK(sphere.contains(BSphere(Vec3(1, 2, 3), 0.1f)));
        CHECK(!sphere.contains(BSphere(Vec3(11, 2, 3), 0.1f)));
        CHECK(!sphere.contains(BSphere(Vec3(1, 2, 3), 10.f))); }
    {   TEST("tests/sphere.rs:128 test_sphere_contains_aabb");
        BSphere sphere(Vec3(1, 2, 3), 1.f);
        CHECK(sphere.contains(Aabb3(Vec3(1, 2, 3), Vec3(1, 2.2f, 3))));
        CHECK(!sphere.contains(Aabb3(Vec3(11, 2, 3), Vec3(11, 2.2f, 3))));
        CHECK(!sphere.contains(Aabb3(Vec3(1, 2, 3), Vec3(11, 2.2f, 3))));
        CHECK(!sphere.contains(Aabb3(Vec3(1, 1.01f, 3), Vec3(1.99f, 2, 3.99f)))); }
    {   TEST("tests/sphere.rs:146 test_sphere_union_sphere");
        BSphere base(Vec3(0, 0, 0), 5.f);
        BSphere u = base.union_(BSphere(Vec3(1, 1, 1), 1.f));
        CHECK_V3(0.f, 0.f, 0.f, u.center); CHECK_EQF(5.f, u.radius);
        u = base.union_(BSphere(Vec3(8, 8, 8), 1.f));
        CHECK_V3(2.8452997f, 2.8452997f, 2.8452997f, u.center); CHECK_EQF(9.928204f, u.radius);
        u = base.union_(BSphere(Vec3(4, 4, 4), 3.f));
        CHECK_V3(1.4226499f, 1.4226499f, 1.4226499f, u.center); CHECK_EQF(7.464102f, u.radius); }
    {   TEST("tests/sphere.rs:183 test_sphere_union_aabb3");
        BSphere base(Vec3(0, 0, 0), 5.f);
        BSphere u = base.union_(Aabb3(Vec3(0, 0, 0), Vec3(2, 2, 2)));
        CHECK_V3(0.f, 0.f, 0.f, u.center); CHECK_EQF(5.f, u.radius);
        u = base.union_(Aabb3(Vec3(6, 6, 6), Vec3(9, 9, 9)));
        CHECK_V3(3.0566242f, 3.0566242f, 3.0566242f, u.center); CHECK_EQF(10.294229f, u.radius);
        u = base.union_(Aabb3(Vec3(4, 4, 4), Vec3(7, 7, 7)));
        CHECK_V3(2.0566242f, 2.0566242f, 2.0566242f, u.center); CHECK_EQF(8.562178f, u.radius); }
    {   TEST("tests/sphere.rs:213 test_surface_area");
        CHECK_EQF(4.f * 3.14159265358979323846f * 2.f * 2.f, BSphere(Vec3(2, 1, -4), 2.f).surface_area()); }
    {   TEST("tests/aabb.rs:270 test_aabb3_contains_sphere");
        Aabb3 aabb(Vec3(0, 0, 0), Vec3(10, 10, 10));
        CHECK(aabb_contains_sphere(aabb, BSphere(Vec3(5, 5, 5), 1.f)));
        CHECK(!aabb_contains_sphere(aabb, BSphere(Vec3(20, 20, 20), 1.f)));
        CHECK(!aabb_contains_sphere(aabb, BSphere(Vec3(5, 5, 5), 10.f))); }
    {   TEST("tests/aabb.rs:329 test_aabb3_union_sphere");
        Aabb3 base(Vec3(0, 0, 0), Vec3(10, 10, 10));
        Aabb3 u = aabb_union_sphere(base, BSphere(Vec3(5, 5, 5), 1.f));
        CHECK_V3(0.f, 0.f, 0.f, u.min); CHECK_V3(10.f, 10.f, 10.f, u.max);
        u = aabb_union_sphere(base, BSphere(Vec3(14, 14, 14), 1.f));
        CHECK_V3(0.f, 0.f, 0.f, u.min); CHECK_V3(15.f, 15.f, 15.f, u.max);
        u = aabb_union_sphere(base, BSphere(Vec3(8, 8, 8), 4.f));
        CHECK_V3(0.f, 0.f, 0.f, u.min); CHECK_V3(12.f, 12.f, 12.f, u.max); }
    {   TEST("tests/frustum.rs:8 test_contains");
        // As written the test builds a 1-DEGREE frustum (`1_f32.to_radians()`), in which a unit sphere 5 units away cannot be `In`:
        // the reference's code gives Cross / Out / Cross for the three spheres, so the test cannot pass as written (third documented
        // discrepancy). With a field of view of 1 RADIAN the same code gives exactly the expected In / Cross / Out: pinned on that.
        Frustum f; CHECK(Frustum::from_mat4(Mat4::perspective_rh_gl(to_radians(1.f), 1.f, 1.f, 10.f), f));
        const char* names[3] = {"In", "Cross", "Out"};
        std::printf("DISCREPANCY tests/frustum.rs:8 test_contains expects In, Cross, Out for a 1-degree frustum; the reference's code gives %s, %s, %s\n",
                    names[f.contains(BSphere(Vec3(0, 0, -5), 1.f))], names[f.contains(BSphere(Vec3(0, 3, -5), 1.f))], names[f.contains(BSphere(Vec3(0, 0, 5), 1.f))]);
        Frustum g; CHECK(Frustum::from_mat4(Mat4::perspective_rh_gl(1.f, 1.f, 1.f, 10.f), g));
        CHECK(g.contains(BSphere(Vec3(0, 0, -5), 1.f)) == In);
        CHECK(g.contains(BSphere(Vec3(0, 3, -5), 1.f)) == Cross);
        CHECK(g.contains(BSphere(Vec3(0, 0, 5), 1.f)) == Out); }
    {   TEST("SweepAndPrune3<volume::Sphere>: touching spheres are reported, separated ones are not (sphere.rs:82-91)");
        std::vector<BSphere> b = {BSphere(Vec3(5, 0, 0), 0.5f), BSphere(Vec3(0, 0, 0), 1.f), BSphere(Vec3(2, 0, 0), 1.f), BSphere(Vec3(2, 2.1f, 0), 1.f)};
        SweepAndPrune3Sphere sap;
        std::vector<uint32_t> perm;
        auto pairs = sap.find_collider_pairs(b, &perm);
        // sorted by centre.x - r: (0,0,0) r1 [-1], (2,0,0) r1 [1], (2,2.1,0) r1 [1, same max], (5,0,0) r.5 [4.5]
        CHECK(perm.size() == 4 && perm[0] == 1 && perm[1] == 2 && perm[2] == 3 && perm[3] == 0);
        CHECK(pairs.size() == 1 && pairs[0].first == 0 && pairs[0].second == 1); }
}

int main() {
    bounding_sphere_tests();
    simplex_tests();
    epa_tests();
    gjk_tests();
    support_tests();
    aabb_tests();
    plane_point_tests();
    dbvt_tests();
    bound_tests();
    ray_tests();
    std::printf("KAT summary: %d checks passed, %d failed\n", g_pass, g_fail);
    return g_fail == 0 ? 0 : 1;
}
