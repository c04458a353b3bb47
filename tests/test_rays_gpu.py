"""GPU parity tests for the ray casts against primitives: the reference's own ray tests through the
GPU API with their exact expected values, then differential tests against the CPU oracle, bit for bit."""
import os

import numpy as np
import pytest

import bam3d_b200 as b3
from bam3d_b200 import scenes
from tests.golden import make_golden
from tests.test_narrow_gpu import assert_f32_equal, shapes_of

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def unit(x, y, z):
    v = np.array([x, y, z], dtype=np.float32)
    return v * (np.float32(1.0) / np.sqrt(np.float32((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]), dtype=np.float32))


def test_kat_sphere_and_cuboid_rays(ctx, oracle):
    """sphere.rs:134-183, cuboid.rs test_ray_* (incl. the 0.3 degree rotation expecting x = 5.000068)"""
    I = oracle.transform_z(0., 0., 0., 0.)
    sph = b3.Sphere(10.)
    hit = b3.Ray([20., 0., 0.], [-1., 0., 0.])
    assert b3.intersects_transformed(sph, hit, I, ctx)
    assert not b3.intersects_transformed(sph, b3.Ray([20., 0., 0.], [1., 0., 0.]), I, ctx)
    assert not b3.intersects_transformed(sph, hit, oracle.transform_z(0., 15., 0., 0.), ctx)
    assert b3.intersection_transformed(sph, hit, I, ctx).tolist() == [10., 0., 0.]
    assert b3.intersection_transformed(sph, b3.Ray([20., 0., 0.], [1., 0., 0.]), I, ctx) is None
    assert b3.intersection_transformed(sph, hit, oracle.transform_z(0., 15., 0., 0.), ctx) is None
    cub = b3.Cuboid(10., 10., 10.)
    r = b3.Ray([10., 0., 0.], [-1., 0., 0.])
    t = oracle.transform_z(0., 1., 0., 0.)
    assert b3.intersects_transformed(cub, r, t, ctx) and not b3.intersects_transformed(cub, b3.Ray([10., 0., 0.], [1., 0., 0.]), t, ctx)
    assert b3.intersects_transformed(cub, r, oracle.transform_z(0., 1., 0., 0.3), ctx)
    assert b3.intersection_transformed(cub, r, t, ctx).tolist() == [5., 0., 0.]
    p = b3.intersection_transformed(cub, r, oracle.transform_z(0., 0., 0., 0.3), ctx)
    eps = np.finfo(np.float32).eps
    assert abs(p[0] - np.float32(5.000068)) < eps and abs(p[1]) < eps and abs(p[2]) < eps


def test_kat_cylinder_and_capsule_rays(ctx, oracle):
    """cylinder.rs test_discrete_1..4 / test_continuous_1..5, capsule.rs test_discrete_1..5 / test_continuous_1..5"""
    I = oracle.transform_z(0., 0., 0., 0.)
    cyl, cap = b3.Cylinder(2., 1.), b3.Capsule(2., 1.)
    side_in, side_out = b3.Ray([-3., 0., 0.], [1., 0., 0.]), b3.Ray([-3., 0., 0.], [-1., 0., 0.])
    down, up = b3.Ray([0., 3., 0.], unit(0.1, -1., 0.1)), b3.Ray([0., 3., 0.], unit(0.1, 1., 0.1))
    for prim in (cyl, cap):
        assert b3.intersects_transformed(prim, side_in, I, ctx) and not b3.intersects_transformed(prim, side_out, I, ctx)
        assert b3.intersects_transformed(prim, down, I, ctx) and not b3.intersects_transformed(prim, up, I, ctx)
        assert b3.intersection_transformed(prim, side_in, I, ctx).tolist() == [-1., 0., 0.]
        assert b3.intersection_transformed(prim, side_out, I, ctx) is None and b3.intersection_transformed(prim, up, I, ctx) is None
    assert b3.intersects_transformed(cap, b3.Ray([0., 3., 0.], [0., -1., 0.]), I, ctx)
    assert b3.intersection_transformed(cyl, down, I, ctx).tolist() == [np.float32(0.1), 2., np.float32(0.1)]
    assert b3.intersection_transformed(cyl, b3.Ray([0., 3., 0.], [0., -1., 0.]), I, ctx).tolist() == [0., 2., 0.]
    p = b3.intersection_transformed(cap, b3.Ray([0., 4., 0.], unit(0.1, -1., 0.1)), I, ctx)
    assert p.tolist() == [np.float32(0.10102587), np.float32(2.9897413), np.float32(0.10102587)]
    assert b3.intersection_transformed(cap, b3.Ray([0., 4., 0.], [0., -1., 0.]), I, ctx).tolist() == [0., 3., 0.]


@pytest.mark.parametrize("query", [b3.RAY_INTERSECTS, b3.RAY_INTERSECTION, b3.PARTICLE_SWEEP])
def test_ray_cast_parity_all_kinds(ctx, oracle, query):
    """Every analytic primitive under random rigid transforms; aimed, random, axis-parallel and unnormalised rays."""
    scene = scenes.all_kinds_pairs(50_000, s=1.5)
    rays, segments, casts = scenes.primitive_rays(scene, 400_000)
    src = segments if query == b3.PARTICLE_SWEEP else rays
    pts, st = b3.ray_cast_batch(shapes_of(ctx, scene), src, casts, query)
    opts, ost = oracle.ray_cast(scene, src, casts, query, nthreads=8)
    assert np.array_equal(st, ost), f"{int((st != ost).sum())} status mismatches"
    assert 0.02 < (st == 0).mean() < 0.9
    for k in range(6):  # every kind both hits and misses
        m = scene["kind"][casts[:, 1]] == k
        assert (st[m] == 0).any() and (st[m] == 1).any(), k
    assert_f32_equal("hit point", pts, opts)


def test_ray_cast_vertex_only_polyhedron_unsupported_and_empty(ctx):
    lib = scenes.mesh_library(4)
    ps = scenes.polyhedron_pairs(8, library=lib)
    shapes = b3.ShapeSet(ctx, ps["kind"], ps["params"], ps["xform"], ps["mesh_id"], b3.MeshLibrary(ctx, *lib))
    pts, st = b3.ray_cast_batch(shapes, [[0, 0, 0, 1, 0, 0]], [[0, 0], [0, 3]])
    assert (st == b3.STATUS_UNSUPPORTED).all() and not pts.any()
    pts, st = b3.ray_cast_batch(shapes, np.zeros((0, 6), np.float32), np.zeros((0, 2), np.uint32))
    assert len(st) == 0
    with pytest.raises(b3.Bam3dError):
        b3.ray_cast_batch(shapes, [[0, 0, 0, 1, 0, 0]], [[1, 0]])  # ray index out of range


def test_rays_against_golden(ctx):
    sc, rays, segments, casts = make_golden.scenes_for_rays()
    g = np.load(os.path.join(HERE, "golden", "rays_golden.npz"))
    shapes = shapes_of(ctx, sc)
    for name, q, r in (("intersects", 0, rays), ("intersection", 1, rays), ("sweep", 2, segments)):
        pts, st = b3.ray_cast_batch(shapes, r, casts, q)
        assert np.array_equal(st, g[name + "_status"])
        assert_f32_equal(name, pts, g[name + "_points"])
