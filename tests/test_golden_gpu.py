"""GPU path against the COMMITTED golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py): these
do not need the oracle at run time."""
import os

import numpy as np
import pytest

import bam3d_b200 as b3
from bam3d_b200 import broad
from tests.golden import make_golden

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def same(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == "f":
        return a.shape == b.shape and bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())
    return np.array_equal(a, b)


def test_narrow_against_golden(ctx):
    narrow, _, _, _ = make_golden.scenes_for_golden()
    g = np.load(os.path.join(HERE, "golden", "narrow_golden.npz"))
    meshes = b3.MeshLibrary(ctx, narrow["mesh_verts"], narrow["mesh_offsets"])
    s0 = b3.ShapeSet(ctx, narrow["kind"], narrow["params"], narrow["xform"], narrow["mesh_id"], meshes)
    s1 = b3.ShapeSet(ctx, narrow["kind"], narrow["params"], narrow["xform_end"], narrow["mesh_id"], meshes)
    gjk = b3.GJK.new(ctx)
    assert same(gjk.intersect_batch(s0, narrow["pairs"]), g["intersect_status"])
    c, st = gjk.intersection_batch(b3.CollisionStrategy.FullResolution, s0, narrow["pairs"])
    assert same(st, g["contact_status"]) and same(c[:, :7], g["contact"][:, :7])
    ref, dist, dst = gjk.distance_batch(s0, narrow["pairs"])
    assert same(dst, g["distance_status"]) and same(ref, g["distance_ref"]) and same(dist, g["distance_geo"])
    tc, tst = gjk.intersection_time_of_impact_batch(s0, s1, narrow["pairs"])
    hit = tst == 0
    assert same(tst, g["toi_status"]) and same(tc[hit][:, [0, 1, 2, 3, 7]], g["toi"][hit][:, [0, 1, 2, 3, 7]])
    assert np.allclose(tc[hit][:, 4:7], g["toi"][hit][:, 4:7], rtol=1e-4, atol=1e-4, equal_nan=True)
    assert same(broad.world_aabbs(s0), g["world_aabbs"])


def test_broad_against_golden(ctx):
    _, boxes, rays, mats = make_golden.scenes_for_golden()
    g = np.load(os.path.join(HERE, "golden", "broad_golden.npz"))
    sap = broad.SweepAndPrune3.new(ctx)
    pairs, perm = sap.find_collider_pairs(boxes)
    assert same(pairs, g["sap_pairs"]) and same(perm, g["sap_perm"]) and sap.get_sweep_axis() == int(g["sap_next_axis"][0])
    bvh = broad.Bvh(ctx, boxes)
    off, vals, pts = bvh.query_ray(rays)
    assert same(off, g["ray_offsets"])
    for q in range(len(rays)):
        a, b = int(off[q]), int(off[q + 1])
        o = np.argsort(vals[a:b], kind="stable")
        assert same(vals[a:b][o], g["ray_values"][a:b]) and same(pts[a:b][o], g["ray_points"][a:b])
    planes = np.stack([broad.frustum_from_mat4(m).reshape(24) for m in mats])
    assert same(planes, g["frustum_planes"])
    off, vals, rel = bvh.query_frustum(planes)
    assert same(off, g["frustum_offsets"])
    for q in range(len(mats)):
        a, b = int(off[q]), int(off[q + 1])
        o = np.argsort(vals[a:b], kind="stable")
        assert same(vals[a:b][o], g["frustum_values"][a:b]) and same(rel[a:b][o], g["frustum_relation"][a:b])
    dirty = (np.arange(len(boxes)) % 3 == 0).astype(np.uint8)
    assert same(bvh.find_collider_pairs(dirty), g["collider_pairs"])


def test_sphere_bounds_against_golden(ctx):
    narrow, _, _, _ = make_golden.scenes_for_golden()
    g = np.load(os.path.join(HERE, "golden", "spheres_golden.npz"))
    meshes = b3.MeshLibrary(ctx, narrow["mesh_verts"], narrow["mesh_offsets"])
    s0 = b3.ShapeSet(ctx, narrow["kind"], narrow["params"], narrow["xform"], narrow["mesh_id"], meshes)
    assert same(broad.world_spheres(s0), g["world_spheres"])
    for axis in (0, 1, 2):
        sap = broad.SweepAndPrune3Sphere.with_sweep_axis(axis, ctx)
        pairs, perm = sap.find_collider_pairs(g["sap_spheres"])
        assert len(g[f"pairs_axis{axis}"]) > 500
        assert same(pairs, g[f"pairs_axis{axis}"]) and same(perm, g[f"perm_axis{axis}"]) and sap.get_sweep_axis() == int(g[f"next_axis{axis}"][0])
