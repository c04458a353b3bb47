"""GPU parity tests for ConvexPolyhedron::new_with_faces (SURVEY.md 8f-4 and the polyhedron part of 8f-1): hill-climbing
support over the half-edge structure inside GJK / EPA / distance / TOI, and the ray walk over the faces — against the CPU
oracle's restatement of primitive/polyhedron.rs, plus the reference's own polyhedron tests through the GPU API."""
import numpy as np
import pytest

import bam3d_b200 as b3
from bam3d_b200 import scenes
from tests.test_narrow_gpu import assert_f32_equal

pytestmark = pytest.mark.gpu

TETRA_V = [[1., 0., 0.], [0., 1., 0.], [0., 0., 1.], [0., 0., 0.]]
TETRA_F = [(1, 3, 2), (3, 1, 0), (2, 0, 1), (0, 2, 3)]


@pytest.fixture
def faced_library(oracle):
    """A mesh library where every second mesh has faces (HalfEdge mode) and the others stay VertexOnly."""
    verts, offsets = scenes.mesh_library(64, vmin=8, vmax=200, seed=0xB3D0FACE)
    faces, foffs = scenes.mesh_faces(verts, offsets)
    keep = np.arange(len(offsets) - 1) % 2 == 0
    sel, new_foffs = [], [0]
    for m in range(len(offsets) - 1):
        if keep[m]:
            sel.append(faces[foffs[m]:foffs[m + 1]])
        new_foffs.append(new_foffs[-1] + (int(foffs[m + 1] - foffs[m]) if keep[m] else 0))
    faces, foffs = np.concatenate(sel), np.array(new_foffs, dtype=np.uint32)
    oracle.set_mesh_faces(verts, offsets, faces, foffs)
    yield verts, offsets, faces, foffs
    oracle.set_mesh_faces(None, None, None, None)


def test_kat_polyhedron_half_edge_and_rays(ctx, oracle):
    """polyhedron.rs:533-700: test_polytope_half_edge, test_ray_discrete(_transformed), test_ray_continuous(_transformed)"""
    I = oracle.transform_z(0., 0., 0., 0.)
    eps = np.finfo(np.float32).eps
    poly = b3.ConvexPolyhedron.new_with_faces(TETRA_V, TETRA_F)
    down, up = b3.Ray([0.25, 5., 0.25], [0., -1., 0.]), b3.Ray([0.5, 5., 0.5], [0., 1., 0.])
    assert b3.intersects_transformed(poly, down, I, ctx) and not b3.intersects_transformed(poly, up, I, ctx)
    assert b3.intersects_transformed(poly, down, oracle.transform_z(0., 1., 0., 0.), ctx)
    assert b3.intersects_transformed(poly, down, oracle.transform_z(0., 0., 0., 0.3), ctx)
    for t, want in ((I, (0.25000018, 0.4999997, 0.25000018)), (oracle.transform_z(0., 1., 0., 0.), (0.25000018, 1.4999998, 0.25000018)),
                    (oracle.transform_z(0., 0., 0., 0.3), (0.2500002, 0.4987077, 0.25000012))):
        p = b3.intersection_transformed(poly, down, t, ctx)
        assert all(abs(p[k] - np.float32(want[k])) < eps for k in range(3)), (p, want)
    assert b3.intersection_transformed(poly, up, I, ctx) is None
    assert b3.intersection_transformed(poly, b3.Ray([0., 5., 0.], [0., -1., 0.]), I, ctx).tolist() == [0., 1., 0.]
    b3.intersection_transformed(poly, b3.Ray([1., -1., 1.], [0., 1., 0.]), I, ctx)  # test_intersect_face: must not fail
    # a VertexOnly polyhedron has no faces to walk
    with pytest.raises(NotImplementedError):
        b3.intersects_transformed(b3.ConvexPolyhedron.new(TETRA_V), down, I, ctx)
    # support points through GJK: the two modes agree on the tetrahedron (test_polytope_half_edge)
    gjk = b3.GJK.new(ctx)
    sph = b3.Sphere(0.25)
    for prim in (poly, b3.ConvexPolyhedron.new(TETRA_V)):
        assert gjk.intersect(prim, I, sph, oracle.transform_z(1.1, 0., 0., 0.))
        assert not gjk.intersect(prim, I, sph, oracle.transform_z(1.3, 0., 0., 0.))
    # a degenerate face: the reference panics in Plane::from_points(..).unwrap()
    with pytest.raises(b3.Bam3dError):
        b3.MeshLibrary(ctx, TETRA_V, [0, 4], [(0, 0, 1)], [0, 1])


def test_half_edge_narrow_phase_parity(ctx, oracle, faced_library):
    """intersect / contact / distance / TOI with hill-climbing supports (HalfEdge meshes) and brute-force supports
    (VertexOnly meshes) mixed in one library, polyhedron-polyhedron pairs (warp kernels)."""
    verts, offsets, faces, foffs = faced_library
    ps = scenes.polyhedron_pairs(30_000, library=(verts, offsets), s=1.6)
    meshes = b3.MeshLibrary(ctx, verts, offsets, faces, foffs)
    shapes = b3.ShapeSet(ctx, ps["kind"], ps["params"], ps["xform"], ps["mesh_id"], meshes)
    gjk = b3.GJK.new(ctx)
    hit, ost = oracle.intersect(ps, nthreads=8)
    assert np.array_equal(gjk.intersect_batch(shapes, ps["pairs"]), ost)
    c, st = gjk.intersection_batch(b3.CollisionStrategy.FullResolution, shapes, ps["pairs"])
    oc, _, ocst = oracle.contact(ps, strategy=0, nthreads=8)
    assert np.array_equal(st, ocst) and (st == 0).sum() > 3000
    assert_f32_equal("contact", c[:, :7], oc[:, :7], st == 0)
    ref, dist, dst = gjk.distance_batch(shapes, ps["pairs"])
    oref, ogeo, _, odst = oracle.distance(ps, nthreads=8)
    assert np.array_equal(dst, odst)
    assert_f32_equal("distance", np.stack([ref, dist], 1), np.stack([oref, ogeo], 1), dst == 0)
    x1 = scenes.end_transforms(ps)
    shapes_end = b3.ShapeSet(ctx, ps["kind"], ps["params"], x1, ps["mesh_id"], meshes)
    tc, tst = gjk.intersection_time_of_impact_batch(shapes, shapes_end, ps["pairs"])
    otc, _, otst = oracle.toi(ps, x1, nthreads=8)
    assert np.array_equal(tst, otst)
    assert_f32_equal("toi", tc[:, [0, 1, 2, 3, 7]], otc[:, [0, 1, 2, 3, 7]], tst == 0)
    # the modes really differ somewhere: with every mesh VertexOnly some answers change (hill climbing stops on plateaus)
    oracle.set_mesh_faces(None, None, None, None)
    oc2, _, ocst2 = oracle.contact(ps, strategy=0, nthreads=8)
    assert not np.array_equal(oc2[:, :7], oc[:, :7]) or not np.array_equal(ocst2, ocst)


def test_half_edge_against_analytic_shapes_thread_path(ctx, oracle, faced_library):
    """Polyhedra against spheres / cuboids / cylinders in one batch: the binned thread + warp plan with both polyhedron modes."""
    verts, offsets, faces, foffs = faced_library
    n = 20_000
    a = scenes.mixed_pairs(n, s=1.4, seed=0xB3D0FAC1)
    p = scenes.polyhedron_pairs(n, library=(verts, offsets), s=1.4, seed=0xB3D0FAC1)
    rng = np.random.default_rng(5)
    scene = dict(a)
    scene["kind"] = np.where(rng.random(2 * n) < 0.5, p["kind"], a["kind"]).astype(np.uint8)
    scene["mesh_id"], scene["mesh_verts"], scene["mesh_offsets"] = p["mesh_id"], verts, offsets
    meshes = b3.MeshLibrary(ctx, verts, offsets, faces, foffs)
    shapes = b3.ShapeSet(ctx, scene["kind"], scene["params"], scene["xform"], scene["mesh_id"], meshes)
    c, st = b3.GJK.new(ctx).intersection_batch(b3.CollisionStrategy.FullResolution, shapes, scene["pairs"])
    oc, _, ost = oracle.contact(scene, strategy=0, nthreads=8)
    assert np.array_equal(st, ost)
    assert_f32_equal("contact", c[:, :7], oc[:, :7], st == 0)


@pytest.mark.parametrize("query", [b3.RAY_INTERSECTS, b3.RAY_INTERSECTION, b3.PARTICLE_SWEEP])
def test_polyhedron_ray_cast_parity(ctx, oracle, faced_library, query):
    """The face walk (find_intersecting_face / next_face_classify) on hulls with 12..396 faces, both below and above the
    10-face switch of next_face_classify; VertexOnly meshes report unsupported."""
    verts, offsets, faces, foffs = faced_library
    ps = scenes.polyhedron_pairs(4000, library=(verts, offsets), s=1.0)
    small = np.array(TETRA_V, dtype=np.float32)  # 4 faces: the `faces.len() < 10` branch
    verts2 = np.concatenate([verts, small])
    offsets2 = np.concatenate([offsets, [offsets[-1] + 4]]).astype(np.uint32)
    faces2 = np.concatenate([faces, np.array(TETRA_F, dtype=np.uint32)])
    foffs2 = np.concatenate([foffs, [foffs[-1] + 4]]).astype(np.uint32)
    oracle.set_mesh_faces(verts2, offsets2, faces2, foffs2)
    scene = dict(ps)
    scene["mesh_verts"], scene["mesh_offsets"] = verts2, offsets2
    scene["mesh_id"] = ps["mesh_id"].copy()
    scene["mesh_id"][::7] = len(offsets2) - 2  # the tetrahedron
    rays, segments, casts = scenes.primitive_rays(scene, 200_000, seed=0xB3D0FAC2)
    src = segments if query == b3.PARTICLE_SWEEP else rays
    meshes = b3.MeshLibrary(ctx, verts2, offsets2, faces2, foffs2)
    shapes = b3.ShapeSet(ctx, scene["kind"], scene["params"], scene["xform"], scene["mesh_id"], meshes)
    pts, st = b3.ray_cast_batch(shapes, src, casts, query)
    opts, ost = oracle.ray_cast(scene, src, casts, query, nthreads=8)
    assert np.array_equal(st, ost), f"{int((st != ost).sum())} status mismatches"
    assert (st == 0).sum() > 5000 and (st == 1).sum() > 5000 and (st == b3.STATUS_UNSUPPORTED).sum() > 5000
    assert_f32_equal("hit point", pts, opts)
