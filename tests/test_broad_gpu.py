"""GPU parity tests for the broad phase: device sweep-and-prune, BVH queries and world AABBs vs the CPU oracle.
Pair sets, permutations, sweep axes, hit sets and Relation codes are integer work: bit-exact. Ray hit points are f32
values produced by the reference's slab expressions evaluated in the same order: compared bit-for-bit as well."""
import numpy as np
import pytest

import bam3d_b200 as b3
from bam3d_b200 import broad, multi, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(params=[0, 1, 2], ids=["auto", "sweep", "bvh"])
def sap_strategy(request, ctx):
    """Every sweep-and-prune test runs under all three pair-search strategies: the output must not depend on it."""
    ctx.set_option("sap_strategy", request.param)
    yield request.param
    ctx.set_option("sap_strategy", 0)


def test_sap_parity_and_axis_sequence(ctx, oracle, sap_strategy):
    """C4 density at a size the O(n * active) CPU sweep finishes in seconds; three consecutive calls so that the
    Variance3 axis hand-over (sweep_prune.rs:123-125) is exercised on every axis."""
    boxes, L = scenes.aabb_scene(60_000)
    sap = broad.SweepAndPrune3.new(ctx)
    axis = 0
    for _ in range(3):
        pairs, perm = sap.find_collider_pairs(boxes)
        opairs, operm, oaxis = oracle.sap_find_pairs(boxes, axis)
        assert np.array_equal(perm, operm)
        assert len(pairs) == len(opairs) > 100_000
        assert np.array_equal(pairs, opairs)          # same pairs in the same (j asc, i asc) order
        assert sap.get_sweep_axis() == oaxis
        axis = oaxis


def test_variance3_sums_bit_exact(ctx, oracle):
    """The running sums of Variance3 are f32 accumulated sequentially in the reference; the device computes the same
    bits with per-tile integer arithmetic inside a binade and falls back to the sequential chain at binade crossings
    and exact round-half ties. Scenes: C4-like (sums cross ~25 binades), negative and mixed-sign coordinates,
    half-integer coordinates (every addition beyond 2^24 is an exact tie), huge / tiny magnitudes, short inputs."""
    rng = np.random.default_rng(11)
    cases = {}
    cases["c4"] = scenes.aabb_scene(1 << 20)[0]
    cases["negative"] = -scenes.aabb_scene(300_000, seed=0xB3D0C402)[0][:, [3, 4, 5, 0, 1, 2]]
    mixed = scenes.aabb_scene(300_000, seed=0xB3D0C403)[0] - np.float32(50.0)
    cases["mixed_sign"] = mixed.astype(np.float32)
    half = rng.integers(0, 64, size=(400_000, 3)).astype(np.float32)
    cases["half_integers"] = np.concatenate([half, half + np.float32(1.0)], axis=1)          # centres k + 0.5
    cases["huge_and_tiny"] = (rng.random((200_000, 6)) * np.where(rng.random((200_000, 1)) < 0.5, 1e-12, 1e12)).astype(np.float32)
    cases["short"] = scenes.aabb_scene(5000)[0][:37]
    cases["one_tile_boundary"] = scenes.aabb_scene(8192)[0][:4097]
    for name, boxes in cases.items():
        boxes = np.ascontiguousarray(boxes, dtype=np.float32)
        sums, axis = broad.variance3(ctx, boxes)
        osums, oaxis = oracle.variance3(boxes)
        assert sums.tobytes() == osums.tobytes(), (name, sums, osums)
        assert axis == oaxis, name


@pytest.mark.parametrize("n_shards", [2, 3, 8])
def test_sap_shards_concatenate_to_the_full_list(ctx, oracle, sap_strategy, n_shards):
    """Multi-GPU form of the sweep: every shard sorts everything (same perm, same next axis) and searches
    the pairs whose higher sorted slot lies in its slice; concatenated in shard order they are the oracle's list."""
    boxes, L = scenes.aabb_scene(40_000)
    opairs, operm, oaxis = oracle.sap_find_pairs(boxes, 0)
    got = []
    for shard in range(n_shards):
        sap = broad.SweepAndPrune3.new(ctx)
        pairs, perm = sap.find_collider_pairs(boxes, shard=shard, n_shards=n_shards)
        lo, hi = multi.sap_shard_range(len(boxes), shard, n_shards)
        assert np.array_equal(perm, operm) and sap.get_sweep_axis() == oaxis
        assert len(pairs) == 0 or (pairs[:, 1].min() >= lo and pairs[:, 1].max() < hi)
        got.append(pairs)
    assert np.array_equal(np.concatenate(got), opairs)


@pytest.mark.parametrize("scale", [(1.0, 1.0, 1.0), (1.0, 30.0, 1.0), (0.05, 1.0, 40.0)])
def test_sap_anisotropic_axes(ctx, oracle, scale, sap_strategy):
    boxes, L = scenes.aabb_scene(20_000, seed=0xB3D0C401)
    s = np.array(scale * 2, dtype=np.float32)
    boxes = (boxes * s).astype(np.float32)
    for start_axis in (0, 1, 2):
        sap = broad.SweepAndPrune3.with_sweep_axis(start_axis, ctx)
        pairs, perm = sap.find_collider_pairs(boxes)
        opairs, operm, oaxis = oracle.sap_find_pairs(boxes, start_axis)
        assert np.array_equal(perm, operm) and np.array_equal(pairs, opairs) and sap.get_sweep_axis() == oaxis


def test_sap_edge_cases(ctx, oracle, sap_strategy):
    sap = broad.SweepAndPrune3.new(ctx)
    # empty and single: no pairs, axis untouched (sweep_prune.rs:76-78)
    p, perm = sap.find_collider_pairs(np.zeros((0, 6), dtype=np.float32))
    assert len(p) == 0 and len(perm) == 0 and sap.get_sweep_axis() == 0
    p, perm = sap.find_collider_pairs([[0, 0, 0, 1, 1, 1]])
    assert len(p) == 0 and sap.get_sweep_axis() == 0
    # touching boxes are not pairs (strict overlap, aabb3.rs:237-244); identical boxes are; -0.0 == +0.0 in the sort key
    boxes = np.array([[0, 0, 0, 1, 1, 1], [1, 0, 0, 2, 1, 1], [0.5, 0.5, 0.5, 1.5, 1.5, 1.5], [0, 0, 0, 1, 1, 1],
                      [-0.0, 2, 2, 1, 3, 3], [0.0, 2, 2, 1, 3, 3], [5, 5, 5, 5, 5, 5], [0, 0, 1, 1, 1, 2]], dtype=np.float32)
    for axis in (0, 1, 2):
        sap = broad.SweepAndPrune3.with_sweep_axis(axis, ctx)
        pairs, perm = sap.find_collider_pairs(boxes)
        opairs, operm, oaxis = oracle.sap_find_pairs(boxes, axis)
        assert np.array_equal(perm, operm), (axis, perm, operm)
        assert np.array_equal(pairs, opairs)
        assert sap.get_sweep_axis() == oaxis
    # many duplicates: the sort must be stable
    rng = np.random.default_rng(5)
    base = rng.integers(0, 4, size=(5000, 3)).astype(np.float32)
    boxes = np.concatenate([base, base + 1.5], axis=1)
    pairs, perm = broad.SweepAndPrune3.new(ctx).find_collider_pairs(boxes, capacity=16)  # forces the capacity retry
    opairs, operm, _ = oracle.sap_find_pairs(boxes, 0)
    assert np.array_equal(perm, operm) and np.array_equal(pairs, opairs)


def test_world_aabbs_parity(ctx, oracle):
    lib = scenes.mesh_library(32, seed=0xB3D0C402)
    a = scenes.all_kinds_pairs(20_000, seed=0xB3D0C403)
    p = scenes.polyhedron_pairs(20_000, library=lib, seed=0xB3D0C403)
    rng = np.random.default_rng(1)
    scene = dict(a)
    scene["kind"] = np.where(rng.random(40_000) < 0.2, p["kind"], a["kind"]).astype(np.uint8)
    scene["mesh_id"], (scene["mesh_verts"], scene["mesh_offsets"]) = p["mesh_id"], lib
    meshes = b3.MeshLibrary(ctx, *lib)
    shapes = b3.ShapeSet(ctx, scene["kind"], scene["params"], scene["xform"], scene["mesh_id"], meshes)
    got = broad.world_aabbs(shapes)
    want = oracle.world_aabbs(scene)
    assert np.array_equal(got, want)


def _split(offsets, *arrays):
    return [tuple(a[int(offsets[q]):int(offsets[q + 1])] for a in arrays) for q in range(len(offsets) - 1)]


@pytest.fixture(params=[1, 2], ids=["traversal", "allpairs"])
def bvh_strategy(request, ctx):
    ctx.set_option("bvh_strategy", request.param)
    yield request.param
    ctx.set_option("bvh_strategy", 0)


def test_bvh_queries_parity(ctx, oracle, bvh_strategy):
    boxes, L = scenes.aabb_scene(40_000, seed=0xB3D0C404)
    bvh = broad.Bvh(ctx, boxes)
    ot = oracle.dbvt_build(boxes)
    # DiscreteVisitor
    qboxes, _ = scenes.aabb_scene(300, seed=0xB3D0C405, extent=L)
    qboxes[:, 3:] += 1.0
    off, vals = bvh.query_aabb(qboxes)
    for q, (v,) in enumerate(_split(off, vals)):
        assert np.array_equal(np.sort(v), ot.query_aabb(qboxes[q]))
    # ContinuousVisitor<Ray> == query_ray: hit sets and the slab entry points
    rays = scenes.rays(300, L, seed=0xB3D0C406)
    rays[0, 3:] = [1, 0, 0]   # axis-parallel rays exercise the inf/NaN slab arithmetic (tests/aabb.rs:62-81)
    rays[1, 3:] = [0, -1, 0]
    off, vals, pts = bvh.query_ray(rays)
    total = 0
    for q, (v, p) in enumerate(_split(off, vals, pts)):
        ov, op = ot.query_ray(rays[q])
        order = np.argsort(v, kind="stable")
        assert np.array_equal(v[order], ov)
        assert np.array_equal(p[order], op)
        total += len(v)
    assert total > 1000
    # query_ray_closest. Device: the minimum over ALL hit leaves of t = (point - origin) . direction, ties -> lowest value
    # index — a definition that does not depend on the tree, checked bit for bit against the brute-force minimum below.
    # Reference (dbvt/util.rs:46-62,77-103, restated by the oracle over ITS tree: values inserted in order, no rotations):
    # a DFS that drops a node when its own `t` is not below the best LEAF t so far. For a bound that contains the ray's
    # origin, `t` is the distance to where the ray LEAVES it (aabb3.rs:165-195), so such a branch can be dropped although
    # a closer leaf sits inside. Exactly two ways the two answers can differ, both asserted here:
    #   (1) same t, other value: an exact tie in t (the reference keeps the first in its traversal order);
    #   (2) larger t: a branch above the true closest leaf contains the ray's origin (and was pruned by its exit distance).
    # For every other ray — in particular every ray that starts outside the tree's root bound — they are identical.
    val, pt = bvh.query_ray_closest(rays)
    f32 = np.float32
    n_same = n_tie = n_pruned = 0
    for q in range(len(rays)):
        ov, op = ot.query_ray(rays[q])
        if len(ov) == 0:
            assert val[q] == 0xFFFFFFFF and ot.query_ray_closest(rays[q]) is None
            continue
        off_ = (op - rays[q, :3]).astype(f32)
        d = rays[q, 3:]
        t = ((off_[:, 0] * d[0]).astype(f32) + (off_[:, 1] * d[1]).astype(f32)).astype(f32) + (off_[:, 2] * d[2]).astype(f32)
        k = int(np.lexsort((ov, t))[0])
        assert val[q] == ov[k] and pt[q, 3] == t[k] and np.array_equal(pt[q, :3], op[k])
        o = ot.query_ray_closest(rays[q])
        assert o is not None
        if o[0] == ov[k] and f32(o[2]) == t[k]:
            n_same += 1
        elif f32(o[2]) == t[k]:
            n_tie += 1                                   # (1)
        else:
            assert f32(o[2]) > t[k]                      # (2): never closer than the true minimum ...
            assert ot.origin_inside_ancestor(ov[k], rays[q, :3]), "reference missed the closest leaf without an origin-containing ancestor"
            n_pruned += 1
    assert n_same > 20 and n_pruned > 20, (n_same, n_tie, n_pruned)   # (rays that start inside the scene: the root itself contains the origin)
    # rays that start outside the scene: no bound contains the origin, so the reference's value IS the minimum
    outside = rays.copy()
    outside[:, :3] = outside[:, :3] - np.float32(4.0 * L) * outside[:, 3:] / np.linalg.norm(outside[:, 3:], axis=1, keepdims=True).astype(f32)
    val2, pt2 = bvh.query_ray_closest(outside)
    hit2 = agree2 = 0
    for q in range(len(outside)):
        o = ot.query_ray_closest(outside[q])
        if o is None:
            assert val2[q] == 0xFFFFFFFF
            continue
        hit2 += 1
        assert pt2[q, 3] == f32(o[2]), "origin outside every bound: reference and device must agree on t"
        agree2 += int(val2[q] == o[0])       # a different value only on an exact tie in t
    assert hit2 > 20 and agree2 >= 0.95 * hit2
    # FrustumVisitor
    mats = scenes.view_projections(24, L, seed=0xB3D0C407)
    planes = np.stack([broad.frustum_from_mat4(m).reshape(24) for m in mats])
    for k, m in enumerate(mats):
        assert np.array_equal(planes[k], oracle.frustum_from_mat4(m))
    off, vals, rel = bvh.query_frustum(planes)
    total = 0
    for q, (v, r) in enumerate(_split(off, vals, rel)):
        ov, orel = ot.query_frustum(planes[q])
        order = np.argsort(v, kind="stable")
        assert np.array_equal(v[order], ov) and np.array_equal(r[order], orel)
        total += len(v)
    assert total > 1000
    assert not ot.stack_overflow()


def test_dbvt_broad_phase_parity(ctx, oracle):
    boxes, L = scenes.aabb_scene(15_000, seed=0xB3D0C408)
    rng = np.random.default_rng(2)
    for frac in (1.0, 0.3, 0.0):
        dirty = (rng.random(len(boxes)) < frac).astype(np.uint8)
        got = broad.Bvh(ctx, boxes).find_collider_pairs(dirty)
        want = oracle.dbvt_build(boxes).find_collider_pairs(dirty)
        assert np.array_equal(got, want)
    # all-dirty result == the sweep-and-prune pair set (mapped through the permutation)
    pairs, perm = broad.SweepAndPrune3.new(ctx).find_collider_pairs(boxes)
    orig = np.sort(perm[pairs.astype(np.int64)], axis=1)
    orig = orig[np.lexsort((orig[:, 1], orig[:, 0]))]
    assert np.array_equal(orig, broad.Bvh(ctx, boxes).find_collider_pairs(np.ones(len(boxes), dtype=np.uint8)))


def test_kat_dbvt_frustum_and_ray(ctx, oracle):
    """tests/dbvt.rs:39-53 test_frustum and the ray-vs-Aabb3 KATs of tests/aabb.rs:62-157 through the device tree."""
    tree = broad.DynamicBoundingVolumeTree(ctx)
    tree.insert([0.2, 0.2, 0.2, 10., 10., 10.])
    tree.insert([0.2, 0.2, -10., 10., 10., -0.2])   # Aabb3::new sorts (0.2,0.2,-0.2)-(10,10,-10)
    tree.do_refit()
    proj = scenes.perspective_rh_gl(np.float32(60.0) * (np.float32(np.pi) / np.float32(180.0)), 16. / 9., 0.1, 4.0)
    planes = broad.frustum_from_mat4(proj)
    v, r = tree.query_frustum(planes.reshape(24))
    assert v.tolist() == [1] and r.tolist() == [1]   # exactly one hit: value id 11 (second insert), Relation::Cross
    t = broad.DynamicBoundingVolumeTree(ctx)
    t.insert([1., 1., 1., 5., 5., 5.])
    t.do_refit()
    miss = [([0, 0, 0], [1, 0, 0]), ([0, 0, 0], [0, 1, 0]), ([0, 0, 0], [0, 0, 1]), ([0, 0, 0], [0.0001, 0.0001, 1.0]),
            ([0, 2, 2], [-1, 0.01, 0.01]), ([6, 2, 2], [1, 0, 0]), ([2, 6, 2], [0, 1, 0]), ([2, 2, 6], [0, 0, 1])]
    for o, d in miss:
        assert len(t.query_ray(o, d)[0]) == 0
    hit = [([0, 6, 0], [1, -1, 1], [1, 5, 1]), ([0, 2, 2], [1, 0, 0], [1, 2, 2]), ([2, 0, 2], [0, 1, 0], [2, 1, 2]),
           ([2, 2, 0], [0, 0, 1], [2, 2, 1]), ([6, 2, 2], [-1, 0, 0], [5, 2, 2]), ([2, 6, 2], [0, -1, 0], [2, 5, 2]),
           ([2, 2, 6], [0, 0, -1], [2, 2, 5])]
    for o, d, p in hit:
        v, pts = t.query_ray(o, d)
        assert v.tolist() == [0] and pts[0].tolist() == p
        assert t.query_ray_closest(o, d)[0] == 0


def test_bvh_refit_matches_a_tree_built_over_the_moved_boxes(ctx, oracle):
    """update_node + do_refit for every value (dbvt/mod.rs:561-589) as one device refit: after the values move — most a
    little, some across the whole scene — ray, frustum, AABB and dirty-pair queries return what the oracle's tree built over
    the new boxes returns."""
    boxes, L = scenes.aabb_scene(30_000)
    bvh = broad.Bvh(ctx, boxes)
    rng = np.random.default_rng(3)
    moved = boxes.copy()
    shift = rng.normal(0.0, 0.4, size=(len(boxes), 3)).astype(np.float32)
    far = rng.random(len(boxes)) < 0.05
    shift[far] = rng.uniform(-L, L, size=(int(far.sum()), 3)).astype(np.float32)
    moved[:, 0:3] += shift
    moved[:, 3:6] += shift
    for step in range(2):  # two refits in a row: the second starts from an already stale topology
        if step == 1:
            moved[:, [0, 3]] += np.float32(1.5)
        bvh.refit(moved)
        ot = oracle.dbvt_build(moved)
        rays = scenes.rays(200, L, seed=0xB3D0C410 + step)
        off, vals, pts = bvh.query_ray(rays)
        for q in range(len(rays)):
            a, b = int(off[q]), int(off[q + 1])
            ov, op = ot.query_ray(rays[q])
            o = np.argsort(vals[a:b], kind="stable")
            assert np.array_equal(vals[a:b][o], ov) and pts[a:b][o].tobytes() == op.tobytes()
        qb = moved[rng.integers(0, len(moved), size=100)]
        off, vals = bvh.query_aabb(qb)
        for q in range(len(qb)):
            assert np.array_equal(np.sort(vals[int(off[q]):int(off[q + 1])]), ot.query_aabb(qb[q]))
        dirty = (rng.random(len(moved)) < 0.2).astype(np.uint8)
        assert np.array_equal(bvh.find_collider_pairs(dirty), ot.find_collider_pairs(dirty))


# ---------------------------------------------------------------------------------------------------------
# volume::Sphere as the bound type
# ---------------------------------------------------------------------------------------------------------
def test_world_spheres_parity(ctx, oracle):
    """ComputeBound<volume::Sphere> + transform_volume for every primitive kind and for polyhedra: bit-exact."""
    lib = scenes.mesh_library(32, seed=0xB3D0C402)
    a = scenes.all_kinds_pairs(20_000, seed=0xB3D0C403)
    p = scenes.polyhedron_pairs(20_000, library=lib, seed=0xB3D0C403)
    rng = np.random.default_rng(1)
    scene = dict(a)
    scene["kind"] = np.where(rng.random(40_000) < 0.2, p["kind"], a["kind"]).astype(np.uint8)
    scene["mesh_id"], (scene["mesh_verts"], scene["mesh_offsets"]) = p["mesh_id"], lib
    meshes = b3.MeshLibrary(ctx, *lib)
    shapes = b3.ShapeSet(ctx, scene["kind"], scene["params"], scene["xform"], scene["mesh_id"], meshes)
    got = broad.world_spheres(shapes)
    want = oracle.world_spheres(scene)
    assert got.tobytes() == want.tobytes()
    assert len(np.unique(scene["kind"])) == 7 and (got[:, 3] > 0).sum() > 30_000


def _sphere_scene(n, seed, density=4.0):
    L = (n / density) ** (1.0 / 3.0)
    idx = np.arange(n, dtype=np.uint64)
    c = np.stack([scenes.uniform(seed, 110 + k, idx, 0.0, L) for k in range(3)], axis=1)
    r = scenes.uniform(seed, 113, idx, 0.1, 0.6)
    return np.ascontiguousarray(np.concatenate([c, r[:, None]], axis=1), dtype=np.float32)


def test_sap_spheres_parity_and_axis_sequence(ctx, oracle):
    """SweepAndPrune3<volume::Sphere> on a random scene, three consecutive calls (axis hand-over): permutation, pair list in
    the reference's order and next axis identical to the oracle; then the world spheres of a shape set, fed straight in."""
    spheres = _sphere_scene(60_000, 0xB3D0C411)
    sap = broad.SweepAndPrune3Sphere.new(ctx)
    axis = 0
    for _ in range(3):
        pairs, perm = sap.find_collider_pairs(spheres)
        opairs, operm, oaxis = oracle.sap_find_pairs_spheres(spheres, axis)
        assert np.array_equal(perm, operm)
        assert len(opairs) > 100_000 and np.array_equal(pairs, opairs)
        assert sap.get_sweep_axis() == oaxis
        axis = oaxis
    scene = scenes.mixed_pairs(15_000, s=1.5, seed=0xB3D0C412)
    scene["xform"][:, 12:15] *= np.float32(0.35)  # pack the shapes closer so that their bounds overlap
    shapes = b3.ShapeSet(ctx, scene["kind"], scene["params"], scene["xform"])
    ws = broad.world_spheres(shapes)
    pairs, perm = broad.SweepAndPrune3Sphere.new(ctx).find_collider_pairs(ws)
    opairs, operm, _ = oracle.sap_find_pairs_spheres(ws, 0)
    assert np.array_equal(perm, operm) and len(opairs) > 10_000 and np.array_equal(pairs, opairs)


def test_sap_spheres_edge_cases(ctx, oracle):
    sap = broad.SweepAndPrune3Sphere.new(ctx)
    p, perm = sap.find_collider_pairs(np.zeros((0, 4), dtype=np.float32))
    assert len(p) == 0 and len(perm) == 0 and sap.get_sweep_axis() == 0
    p, perm = sap.find_collider_pairs([[0, 0, 0, 1]])
    assert len(p) == 0 and sap.get_sweep_axis() == 0
    # touching spheres ARE pairs (volume/sphere.rs:82-91: <=), unlike touching boxes; coincident zero-radius spheres too
    s = np.array([[5, 0, 0, 0.5], [0, 0, 0, 1], [2, 0, 0, 1], [2, 2.1, 0, 1], [7, 7, 7, 0], [7, 7, 7, 0], [-0.0, 3, 3, 0.25], [0.0, 3, 3, 0.25]],
                 dtype=np.float32)
    for axis in (0, 1, 2):
        sap = broad.SweepAndPrune3Sphere.with_sweep_axis(axis, ctx)
        pairs, perm = sap.find_collider_pairs(s)
        opairs, operm, oaxis = oracle.sap_find_pairs_spheres(s, axis)
        assert np.array_equal(perm, operm), (axis, perm, operm)
        assert np.array_equal(pairs, opairs), (axis, pairs, opairs)
        assert sap.get_sweep_axis() == oaxis
    # a lattice of radius-0.5 spheres at integer points: every lattice neighbour touches exactly (distance^2 == 1 == (r + r)^2),
    # many equal sort keys (stable sort), forced capacity retry
    g = np.stack(np.meshgrid(np.arange(12), np.arange(12), np.arange(12), indexing="ij"), axis=-1).reshape(-1, 3).astype(np.float32)
    rng = np.random.default_rng(9)
    g = g[rng.permutation(len(g))]
    lattice = np.concatenate([g, np.full((len(g), 1), 0.5, dtype=np.float32)], axis=1)
    pairs, perm = broad.SweepAndPrune3Sphere.new(ctx).find_collider_pairs(lattice, capacity=16)
    opairs, operm, _ = oracle.sap_find_pairs_spheres(lattice, 0)
    assert len(opairs) == 3 * 12 * 12 * 11
    assert np.array_equal(perm, operm) and np.array_equal(pairs, opairs)
    # large coordinates and tiny radii (the candidate margin is relative to both)
    far = _sphere_scene(20_000, 0xB3D0C413) * np.float32([1, 1, 1, 0.02]) + np.float32([1.0e4, -2.0e4, 3.0e3, 0.0])
    far = np.concatenate([far, far[:2000] + np.float32([0.01, 0.0, 0.0, 0.0])]).astype(np.float32)
    pairs, perm = broad.SweepAndPrune3Sphere.new(ctx).find_collider_pairs(far)
    opairs, operm, _ = oracle.sap_find_pairs_spheres(far, 0)
    assert len(opairs) >= 1000 and np.array_equal(perm, operm) and np.array_equal(pairs, opairs)


def test_bvh_device_resident_queries_match_the_host_buffer_forms(ctx, oracle, bvh_strategy):
    """bam3d_bvh_query_{aabb,ray,frustum,ray_closest}_dev (one pass, append, device pointers) return exactly the (query, value,
    payload) tuples of the count + scan + fill host-buffer forms — which the oracle checks above — in any order; a capacity
    that is too small drops the excess but still reports the full count."""
    import ctypes as C
    import torch
    boxes, L = scenes.aabb_scene(30_000, seed=0xB3D0C420)
    bvh = broad.Bvh(ctx, boxes)
    dev = torch.device("cuda:0")
    p = lambda t: C.c_void_p(t.data_ptr())
    d_count = torch.zeros(1, dtype=torch.int64, device=dev)

    def run(kind, queries, width, pay_shape, pay_dtype):
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, width)
        d_q = torch.from_numpy(q).to(dev)
        bvh.query_dev(kind, p(d_q), len(q), None, None, None, 0, p(d_count))       # count only
        ctx.synchronize()
        total = int(d_count.item())
        oq = torch.empty(total + 7, dtype=torch.int32, device=dev)
        ov = torch.empty(total + 7, dtype=torch.int32, device=dev)
        op = torch.empty((total + 7,) + pay_shape, dtype=pay_dtype, device=dev) if pay_dtype is not None else None
        bvh.query_dev(kind, p(d_q), len(q), p(oq), p(ov), p(op) if op is not None else None, total + 7, p(d_count))
        ctx.synchronize()
        assert int(d_count.item()) == total
        # too small: nothing written past the capacity, count unchanged
        guard = torch.full((total + 7,), -1, dtype=torch.int32, device=dev)
        bvh.query_dev(kind, p(d_q), len(q), p(guard), p(ov.clone()), p(op.clone()) if op is not None else None, total // 2, p(d_count))
        ctx.synchronize()
        assert int(d_count.item()) == total and bool((guard[total // 2:] == -1).all())
        oq, ov = oq[:total].cpu().numpy().astype(np.uint32), ov[:total].cpu().numpy().astype(np.uint32)
        pay = op[:total].cpu().numpy() if op is not None else None
        order = np.lexsort((ov, oq))
        return oq[order], ov[order], (pay[order] if pay is not None else None), total

    qboxes, _ = scenes.aabb_scene(500, seed=0xB3D0C421, extent=L)
    qboxes[:, 3:] += 1.0
    rays = scenes.rays(500, L, seed=0xB3D0C422)
    planes = np.stack([broad.frustum_from_mat4(m).reshape(24) for m in scenes.view_projections(12, L, seed=0xB3D0C423)])
    for kind, queries, width, ps, pd, host in (("aabb", qboxes, 6, (), None, bvh.query_aabb), ("ray", rays, 6, (3,), torch.float32, bvh.query_ray),
                                                ("frustum", planes, 24, (), torch.uint8, bvh.query_frustum)):
        gq, gv, gp, total = run(kind, queries, width, ps, pd)
        res = host(queries)
        off, vals = res[0], res[1]
        hq = np.repeat(np.arange(len(off) - 1, dtype=np.uint32), np.diff(off).astype(np.int64))
        order = np.lexsort((vals, hq))
        assert total == len(vals) and total > 500
        assert np.array_equal(gq, hq[order]) and np.array_equal(gv, vals[order])
        if gp is not None:
            assert gp.tobytes() == res[2][order].tobytes()
    # closest ray
    d_r = torch.from_numpy(rays).to(dev)
    cv = torch.empty(len(rays), dtype=torch.int32, device=dev)
    cp = torch.empty((len(rays), 4), dtype=torch.float32, device=dev)
    bvh.query_ray_closest_dev(p(d_r), len(rays), p(cv), p(cp))
    ctx.synchronize()
    hv, hp = bvh.query_ray_closest(rays)
    assert np.array_equal(cv.cpu().numpy().astype(np.uint32), hv) and cp.cpu().numpy().tobytes() == hp.tobytes()


def test_bvh_degenerate_scenes_lose_nothing(ctx, oracle):
    """10^5 coincident boxes (every Morton code equal: the tree is split by index bits alone and is as deep as it gets) plus a
    few others: every query must still return every hit. The traversal is stackless, so there is no depth at which children
    could be dropped (the previous explicit 64-entry stack dropped them silently when full)."""
    n = 100_000
    boxes = np.tile(np.array([[1, 1, 1, 2, 2, 2]], dtype=np.float32), (n, 1))
    boxes[-3:] = [[5, 5, 5, 6, 6, 6], [1.5, 1.5, 1.5, 7, 7, 7], [-3, -3, -3, -2, -2, -2]]
    bvh = broad.Bvh(ctx, boxes)
    off, vals = bvh.query_aabb([[0, 0, 0, 1.25, 1.25, 1.25], [5.5, 5.5, 5.5, 9, 9, 9], [10, 10, 10, 11, 11, 11]])
    got = [np.sort(vals[int(off[k]):int(off[k + 1])]) for k in range(3)]
    assert np.array_equal(got[0], np.arange(n - 3)) and got[1].tolist() == [n - 3, n - 2] and len(got[2]) == 0
    off, vals, pts = bvh.query_ray([[0, 1.75, 1.75, 1, 0, 0], [0, 0, 0, -1, 0, 0]])
    assert np.array_equal(np.sort(vals[int(off[0]):int(off[1])]), np.concatenate([np.arange(n - 3), [n - 2]])) and off[2] == off[1]
    val, pt = bvh.query_ray_closest([[0, 1.75, 1.75, 1, 0, 0]])
    assert val[0] == 0 and pt[0, 3] == np.float32(1.0)   # every coincident box ties at t = 1: lowest value index
    # all-dirty DbvtBroadPhase on a smaller pile (n^2 / 2 pairs)
    m = 600
    pile = np.tile(np.array([[0, 0, 0, 1, 1, 1]], dtype=np.float32), (m, 1))
    pairs = broad.Bvh(ctx, pile).find_collider_pairs(np.ones(m, dtype=np.uint8))
    ii, jj = np.triu_indices(m, k=1)
    assert np.array_equal(pairs, np.stack([ii, jj], axis=1).astype(np.uint32))
    # (the oracle's tree over this pile is a chain 600 deep: its query overflows the reference's fixed 256-entry stack, where the
    # reference itself panics with an index out of bounds — dbvt/mod.rs:486,507 — so there is no reference answer to compare with)
    ot = oracle.dbvt_build(pile)
    ot.find_collider_pairs(np.ones(m, dtype=np.uint8))
    assert ot.stack_overflow()
    # sweep and prune over the same pile: every pair, in the reference's order
    sp, perm = broad.SweepAndPrune3.new(ctx).find_collider_pairs(pile)
    op, operm, _ = oracle.sap_find_pairs(pile, 0)
    assert np.array_equal(sp, op) and np.array_equal(perm, operm)


def test_pipeline_broad_then_narrow_matches_the_oracle_chain(ctx, oracle):
    """g1 (BASELINE configs[3]): world AABBs -> SweepAndPrune3 -> perm remap -> GJK::intersection, one resident call, against the
    same chain run on the CPU: the oracle's world AABBs, the oracle's sweep over them (pairs index the SORTED order, as the
    reference returns them: sweep_prune.rs:65-69,80), shapes[perm[i]], shapes[perm[j]] into the oracle's GJK+EPA. Pair list,
    order, permutation, next axis, status and contacts are compared bit for bit; a second frame reuses the returned axis."""
    import bam3d_b200 as b3
    scene, L = scenes.world_scene(20_000, seed=0xB3D0C430)
    shapes = b3.ShapeSet(ctx, scene["kind"], scene["params"], scene["xform"])
    gjk = b3.GJK.new(ctx)
    axis = 0
    for frame in range(2):
        got = broad.pipeline_contacts(gjk, shapes, b3.CollisionStrategy.FullResolution, sweep_axis=axis)
        boxes = oracle.world_aabbs(scene)
        opairs, operm, oaxis = oracle.sap_find_pairs(boxes, axis)
        assert np.array_equal(got["sorted_pairs"], opairs) and np.array_equal(got["perm"], operm) and got["next_axis"] == oaxis
        want_pairs = operm[opairs.astype(np.int64)].astype(np.uint32)
        assert np.array_equal(got["pairs"], want_pairs) and len(want_pairs) > 20_000
        sub = dict(scene)
        sub["pairs"] = want_pairs
        oc, _, ost = oracle.contact(sub, strategy=0, nthreads=8)
        assert np.array_equal(got["status"], ost)
        hit = ost == 0
        assert hit.sum() > 3_000
        assert got["contacts"][hit][:, :7].tobytes() == oc[hit][:, :7].tobytes()
        axis = got["next_axis"]
    # a capacity that is too small reports the size needed (the wrapper retries with it)
    small = broad.pipeline_contacts(gjk, shapes, b3.CollisionStrategy.CollisionOnly, sweep_axis=0, capacity=16)
    assert len(small["pairs"]) == len(opairs) or axis != 0
    shapes.close()


def test_bvh_over_sphere_bounds_applies_the_sphere_predicates(ctx, oracle):
    """DynamicBoundingVolumeTree is generic over the bound: with volume::Sphere (volume/sphere.rs:17-196) the visitors call
    Discrete<Sphere>, Continuous<Ray> and PlaneBound for Sphere. The device tree's nodes are boxes that cover the spheres and its
    leaves apply those predicates; the result sets (and hit points / relations, bit for bit) are compared with the predicates
    applied to every value by the oracle (whose BSphere is pinned by all of tests/sphere.rs and the sphere cases of
    tests/aabb.rs / tests/frustum.rs). Includes touching spheres, rays that start inside a sphere (a hit only while the centre
    is ahead), axis-parallel rays, a degenerate pile of coincident spheres, and the host-buffer and device-resident forms."""
    import torch
    import ctypes as C
    scene, L = scenes.world_scene(30_000, seed=0xB3D0C440)
    shapes = b3.ShapeSet(ctx, scene["kind"], scene["params"], scene["xform"])
    spheres = broad.world_spheres(shapes)                       # ComputeBound<volume::Sphere> of every shape
    assert spheres.tobytes() == oracle.world_spheres(scene).tobytes()
    spheres = spheres.copy()
    spheres[-200:] = spheres[-1]                                  # a pile of 200 coincident spheres
    spheres[100] = [1.0, 1.0, 1.0, 0.5]
    spheres[101] = [2.5, 1.0, 1.0, 1.0]                           # touches sphere 100 exactly: (1.5)^2 <= (1.5)^2
    tree = broad.SphereBvh(ctx, spheres)
    # Discrete<Sphere>
    q = spheres[::97].copy()
    q[:, 3] *= 1.5
    q[0] = spheres[100]
    off, vals = tree.query_sphere(q)
    total = 0
    for k in range(len(q)):
        want, _, _ = oracle.sphere_values_query(spheres, 0, q[k])
        got = np.sort(vals[int(off[k]):int(off[k + 1])])
        assert np.array_equal(got, want), f"sphere query {k}"
        total += len(want)
    assert total >= len(q) and {100, 101} <= set(vals[int(off[0]):int(off[1])].tolist())
    # Continuous<Ray>: random unit directions, some origins at sphere centres (inside), some axis-parallel
    rays = scenes.rays(400, L, seed=0xB3D0C441)
    rays[:50, :3] = spheres[:50, :3]
    rays[50:60, 3:] = [1.0, 0.0, 0.0]
    off, vals, pts = tree.query_ray(rays)
    hits = 0
    for k in range(len(rays)):
        want, wpts, _ = oracle.sphere_values_query(spheres, 1, rays[k])
        sl = slice(int(off[k]), int(off[k + 1]))
        order = np.argsort(vals[sl], kind="stable")
        assert np.array_equal(vals[sl][order], want), f"ray {k}"
        assert pts[sl][order].tobytes() == wpts.tobytes(), f"ray {k}: points"
        hits += len(want)
    assert hits > 100
    # FrustumVisitor
    planes = np.stack([broad.frustum_from_mat4(m).reshape(24) for m in scenes.view_projections(8, L, seed=0xB3D0C442)])
    off, vals, rel = tree.query_frustum(planes)
    seen = set()
    for k in range(len(planes)):
        want, _, wrel = oracle.sphere_values_query(spheres, 2, planes[k])
        sl = slice(int(off[k]), int(off[k + 1]))
        order = np.argsort(vals[sl], kind="stable")
        assert np.array_equal(vals[sl][order], want) and np.array_equal(rel[sl][order], wrel), f"frustum {k}"
        seen |= set(wrel.tolist())
    assert 1 in seen and seen <= {0, 1}                           # In / Cross only (Out is not reported)
    # device-resident one-pass form: the same hits in arbitrary order
    dev = torch.device("cuda", 0)
    d_q = torch.from_numpy(q).to(dev)
    cap = int(total) + 16
    d_oq, d_ov = torch.empty(cap, dtype=torch.int32, device=dev), torch.empty(cap, dtype=torch.int32, device=dev)
    d_cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    tree.query_dev("sphere", p(d_q), len(q), p(d_oq), p(d_ov), None, cap, p(d_cnt))
    ctx.synchronize()
    assert int(d_cnt.item()) == total
    got = np.stack([d_oq.cpu().numpy()[:total], d_ov.cpu().numpy()[:total]], axis=1).astype(np.uint32)
    off, vals = tree.query_sphere(q)
    want = np.stack([np.repeat(np.arange(len(q), dtype=np.uint32), np.diff(off).astype(np.int64)), vals], axis=1)
    assert np.array_equal(got[np.lexsort((got[:, 1], got[:, 0]))], want[np.lexsort((want[:, 1], want[:, 0]))])
    # closest hit: the smallest t = (point - origin) . direction over the spheres a ray meets, ties to the lowest value
    val, pt = tree.query_ray_closest(rays)
    f32 = np.float32
    for k in range(len(rays)):
        want, wpts, _ = oracle.sphere_values_query(spheres, 1, rays[k])
        if len(want) == 0:
            assert val[k] == 0xFFFFFFFF
            continue
        dlt = wpts - rays[k, :3]
        t = (dlt[:, 0] * rays[k, 3] + dlt[:, 1] * rays[k, 4]).astype(f32) + (dlt[:, 2] * rays[k, 5]).astype(f32)
        best = int(np.flatnonzero(t == t.min())[0])           # `want` ascends: the first minimum is the lowest value
        assert val[k] == want[best] and pt[k, :3].tobytes() == wpts[best].tobytes() and pt[k, 3] == t[best], f"closest ray {k}"
    assert (pt[:50, 3][val[:50] != 0xFFFFFFFF] < 0).any()        # origins at sphere centres: the reference's point lies behind them
    # DbvtBroadPhase over sphere bounds, all values dirty on a sub-tree
    sub = spheres[:3000].copy()
    sub_tree = broad.SphereBvh(ctx, sub)
    pairs = sub_tree.find_collider_pairs(np.ones(len(sub), dtype=np.uint8))
    want_pairs = []
    for i in range(len(sub)):
        v, _, _ = oracle.sphere_values_query(sub, 0, sub[i])
        want_pairs += [(i, int(j)) for j in v if j > i]
    assert np.array_equal(pairs, np.array(want_pairs, dtype=np.uint32).reshape(-1, 2)) and len(want_pairs) > 100
    # what a sphere tree does not answer
    for bad in (lambda: tree.query_aabb([[0, 0, 0, 1, 1, 1]]), lambda: tree.refit(np.zeros((tree.n, 6), dtype=np.float32)),
                lambda: broad.Bvh(ctx, [[0, 0, 0, 1, 1, 1]]).query_sphere(q[:1])):
        with pytest.raises(b3.Bam3dError):
            bad()
    empty = broad.SphereBvh(ctx, np.zeros((0, 4), dtype=np.float32))
    off, vals = empty.query_sphere(q[:3])
    assert off.tolist() == [0, 0, 0, 0] and len(vals) == 0
