"""Parity at BASELINE.json's full batch sizes (2^20 pairs, 2^22 boxes), where the CPU oracle is too slow to redo the whole
batch: size-independent properties the domain offers, plus the oracle on a random sample of the full batch.

  * the answer for a pair is a function of that pair only: permuting the batch permutes the answers, splitting it in
    uneven pieces and concatenating gives the same bytes, running it twice gives the same bytes (binning, queue order,
    warp-aggregated appends and the side-stream overflow tier must not leak into results);
  * the entry points agree with each other the way the reference's methods do: `intersection` is Some exactly where
    `intersect` is (gjk/mod.rs:317-341 calls it first), CollisionOnly hits == FullResolution hits;
  * contacts are well-formed: unit normal, nothing written for misses;
  * sweep and prune: the tiled sweep and the BVH overlap query are two independent searches and must emit the same list;
    every emitted pair strictly overlaps; the list is in the reference's order (j ascending, then i) without duplicates.
"""
import numpy as np
import pytest

import bam3d_b200 as b3
from bam3d_b200 import broad, scenes

pytestmark = pytest.mark.gpu

N = 1 << 20
FULL = b3.CollisionStrategy.FullResolution


def shapes_of(ctx, scene):
    meshes = None
    if scene.get("mesh_verts") is not None:
        meshes = b3.MeshLibrary(ctx, scene["mesh_verts"], scene["mesh_offsets"])
    return b3.ShapeSet(ctx, scene["kind"], scene["params"], scene["xform"], scene.get("mesh_id"), meshes)


def sample_scene(scene, idx):
    sub = dict(scene)
    sub["pairs"] = np.ascontiguousarray(scene["pairs"][idx])
    return sub


def same_bits(a, b):
    """Identical values: bytes for integer data; for f32 every value equal as a number (NaN == NaN, as in
    tests/test_narrow_gpu.py::assert_f32_equal — the sign / payload of a NaN is not part of the contract)."""
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape:
        return False
    if a.dtype.kind == "f":
        return bool(((a == b) | (np.isnan(a) & np.isnan(b))).all())
    return np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_c1_full_size_intersect(ctx, oracle):
    scene = scenes.cuboid_pairs(N, s=2.0)
    shapes = shapes_of(ctx, scene)
    gjk = b3.GJK.new(ctx)
    pairs = scene["pairs"]
    st = gjk.intersect_batch(shapes, pairs)
    assert 0.25 < (st == 0).mean() < 0.45
    assert same_bits(st, gjk.intersect_batch(shapes, pairs)), "not deterministic"
    rng = np.random.default_rng(1)
    perm = rng.permutation(N)
    assert same_bits(gjk.intersect_batch(shapes, pairs[perm]), st[perm]), "depends on the pair's position in the batch"
    cuts = [0, 1, 33, 100_001, 777_777, N]
    pieces = [gjk.intersect_batch(shapes, pairs[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    assert same_bits(np.concatenate(pieces), st), "depends on how the batch is split"
    idx = np.sort(rng.choice(N, 40_000, replace=False))
    _, ost = oracle.intersect(sample_scene(scene, idx), nthreads=8)
    assert np.array_equal(st[idx], ost)


def test_c2_full_size_contact(ctx, oracle):
    scene = scenes.mixed_pairs(N, s=1.5)
    shapes = shapes_of(ctx, scene)
    gjk = b3.GJK.new(ctx)
    pairs = scene["pairs"]
    c, st = gjk.intersection_batch(FULL, shapes, pairs)
    hit = st == 0
    assert 0.4 < hit.mean() < 0.6
    # agreement between the entry points
    ist = gjk.intersect_batch(shapes, pairs)
    assert (st != b3.STATUS_CAPACITY).all()  # (round 1 left 27 pairs here; the third EPA tier finishes them as the reference does)
    assert np.array_equal(hit, ist == 0)
    from tests.pinched import ALL_PINCHED, pinched_scene
    pidx = np.array(ALL_PINCHED)
    want, wst, _, _, _ = oracle.contact_face_limit(pinched_scene(pidx), 1 << 18, in_phases=True, nthreads=8)
    assert np.array_equal(st[pidx], wst) and same_bits(c[pidx][:, 0:7], want[:, 0:7])
    c1, st1 = gjk.intersection_batch(b3.CollisionStrategy.CollisionOnly, shapes, pairs)
    assert np.array_equal(st1 == 0, ist == 0) and not c1.any()
    # (the reference's `distance` is not the complement of `intersect`: its loop can meet `dp - d0 < tolerance` on an
    # overlapping pair and return Some, gjk/mod.rs:262-276 — so it is checked against the oracle on a sample instead)
    dref, ddist, dst = gjk.distance_batch(shapes, pairs)
    didx = np.sort(np.random.default_rng(12).choice(N, 40_000, replace=False))
    oref, odist, _, odst = oracle.distance(sample_scene(scene, didx), nthreads=8)
    assert np.array_equal(dst[didx], odst)
    assert same_bits(dref[didx][odst == 0], oref[odst == 0]) and same_bits(ddist[didx][odst == 0], odist[odst == 0])
    # well-formed contacts
    assert not c[~hit].any()
    n = c[hit, 0:3].astype(np.float64)
    finite = np.isfinite(n).all(axis=1)
    assert (~finite).sum() < 50  # NaN normals of degenerate polytopes are reproduced, not hidden
    assert np.abs(np.linalg.norm(n[finite], axis=1) - 1.0).max() < 1e-5
    assert (c[hit, 3][finite] < 0).mean() < 1e-3  # (a pinched polytope's closest face can lie behind the origin: reproduced, rare)
    # position / split / repetition independence, bit for bit
    c2, st2 = gjk.intersection_batch(FULL, shapes, pairs)
    assert same_bits(st, st2) and same_bits(c, c2), "not deterministic"
    rng = np.random.default_rng(2)
    perm = rng.permutation(N)
    cp, stp = gjk.intersection_batch(FULL, shapes, pairs[perm])
    assert same_bits(stp, st[perm]) and same_bits(cp, c[perm]), "depends on the pair's position in the batch"
    cuts = [0, 5, 250_000, 250_031, N]
    parts = [gjk.intersection_batch(FULL, shapes, pairs[a:b]) for a, b in zip(cuts[:-1], cuts[1:])]
    assert same_bits(np.concatenate([p[1] for p in parts]), st) and same_bits(np.concatenate([p[0] for p in parts]), c)
    # the oracle on a random sample of the full batch
    idx = np.sort(rng.choice(N, 30_000, replace=False))
    oc, _, ost = oracle.contact(sample_scene(scene, idx), strategy=0, nthreads=8)
    assert np.array_equal(st[idx], ost)
    ok = ost == 0
    assert same_bits(c[idx][ok][:, 0:7], oc[ok][:, 0:7])


def test_c3_full_size_polyhedra(ctx, oracle):
    scene = scenes.polyhedron_pairs(N)
    shapes = shapes_of(ctx, scene)
    gjk = b3.GJK.new(ctx)
    pairs = scene["pairs"]
    c, st = gjk.intersection_batch(FULL, shapes, pairs)
    hit = st == 0
    assert 0.3 < hit.mean() < 0.7
    ist = gjk.intersect_batch(shapes, pairs)
    settled = st != b3.STATUS_CAPACITY
    assert np.array_equal(hit[settled], (ist == 0)[settled])
    rng = np.random.default_rng(3)
    perm = rng.permutation(N)[: N // 4]
    cp, stp = gjk.intersection_batch(FULL, shapes, pairs[perm])
    assert same_bits(stp, st[perm]) and same_bits(cp, c[perm])
    idx = np.sort(rng.choice(N, 6_000, replace=False))
    oc, _, ost = oracle.contact(sample_scene(scene, idx), strategy=0, nthreads=8)
    assert np.array_equal(st[idx], ost)
    ok = ost == 0
    assert same_bits(c[idx][ok][:, 0:7], oc[ok][:, 0:7])


def test_toi_full_size(ctx, oracle):
    scene = scenes.mixed_pairs(N, s=1.5, seed=scenes.SEED_C5)
    x1 = scenes.end_transforms(scene)
    s0 = shapes_of(ctx, scene)
    s1 = b3.ShapeSet(ctx, scene["kind"], scene["params"], x1)
    gjk = b3.GJK.new(ctx)
    pairs = scene["pairs"]
    c, st = gjk.intersection_time_of_impact_batch(s0, s1, pairs)
    hit = st == 0
    assert hit.sum() > 10_000
    toi = c[hit, 7]
    assert ((toi >= 0) & (toi <= 1)).all()
    assert not c[st == b3.STATUS_MISS].any()
    rng = np.random.default_rng(4)
    perm = rng.permutation(N)
    cp, stp = gjk.intersection_time_of_impact_batch(s0, s1, pairs[perm])
    assert same_bits(stp, st[perm]) and same_bits(cp, c[perm])
    idx = np.sort(rng.choice(N, 40_000, replace=False))
    oc, _, ost = oracle.toi(sample_scene(scene, idx), x1, nthreads=8)
    assert np.array_equal(st[idx], ost)
    ok = ost == 0
    assert same_bits(c[idx][ok][:, 7], oc[ok][:, 7])
    assert same_bits(c[idx][ok][:, 0:4], oc[ok][:, 0:4])


def test_c4_full_size_sweep_and_prune(ctx):
    boxes, _ = scenes.aabb_scene(1 << 22)
    sap = broad.SweepAndPrune3.new(ctx)
    pairs, perm = sap.find_collider_pairs(boxes)  # auto strategy: BVH overlap query at this density
    n = len(boxes)
    assert 14_000_000 < len(pairs) < 16_500_000
    # perm is a permutation that sorts by (min, max) on the sweep axis, stable
    assert np.array_equal(np.sort(perm), np.arange(n, dtype=perm.dtype))
    sb = boxes[perm]
    key = sb[:, 0].astype(np.float64) * 4096.0 + sb[:, 3]
    assert (np.diff(sb[:, 0]) >= 0).all()
    ties = np.diff(sb[:, 0]) == 0
    assert (np.diff(sb[:, 3])[ties] >= 0).all()
    del key
    # every emitted pair strictly overlaps on all three axes (aabb3.rs:237-244), i < j, reference order, no duplicates
    i, j = pairs[:, 0].astype(np.int64), pairs[:, 1].astype(np.int64)
    assert (i < j).all()
    a, b = sb[i], sb[j]
    assert ((a[:, 0:3] < b[:, 3:6]) & (b[:, 0:3] < a[:, 3:6])).all()
    order_key = j * n + i
    assert (np.diff(order_key) > 0).all()
    # an independent count: boxes overlapping a sample of boxes, by brute force in numpy
    rng = np.random.default_rng(5)
    degree = np.bincount(np.concatenate([i, j]), minlength=n)
    for k in rng.choice(n, 64, replace=False):
        o = ((sb[k, 0:3] < sb[:, 3:6]) & (sb[:, 0:3] < sb[k, 3:6])).all(axis=1)
        assert o.sum() - 1 == degree[k]
    # the sweep axis Variance3 picked is the one the oracle-checked kernel reports on its own
    _, axis = broad.variance3(ctx, sb)
    assert sap.get_sweep_axis() == axis


def test_sap_strategies_agree_at_2_20(ctx):
    boxes, _ = scenes.aabb_scene(1 << 20)
    out = []
    for strategy in (1, 2):
        ctx.set_option("sap_strategy", strategy)
        try:
            out.append(broad.SweepAndPrune3.new(ctx).find_collider_pairs(boxes))
        finally:
            ctx.set_option("sap_strategy", 0)
    assert np.array_equal(out[0][1], out[1][1])
    assert len(out[0][0]) > 3_000_000 and np.array_equal(out[0][0], out[1][0])
