"""Writes the golden fixtures under tests/golden/ from the CPU oracle (which is itself pinned by the reference's
known-answer tests, oracle/kat_main.cpp). Small on purpose: every primitive kind, every query, a few hundred items.
Run `python -m tests.golden.make_golden` to regenerate after a deliberate oracle change."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def scenes_for_golden():
    from bam3d_b200 import scenes
    lib = scenes.mesh_library(8, vmin=4, vmax=40, seed=0xB3D060D0)
    a = scenes.all_kinds_pairs(768, s=1.2, seed=0xB3D060D1)
    p = scenes.polyhedron_pairs(768, library=lib, s=1.2, seed=0xB3D060D1)
    rng = np.random.default_rng(42)
    narrow = dict(a)
    narrow["kind"] = np.where(rng.random(1536) < 0.15, p["kind"], a["kind"]).astype(np.uint8)
    narrow["mesh_id"], narrow["mesh_verts"], narrow["mesh_offsets"] = p["mesh_id"], lib[0], lib[1]
    narrow["xform_end"] = scenes.end_transforms(narrow, seed=0xB3D060D2)
    boxes, L = scenes.aabb_scene(2048, seed=0xB3D060D3)
    rays = scenes.rays(48, L, seed=0xB3D060D4)
    mats = scenes.view_projections(6, L, seed=0xB3D060D5)
    return narrow, boxes, rays, mats


def generate(oracle):
    narrow, boxes, rays, mats = scenes_for_golden()
    out = {}
    c, hit, st = oracle.contact(narrow, strategy=0)
    ref, geo, some, dst = oracle.distance(narrow)
    tc, thit, tst = oracle.toi(narrow, narrow["xform_end"])
    ihit, ist = oracle.intersect(narrow)
    out["narrow_golden"] = dict(contact=c, contact_status=st, intersect_status=ist, distance_ref=ref, distance_geo=geo,
                                distance_status=dst, toi=tc, toi_status=tst, world_aabbs=oracle.world_aabbs(narrow))
    pairs, perm, axis = oracle.sap_find_pairs(boxes, 0)
    tree = oracle.dbvt_build(boxes)
    rv, rp, roff = [], [], [0]
    for r in rays:
        v, p = tree.query_ray(r)
        rv.append(v); rp.append(p); roff.append(roff[-1] + len(v))
    fv, fr, foff, planes = [], [], [0], []
    for m in mats:
        pl = oracle.frustum_from_mat4(m)
        planes.append(pl)
        v, r = tree.query_frustum(pl)
        fv.append(v); fr.append(r); foff.append(foff[-1] + len(v))
    dirty = (np.arange(len(boxes)) % 3 == 0).astype(np.uint8)
    out["broad_golden"] = dict(sap_pairs=pairs, sap_perm=perm, sap_next_axis=np.array([axis], dtype=np.int32),
                               ray_values=np.concatenate(rv), ray_points=np.concatenate(rp), ray_offsets=np.array(roff, dtype=np.uint64),
                               frustum_planes=np.array(planes, dtype=np.float32), frustum_values=np.concatenate(fv),
                               frustum_relation=np.concatenate(fr), frustum_offsets=np.array(foff, dtype=np.uint64),
                               collider_pairs=tree.find_collider_pairs(dirty))
    prim = scenes_for_rays()
    rays, segments, casts = prim[1:]
    rg = {}
    for name, q, r in (("intersects", 0, rays), ("intersection", 1, rays), ("sweep", 2, segments)):
        pts, st = oracle.ray_cast(prim[0], r, casts, q)
        rg[name + "_points"], rg[name + "_status"] = pts, st
    out["rays_golden"] = rg
    # volume::Sphere as the bound type: world bounding spheres of the narrow scene's shapes, and
    # SweepAndPrune3<volume::Sphere> over them (translations scaled by 0.8 so that more bounds overlap) on every sweep axis
    ws = oracle.world_spheres(narrow)
    packed = ws * np.array([0.8, 0.8, 0.8, 1.0], dtype=np.float32)
    sg = dict(world_spheres=ws, sap_spheres=packed.astype(np.float32))
    for axis in (0, 1, 2):
        spairs, sperm, snext = oracle.sap_find_pairs_spheres(sg["sap_spheres"], axis)
        sg[f"pairs_axis{axis}"], sg[f"perm_axis{axis}"], sg[f"next_axis{axis}"] = spairs, sperm, np.array([snext], dtype=np.int32)
    out["spheres_golden"] = sg
    return out


def scenes_for_rays():
    from bam3d_b200 import scenes
    sc = scenes.all_kinds_pairs(600, s=1.2, seed=0xB3D060D6)
    rays, segments, casts = scenes.primitive_rays(sc, 3000, seed=0xB3D060D7)
    return sc, rays, segments, casts


if __name__ == "__main__":
    import subprocess
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    from tests.oracle_ffi import Oracle
    for name, arrays in generate(Oracle()).items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **arrays)
        print(name, {k: v.shape for k, v in arrays.items()})
