"""Development tool: time one workload's resident step for several builds of libbam3d_gpu.so (A/B kernel variants).

  python tests/tools/ab_time.py c2 bam3d_b200/libbam3d_gpu.so gpurun_out/lib_x.so ...

Each library runs in its own process (BAM3D_LIB selects it). Prints ms per step (min / median over --steps), the
event-timed main-kernel ms, and a CRC of (status, contacts) so that variants can be checked for identical output;
the first variant's output is also checked against the CPU oracle on a prefix (--check pairs).
"""
import argparse
import ctypes as C
import os
import subprocess
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def child(args):
    import numpy as np
    import torch
    import bam3d_b200 as b3
    import bench
    wl, n = args.workload, args.pairs
    # c3i / c3d / c3t: GJK intersect / distance / time of impact on the C3 polyhedron scene (the warp-per-pair kernels)
    poly_mode = {"c3i": "intersect", "c3d": "distance", "c3t": "toi"}.get(wl)
    if poly_mode:
        from bam3d_b200 import scenes
        scene = bench.make_scene("c3", n, 0)
        scene["xform_end"] = scenes.end_transforms(scene)
        wl = {"intersect": "c1", "distance": "dist", "toi": "toi"}[poly_mode]
    else:
        scene = bench.make_scene(wl, n, 0)
    ctx = b3.Context(0)
    gjk = b3.GJK.new(ctx)
    dev = torch.device("cuda:0")
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    meshes = b3.MeshLibrary(ctx, scene["mesh_verts"], scene["mesh_offsets"]) if scene.get("mesh_verts") is not None else None
    shapes = b3.ShapeSet(ctx, scene["kind"], scene["params"], scene["xform"], scene.get("mesh_id"), meshes)
    shapes_end = b3.ShapeSet(ctx, scene["kind"], scene["params"], scene["xform_end"], scene.get("mesh_id"), meshes) if wl == "toi" else None
    d_pairs = torch.from_numpy(scene["pairs"].astype(np.int32).reshape(-1)).to(dev)
    d_status = torch.zeros(n, dtype=torch.uint8, device=dev)
    d_contacts = torch.zeros(n * 8, dtype=torch.float32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())

    def step():
        if wl == "c1":
            gjk.intersect_dev(shapes, p(d_pairs), n, p(d_status))
        elif wl == "dist":
            gjk.distance_dev(shapes, p(d_pairs), n, p(d_contacts), C.c_void_p(d_contacts.data_ptr() + 4 * n), p(d_status))
        elif wl == "toi":
            gjk.intersection_time_of_impact_dev(shapes, shapes_end, p(d_pairs), n, p(d_contacts), p(d_status))
        else:
            gjk.intersection_dev(b3.CollisionStrategy.FullResolution, shapes, p(d_pairs), n, p(d_contacts), p(d_status))

    for _ in range(2):
        step()
    ctx.synchronize()
    ctx.set_option("timing", 1)
    ctx.read_timing()
    times = []
    for _ in range(args.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        step()
        e1.record(ext)
        ctx.synchronize()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    kms, kcalls = ctx.read_timing()
    st = d_status.cpu().numpy()
    ct = d_contacts.cpu().numpy().reshape(n, 8)
    crc = zlib.crc32(ct.tobytes(), zlib.crc32(st.tobytes()))
    tag = os.path.basename(os.environ.get('BAM3D_LIB', 'default')) + "".join(f" {k}={v}" for k, v in os.environ.items() if k.startswith("BAM3D_") and k != "BAM3D_LIB")
    line = f"{tag:28s} {wl} n={n} min {min(times):8.3f} ms  med {sorted(times)[len(times) // 2]:8.3f} ms  kernels {kms / max(kcalls, 1):8.3f} ms  crc {crc:08x}"
    if args.check:
        from tests.oracle_ffi import Oracle
        m = args.check
        sub = dict(scene)
        sub["pairs"] = scene["pairs"][:m]
        orc = Oracle()
        if wl == "c1":
            ok = np.array_equal(orc.intersect(sub, nthreads=16)[1], st[:m])
        elif wl == "dist":
            oref, ogeo, _, ost = orc.distance(sub, nthreads=16)
            flat = d_contacts.cpu().numpy()
            ok = np.array_equal(ost, st[:m]) and np.array_equal(oref[ost == 0], flat[:m][ost == 0]) and np.array_equal(ogeo[ost == 0], flat[n:n + m][ost == 0])
        elif wl == "toi":
            oc, _, ost = orc.toi(sub, scene["xform_end"], nthreads=16)
            ok = np.array_equal(ost, st[:m]) and np.array_equal(oc[ost == 0][:, 7], ct[:m][ost == 0][:, 7])
        else:
            oc, _, ost = orc.contact(sub, strategy=0, nthreads=16)
            ok = np.array_equal(ost, st[:m]) and np.array_equal(oc[ost == 0][:, :7], ct[:m][ost == 0][:, :7])
        line += f"  oracle[{m}] {'OK' if ok else 'MISMATCH'}"
    print(line, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload")
    ap.add_argument("libs", nargs="*")
    ap.add_argument("--pairs", type=int, default=1 << 20)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--check", type=int, default=0)
    ap.add_argument("--child", action="store_true")
    args = ap.parse_args()
    if args.child:
        return child(args)
    for i, lib in enumerate(args.libs or [os.path.join(ROOT, "bam3d_b200", "libbam3d_gpu.so")]):
        lib, _, extra = lib.partition("@")  # path[@KEY=VALUE[,KEY=VALUE]]: environment switches for this run (e.g. BAM3D_TOI_REFILL=1)
        env = dict(os.environ, BAM3D_LIB=os.path.abspath(lib))
        env.update(kv.split("=", 1) for kv in extra.split(",") if kv)
        cmd = [sys.executable, os.path.abspath(__file__), args.workload, "--child", "--pairs", str(args.pairs), "--steps", str(args.steps),
               "--check", str(args.check if i == 0 else 0)]
        subprocess.run(cmd, env=env, check=False)


if __name__ == "__main__":
    main()
