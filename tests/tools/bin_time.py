"""Development tool: time the contact path of the C2 scene one (kindL, kindR) bin at a time (same scene, pair subsets)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bam3d_b200 as b3
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
only = sys.argv[2].split(",") if len(sys.argv) > 2 else None   # e.g. "sph-sph,all"
scene = bench.make_scene("c2", n, 0)
ctx = b3.Context(0)
gjk = b3.GJK.new(ctx)
dev = torch.device("cuda:0")
ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
shapes = b3.ShapeSet(ctx, scene["kind"], scene["params"], scene["xform"], None, None)
k = scene["kind"][scene["pairs"]]
names = {2: "sph", 3: "cub", 4: "cyl"}
p = lambda t: C.c_void_p(t.data_ptr())
total = 0.0
for sel_name, sel in [("all", np.ones(n, bool))] + [(f"{names[a]}-{names[b]}", (k[:, 0] == a) & (k[:, 1] == b)) for a in (2, 3, 4) for b in (2, 3, 4)]:
    if only and sel_name not in only:
        continue
    pairs = np.ascontiguousarray(scene["pairs"][sel])
    m = len(pairs)
    d_pairs = torch.from_numpy(pairs.astype(np.int32).reshape(-1)).to(dev)
    d_status = torch.zeros(m, dtype=torch.uint8, device=dev)
    d_contacts = torch.zeros(m * 8, dtype=torch.float32, device=dev)
    times = []
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        gjk.intersection_dev(b3.CollisionStrategy.FullResolution, shapes, p(d_pairs), m, p(d_contacts), p(d_status))
        e1.record(ext)
        ctx.synchronize()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    st = d_status.cpu().numpy()
    t = min(times[1:])
    if sel_name != "all":
        total += t
    tag = " ".join(f"{k_}={v_}" for k_, v_ in os.environ.items() if k_.startswith("BAM3D_"))
    print(f"{tag} {sel_name:8s} pairs {m:8d} hits {(st == 0).sum():7d} cap {(st == 5).sum():3d}  {t:8.3f} ms", flush=True)
print(f"sum of bins {total:.3f} ms")
