"""Development tool: time bam3d_ray_cast_dev (resident) for the three queries on 2^22 casts against 2^20 mixed shapes."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bam3d_b200 as b3
from bam3d_b200 import scenes

n_shapes, n = 1 << 20, 1 << 22
scene = scenes.all_kinds_pairs(n_shapes // 2, s=1.5)
rays, segments, casts = scenes.primitive_rays(scene, n)
ctx = b3.Context(0)
shapes = b3.ShapeSet(ctx, scene["kind"], scene["params"], scene["xform"])
dev = torch.device("cuda:0")
ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
d_rays, d_seg = torch.from_numpy(rays).to(dev), torch.from_numpy(segments).to(dev)
d_casts = torch.from_numpy(casts.astype(np.int32)).to(dev)
d_pts = torch.empty(n * 3, dtype=torch.float32, device=dev)
d_st = torch.empty(n, dtype=torch.uint8, device=dev)
p = lambda t: C.c_void_p(t.data_ptr())
for q, src in ((0, d_rays), (1, d_rays), (2, d_seg)):
    ts = []
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        ctx.check(ctx._lib.bam3d_ray_cast_dev(ctx.handle, shapes.handle, p(src), n, p(d_casts), n, q, p(d_pts), p(d_st)))
        e1.record(ext)
        ctx.synchronize(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = min(ts[1:])
    # algorithmic bytes per cast: ray 24 + cast 8 + shape record 96 + meta 4 in, point 12 + status 1 out
    print(f"query {q}: {t:.3f} ms, {n / t * 1e-6:.2f} Gcasts/s, {n * 145 / t * 1e-6:.0f} GB/s algorithmic, hits {(d_st == 0).float().mean().item():.3f}")
