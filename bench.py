#!/usr/bin/env python
"""bench.py — the hot path of bam3d on B200: batched GJK+EPA contact generation (pairs/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c1|c2|c3|toi|c4|c5]

A step is one pass of the narrow phase over one batch of synthetic pairs:
  c2 (default, BASELINE.json configs[1]): GJK3+EPA3 Contact (FullResolution) on 2^20 mixed Sphere/Cuboid/Cylinder
     pairs per GPU; c1: GJK intersect on 2^20 Cuboid-Cuboid pairs; c3: GJK+EPA on 2^20 ConvexPolyhedron pairs
     (64-256 vertices); toi: time of impact on 2^20 mixed pairs.
`value`  = pairs/s over all ranks with everything resident in HBM (device-pointer C-ABI entry points, CUDA events
           on the library's stream, max over ranks).
`e2e`    = the same through the host-buffer C-ABI call a user of the crate would make each frame: pinned host
           transforms + pair list in, contacts + status out, copies inside the timed region.
`roofline` = algorithmic HBM bytes of the dominant kernel / its event-timed duration vs MEASURED_PEAKS.json.
`cpu_baseline` = the CPU oracle (strict-f32 C++ restatement of the reference; the Rust crate cannot be built in this
           image) on all host cores, same workload. `--impl reference` times only that.
Multi-GPU (torchrun, one rank per GPU): pairs shard contiguously (weak scaling: 2^20 pairs per GPU); every rank
compacts its hits and the compacted Contact40 records are all-gathered over NCCL each step.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALGO_BYTES = {  # algorithmic HBM bytes per pair
    "c1": 129.0,   # 2 x (48 B affine + 16 B params) + 1 B status
    "c2": 148.0,   # 128 B in + 40 B contact x ~0.5 hit rate
    "c3": 124.0,   # 2 x 48 B + 2 x 4 B mesh id + 40 B x hit rate (vertex reads are L2 traffic)
    "toi": 264.0,  # 4 x 48 + 2 x 16 + 40 B out
    "c4": 220.0,   # per box: 24 B read + 4 B perm + 8 radix passes x 2 x 12 B; plus 8 B per emitted pair (added below)
    "c5": 0.4 * 148.0 + 0.2 * 129.0 + 0.2 * 124.0 + 0.2 * 264.0,  # the C5 mix of the four rows above (162.6 B/pair)
}
# DRAM traffic of the dominant kernel per launch (bytes), from the committed ncu captures (profiles/r1_*_ncu_full.txt);
# for C2 it is far above the algorithmic bytes because the EPA polytopes (7.5 KB slab record per resident thread) do not
# all stay in L2 (DESIGN.md §11 records what was tried against it this round: 24 GB -> 5.4 GB).
NCU_DRAM_BYTES = {"c1": 223.0e6, "c2": 5.36e9, "c3": 261.8e6, "toi": 636.5e6}  # c2: profiles/r1c_c2_ncu_full.txt, c3: profiles/r1f_c3_ncu_full.txt, toi: profiles/r1h_toi_ncu_full.txt
WORKLOAD_DESC = {
    "c1": "C1: GJK intersect, 2^20 random Cuboid-Cuboid pairs per GPU (f32, random transforms, offset U[-2,2]^3)",
    "c2": "C2: GJK+EPA Contact (FullResolution), 2^20 mixed Sphere/Cuboid/Cylinder pairs per GPU (offset U[-1.5,1.5]^3)",
    "c3": "C3: GJK+EPA Contact, 2^20 ConvexPolyhedron pairs per GPU (4096-mesh library, 64-256 vertices, warp-cooperative support)",
    "toi": "TOI: GJK time of impact, 2^20 mixed Sphere/Cuboid/Cylinder pairs per GPU (end = start + U[-3,3]^3)",
    "c5": "C5: mixed convex pairs per GPU slice: 40 % analytic GJK+EPA contact (C2 mix), 20 % GJK intersect only, 20 % "
          "ConvexPolyhedron GJK+EPA contact (C3 library), 20 % GJK time of impact; contacts compacted and all-gathered (NCCL) "
          "every step; 2^23 pairs per GPU = 64M pairs on 8 GPUs",
    "c4": "C4: SweepAndPrune3 broad phase over 2^22 Aabb3 per GPU (centres U[0,100]^3, half-extents U[0.1,0.5]^3), "
          "plus BVH build and 2^16 ray / 64 frustum queries reported in `extra`",
}
METRIC = {"c1": "gjk_intersect_pairs_per_sec", "c2": "gjk_epa_contact_pairs_per_sec", "c3": "gjk_epa_contact_pairs_per_sec",
          "toi": "gjk_toi_pairs_per_sec", "c4": "sap_candidate_pairs_per_sec", "c5": "mixed_narrow_phase_pairs_per_sec"}


def make_scene(workload, n, rank):
    from bam3d_b200 import scenes
    start = rank * n  # every rank generates exactly its own slice of the global scene
    if workload == "c1":
        return scenes.cuboid_pairs(n, s=2.0, start=start)
    if workload in ("c2", "toi"):
        sc = scenes.mixed_pairs(n, s=1.5 if workload == "c2" else 4.0, start=start)
        if workload == "toi":
            sc["xform_end"] = scenes.end_transforms(sc, start=2 * start)
        return sc
    if workload == "c3":
        return scenes.polyhedron_pairs(n, n_meshes=4096, s=2.0, start=start)
    raise SystemExit(f"unknown workload {workload}")


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle on all host cores (the reference crate itself cannot be built here: no cargo / rustc)
# ---------------------------------------------------------------------------------------------------------
def oracle_run(workload, scene, nthreads):
    from tests.oracle_ffi import Oracle
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    orc = Oracle()

    def once():
        if workload == "c1":
            orc.intersect(scene, nthreads=nthreads)
        elif workload == "toi":
            orc.toi(scene, scene["xform_end"], nthreads=nthreads)
        else:
            orc.contact(scene, strategy=0, nthreads=nthreads)
    return once


def cpu_sample(scene, workload, n_sample):
    sc = dict(scene)
    sc["pairs"] = scene["pairs"][:n_sample]
    return sc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if args.workload == "c4":
        return run_reference_c4(args)
    if args.workload == "c5":
        return run_reference_c5(args)
    n = args.pairs
    cores = os.cpu_count() or 1
    scene = make_scene(args.workload, n, 0)
    n_sample = min(n, args.cpu_pairs)
    once = oracle_run(args.workload, cpu_sample(scene, args.workload, n_sample), cores)
    for _ in range(args.warmup):
        once()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        once()
    dt = time.perf_counter() - t0
    value = n_sample * args.steps / dt
    unit = "pairs/s"
    line = {
        "impl": "reference", "metric": METRIC[args.workload], "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC[args.workload], "pairs_per_step": n_sample,
                   "note": "reference crate (Rust) cannot be built in this image; this arm times the strict-f32 C++ restatement of it"},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port",
                         "sample": f"{n_sample} pairs of the workload per step, std::thread static partition over {cores} threads"},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def window(self, t0, t1):
        """Keep only the samples taken inside [t0, t1] (the timed region)."""
        self.rows = [r for r in self.rows if t0 - 0.05 <= r[0] <= t1 + 0.05]

    def count(self, t0, t1):
        return sum(1 for r in self.rows if t0 - 0.05 <= r[0] <= t1 + 0.05)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for _, r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nme in enumerate(names):
                if f[2 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def run_reference_c4(args):
    """CPU arm of C4: the oracle's sequential sweep (the reference is single-threaded here: one sort + one sweep)."""
    from bam3d_b200 import scenes
    from tests.oracle_ffi import Oracle
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    orc = Oracle()
    n = min(args.boxes, args.cpu_boxes)
    boxes, L = scenes.aabb_scene(n, extent=100.0 * (n / float(1 << 22)) ** (1.0 / 3.0))
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        pairs, perm, axis = orc.sap_find_pairs(boxes, 0)
    dt = time.perf_counter() - t0
    value = len(pairs) * steps / dt
    line = {"impl": "reference", "metric": METRIC["c4"], "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": 0, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC["c4"], "boxes_per_step": n, "pairs_per_step": int(len(pairs)),
                       "note": "reference crate (Rust) cannot be built in this image; this arm times the strict-f32 C++ restatement; "
                               "the reference's sweep is sequential, so 1 thread"},
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": 1, "kind": "port", "sample": f"{n} boxes at C4 density, {steps} sweeps"},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def run_ours_c4(args):
    """C4: device sweep-and-prune over 2^22 boxes per GPU (candidate pairs/s), BVH build + queries as extras."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; bam3d_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import bam3d_b200 as b3
    from bam3d_b200 import broad, build, scenes
    build.build()
    n = args.boxes
    # N > 1: every rank holds the SAME box set and sorts all of it (replicated sort); the pair search is
    # sharded by the pairs' higher sorted slot, so the ranks' pair slices are disjoint and concatenate to the full list
    boxes, L = scenes.aabb_scene(n, start=0, extent=100.0 * (n / float(1 << 22)) ** (1.0 / 3.0))
    ctx = b3.Context(local_rank)
    lib, h = ctx._lib, ctx.handle
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    d_boxes = torch.from_numpy(boxes).to(dev)
    cap = 6 * n
    d_pairs = torch.empty(cap * 2, dtype=torch.int32, device=dev)
    d_perm = torch.empty(n, dtype=torch.int32, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    cnt = C.c_uint64(0)

    def step_resident():
        axis = C.c_int32(0)
        ctx.check(lib.bam3d_sap_find_pairs_shard_dev(h, p(d_boxes), n, C.byref(axis), p(d_perm), p(d_pairs), cap, C.byref(cnt), rank, world))

    def sync_all():
        ctx.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        step_resident()
    sync_all()
    ctx.set_option("timing", 1); ctx.read_timing()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time(); ev0.record(ext)
    for _ in range(args.steps):
        step_resident()
    ev1.record(ext); sync_all(); w1 = time.time()
    ms = ev0.elapsed_time(ev1)
    kernel_ms, kernel_calls = ctx.read_timing(); ctx.set_option("timing", 0)
    launches = ctx.launch_count - launches0
    npairs = int(cnt.value)
    clocks = None
    if sampler:
        sampler.window(w0, w1); clocks = sampler.stop(); clocks["note"] = "sampled during the timed region"
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    tot = torch.tensor([npairs], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    npairs_total = int(tot.item())  # pairs of the whole box set: the ranks' shards are disjoint
    value = npairs_total * args.steps / (ms_max * 1e-3)

    # e2e: host boxes in, perm + pairs out through the host-buffer call
    h_boxes = torch.from_numpy(boxes).pin_memory()
    h_pairs = torch.empty((cap, 2), dtype=torch.int32).pin_memory()
    h_perm = torch.empty(n, dtype=torch.int32).pin_memory()
    hp = lambda t: C.c_void_p(t.data_ptr())

    def step_e2e():
        axis = C.c_int32(0)
        ctx.check(lib.bam3d_sap_find_pairs_shard(h, hp(h_boxes), n, C.byref(axis), hp(h_perm), hp(h_pairs), cap, C.byref(cnt), rank, world))
    e2e_steps = max(3, min(args.steps, 5))
    step_e2e(); sync_all()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = npairs_total * e2e_steps / float(t.item())

    extra = {}
    if rank == 0:
        # BVH build + queries (DBVT query path), timed with wall clock around synchronous host-buffer calls
        t0 = time.perf_counter(); bvh = broad.Bvh(ctx, boxes); extra["bvh_build_ms"] = 1e3 * (time.perf_counter() - t0)
        rays = scenes.rays(1 << 16, L)
        bvh.query_ray(rays)
        t0 = time.perf_counter(); off, vals, pts = bvh.query_ray(rays); dt = time.perf_counter() - t0
        extra["ray_queries_per_s"] = len(rays) / dt; extra["ray_hits"] = int(len(vals))
        t0 = time.perf_counter(); bvh.query_ray_closest(rays); dt = time.perf_counter() - t0
        extra["ray_closest_queries_per_s"] = len(rays) / dt
        planes = np.stack([broad.frustum_from_mat4(m).reshape(24) for m in scenes.view_projections(64, L)])
        bvh.query_frustum(planes)
        t0 = time.perf_counter(); off, vals, rel = bvh.query_frustum(planes); dt = time.perf_counter() - t0
        extra["frustum_queries_per_s"] = len(planes) / dt; extra["frustum_hits"] = int(len(vals))
        peak, peak_src = load_peaks()
        per_call_ms = kernel_ms / max(1, args.steps)
        algo_bytes = ALGO_BYTES["c4"] * n + 8.0 * npairs
        achieved = algo_bytes / (per_call_ms * 1e-3) / 1e9
        # CPU baseline: the oracle's sequential sweep on a bounded sample at the same density
        from tests.oracle_ffi import Oracle
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        orc = Oracle()
        ns = min(n, args.cpu_boxes)
        sboxes, _ = scenes.aabb_scene(ns, extent=100.0 * (ns / float(1 << 22)) ** (1.0 / 3.0))
        t0 = time.perf_counter(); opairs, _, _ = orc.sap_find_pairs(sboxes, 0); cpu_dt = time.perf_counter() - t0
        line = {"metric": METRIC["c4"], "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD_DESC["c4"], "boxes": n, "pairs_total": npairs_total, "pairs_rank0": npairs,
                           "l2": "inputs larger than L2 (96 MB boxes + 190 MB sort/sorted-box scratch + pair keys); no flush",
                           "multi_gpu": "replicated sort, pair search sharded by the pairs' higher sorted slot (no collective)" if world > 1 else "single GPU"},
                "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": n * 24, "d2h_bytes_per_step": n * 4 + npairs * 8,
                        "steps": e2e_steps, "path": "host-buffer bam3d_sap_find_pairs, pinned host memory"},
                "gpu_launches": launches, "clocks": clocks,
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                             "kernel": "bvh_overlap_append_kernel (auto strategy at this density) + CUB radix sorts + LBVH build; sap_variance_chain_kernel on the side stream", "kernel_ms_per_launch": per_call_ms,
                             "algorithmic_bytes_per_box": ALGO_BYTES["c4"], "peak_source": peak_src,
                             "note": "a 1-axis sweep in a 3-D uniform scene would do n x ~24k candidate tests, so the pair search runs as a BVH overlap query over the sorted boxes: traversal-latency bound, not HBM bound"},
                "cpu_baseline": {"value": len(opairs) / cpu_dt, "unit": "pairs/s", "cores": 1, "kind": "port",
                                 "sample": f"{ns} boxes at C4 density, one sequential sweep (the reference's algorithm is sequential)"},
                "extra": extra}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    return 0



# ---------------------------------------------------------------------------------------------------------
# C5: the 64M mixed configuration, one contiguous slice per GPU
# ---------------------------------------------------------------------------------------------------------
C5_PARTS = (("contact", 0.4), ("intersect", 0.2), ("poly", 0.2), ("toi", 0.2))


def c5_scenes(n, rank):
    """The four sub-batches of one rank's C5 slice (seed ...0005). Sizes are multiples of 1024."""
    from bam3d_b200 import scenes
    sizes = {k: int(n * f) // 1024 * 1024 for k, f in C5_PARTS}
    sizes["contact"] += n - sum(sizes.values())
    seed = scenes.SEED_C5
    out = {}
    out["contact"] = scenes.mixed_pairs(sizes["contact"], s=1.5, seed=seed, start=rank * sizes["contact"])
    out["intersect"] = scenes.mixed_pairs(sizes["intersect"], s=1.5, seed=seed + 1, start=rank * sizes["intersect"])
    out["poly"] = scenes.polyhedron_pairs(sizes["poly"], n_meshes=4096, s=2.0, start=rank * sizes["poly"])
    t = scenes.mixed_pairs(sizes["toi"], s=4.0, seed=seed + 2, start=rank * sizes["toi"])
    t["xform_end"] = scenes.end_transforms(t, start=2 * rank * sizes["toi"])
    out["toi"] = t
    return out, sizes


def c5_oracle_run(parts, nthreads):
    from tests.oracle_ffi import Oracle
    orc = Oracle()

    def run():
        orc.contact(parts["contact"], strategy=0, nthreads=nthreads)
        orc.intersect(parts["intersect"], nthreads=nthreads)
        orc.contact(parts["poly"], strategy=0, nthreads=nthreads)
        orc.toi(parts["toi"], parts["toi"]["xform_end"], nthreads=nthreads)
    return run


def c5_sample(parts, n_sample):
    """The same 40/20/20/20 mix on the first pairs of every sub-batch."""
    out = {}
    for k, f in C5_PARTS:
        sub = dict(parts[k])
        sub["pairs"] = parts[k]["pairs"][: max(1, int(n_sample * f))]
        out[k] = sub
    return out


def run_reference_c5(args):
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    cores = os.cpu_count() or 1
    n_sample = min(args.pairs, args.cpu_pairs)
    parts, _ = c5_scenes(max(n_sample, 8192), 0)
    sample = c5_sample(parts, n_sample)
    n_done = sum(len(v["pairs"]) for v in sample.values())
    once = c5_oracle_run(sample, cores)
    for _ in range(args.warmup):
        once()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        once()
    dt = time.perf_counter() - t0
    value = n_done * args.steps / dt
    print(json.dumps({
        "impl": "reference", "metric": METRIC["c5"], "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD_DESC["c5"], "pairs_per_step": n_done,
                   "note": "reference crate (Rust) cannot be built in this image; this arm times the strict-f32 C++ restatement of its code path"},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"{n_done} pairs of the C5 mix per step, std::thread static partition over {cores} threads"},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
    return 0


def run_ours_c5(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0")); local_rank = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; bam3d_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import bam3d_b200 as b3
    from bam3d_b200 import build, multi
    build.build()
    n = args.pairs
    parts, sizes = c5_scenes(n, rank)
    ctx = b3.Context(local_rank)
    gjk = b3.GJK.new(ctx)
    lib, h = ctx._lib, ctx.handle
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    meshes = b3.MeshLibrary(ctx, parts["poly"]["mesh_verts"], parts["poly"]["mesh_offsets"])
    shp = {k: b3.ShapeSet(ctx, v["kind"], v["params"], v["xform"], v.get("mesh_id"), meshes if k == "poly" else None) for k, v in parts.items()}
    shp_end = b3.ShapeSet(ctx, parts["toi"]["kind"], parts["toi"]["params"], parts["toi"]["xform_end"])
    d_pairs = {k: torch.from_numpy(v["pairs"].astype(np.int32).reshape(-1)).to(dev) for k, v in parts.items()}
    # dense outputs of the three contact-producing sub-batches share one buffer so that one compaction feeds the gather
    order = ("contact", "poly", "toi")
    off = {}
    acc = 0
    for k in order:
        off[k] = acc
        acc += sizes[k]
    n_dense = acc
    d_status = torch.empty(n_dense, dtype=torch.uint8, device=dev)
    d_contacts = torch.empty(n_dense * 8, dtype=torch.float32, device=dev)
    d_status_i = torch.empty(sizes["intersect"], dtype=torch.uint8, device=dev)
    d_compact = torch.empty(n_dense * 10, dtype=torch.int32, device=dev)
    d_count = torch.zeros(1, dtype=torch.int64, device=dev)
    gathered = torch.empty(world * n_dense * 10, dtype=torch.int32, device=dev) if world > 1 else None
    counts = torch.zeros(world, dtype=torch.int64, device=dev) if world > 1 else None
    pc = lambda t, o=0: C.c_void_p(t.data_ptr() + o)
    FR = b3.CollisionStrategy.FullResolution

    def step_resident():
        gjk.intersection_dev(FR, shp["contact"], pc(d_pairs["contact"]), sizes["contact"], pc(d_contacts, 32 * off["contact"]), pc(d_status, off["contact"]))
        gjk.intersect_dev(shp["intersect"], pc(d_pairs["intersect"]), sizes["intersect"], pc(d_status_i))
        gjk.intersection_dev(FR, shp["poly"], pc(d_pairs["poly"]), sizes["poly"], pc(d_contacts, 32 * off["poly"]), pc(d_status, off["poly"]))
        gjk.intersection_time_of_impact_dev(shp["toi"], shp_end, pc(d_pairs["toi"]), sizes["toi"], pc(d_contacts, 32 * off["toi"]), pc(d_status, off["toi"]))
        gjk.compact_contacts_dev(pc(d_contacts), pc(d_status), n_dense, rank * n_dense, pc(d_compact), n_dense, pc(d_count))
        if world > 1:
            with torch.cuda.stream(ext):
                multi.gather_contacts(d_compact, d_count, counts_buf=counts, out=gathered)

    def sync_all():
        ctx.synchronize(); torch.cuda.synchronize()
        if world > 1:
            dist.barrier(); torch.cuda.synchronize()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(args.warmup):
        step_resident()
    sync_all()
    ctx.set_option("timing", 1); ctx.read_timing()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time(); ev0.record(ext)
    for _ in range(args.steps):
        step_resident()
    ev1.record(ext); sync_all(); w1 = time.time()
    ms = ev0.elapsed_time(ev1)
    kernel_ms, kernel_calls = ctx.read_timing(); ctx.set_option("timing", 0)
    launches = ctx.launch_count - launches0
    clocks = None
    if sampler:
        sampler.window(w0, w1); clocks = sampler.stop(); clocks["note"] = "sampled during the timed region"
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n * args.steps / (ms_max * 1e-3)
    n_hits = int(d_count.item())

    # e2e: host transforms + pairs in, status (+ contacts) out, through the host-buffer C ABI, per sub-batch
    hx = {k: torch.from_numpy(v["xform"]).pin_memory() for k, v in parts.items()}
    hx_end = torch.from_numpy(parts["toi"]["xform_end"]).pin_memory()
    hpairs = {k: torch.from_numpy(v["pairs"].astype(np.int32)).pin_memory() for k, v in parts.items()}
    hst = {k: torch.empty(sizes[k], dtype=torch.uint8).pin_memory() for k in sizes}
    hct = {k: torch.empty((sizes[k], 8), dtype=torch.float32).pin_memory() for k in order}
    hp = lambda t: C.c_void_p(t.data_ptr())
    S = C.byref(gjk.settings)

    def step_e2e():
        for k in sizes:
            ctx.check(lib.bam3d_shapes_set_transforms(h, shp[k].handle, hp(hx[k])))
        ctx.check(lib.bam3d_shapes_set_transforms(h, shp_end.handle, hp(hx_end)))
        ctx.check(lib.bam3d_gjk_contact(h, shp["contact"].handle, hp(hpairs["contact"]), sizes["contact"], S, 0, hp(hct["contact"]), hp(hst["contact"])))
        ctx.check(lib.bam3d_gjk_intersect(h, shp["intersect"].handle, hp(hpairs["intersect"]), sizes["intersect"], S, hp(hst["intersect"])))
        ctx.check(lib.bam3d_gjk_contact(h, shp["poly"].handle, hp(hpairs["poly"]), sizes["poly"], S, 0, hp(hct["poly"]), hp(hst["poly"])))
        ctx.check(lib.bam3d_gjk_time_of_impact(h, shp["toi"].handle, shp_end.handle, hp(hpairs["toi"]), sizes["toi"], S, hp(hct["toi"]), hp(hst["toi"])))
    h2d = sum(len(v["kind"]) * 64 + len(v["pairs"]) * 8 for v in parts.values()) + len(parts["toi"]["kind"]) * 64
    d2h = n + n_dense * 32
    e2e_steps = max(3, min(args.steps, 5))

    def timed(fn):
        sync_all()
        t0 = time.perf_counter()
        fn()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * n * e2e_steps / float(t.item())

    step_e2e()
    e2e_sync = timed(lambda: [step_e2e() for _ in range(e2e_steps)])
    status_sync = {k: hst[k].clone() for k in sizes}

    # The same frames through one frame queue per sub-batch (bam3d_frames_*), two steps in flight: step k+1's transforms and
    # pairs are copied in, and step k-1's results copied out, while step k's kernels run.
    DEPTH = 2
    query = {"contact": 1, "intersect": 0, "poly": 1, "toi": 2}
    queues = {}
    for k, v in parts.items():
        q = C.c_void_p()
        kind, params = np.ascontiguousarray(v["kind"], dtype=np.uint8), np.ascontiguousarray(v["params"], dtype=np.float32)
        mesh_id = None if v.get("mesh_id") is None else np.ascontiguousarray(v["mesh_id"], dtype=np.uint32)
        ctx.check(lib.bam3d_frames_create(h, C.c_void_p(kind.ctypes.data), C.c_void_p(params.ctypes.data), hp(hx[k]),
                                          C.c_void_p(mesh_id.ctypes.data) if mesh_id is not None else None, len(kind),
                                          meshes.handle if k == "poly" else None, sizes[k], DEPTH, 1 if k == "toi" else 0, C.byref(q)))
        queues[k] = q
    qst = [{k: torch.empty(sizes[k], dtype=torch.uint8).pin_memory() for k in sizes} for _ in range(DEPTH)]
    qct = [{k: torch.empty((sizes[k], 8), dtype=torch.float32).pin_memory() for k in order} for _ in range(DEPTH)]
    submitted = {"n": 0}

    def run_queues(steps):
        ticket = C.c_uint64(0)
        first = submitted["n"]
        for s_ in range(steps):
            f = first + s_
            if s_ >= DEPTH:
                for k in sizes:
                    ctx.check(lib.bam3d_frames_wait(h, queues[k], f - DEPTH))
            b = f % DEPTH
            for k in sizes:
                ctx.check(lib.bam3d_frames_submit(h, queues[k], query[k], hp(hx[k]), hp(hx_end) if k == "toi" else None, hp(hpairs[k]), sizes[k], S, 0,
                                                  hp(qct[b][k]) if k in qct[b] else None, hp(qst[b][k]), C.byref(ticket)))
        for f in range(first + max(0, steps - DEPTH), first + steps):
            for k in sizes:
                ctx.check(lib.bam3d_frames_wait(h, queues[k], f))
        submitted["n"] = first + steps

    run_queues(2)
    e2e_value = timed(lambda: run_queues(e2e_steps))
    for b in range(min(DEPTH, e2e_steps)):
        for k in sizes:
            if not torch.equal(qst[b][k], status_sync[k]):
                raise SystemExit("bench.py: frame-queue results differ from the synchronous calls")
    for q in queues.values():
        lib.bam3d_frames_free(h, q)

    if rank == 0:
        peak, peak_src = load_peaks()
        step_kernel_ms = kernel_ms / max(1, args.steps)
        achieved = ALGO_BYTES["c5"] * n / (step_kernel_ms * 1e-3) / 1e9
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        cores = os.cpu_count() or 1
        sample = c5_sample(parts, min(n, args.cpu_pairs))
        n_sample = sum(len(v["pairs"]) for v in sample.values())
        once = c5_oracle_run(sample, cores)
        once()
        best = 1e30
        for _ in range(2):
            t0 = time.perf_counter(); once(); best = min(best, time.perf_counter() - t0)
        print(json.dumps({
            "metric": METRIC["c5"], "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC["c5"], "pairs_per_gpu": n, "sub_batches": sizes, "contacts_per_gpu_rank0": n_hits,
                       "l2": "inputs larger than L2 (shape records 96 B x 2 per pair); no flush",
                       "multi_gpu": "contiguous pair slices, per-rank compaction, NCCL all-gather of Contact40 records each step" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "path": f"one frame queue per sub-batch (bam3d_frames_submit / bam3d_frames_wait), {DEPTH} steps in flight: every step copies "
                            "its transforms + pairs in from pinned host memory and its status (+ contacts) out to pinned host memory",
                    "one_call_after_the_other": {"value": e2e_sync, "unit": "pairs/s",
                                                 "path": "bam3d_shapes_set_transforms + host-buffer bam3d_gjk_* calls per sub-batch, synchronous"}},
            "gpu_launches": launches, "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                         "kernel": "epa_refill_kernel and gjk_contact_kernel<true> dominate (see the c2 / c3 lines and profiles/)",
                         "kernel_ms_per_step": step_kernel_ms, "algorithmic_bytes_per_pair": ALGO_BYTES["c5"], "peak_source": peak_src,
                         "note": "plain-FP32 SIMT, compute/latency bound"},
            "cpu_baseline": {"value": n_sample / best, "unit": "pairs/s", "cores": cores, "kind": "port",
                             "sample": f"{n_sample} pairs with the same 40/20/20/20 mix, best of 2, std::thread static partition"}}), flush=True)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    if args.workload == "c4":
        return run_ours_c4(args)
    if args.workload == "c5":
        return run_ours_c5(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; bam3d_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    import bam3d_b200 as b3
    from bam3d_b200 import build, multi
    build.build()

    wl, n = args.workload, args.pairs
    scene = make_scene(wl, n, rank)
    ctx = b3.Context(local_rank)
    gjk = b3.GJK.new(ctx)
    ext = torch.cuda.ExternalStream(ctx.stream, device=dev)
    meshes = b3.MeshLibrary(ctx, scene["mesh_verts"], scene["mesh_offsets"]) if scene.get("mesh_verts") is not None else None
    shapes = b3.ShapeSet(ctx, scene["kind"], scene["params"], scene["xform"], scene.get("mesh_id"), meshes)
    shapes_end = b3.ShapeSet(ctx, scene["kind"], scene["params"], scene["xform_end"], scene.get("mesh_id"), meshes) if wl == "toi" else None

    # resident buffers (torch owns the memory; the library gets raw device pointers)
    d_pairs = torch.from_numpy(scene["pairs"].astype(np.int32).reshape(-1)).to(dev)
    d_status = torch.empty(n, dtype=torch.uint8, device=dev)
    d_contacts = torch.empty(n * 8, dtype=torch.float32, device=dev)
    d_compact = torch.empty(n * 10, dtype=torch.int32, device=dev)  # Contact40 records
    d_count = torch.zeros(1, dtype=torch.int64, device=dev)
    gathered = torch.empty(world * n * 10, dtype=torch.int32, device=dev) if world > 1 else None
    counts = torch.zeros(world, dtype=torch.int64, device=dev) if world > 1 else None
    p = lambda t: C.c_void_p(t.data_ptr())

    def step_resident():
        if wl == "c1":
            gjk.intersect_dev(shapes, p(d_pairs), n, p(d_status))
            return
        if wl == "toi":
            gjk.intersection_time_of_impact_dev(shapes, shapes_end, p(d_pairs), n, p(d_contacts), p(d_status))
        else:
            gjk.intersection_dev(b3.CollisionStrategy.FullResolution, shapes, p(d_pairs), n, p(d_contacts), p(d_status))
        if world > 1:
            # compact hits, then all-gather the compacted Contact40 records over NCCL (variable length: counts first)
            gjk.compact_contacts_dev(p(d_contacts), p(d_status), n, rank * n, p(d_compact), n, p(d_count))
            with torch.cuda.stream(ext):
                multi.gather_contacts(d_compact, d_count, counts_buf=counts, out=gathered)

    def sync_all():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- resident throughput ------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None  # nvidia-smi needs ~1 s to start: launch it before warm-up
    for _ in range(args.warmup):
        step_resident()
    sync_all()
    ctx.set_option("timing", 1)
    ctx.read_timing()
    launches0 = ctx.launch_count
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.time()
    ev0.record(ext)
    for _ in range(args.steps):
        step_resident()
    ev1.record(ext)
    sync_all()
    w1 = time.time()
    ms = ev0.elapsed_time(ev1)
    kernel_ms, kernel_calls = ctx.read_timing()
    ctx.set_option("timing", 0)
    launches = ctx.launch_count - launches0
    clocks = None
    # If the timed region was shorter than nvidia-smi's sampling period, keep every GPU under the same load for ~1.5 s
    # (untimed) so that the clocks reported are clocks under this workload. The steps contain collectives when
    # world > 1, so the decision and the repeat count come from rank 0 and EVERY rank runs the same number of steps.
    repeat = torch.zeros(1, dtype=torch.int64, device=dev)
    if sampler and sampler.count(w0, w1) < 3:
        repeat[0] = max(1, int(1.5 / max(ms * 1e-3 / args.steps, 1e-5)))
    if world > 1:
        dist.broadcast(repeat, src=0)
    reps = int(repeat.item())
    if reps:
        w0 = time.time()
        for _ in range(reps):
            step_resident()
        sync_all()
        w1 = time.time()
    if sampler:
        sampler.window(w0, w1)
        clocks = sampler.stop()
        clocks["note"] = ("timed region shorter than the sampling period: sampled during an untimed repeat of the same steps"
                          if reps else "sampled during the timed region")
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * n * args.steps / (ms_max * 1e-3)

    # ---- end to end through the host-buffer C ABI ----------------------------------------------------------
    # per step: pinned host transforms (n_shapes x 64 B) + pair list (n x 8 B) in; status (+ contacts) out
    n_shapes = len(scene["kind"])
    h_xform = torch.from_numpy(scene["xform"]).pin_memory()
    h_pairs = torch.from_numpy(scene["pairs"].astype(np.int32)).pin_memory()
    h_status = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_contacts = torch.empty((n, 8), dtype=torch.float32).pin_memory()
    h_xform_end = torch.from_numpy(scene["xform_end"]).pin_memory() if wl == "toi" else None
    hp = lambda t: C.c_void_p(t.data_ptr())
    lib, h = ctx._lib, ctx.handle

    def step_e2e():
        ctx.check(lib.bam3d_shapes_set_transforms(h, shapes.handle, hp(h_xform)))
        if wl == "c1":
            ctx.check(lib.bam3d_gjk_intersect(h, shapes.handle, hp(h_pairs), n, C.byref(gjk.settings), hp(h_status)))
        elif wl == "toi":
            ctx.check(lib.bam3d_shapes_set_transforms(h, shapes_end.handle, hp(h_xform_end)))
            ctx.check(lib.bam3d_gjk_time_of_impact(h, shapes.handle, shapes_end.handle, hp(h_pairs), n, C.byref(gjk.settings), hp(h_contacts), hp(h_status)))
        else:
            ctx.check(lib.bam3d_gjk_contact(h, shapes.handle, hp(h_pairs), n, C.byref(gjk.settings), 0, hp(h_contacts), hp(h_status)))

    h2d = n_shapes * 64 * (2 if wl == "toi" else 1) + n * 8
    d2h = n * 1 + (0 if wl == "c1" else n * 32)
    e2e_steps = max(3, min(args.steps, 10))

    def timed(fn):
        sync_all()
        t0 = time.perf_counter()
        fn()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * n * e2e_steps / float(t.item())

    # (1) one call after the other: set_transforms, then the host-buffer call (synchronous on return)
    for _ in range(3):
        step_e2e()
    e2e_sync = timed(lambda: [step_e2e() for _ in range(e2e_steps)])
    status_sync = h_status.clone()

    # (2) the same frames through the frame queue (bam3d_frames_*): every frame still copies its transforms and pairs in
    # from pinned host memory and its results out to pinned host memory, but 3 frames are in flight, so the PCIe copies
    # of neighbouring frames run beside the kernels. Every frame has its own result buffers; the timed region ends when
    # the last frame's results are in host memory.
    DEPTH = 3
    frames = C.c_void_p()
    np_kind, np_params = np.ascontiguousarray(scene["kind"], dtype=np.uint8), np.ascontiguousarray(scene["params"], dtype=np.float32)
    np_mesh = None if scene.get("mesh_id") is None else np.ascontiguousarray(scene["mesh_id"], dtype=np.uint32)
    ctx.check(lib.bam3d_frames_create(h, C.c_void_p(np_kind.ctypes.data), C.c_void_p(np_params.ctypes.data), hp(h_xform),
                                      C.c_void_p(np_mesh.ctypes.data) if np_mesh is not None else None, n_shapes,
                                      meshes.handle if meshes is not None else None, n, DEPTH, 1 if wl == "toi" else 0, C.byref(frames)))
    q_status = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(DEPTH)]
    q_contacts = [torch.empty((n, 8), dtype=torch.float32).pin_memory() for _ in range(DEPTH)]
    query = {"c1": 0, "toi": 2}.get(wl, 1)

    def run_queue(steps):
        ticket = C.c_uint64(0)
        for k in range(steps):
            if k >= DEPTH:
                ctx.check(lib.bam3d_frames_wait(h, frames, first + k - DEPTH))  # frees result buffers k % DEPTH
            ctx.check(lib.bam3d_frames_submit(h, frames, query, hp(h_xform), hp(h_xform_end) if wl == "toi" else None, hp(h_pairs), n,
                                              C.byref(gjk.settings), 0, hp(q_contacts[k % DEPTH]), hp(q_status[k % DEPTH]), C.byref(ticket)))
        for k in range(max(0, steps - DEPTH), steps):
            ctx.check(lib.bam3d_frames_wait(h, frames, first + k))

    first = 0
    run_queue(3)
    first = 3
    e2e_value = timed(lambda: run_queue(e2e_steps))
    for b in q_status[:min(DEPTH, e2e_steps)]:
        if not torch.equal(b, status_sync):
            raise SystemExit("bench.py: frame-queue results differ from the synchronous call")
    lib.bam3d_frames_free(h, frames)
    hit_rate = float((h_status == 0).float().mean().item()) if wl != "c1" else float((h_status == 0).float().mean().item())

    if rank == 0:
        peak, peak_src = load_peaks()
        per_launch_ms = kernel_ms / max(1, kernel_calls)
        achieved = ALGO_BYTES[wl] * n / (per_launch_ms * 1e-3) / 1e9
        cores = os.cpu_count() or 1
        n_sample = min(n, args.cpu_pairs)
        once = oracle_run(wl, cpu_sample(scene, wl, n_sample), cores)
        once()
        best = 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            once()
            best = min(best, time.perf_counter() - t0)
        line = {
            "metric": METRIC[wl], "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD_DESC[wl], "pairs_per_gpu": n, "shapes_per_gpu": n_shapes, "hit_rate": hit_rate,
                       "l2": "inputs larger than L2 (shape records 96 B x 2 per pair + pairs + outputs > 126 MB); no flush",
                       "multi_gpu": "contiguous pair slices, per-rank compaction, NCCL all-gather of Contact40 records each step" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "path": f"bam3d_frames_submit / bam3d_frames_wait, {DEPTH} frames in flight: every frame copies its transforms + pairs in "
                            "from pinned host memory and its status (+ contacts) out to pinned host memory; copies of neighbouring "
                            "frames overlap the kernels",
                    "one_call_after_the_other": {"value": e2e_sync, "unit": "pairs/s",
                                                 "path": "bam3d_shapes_set_transforms + host-buffer bam3d_gjk_* call, synchronous on return"}},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": NCU_DRAM_BYTES.get(wl),
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture in profiles/"
                                           if NCU_DRAM_BYTES.get(wl) else None,
                         "kernel": {"c1": "gjk_intersect_kernel<false>", "c2": "epa_refill_kernel (+ gjk_phase_kernel, gjk_contact_overflow_kernel)",
                                    "c3": "gjk_contact_kernel<true>", "toi": "gjk_toi_refill_kernel"}[wl],
                         "kernel_ms_per_launch": per_launch_ms, "algorithmic_bytes_per_pair": ALGO_BYTES[wl], "peak_source": peak_src,
                         "note": "plain-FP32 SIMT, compute/latency bound: see profiles/ for FP32-pipe and branch efficiency"},
            "cpu_baseline": {"value": n_sample / best, "unit": "pairs/s", "cores": cores, "kind": "port",
                             "sample": f"first {n_sample} pairs of rank 0's batch, best of 3, std::thread static partition"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(ALGO_BYTES))
    ap.add_argument("--pairs", type=int, default=None, help="pairs per GPU per step (default 2^20; c5: 2^23)")
    ap.add_argument("--cpu-pairs", type=int, default=1 << 20, help="pairs in the CPU sample")
    ap.add_argument("--boxes", type=int, default=1 << 22, help="c4: boxes per GPU")
    ap.add_argument("--cpu-boxes", type=int, default=1 << 17, help="c4: boxes in the CPU sample (the CPU sweep is O(n^(5/3)))")
    args = ap.parse_args()
    if args.pairs is None:
        args.pairs = (1 << 23) if (args.workload == "c5" and args.impl == "ours") else (1 << 20)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
