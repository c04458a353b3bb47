#!/usr/bin/env python
"""bench.py — the hot path of bam3d on B200: batched GJK / GJK+EPA / time of impact / sweep-and-prune (pairs/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c5|c1|c2|c3|toi|dist|c4] [--no-extra]

The default run measures the configuration BASELINE.json's targets are quoted on and, in `extra`, every other config
at its stated size:
  main   C5: 2^23 mixed convex pairs per GPU (= 64M on 8 GPUs): 40 % analytic GJK+EPA contact, 20 % GJK intersect,
         20 % ConvexPolyhedron GJK+EPA contact, 20 % time of impact; hits compacted and gathered over NCCL every step.
         `intersect_all` / `contact_all` in the same line: GJK::intersect resp. GJK+EPA over ALL of those pairs.
  extra  c1 (2^20 Cuboid pairs, intersect), c2 (2^20 mixed pairs, contact), c3 (2^22 polyhedron pairs, contact),
         toi (2^20), dist (2^20, GJK::distance), c4 (SweepAndPrune3 over 2^22 boxes, BVH build, 2^20 ray / 1024 frustum
         queries, and the broad -> narrow pipeline) — each with its own roofline, cpu_baseline and e2e.
A step is one pass of the path over one batch.
`value`  = pairs/s over all ranks with everything resident in HBM (device-pointer C-ABI calls, CUDA events, max over
           ranks). Consecutive steps alternate between two contexts (own stream + scratch each), so step k+1 starts
           while step k's persistent EPA kernel drains and while step k's contacts travel over NCCL.
`e2e`    = the same through the host-buffer frame queue a user of the crate calls each frame: pinned host transforms +
           pair list in, status + contacts out (N > 1: the gathered hits of all ranks), copies inside the timed region.
`roofline` = algorithmic HBM bytes of the dominant kernel / its event-timed duration (one context alone, right after
           the timed region) vs MEASURED_PEAKS.json; `ncu` = the counters of the committed capture of the same kernel.
`cpu_baseline` = the CPU oracle (strict-f32 C++ restatement of the reference; the Rust crate cannot be built in this
           image) on the host cores, bounded sample. `--impl reference` times only that.
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
    "dist": 137.0,  # 128 B in + 2 x 4 B + 1 B out
    "c4": 220.0,   # per box: 24 B read + 4 B perm + 8 radix passes x 2 x 12 B; plus 8 B per emitted pair (added below)
}
ALGO_BYTES["c5"] = 0.4 * ALGO_BYTES["c2"] + 0.2 * ALGO_BYTES["c1"] + 0.2 * ALGO_BYTES["c3"] + 0.2 * ALGO_BYTES["toi"]
SIZES = {"c1": 1 << 20, "c2": 1 << 20, "c3": 1 << 22, "toi": 1 << 20, "dist": 1 << 20, "c5": 1 << 23}
WORKLOAD_DESC = {
    "c1": "C1: GJK intersect, 2^20 random Cuboid-Cuboid pairs per GPU (f32, random transforms, offset U[-2,2]^3)",
    "c2": "C2: GJK+EPA Contact (FullResolution), 2^20 mixed Sphere/Cuboid/Cylinder pairs per GPU (offset U[-1.5,1.5]^3)",
    "c3": "C3: GJK+EPA Contact, 2^22 ConvexPolyhedron pairs per GPU (4096-mesh library, 64-256 vertices, warp-cooperative support)",
    "toi": "TOI: GJK time of impact, 2^20 mixed Sphere/Cuboid/Cylinder pairs per GPU (end = start + U[-3,3]^3)",
    "dist": "DIST: GJK distance, 2^20 mixed Sphere/Cuboid/Cylinder pairs per GPU (the C2 scene)",
    "c5": "C5: 2^23 mixed convex pairs per GPU (64M on 8 GPUs): 40 % analytic GJK+EPA contact (C2 mix), 20 % GJK intersect only, "
          "20 % ConvexPolyhedron GJK+EPA contact (C3 library), 20 % GJK time of impact; hits compacted and gathered over NCCL every step",
    "c4": "C4: SweepAndPrune3 broad phase over 2^22 Aabb3 per GPU (centres U[0,100]^3, half-extents U[0.1,0.5]^3), BVH build, "
          "2^20 ray and 1024 frustum queries, and the broad -> narrow pipeline, in `extra`",
}
METRIC = {"c1": "gjk_intersect_pairs_per_sec", "c2": "gjk_epa_contact_pairs_per_sec", "c3": "gjk_epa_contact_pairs_per_sec",
          "toi": "gjk_toi_pairs_per_sec", "dist": "gjk_distance_pairs_per_sec", "c4": "sap_candidate_pairs_per_sec",
          "c5": "mixed_narrow_phase_pairs_per_sec"}
KERNEL = {"c1": "gjk_intersect_kernel<false>", "c2": "epa_refill_kernel (+ gjk_phase_kernel, gjk_contact_overflow_kernel)",
          "c3": "gjk_contact_kernel<true>", "toi": "gjk_toi_refill_kernel", "dist": "gjk_distance_refill_kernel",
          "c5": "epa_refill_kernel and gjk_contact_kernel<true> dominate (see extra.c2 / extra.c3)"}
REF_NOTE = "reference crate (Rust) cannot be built in this image; this arm times the strict-f32 C++ restatement of its code path"


def config_for(wl, n):
    """The `config` object — identical in both arms (it names the workload, not the sample an arm measured on)."""
    return {"workload": WORKLOAD_DESC[wl], "pairs_per_gpu": int(n),
            "l2": "inputs larger than L2 (shape records 96 B x 2 per pair + pairs + outputs > 126 MB); no flush"}


def make_scene(workload, n, rank):
    from bam3d_b200 import scenes
    start = rank * n  # every rank generates exactly its own slice of the global scene
    if workload == "c1":
        return scenes.cuboid_pairs(n, s=2.0, start=start)
    if workload in ("c2", "toi", "dist"):
        sc = scenes.mixed_pairs(n, s=1.5 if workload != "toi" else 4.0, start=start)
        if workload == "toi":
            sc["xform_end"] = scenes.end_transforms(sc, start=2 * start)
        return sc
    if workload == "c3":
        return scenes.polyhedron_pairs(n, n_meshes=4096, s=2.0, start=start)
    raise SystemExit(f"unknown workload {workload}")


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


def c5_sample(parts, n_sample):
    """The same 40/20/20/20 mix on the first pairs of every sub-batch."""
    out = {}
    for k, f in C5_PARTS:
        sub = dict(parts[k])
        sub["pairs"] = parts[k]["pairs"][: max(1, int(n_sample * f))]
        out[k] = sub
    return out


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores (the reference crate itself cannot be built here: no cargo / rustc)
# ---------------------------------------------------------------------------------------------------------
_ORACLE = None


def oracle():
    global _ORACLE
    if _ORACLE is None:
        from tests.oracle_ffi import Oracle
        subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
        _ORACLE = Oracle()
    return _ORACLE


def oracle_once(workload, scene, nthreads):
    orc = oracle()
    if workload == "c5":
        def run():
            orc.contact(scene["contact"], strategy=0, nthreads=nthreads)
            orc.intersect(scene["intersect"], nthreads=nthreads)
            orc.contact(scene["poly"], strategy=0, nthreads=nthreads)
            orc.toi(scene["toi"], scene["toi"]["xform_end"], nthreads=nthreads)
        return run
    if workload == "c1":
        return lambda: orc.intersect(scene, nthreads=nthreads)
    if workload == "toi":
        return lambda: orc.toi(scene, scene["xform_end"], nthreads=nthreads)
    if workload == "dist":
        return lambda: orc.distance(scene, nthreads=nthreads)
    return lambda: orc.contact(scene, strategy=0, nthreads=nthreads)


def cpu_time(workload, scene, n_sample, budget_s, repeats=3):
    """Oracle throughput on the first n_sample pairs (c5: the same mix), best of `repeats`, within about budget_s seconds."""
    cores = os.cpu_count() or 1
    if workload == "c5":
        sample = c5_sample(scene, n_sample)
        n_done = sum(len(v["pairs"]) for v in sample.values())
    else:
        sample = dict(scene)
        sample["pairs"] = scene["pairs"][:n_sample]
        n_done = len(sample["pairs"])
    once = oracle_once(workload, sample, cores)
    runs = [timed_cpu(once)]
    for _ in range(repeats - 1):
        if runs[0][0] * 2 > budget_s:
            break
        runs.append(timed_cpu(once))
    wall, cpu_s = min(runs)
    return cpu_baseline_object(n_done, wall, cpu_s, cores, f"first {n_done} pairs of rank 0's batch" +
                               (" (same 40/20/20/20 mix)" if workload == "c5" else "") + f", best of {len(runs)}")


def timed_cpu(fn):
    """-> (wall seconds, CPU seconds summed over the process's threads) of one call."""
    w0, c0 = time.perf_counter(), time.process_time()
    fn()
    return time.perf_counter() - w0, time.process_time() - c0


CPU_VALUE_IS = ("pairs / (CPU seconds summed over the threads / cores): the rate of a perfectly balanced run. A bounded sample is not balanced — "
                "one pinched EPA polytope (10^5 faces; the reference's add() is quadratic in it) runs 3 s on one thread while the other "
                "threads have finished the rest of a 1M-pair sample in 0.3 s — but the full workload is (8 such samples per GPU and "
                "step), so the wall-clock rate of the sample (`wall_value`) understates what these cores would do on it")


def cpu_baseline_object(n_done, wall, cpu_s, cores, what):
    balanced = max(cpu_s / cores, 1e-9)
    return {"value": n_done / balanced if cpu_s > 0 else n_done / wall, "unit": "pairs/s", "cores": cores, "kind": "port",
            "wall_value": n_done / wall, "wall_s": wall, "cpu_s": cpu_s, "value_is": CPU_VALUE_IS,
            "sample": what + f", std::thread, 256-pair chunks from a shared counter, {cores} threads"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = args.workload
    cores = os.cpu_count() or 1
    if wl == "c4":
        return run_reference_c4(args)
    n = args.pairs
    n_sample = min(n, args.cpu_pairs)
    if wl == "c5":
        parts, _ = c5_scenes(max(n_sample, 8192), 0)
        scene = c5_sample(parts, n_sample)
        n_done = sum(len(v["pairs"]) for v in scene.values())
    else:
        scene = make_scene(wl, n_sample, 0)
        n_done = n_sample
    once = oracle_once(wl, scene, cores)
    for _ in range(args.warmup):
        once()
    wall, cpu_s = timed_cpu(lambda: [once() for _ in range(args.steps)])
    base = cpu_baseline_object(n_done * args.steps, wall, cpu_s, cores, f"{n_done} pairs of the workload per step")
    value = base["value"]
    cfg = config_for(wl, n)
    line = {
        "impl": "reference", "metric": METRIC[wl], "value": value, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * n_done / value, "wall_ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "note": REF_NOTE,
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    if wl == "c5" and not args.no_extra:
        # the other metrics BASELINE.json names, on the same cores: GJK intersect (C1), GJK+EPA (C2), SaP candidate pairs
        extra = {}
        for sub, ns in (("c1", 1 << 19), ("c2", 1 << 18)):
            extra[sub] = cpu_time(sub, make_scene(sub, ns, 0), ns, 6.0)
        extra["c4"] = cpu_c4(min(args.boxes, args.cpu_boxes))
        line["extra"] = extra
    print(json.dumps(line), flush=True)
    return 0


def cpu_c4(ns):
    from bam3d_b200 import scenes
    orc = oracle()
    sboxes, _ = scenes.aabb_scene(ns, extent=100.0 * (ns / float(1 << 22)) ** (1.0 / 3.0))
    t0 = time.perf_counter(); opairs, _, _ = orc.sap_find_pairs(sboxes, 0); dt = time.perf_counter() - t0
    return {"value": len(opairs) / dt, "unit": "pairs/s", "cores": 1, "kind": "port",
            "sample": f"{ns} boxes at C4 density, one sequential sweep (the reference's algorithm is sequential)"}


def run_reference_c4(args):
    """CPU arm of C4: the oracle's sequential sweep (the reference is single-threaded here: one sort + one sweep)."""
    n = min(args.boxes, args.cpu_boxes)
    steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(steps):
        cb = cpu_c4(n)
    dt = time.perf_counter() - t0
    line = {"impl": "reference", "metric": METRIC["c4"], "value": cb["value"], "unit": "pairs/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": 0, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD_DESC["c4"], "boxes_per_gpu": args.boxes}, "note": REF_NOTE,
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def summary(self, windows):
        """Clocks over the samples taken inside any of the [t0, t1] windows (the timed regions)."""
        sm, mx, reasons = [], [], set()
        for t, r in self.rows:
            if not any(a - 0.05 <= t <= b + 0.05 for a, b in windows):
                continue
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for k, nme in enumerate(self.NAMES):
                if f[2 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "note": "sampled during the timed regions of this line"}

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_ncu():
    """Counters of the committed ncu --set full captures (profiles/ncu_counters.json, written by profiles/summarize.py from
    the .ncu-rep of the build named in it). Nothing is hard-coded here: no file, no counters."""
    p = os.path.join(ROOT, "profiles", "ncu_counters.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def roofline_for(wl, algo_bytes, kernel_ms, units_note):
    peak, peak_src = load_peaks()
    achieved = algo_bytes / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
    ncu = load_ncu().get(wl)
    r = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
         "traffic": (ncu or {}).get("dram_bytes"), "kernel": KERNEL.get(wl, wl), "kernel_ms_per_launch": kernel_ms,
         "algorithmic_bytes": algo_bytes, "algorithmic_note": units_note, "peak_source": peak_src,
         "timing": "CUDA events around the bracketed main kernels (bam3d_ctx_set_option(\"timing\")), one context alone, "
                   "in the same process right after the timed region",
         "note": "plain-FP32 SIMT, compute/latency bound: the ncu object holds FP32-pipe, branch and lane efficiency"}
    if ncu:
        r["ncu"] = ncu  # fp32_pipe_pct, branch_uniform_pct, lanes_active, dram_gbs, source file under profiles/
        if ncu.get("duration_ms_under_ncu"):
            # the live bracket covers every main kernel of the call (for C2: GJK phase + EPA + both overflow tiers, serialised on one
            # context); the dominant kernel alone, from the committed capture of this build — not measured by this run:
            r["dominant_kernel_alone"] = {"kernel": ncu.get("kernel"), "ms_under_ncu": ncu["duration_ms_under_ncu"],
                                          "frac": algo_bytes / (ncu["duration_ms_under_ncu"] * 1e-3) / 1e9 / peak, "source": ncu.get("source")}
    return r


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
class Env:
    """Process-wide state of the GPU arm: torch.distributed (rendezvous and scalar reductions only), two library
    contexts (the compute lanes) and the library's own NCCL communicator for the contact exchange."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0")); self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; bam3d_b200 has no CPU fallback")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        import bam3d_b200 as b3
        from bam3d_b200 import build, multi
        build.build()
        self.b3 = b3
        # compute lanes: library contexts with their own stream and scratch; step k runs on lane k % lanes, so the steps in
        # flight overlap each other's serial tails (EPA drain, overflow tiers) — BAM3D_BENCH_LANES (default 2)
        self.lanes = max(1, int(os.environ.get("BAM3D_BENCH_LANES", "2")))
        self.ctxs = [b3.Context(self.local_rank) for _ in range(self.lanes)]
        self.gjks = [b3.GJK.new(c) for c in self.ctxs]
        self.lib = self.ctxs[0]._lib
        self.ext = [torch.cuda.ExternalStream(c.stream, device=self.dev) for c in self.ctxs]
        # development knobs: BAM3D_BENCH_NO_EXCHANGE=1 times the ranks without the contact exchange (what the exchange costs),
        # BAM3D_BENCH_SLICE=k generates slice k of the scene on a single GPU (how much the slices differ)
        self.comm = multi.Comm(self.local_rank, self.rank, self.world) if self.world > 1 and not os.environ.get("BAM3D_BENCH_NO_EXCHANGE") else None
        self.slice = int(os.environ.get("BAM3D_BENCH_SLICE", self.rank))
        self.sampler = ClockSampler(self.local_rank) if self.rank == 0 else None
        self.windows = []

    def sync_all(self):
        for c in self.ctxs:
            c.synchronize()
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, x):
        t = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def launches(self):
        return sum(c.launch_count for c in self.ctxs)

    def timed(self, enqueue, steps, warmup):
        """`warmup` untimed steps, then exactly `steps` steps bracketed by barrier + synchronize on both sides, timed with
        CUDA events on lane 0's stream (the end event waits for lane 1 and the exchange). enqueue(k) enqueues step k on lane
        k % lanes and returns nothing; enqueue.finish() completes whatever is still pending (exchange). -> (ms max over ranks, launches)"""
        torch = self.torch
        for k in range(warmup):
            enqueue(k)
        enqueue.finish()
        self.sync_all()
        l0 = self.launches()
        ev0, ev1, evl = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event()
        w0 = time.time()
        ev0.record(self.ext[0])
        for k in range(steps):
            enqueue(warmup + k)
        enqueue.finish()
        for x in self.ext[1:]:
            evl.record(x)
            self.ext[0].wait_event(evl)
        if self.comm is not None:
            self.comm.join(self.ctxs[0])
        ev1.record(self.ext[0])
        self.sync_all()
        w1 = time.time()
        self.windows.append((w0, w1))
        return self.max_over_ranks(ev0.elapsed_time(ev1)), self.launches() - l0

    def serial_kernel_ms(self, call, reps=5):
        """Event-timed duration of the bracketed main kernels of `call(ctx0 lane)` run alone, averaged over `reps`."""
        c = self.ctxs[0]
        self.sync_all()
        c.set_option("timing", 1); c.read_timing()
        for _ in range(reps):
            call()
        ms, calls = c.read_timing()
        c.set_option("timing", 0)
        return ms / max(1, reps), calls / max(1, reps)

    def clocks(self):
        if not self.sampler:
            return None
        c = self.sampler.summary(self.windows)
        self.windows = []
        return c


def dptr(t, off=0):
    return C.c_void_p(t.data_ptr() + off)


class Exchange:
    """Per-lane buffers and bookkeeping of the contact exchange for a resident loop. Step k (lane k % L) is: its kernels; then the
    payload of step k - L (the step that last used this lane: by now its counts are on the host, so the call rarely blocks);
    then compaction and the count exchange of step k. The host blocks only when it is L steps ahead of the device, so every
    lane always has its next step queued behind the running one — finishing step k - 1 here instead (one step less latency for
    the payload) leaves the lane idle from the moment its step ends until the host has seen the counts, relaunched and the peers
    have caught up: 44.5 instead of 41.4 ms per C5 step on 2 GPUs."""

    def __init__(self, env, n_dense):
        torch = env.torch
        self.env, self.n = env, n_dense
        L = env.lanes
        self.compact = [torch.empty(n_dense * 10, dtype=torch.int32, device=env.dev) for _ in range(L)]
        self.count = [torch.zeros(1, dtype=torch.int64, device=env.dev) for _ in range(L)]
        self.gathered = [torch.empty(env.world * n_dense * 10, dtype=torch.int32, device=env.dev) for _ in range(L)] if env.comm is not None else None
        self.ticket = [None] * L   # per lane: the exchange whose payload has not been enqueued yet
        self.total = 0

    def after_step(self, lane, d_contacts, d_status):
        env = self.env
        gjk, ctx = env.gjks[lane], env.ctxs[lane]
        if env.comm is not None and self.ticket[lane] is not None:
            t = self.ticket[lane]
            self.finish_lane(lane)
            env.comm.wait(ctx, t)   # the payload reads this lane's compact buffer
        gjk.compact_contacts_dev(dptr(d_contacts), dptr(d_status), self.n, env.rank * self.n, dptr(self.compact[lane]), self.n, dptr(self.count[lane]))
        if env.comm is not None:
            self.ticket[lane] = env.comm.begin(ctx, dptr(self.count[lane]))

    def finish_lane(self, lane):
        env = self.env
        _, self.total = env.comm.finish(env.ctxs[lane], self.ticket[lane], dptr(self.compact[lane]), dptr(self.gathered[lane]), env.world * self.n)
        self.ticket[lane] = None

    def finish(self):
        """The payloads still owed, oldest first (every rank issues its collectives in the same order)."""
        if self.env.comm is None:
            return
        for lane in sorted((l for l in range(self.env.lanes) if self.ticket[l] is not None), key=lambda l: self.ticket[l]):
            self.finish_lane(lane)


def e2e_frames(env, queues, steps, world_gather):
    """Run `steps` frames through the frame queues (`queues`: list of dicts with handle, query, hx, hx_end, hpairs, n, bufs[depth],
    base), depth 2. Returns seconds (wall, max over ranks) with a barrier + synchronize on both sides."""
    lib, ctx, comm = env.lib, env.ctxs[0], env.comm
    h = ctx.handle
    S = C.byref(env.gjks[0].settings)
    hp = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    DEPTH = 2

    def run(first, count):
        ticket = C.c_uint64(0)
        for s_ in range(count):
            f = first + s_
            if s_ >= DEPTH:
                for q in queues:
                    ctx.check(lib.bam3d_frames_wait(h, q["handle"], f - DEPTH))
            b = f % DEPTH
            for q in queues:
                buf = q["bufs"][b]
                if world_gather and q["query"] != 0:
                    ctx.check(lib.bam3d_frames_submit_gather(h, q["handle"], comm.handle, q["query"], hp(q["hx"]), hp(q.get("hx_end")), hp(q["hpairs"]), q["n"], S, 0,
                                                             q["base"], hp(buf["gathered"]), buf["gathered"].numel() // 10, hp(buf["status"]), C.byref(ticket)))
                else:
                    ctx.check(lib.bam3d_frames_submit(h, q["handle"], q["query"], hp(q["hx"]), hp(q.get("hx_end")), hp(q["hpairs"]), q["n"], S, 0,
                                                      hp(buf.get("contacts")), hp(buf["status"]), C.byref(ticket)))
        for f in range(first + max(0, count - DEPTH), first + count):
            for q in queues:
                ctx.check(lib.bam3d_frames_wait(h, q["handle"], f))
        return first + count

    nxt = run(0, 2)
    env.sync_all()
    w0 = time.time(); t0 = time.perf_counter()
    run(nxt, steps)
    dt = time.perf_counter() - t0
    env.windows.append((w0, time.time()))
    return env.max_over_ranks(dt)


def make_queue(env, scene, meshes, n, query, base, gather_cap):
    """One frame queue over `scene` (+ pinned host inputs and two sets of pinned result buffers). gather_cap: records the
    pinned buffer for the gathered hits of all ranks holds (None: this rank's dense contacts come back instead)."""
    torch, lib, ctx = env.torch, env.lib, env.ctxs[0]
    kind, params = np.ascontiguousarray(scene["kind"], dtype=np.uint8), np.ascontiguousarray(scene["params"], dtype=np.float32)
    mesh_id = None if scene.get("mesh_id") is None else np.ascontiguousarray(scene["mesh_id"], dtype=np.uint32)
    # per-frame transforms in the 48-byte layout (the four columns' xyz: a model transform's fourth row is (0,0,0,1), and the
    # 64-byte form is rejected when it is not); BAM3D_BENCH_MAT4=1 uploads whole Mat4s instead
    from bam3d_b200 import api as b3api, _ffi
    affine = not os.environ.get("BAM3D_BENCH_MAT4")
    conv = (lambda x: b3api.affine3x4(x)) if affine else (lambda x: x)
    x0 = np.ascontiguousarray(scene["xform"], dtype=np.float32)
    hx = torch.from_numpy(conv(x0)).pin_memory()
    q = {"query": query, "n": n, "base": base, "hx": hx, "hpairs": torch.from_numpy(scene["pairs"].astype(np.int32)).pin_memory()}
    if query == 2:
        q["hx_end"] = torch.from_numpy(conv(scene["xform_end"])).pin_memory()
    handle = C.c_void_p()
    ctx.check(lib.bam3d_frames_create(ctx.handle, C.c_void_p(kind.ctypes.data), C.c_void_p(params.ctypes.data), C.c_void_p(x0.ctypes.data),
                                      C.c_void_p(mesh_id.ctypes.data) if mesh_id is not None else None, len(kind),
                                      meshes.handle if meshes is not None else None, n, 2, 1 if query == 2 else 0, C.byref(handle)))
    if affine:
        ctx.check(lib.bam3d_frames_set_transform_layout(ctx.handle, handle, _ffi.XFORM_AFFINE3X4))
    q["handle"] = handle
    q["xform_bytes"] = 48 if affine else 64
    bufs = []
    for _ in range(2):
        b = {"status": torch.empty(n, dtype=torch.uint8).pin_memory()}
        if query != 0:
            if gather_cap is not None:
                b["gathered"] = torch.empty((gather_cap, 10), dtype=torch.int32).pin_memory()
            else:
                b["contacts"] = torch.empty((n, 8), dtype=torch.float32).pin_memory()
        bufs.append(b)
    q["bufs"] = bufs
    q["h2d"] = len(kind) * q["xform_bytes"] * (2 if query == 2 else 1) + n * 8
    return q


def bench_narrow(env, wl, n, steps, warmup, cpu_pairs, cpu_budget):
    """One narrow-phase workload (c1 / c2 / c3 / toi / dist) on every rank's own slice -> the JSON object (rank 0) or None."""
    torch, b3 = env.torch, env.b3
    scene = make_scene(wl, n, env.rank)
    ctx0 = env.ctxs[0]
    meshes = b3.MeshLibrary(ctx0, scene["mesh_verts"], scene["mesh_offsets"]) if scene.get("mesh_verts") is not None else None
    shapes = b3.ShapeSet(ctx0, scene["kind"], scene["params"], scene["xform"], scene.get("mesh_id"), meshes)
    shapes_end = b3.ShapeSet(ctx0, scene["kind"], scene["params"], scene["xform_end"], scene.get("mesh_id"), meshes) if wl == "toi" else None
    n_shapes = len(scene["kind"])
    d_pairs = torch.from_numpy(scene["pairs"].astype(np.int32).reshape(-1)).to(env.dev)
    d_status = [torch.empty(n, dtype=torch.uint8, device=env.dev) for _ in range(env.lanes)]
    has_contacts = wl in ("c2", "c3", "toi")
    d_contacts = [torch.empty(n * 8, dtype=torch.float32, device=env.dev) for _ in range(env.lanes)] if has_contacts else None
    d_f = [torch.empty(2 * n, dtype=torch.float32, device=env.dev) for _ in range(env.lanes)] if wl == "dist" else None
    xch = Exchange(env, n) if (has_contacts and env.world > 1) else None
    FR = b3.CollisionStrategy.FullResolution

    def compute(lane):
        gjk = env.gjks[lane]
        if wl == "c1":
            gjk.intersect_dev(shapes, dptr(d_pairs), n, dptr(d_status[lane]))
        elif wl == "toi":
            gjk.intersection_time_of_impact_dev(shapes, shapes_end, dptr(d_pairs), n, dptr(d_contacts[lane]), dptr(d_status[lane]))
        elif wl == "dist":
            gjk.distance_dev(shapes, dptr(d_pairs), n, dptr(d_f[lane]), dptr(d_f[lane], 4 * n), dptr(d_status[lane]))
        else:
            gjk.intersection_dev(FR, shapes, dptr(d_pairs), n, dptr(d_contacts[lane]), dptr(d_status[lane]))

    def enqueue(k):
        lane = k % env.lanes
        compute(lane)
        if xch is not None:
            xch.after_step(lane, d_contacts[lane], d_status[lane])
    enqueue.finish = (lambda: xch.finish()) if xch is not None else (lambda: None)

    ms, launches = env.timed(enqueue, steps, warmup)
    value = env.world * n * steps / (ms * 1e-3)
    kernel_ms, calls = env.serial_kernel_ms(lambda: compute(0))
    hit_rate = float((d_status[0] == 0).float().mean().item())

    # ---- end to end: the frame queue, host buffers, copies inside the timed region ---------------------------
    gather = has_contacts and env.world > 1
    gcap = int(1.1 * hit_rate * n * env.world) + 65536 if gather else None  # every rank's slice has the same hit rate up to sampling noise
    q = make_queue(env, scene, meshes, n, {"c1": 0, "toi": 2}.get(wl, 1), env.rank * n, gcap) if wl != "dist" else None
    e2e_steps = max(3, steps)  # the same K steps, pipeline fill and drain inside the timed region
    if q is not None:
        dt = e2e_frames(env, [q], e2e_steps, gather)
        e2e_value = env.world * n * e2e_steps / dt
        d2h = n + (0 if wl == "c1" else (int(hit_rate * n * env.world) * 40 if gather else n * 32))
        e2e = {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": q["h2d"], "d2h_bytes_per_step": d2h, "steps": e2e_steps,
               "path": "bam3d_frames_submit" + ("_gather" if gather else "") + " / bam3d_frames_wait, 2 frames in flight on two compute lanes: every frame copies its "
                       "transforms (48 B each: the Mat4 without its constant fourth row) + pairs in from pinned host memory and its status + " + ("the gathered hits of all ranks" if gather else "contacts") +
                       " out to pinned host memory"}
        # the queue's results are the resident path's (bit for bit)
        st_q = q["bufs"][0]["status"]
        if not torch.equal(st_q, d_status[0].cpu()):
            raise SystemExit(f"bench.py [{wl}]: frame-queue status differs from the resident path")
        env.lib.bam3d_frames_free(ctx0.handle, q["handle"])
    else:
        # GJK::distance has no frame-queue form: the host-buffer call, synchronous on return
        h_pairs = torch.from_numpy(scene["pairs"].astype(np.int32)).pin_memory()
        h_x = torch.from_numpy(scene["xform"]).pin_memory()
        h_out = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(2)]
        h_st = torch.empty(n, dtype=torch.uint8).pin_memory()
        hp = lambda t: C.c_void_p(t.data_ptr())

        def step_e2e():
            ctx0.check(env.lib.bam3d_shapes_set_transforms(ctx0.handle, shapes.handle, hp(h_x)))
            ctx0.check(env.lib.bam3d_gjk_distance(ctx0.handle, shapes.handle, hp(h_pairs), n, C.byref(env.gjks[0].settings), hp(h_out[0]), hp(h_out[1]), hp(h_st)))
        step_e2e(); env.sync_all()
        w0 = time.time(); t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        dt = env.max_over_ranks(time.perf_counter() - t0)
        env.windows.append((w0, time.time()))
        e2e = {"value": env.world * n * e2e_steps / dt, "unit": "pairs/s", "h2d_bytes_per_step": n_shapes * 64 + n * 8, "d2h_bytes_per_step": n * 9, "steps": e2e_steps,
               "path": "bam3d_shapes_set_transforms + host-buffer bam3d_gjk_distance, synchronous on return"}

    line = None
    if env.rank == 0:
        cfg = config_for(wl, n)
        cfg.update({"shapes_per_gpu": n_shapes, "hit_rate": hit_rate,
                    "multi_gpu": "contiguous pair slices, per-rank compaction, NCCL gather of Contact40 records each step (library communicator, overlapped with the next step)"
                    if xch is not None else ("contiguous pair slices, no exchange (status only)" if env.world > 1 else "single GPU")})
        line = {"metric": METRIC[wl], "value": value, "unit": "pairs/s", "n_gpus": env.world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg, "e2e": e2e, "gpu_launches": launches, "clocks": env.clocks(),
                "roofline": roofline_for(wl, ALGO_BYTES[wl] * n, kernel_ms, f"{ALGO_BYTES[wl]} B/pair x {n} pairs per launch"),
                "cpu_baseline": cpu_time(wl, scene, min(n, cpu_pairs), cpu_budget)}
    for s_ in (shapes, shapes_end):
        if s_ is not None:
            s_.close()
    if meshes is not None:
        meshes.close()
    del d_pairs, d_status, d_contacts, d_f, xch
    torch.cuda.empty_cache()
    return line


def bench_c5(env, n, steps, warmup, cpu_pairs, cpu_budget):
    """The 64M mixed configuration, one contiguous slice per GPU: the mixed step (`value`), plus GJK::intersect and GJK+EPA
    contact over ALL of the slice's pairs (the two rates the north-star targets are quoted on)."""
    torch, b3 = env.torch, env.b3
    parts, sizes = c5_scenes(n, env.slice)
    ctx0 = env.ctxs[0]
    meshes = b3.MeshLibrary(ctx0, parts["poly"]["mesh_verts"], parts["poly"]["mesh_offsets"])
    shp = {k: b3.ShapeSet(ctx0, v["kind"], v["params"], v["xform"], v.get("mesh_id"), meshes if k == "poly" else None) for k, v in parts.items()}
    shp_end = b3.ShapeSet(ctx0, parts["toi"]["kind"], parts["toi"]["params"], parts["toi"]["xform_end"])
    d_pairs = {k: torch.from_numpy(v["pairs"].astype(np.int32).reshape(-1)).to(env.dev) for k, v in parts.items()}
    # dense outputs of all four sub-batches share one buffer per lane (the mixed step leaves the intersect part's contacts unused)
    order = ("contact", "poly", "toi", "intersect")
    off, acc = {}, 0
    for k in order:
        off[k] = acc
        acc += sizes[k]
    n_dense = acc - sizes["intersect"]  # the three contact-producing sub-batches come first
    d_status = [torch.empty(n, dtype=torch.uint8, device=env.dev) for _ in range(env.lanes)]
    d_contacts = [torch.empty(n * 8, dtype=torch.float32, device=env.dev) for _ in range(env.lanes)]
    FR = b3.CollisionStrategy.FullResolution

    def mixed(lane):
        g = env.gjks[lane]
        c, s = d_contacts[lane], d_status[lane]
        g.intersection_dev(FR, shp["contact"], dptr(d_pairs["contact"]), sizes["contact"], dptr(c, 32 * off["contact"]), dptr(s, off["contact"]))
        g.intersect_dev(shp["intersect"], dptr(d_pairs["intersect"]), sizes["intersect"], dptr(s, off["intersect"]))
        g.intersection_dev(FR, shp["poly"], dptr(d_pairs["poly"]), sizes["poly"], dptr(c, 32 * off["poly"]), dptr(s, off["poly"]))
        g.intersection_time_of_impact_dev(shp["toi"], shp_end, dptr(d_pairs["toi"]), sizes["toi"], dptr(c, 32 * off["toi"]), dptr(s, off["toi"]))

    def intersect_all(lane):
        g = env.gjks[lane]
        for k in order:
            g.intersect_dev(shp[k], dptr(d_pairs[k]), sizes[k], dptr(d_status[lane], off[k]))

    def contact_all(lane):
        g = env.gjks[lane]
        for k in order:
            g.intersection_dev(FR, shp[k], dptr(d_pairs[k]), sizes[k], dptr(d_contacts[lane], 32 * off[k]), dptr(d_status[lane], off[k]))

    def loop(fn, n_x, with_exchange):
        xch = Exchange(env, n_x) if with_exchange else None

        def enqueue(k):
            lane = k % env.lanes
            fn(lane)
            if xch is not None:
                xch.after_step(lane, d_contacts[lane], d_status[lane])
        enqueue.finish = (lambda: xch.finish()) if xch is not None else (lambda: None)
        ms, launches = env.timed(enqueue, steps, warmup)
        hits = int(xch.count[0].item()) if xch is not None else None
        return ms, launches, hits

    # the mixed step: compaction always (it is part of the path), exchange when there is more than one rank
    ms, launches, n_hits = loop(mixed, n_dense, True)
    value = env.world * n * steps / (ms * 1e-3)
    kernel_ms, _ = env.serial_kernel_ms(lambda: mixed(0), reps=3)
    status_mixed = d_status[0].cpu()
    ms_i, launches_i, _ = loop(intersect_all, n, False)
    ms_c, launches_c, hits_c = loop(contact_all, n, True)

    # ---- end to end: one frame queue per sub-batch, 2 frames in flight each ------------------------------------
    gather = env.comm is not None
    query = {"contact": 1, "intersect": 0, "poly": 1, "toi": 2}
    base = {k: env.rank * n + off[k] for k in order}
    part_hits = {k: int((status_mixed[off[k]: off[k] + sizes[k]] == 0).sum().item()) for k in order}
    queues = [make_queue(env, parts[k], meshes if k == "poly" else None, sizes[k], query[k], base[k],
                         int(1.1 * part_hits[k] * env.world) + 65536 if gather else None) for k in ("contact", "intersect", "poly", "toi")]
    e2e_steps = max(3, steps)  # the same K steps, pipeline fill and drain inside the timed region
    dt = e2e_frames(env, queues, e2e_steps, gather)
    e2e_value = env.world * n * e2e_steps / dt
    for q, k in zip(queues, ("contact", "intersect", "poly", "toi")):
        if not torch.equal(q["bufs"][0]["status"], status_mixed[off[k]: off[k] + sizes[k]]):
            raise SystemExit(f"bench.py [c5]: frame-queue status of the {k} sub-batch differs from the resident path")
        env.lib.bam3d_frames_free(ctx0.handle, q["handle"])
    h2d = sum(q["h2d"] for q in queues)
    d2h = n + (int((n_hits or 0) * env.world) * 40 if gather else n_dense * 32)

    line = None
    if env.rank == 0:
        cfg = config_for("c5", n)
        cfg.update({"sub_batches": sizes, "contacts_per_gpu_rank0": n_hits,
                    "multi_gpu": "contiguous pair slices, per-rank compaction, NCCL gather of Contact40 records each step (library communicator, overlapped with the next step)"
                    if env.world > 1 else "single GPU (compaction inside the step, nothing to exchange)"})
        line = {"metric": METRIC["c5"], "value": value, "unit": "pairs/s", "n_gpus": env.world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": cfg,
                "intersect_all": {"metric": "gjk_intersect_pairs_per_sec", "value": env.world * n * steps / (ms_i * 1e-3), "unit": "pairs/s", "ms_per_step": ms_i / steps,
                                  "what": "GJK::intersect over ALL pairs of the C5 slice (analytic + polyhedron), status only, no exchange"},
                "contact_all": {"metric": "gjk_epa_contact_pairs_per_sec", "value": env.world * n * steps / (ms_c * 1e-3), "unit": "pairs/s", "ms_per_step": ms_c / steps,
                                "contacts_per_gpu_rank0": hits_c,
                                "what": "GJK+EPA Contact (FullResolution) over ALL pairs of the C5 slice, compaction + NCCL gather inside"},
                "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                        "path": "one frame queue per sub-batch (bam3d_frames_submit" + ("_gather" if gather else "") + " / bam3d_frames_wait), 2 frames in flight on two compute "
                                "lanes: every step copies its transforms (48 B each: the Mat4 without its constant fourth row) + pairs in from pinned host memory and its status + " +
                                ("the gathered hits of all ranks" if gather else "contacts") + " out to pinned host memory"},
                "gpu_launches": launches + launches_i + launches_c, "clocks": env.clocks(),
                "roofline": roofline_for("c5", ALGO_BYTES["c5"] * n, kernel_ms, f"{ALGO_BYTES['c5']:.1f} B/pair (the mix of the four rows of DESIGN.md §5) x {n} pairs per step"),
                "cpu_baseline": cpu_time("c5", parts, min(n, cpu_pairs), cpu_budget)}
    for s_ in list(shp.values()) + [shp_end]:
        s_.close()
    meshes.close()
    del d_pairs, d_status, d_contacts
    torch.cuda.empty_cache()
    return line


def bench_c4(env, n, steps, warmup, cpu_boxes):
    """C4: device sweep-and-prune over 2^22 boxes per GPU (candidate pairs/s); BVH build + queries and the pipeline as extras."""
    torch, b3 = env.torch, env.b3
    from bam3d_b200 import broad, scenes
    rank, world = env.rank, env.world
    # N > 1: every rank holds the SAME box set and sorts all of it (replicated sort); the pair search is
    # sharded by the pairs' higher sorted slot, so the ranks' pair slices are disjoint and concatenate to the full list
    boxes, L = scenes.aabb_scene(n, start=0, extent=100.0 * (n / float(1 << 22)) ** (1.0 / 3.0))
    ctx = env.ctxs[0]
    lib, h = ctx._lib, ctx.handle
    d_boxes = torch.from_numpy(boxes).to(env.dev)
    cap = 6 * n
    d_pairs = torch.empty(cap * 2, dtype=torch.int32, device=env.dev)
    d_perm = torch.empty(n, dtype=torch.int32, device=env.dev)
    cnt = C.c_uint64(0)

    def step_resident(k=0):
        axis = C.c_int32(0)
        ctx.check(lib.bam3d_sap_find_pairs_shard_dev(h, dptr(d_boxes), n, C.byref(axis), dptr(d_perm), dptr(d_pairs), cap, C.byref(cnt), rank, world))
    step_resident.finish = lambda: None
    ms, launches = env.timed(step_resident, steps, warmup)
    npairs = int(cnt.value)
    npairs_total = int(env.sum_over_ranks(npairs))  # pairs of the whole box set: the ranks' shards are disjoint
    value = npairs_total * steps / (ms * 1e-3)
    kernel_ms, _ = env.serial_kernel_ms(step_resident, reps=3)

    # e2e: host boxes in, perm + pairs out through the host-buffer call
    h_boxes = torch.from_numpy(boxes).pin_memory()
    h_pairs = torch.empty((cap, 2), dtype=torch.int32).pin_memory()
    h_perm = torch.empty(n, dtype=torch.int32).pin_memory()
    hp = lambda t: C.c_void_p(t.data_ptr())

    def step_e2e():
        axis = C.c_int32(0)
        ctx.check(lib.bam3d_sap_find_pairs_shard(h, hp(h_boxes), n, C.byref(axis), hp(h_perm), hp(h_pairs), cap, C.byref(cnt), rank, world))
    e2e_steps = max(3, steps)  # the same K steps, pipeline fill and drain inside the timed region
    step_e2e(); env.sync_all()
    w0 = time.time(); t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    e2e_s = env.max_over_ranks(time.perf_counter() - t0)
    env.windows.append((w0, time.time()))
    e2e_value = npairs_total * e2e_steps / e2e_s

    extra = {}
    if rank == 0:
        extra = c4_queries(env, boxes, L)
    pipe = c4_pipeline(env)
    line = None
    if rank == 0:
        if pipe:
            extra["pipeline"] = pipe
        line = {"metric": METRIC["c4"], "value": value, "unit": "pairs/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms / steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD_DESC["c4"], "boxes_per_gpu": n, "pairs_total": npairs_total, "pairs_rank0": npairs,
                           "l2": "inputs larger than L2 (96 MB boxes + 190 MB sort/sorted-box scratch + pair keys); no flush",
                           "multi_gpu": "replicated sort, pair search sharded by the pairs' higher sorted slot (no collective)" if world > 1 else "single GPU"},
                "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": n * 24, "d2h_bytes_per_step": n * 4 + npairs * 8,
                        "steps": e2e_steps, "path": "host-buffer bam3d_sap_find_pairs, pinned host memory"},
                "gpu_launches": launches, "clocks": env.clocks(),
                "roofline": roofline_for("c4", ALGO_BYTES["c4"] * n + 8.0 * npairs, kernel_ms, f"{ALGO_BYTES['c4']} B/box x {n} boxes + 8 B x {npairs} pairs per call"),
                "cpu_baseline": cpu_c4(min(n, cpu_boxes)), "extra": extra}
        line["roofline"]["kernel"] = "bvh_overlap_append_kernel (auto strategy at this density) + CUB radix sorts + LBVH build; sap_variance_chain_kernel on the side stream"
        line["roofline"]["note"] = ("a 1-axis sweep in a 3-D uniform scene would do n x ~24k candidate tests, so the pair search runs as a BVH overlap query over the "
                                    "sorted boxes: traversal-latency bound, not HBM bound")
    del d_boxes, d_pairs, d_perm
    torch.cuda.empty_cache()
    return line


def c4_queries(env, boxes, L):
    """DBVT query path at the stated C4 sizes: BVH build over 2^22 values, 2^20 ray queries, 2^20 closest-ray queries, 1024 frusta."""
    from bam3d_b200 import broad, scenes
    ctx = env.ctxs[0]
    out = {}
    ctx.synchronize()
    t0 = time.perf_counter(); bvh = broad.Bvh(ctx, boxes); out["bvh_build_ms_host_call"] = 1e3 * (time.perf_counter() - t0)
    rays = scenes.rays(1 << 20, L)
    planes = np.stack([broad.frustum_from_mat4(m).reshape(24) for m in scenes.view_projections(1024, L)])
    bench_q = getattr(broad, "bench_queries", None)
    if bench_q is not None:
        out.update(bench_q(env, bvh, rays, planes))   # device-resident one-pass queries, CUDA-event timed
    else:
        sub = rays[: 1 << 16]
        bvh.query_ray(sub)
        t0 = time.perf_counter(); off, vals, pts = bvh.query_ray(sub); dt = time.perf_counter() - t0
        out["ray_queries_per_s"] = len(sub) / dt; out["ray_hits"] = int(len(vals)); out["rays"] = len(sub)
        t0 = time.perf_counter(); bvh.query_ray_closest(rays); dt = time.perf_counter() - t0
        out["ray_closest_queries_per_s"] = len(rays) / dt
        subp = planes[:64]
        bvh.query_frustum(subp)
        t0 = time.perf_counter(); off, vals, rel = bvh.query_frustum(subp); dt = time.perf_counter() - t0
        out["frustum_queries_per_s"] = len(subp) / dt; out["frustum_hits"] = int(len(vals)); out["frusta"] = len(subp)
    return out


def c4_pipeline(env):
    """g1: world AABBs -> sweep and prune -> perm remap -> narrow phase, resident (bam3d_pipeline_*), when the library has it."""
    from bam3d_b200 import broad
    fn = getattr(broad, "bench_pipeline", None)
    return fn(env) if fn is not None else None


def run_ours(args):
    env = Env()
    wl = args.workload
    lines = {}
    try:
        if wl == "c5":
            main = bench_c5(env, args.pairs, args.steps, args.warmup, args.cpu_pairs, 20.0)
            if not args.no_extra:
                sub_steps = max(3, min(args.steps, 10))
                extra = {}
                for sub in ("c1", "c2", "c3", "toi", "dist"):
                    extra[sub] = bench_narrow(env, sub, SIZES[sub], sub_steps, args.warmup, 1 << 19, 4.0)
                extra["c4"] = bench_c4(env, args.boxes, max(3, min(args.steps, 5)), args.warmup, args.cpu_boxes)
                if main is not None:
                    main["extra"] = extra
                    main["rates"] = {  # the three metrics BASELINE.json names, at this N, each on its stated config
                        "gjk_intersect_pairs_per_sec": {"c1": extra["c1"]["value"], "c5_all_pairs": main["intersect_all"]["value"]},
                        "gjk_epa_contact_pairs_per_sec": {"c2": extra["c2"]["value"], "c3": extra["c3"]["value"], "c5_all_pairs": main["contact_all"]["value"]},
                        "sap_candidate_pairs_per_sec": {"c4": extra["c4"]["value"]}}
        elif wl == "c4":
            main = bench_c4(env, args.boxes, args.steps, args.warmup, args.cpu_boxes)
        else:
            main = bench_narrow(env, wl, args.pairs, args.steps, args.warmup, args.cpu_pairs, 20.0)
        if env.rank == 0:
            print(json.dumps(main), flush=True)
    finally:
        if env.sampler:
            env.sampler.stop()
    if env.world > 1:
        env.dist.barrier()
        if env.comm is not None:
            env.comm.close()
        env.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c5", choices=sorted(ALGO_BYTES))
    ap.add_argument("--pairs", type=int, default=None, help="pairs per GPU per step (default: the workload's stated size; c5: 2^23)")
    ap.add_argument("--cpu-pairs", type=int, default=1 << 20, help="pairs in the CPU sample")
    ap.add_argument("--boxes", type=int, default=1 << 22, help="c4: boxes per GPU")
    ap.add_argument("--cpu-boxes", type=int, default=1 << 17, help="c4: boxes in the CPU sample (the CPU sweep is O(n^(5/3)))")
    ap.add_argument("--no-extra", action="store_true", help="c5: skip the sub-lines of the other configs")
    args = ap.parse_args()
    if args.pairs is None:
        args.pairs = SIZES.get(args.workload, 1 << 20)
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
