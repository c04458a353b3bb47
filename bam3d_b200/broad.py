"""Host-side mirror of bam3d's broad-phase API over the CUDA C ABI.

  SweepAndPrune3            src/algorithm/broad_phase/sweep_prune.rs:27-128 (new, with_sweep_axis, get_sweep_axis,
                            find_collider_pairs)
  Aabb3, Ray, Frustum       src/volume/aabb/aabb3.rs, src/ray.rs, src/frustum.rs (plain value holders here)
  DynamicBoundingVolumeTree src/dbvt/mod.rs:239-515 — the query side (query / query_for_indices with DiscreteVisitor,
                            ContinuousVisitor, FrustumVisitor; query_ray, query_ray_closest of dbvt/util.rs) on a static
                            device BVH; insert() only collects values, do_refit()/tick() (re)build the device tree.
  DbvtBroadPhase            src/algorithm/broad_phase/dbvt.rs:14-63
Aabb3 values are rows of 6 f32 (min.xyz, max.xyz).
"""
import ctypes as C

import numpy as np

from . import _ffi
from .api import _ptr, default_context


def _retry_capacity(call, guess):
    """Two-call pattern of the C ABI: on BAM3D_ERR_CAPACITY the required size is reported; allocate and call again."""
    cap = max(int(guess), 1)
    while True:
        rc, need, result = call(cap)
        if rc == _ffi.OK:
            return result
        if rc == _ffi.ERR_CAPACITY and need > cap:
            cap = int(need)
            continue
        return rc


class SweepAndPrune3:
    """Sweep and prune broad phase (sweep_prune.rs:27-30); keeps `sweep_axis` across calls like the reference."""

    def __init__(self, ctx=None, sweep_axis=0):
        self.ctx = ctx if ctx is not None else default_context()
        self.sweep_axis = int(sweep_axis)

    @classmethod
    def new(cls, ctx=None):  # sweep_prune.rs:37-39
        return cls(ctx, 0)

    @classmethod
    def with_sweep_axis(cls, sweep_axis, ctx=None):  # sweep_prune.rs:42-47
        return cls(ctx, sweep_axis)

    def get_sweep_axis(self):  # sweep_prune.rs:50-52
        return self.sweep_axis

    def find_collider_pairs(self, aabbs, capacity=None, shard=0, n_shards=1):
        """-> (pairs[m, 2], perm[n]). With n_shards > 1 only the pairs whose higher sorted index falls in this shard's
        slice are returned (multi-GPU: the sort is replicated, the sweep is sharded; concatenating the shards' pairs
        in shard order gives the unsharded list). The reference re-sorts `shapes` in place and returns indices into the sorted
        slice (sweep_prune.rs:65-69): pairs index the SORTED order; perm[k] is the original index of sorted slot k, so
        `shapes[:] = shapes[perm]` reproduces the reference's side effect."""
        aabbs = np.ascontiguousarray(aabbs, dtype=np.float32).reshape(-1, 6)
        n = len(aabbs)
        perm = np.empty(n, dtype=np.uint32)
        ctx, lib = self.ctx, self.ctx._lib

        def call(cap):
            pairs = np.empty((cap, 2), dtype=np.uint32)
            axis = C.c_int32(self.sweep_axis)
            cnt = C.c_uint64(0)
            rc = lib.bam3d_sap_find_pairs_shard(ctx.handle, _ptr(aabbs), n, C.byref(axis), _ptr(perm), _ptr(pairs), cap, C.byref(cnt),
                                                int(shard), int(n_shards))
            return rc, cnt.value, (pairs, int(cnt.value), int(axis.value))

        res = _retry_capacity(call, capacity if capacity is not None else 8 * n + 1024)
        if isinstance(res, int):
            ctx.check(res)
        pairs, cnt, axis = res
        self.sweep_axis = axis
        return pairs[:cnt], perm


class SweepAndPrune3Sphere(SweepAndPrune3):
    """SweepAndPrune3<volume::Sphere> (sweep_prune.rs:27-128 with Bound for Sphere, volume/sphere.rs:17-48): bounds are rows
    of 4 f32 (centre.xyz, radius); a pair is reported iff it is still active on the sweep axis and Sphere::intersects holds."""

    def find_collider_pairs(self, spheres, capacity=None):
        spheres = np.ascontiguousarray(spheres, dtype=np.float32).reshape(-1, 4)
        n = len(spheres)
        perm = np.empty(n, dtype=np.uint32)
        ctx, lib = self.ctx, self.ctx._lib

        def call(cap):
            pairs = np.empty((cap, 2), dtype=np.uint32)
            axis = C.c_int32(self.sweep_axis)
            cnt = C.c_uint64(0)
            rc = lib.bam3d_sap_find_pairs_spheres(ctx.handle, _ptr(spheres), n, C.byref(axis), _ptr(perm), _ptr(pairs), cap, C.byref(cnt))
            return rc, cnt.value, (pairs, int(cnt.value), int(axis.value))

        res = _retry_capacity(call, capacity if capacity is not None else 8 * n + 1024)
        if isinstance(res, int):
            ctx.check(res)
        pairs, cnt, axis = res
        self.sweep_axis = axis
        return pairs[:cnt], perm


def world_spheres(shapes):
    """ComputeBound<volume::Sphere> + Bound::transform_volume for every shape of a ShapeSet -> [n, 4] (centre.xyz, radius)."""
    out = np.empty((shapes.n, 4), dtype=np.float32)
    shapes.ctx.check(shapes.ctx._lib.bam3d_shapes_world_spheres(shapes.ctx.handle, shapes.handle, _ptr(out)))
    return out


def variance3(ctx, aabbs):
    """Variance3 (sweep_prune.rs:166-216) over `aabbs` in the given order -> (sums[6] = csum.xyz, csumsq.xyz, next axis)."""
    aabbs = np.ascontiguousarray(aabbs, dtype=np.float32).reshape(-1, 6)
    sums = np.zeros(6, dtype=np.float32)
    axis = C.c_int32(0)
    ctx.check(ctx._lib.bam3d_sap_variance(ctx.handle, _ptr(aabbs), len(aabbs), _ptr(sums), C.byref(axis)))
    return sums, int(axis.value)


def frustum_from_mat4(mat4_cols):
    """Frustum::from_mat4 (frustum.rs:46-75) -> 6 planes x (n.xyz, d): left, right, bottom, top, near, far; None if a
    plane degenerates."""
    m = np.ascontiguousarray(mat4_cols, dtype=np.float32).reshape(16)
    out = np.empty(24, dtype=np.float32)
    rc = _ffi.lib().bam3d_frustum_from_mat4(_ptr(m), _ptr(out))
    return out.reshape(6, 4) if rc == _ffi.OK else None


class Bvh:
    """Device BVH over Aabb3 values (bam3d_bvh)."""

    def __init__(self, ctx, aabbs):
        self.ctx = ctx
        self.aabbs = np.ascontiguousarray(aabbs, dtype=np.float32).reshape(-1, 6)
        h = C.c_void_p()
        ctx.check(ctx._lib.bam3d_bvh_build(ctx.handle, _ptr(self.aabbs), len(self.aabbs), C.byref(h)))
        self.handle = h
        self.n = len(self.aabbs)

    def refit(self, aabbs):
        """New bounds for the same values, topology unchanged (update_node + do_refit for every value, dbvt/mod.rs:561-589)."""
        aabbs = np.ascontiguousarray(aabbs, dtype=np.float32).reshape(-1, 6)
        if len(aabbs) != self.n:
            raise ValueError("refit needs one box per value of the tree")
        self.ctx.check(self.ctx._lib.bam3d_bvh_refit(self.ctx.handle, self.handle, _ptr(aabbs)))
        self.aabbs = aabbs

    def close(self):
        if self.handle:
            self.ctx._lib.bam3d_bvh_free(self.ctx.handle, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _var_query(self, fn, queries, qwidth, payload):
        ctx = self.ctx
        q = np.ascontiguousarray(queries, dtype=np.float32).reshape(-1, qwidth)
        nq = len(q)
        offsets = np.zeros(nq + 1, dtype=np.uint64)

        def call(cap):
            values = np.empty(cap, dtype=np.uint32)
            total = C.c_uint64(0)
            if payload is None:
                rc = fn(ctx.handle, self.handle, _ptr(q), nq, _ptr(offsets), _ptr(values), cap, C.byref(total))
                pay = None
            else:
                pay = np.empty((cap,) + payload[0], dtype=payload[1])
                rc = fn(ctx.handle, self.handle, _ptr(q), nq, _ptr(offsets), _ptr(values), _ptr(pay), cap, C.byref(total))
            return rc, total.value, (values, pay, int(total.value))

        res = _retry_capacity(call, 16 * nq + 1024)
        if isinstance(res, int):
            ctx.check(res)
        values, pay, total = res
        return offsets, values[:total], (None if pay is None else pay[:total])

    def query_aabb(self, queries):
        """DiscreteVisitor per query box -> (offsets[nq+1], values)."""
        off, vals, _ = self._var_query(self.ctx._lib.bam3d_bvh_query_aabb, queries, 6, None)
        return off, vals

    def query_ray(self, rays):
        """query_ray per ray (origin.xyz, direction.xyz) -> (offsets, values, points[total, 3])."""
        return self._var_query(self.ctx._lib.bam3d_bvh_query_ray, rays, 6, ((3,), np.float32))

    def query_frustum(self, planes):
        """FrustumVisitor per frustum (6 planes x 4) -> (offsets, values, relation[total])."""
        return self._var_query(self.ctx._lib.bam3d_bvh_query_frustum, planes, 24, ((), np.uint8))

    def query_ray_closest(self, rays):
        """query_ray_closest per ray -> (value[nq] (0xFFFFFFFF = None), point_t[nq, 4])."""
        r = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        val = np.empty(len(r), dtype=np.uint32)
        pt = np.empty((len(r), 4), dtype=np.float32)
        self.ctx.check(self.ctx._lib.bam3d_bvh_query_ray_closest(self.ctx.handle, self.handle, _ptr(r), len(r), _ptr(val), _ptr(pt)))
        return val, pt

    # ---- device-resident one-pass forms (raw device pointers; asynchronous on ctx.stream; hits in arbitrary order) ----
    def query_dev(self, kind, d_queries, nq, d_out_query, d_out_values, d_out_payload, capacity, d_count):
        """kind: "aabb" | "ray" | "frustum". Appends (query, value[, point | relation]) per hit; *d_count = hits found."""
        lib, h = self.ctx._lib, self.ctx.handle
        if kind == "aabb":
            rc = lib.bam3d_bvh_query_aabb_dev(h, self.handle, d_queries, nq, d_out_query, d_out_values, capacity, d_count)
        elif kind == "sphere":
            rc = lib.bam3d_bvh_query_sphere_dev(h, self.handle, d_queries, nq, d_out_query, d_out_values, capacity, d_count)
        elif kind == "ray":
            rc = lib.bam3d_bvh_query_ray_dev(h, self.handle, d_queries, nq, d_out_query, d_out_values, d_out_payload, capacity, d_count)
        else:
            rc = lib.bam3d_bvh_query_frustum_dev(h, self.handle, d_queries, nq, d_out_query, d_out_values, d_out_payload, capacity, d_count)
        self.ctx.check(rc)

    def query_sphere(self, spheres):
        """DiscreteVisitor per query sphere (centre.xyz, radius) on a tree over sphere bounds -> (offsets[nq+1], values)."""
        off, vals, _ = self._var_query(self.ctx._lib.bam3d_bvh_query_sphere, spheres, 4, None)
        return off, vals

    def query_ray_closest_dev(self, d_rays, nq, d_out_value, d_out_point_t):
        self.ctx.check(self.ctx._lib.bam3d_bvh_query_ray_closest_dev(self.ctx.handle, self.handle, d_rays, nq, d_out_value, d_out_point_t))

    def find_collider_pairs(self, dirty):
        """DbvtBroadPhase::find_collider_pairs -> sorted unique (low, high) pairs."""
        dirty = np.ascontiguousarray(dirty, dtype=np.uint8)
        assert len(dirty) == self.n
        ctx = self.ctx

        def call(cap):
            pairs = np.empty((cap, 2), dtype=np.uint32)
            cnt = C.c_uint64(0)
            rc = ctx._lib.bam3d_bvh_find_collider_pairs(ctx.handle, self.handle, _ptr(dirty), _ptr(pairs), cap, C.byref(cnt))
            return rc, cnt.value, (pairs, int(cnt.value))

        res = _retry_capacity(call, 8 * self.n + 1024)
        if isinstance(res, int):
            ctx.check(res)
        pairs, cnt = res
        return pairs[:cnt]


class SphereBvh(Bvh):
    """The same tree over values whose bound is volume::Sphere (bam3d_bvh_build_spheres): spheres[n, 4] = (centre.xyz, radius).
    query_sphere / query_ray / query_ray_closest / query_frustum / find_collider_pairs apply the sphere's own predicates
    (volume/sphere.rs); query_aabb and refit are not available on it (the library reports BAM3D_ERR_ARG)."""

    def __init__(self, ctx, spheres):
        self.ctx = ctx
        self.spheres = np.ascontiguousarray(spheres, dtype=np.float32).reshape(-1, 4)
        h = C.c_void_p()
        ctx.check(ctx._lib.bam3d_bvh_build_spheres(ctx.handle, _ptr(self.spheres), len(self.spheres), C.byref(h)))
        self.handle = h
        self.n = len(self.spheres)


class DynamicBoundingVolumeTree:
    """The query face of dbvt/mod.rs:239-515. insert() appends a value's (bound, fat bound); do_refit() / tick() build
    the device tree; value indices are insertion order, as in the reference's `values` Vec."""

    def __init__(self, ctx=None):
        self.ctx = ctx if ctx is not None else default_context()
        self._bounds = []
        self._bvh = None

    def insert(self, bound, fat_bound=None):  # dbvt/mod.rs:622 — returns the value index
        self._bounds.append(np.asarray(bound, dtype=np.float32).reshape(6))
        self._bvh = None
        return len(self._bounds) - 1

    def do_refit(self):  # dbvt/mod.rs:837-842
        if self._bvh is not None:
            self._bvh.close()
        self._bvh = Bvh(self.ctx, np.array(self._bounds, dtype=np.float32).reshape(-1, 6))

    tick = do_refit  # dbvt/mod.rs:595-599

    def _tree(self):
        if self._bvh is None:
            self.do_refit()
        return self._bvh

    def query_aabb(self, bound):  # query(&mut DiscreteVisitor::new(&bound))
        return self._tree().query_aabb([bound])[1]

    def query_ray(self, origin, direction):  # dbvt/util.rs:114-129
        _, v, p = self._tree().query_ray([list(origin) + list(direction)])
        return v, p

    def query_ray_closest(self, origin, direction):  # dbvt/util.rs:77-103
        v, pt = self._tree().query_ray_closest([list(origin) + list(direction)])
        return None if v[0] == 0xFFFFFFFF else (int(v[0]), pt[0, :3].copy())

    def query_frustum(self, planes):  # query(&mut FrustumVisitor::new(&frustum))
        _, v, r = self._tree().query_frustum([planes])
        return v, r


class DbvtBroadPhase:  # algorithm/broad_phase/dbvt.rs:14-63
    def find_collider_pairs(self, tree, dirty):
        return tree._tree().find_collider_pairs(dirty)


def world_aabbs(shapes):
    """ComputeBound<Aabb3> + Aabb::transform for every shape of a ShapeSet -> n x 6 f32."""
    out = np.empty((shapes.n, 6), dtype=np.float32)
    shapes.ctx.check(shapes.ctx._lib.bam3d_shapes_world_aabbs(shapes.ctx.handle, shapes.handle, _ptr(out)))
    return out


def pipeline_contacts(gjk, shapes, strategy, sweep_axis=0, capacity=None):
    """bam3d_pipeline_contacts: world AABBs -> SweepAndPrune3 -> perm remap -> GJK::intersection per candidate pair, resident.
    -> dict(pairs[m, 2] shape indices, sorted_pairs[m, 2] as the reference returns them, perm[n], contacts[m, 8], status[m],
    next_axis)."""
    ctx = shapes.ctx
    n = shapes.n

    def call(cap):
        pairs = np.empty((cap, 2), dtype=np.uint32)
        spairs = np.empty((cap, 2), dtype=np.uint32)
        perm = np.empty(max(n, 1), dtype=np.uint32)
        contacts = np.empty((cap, 8), dtype=np.float32)
        status = np.empty(cap, dtype=np.uint8)
        axis = C.c_int32(int(sweep_axis))
        cnt = C.c_uint64(0)
        rc = ctx._lib.bam3d_pipeline_contacts(ctx.handle, shapes.handle, C.byref(axis), C.byref(gjk.settings), int(strategy), _ptr(pairs), _ptr(spairs),
                                              _ptr(perm), _ptr(contacts), _ptr(status), cap, C.byref(cnt))
        m = int(cnt.value)
        return rc, m, dict(pairs=pairs[:m], sorted_pairs=spairs[:m], perm=perm[:n], contacts=contacts[:m], status=status[:m], next_axis=int(axis.value))

    res = _retry_capacity(call, capacity if capacity is not None else 8 * n + 1024)
    if isinstance(res, int):
        ctx.check(res)
    return res


# ---------------------------------------------------------------------------------------------------------
# bench.py hooks (C4 extras): device-resident queries at the stated sizes, CUDA-event timed
# ---------------------------------------------------------------------------------------------------------
def bench_queries(env, bvh, rays, planes):
    """2^20 ray queries, 2^20 closest-ray queries and 1024 frustum queries against `bvh`, device-resident, one pass each.
    Every figure is events on the ctx stream around `reps` back-to-back calls (after one warm-up call)."""
    torch = env.torch
    ctx = bvh.ctx
    ext = env.ext[0]
    dev = env.dev
    p = lambda t: C.c_void_p(t.data_ptr())
    out = {}

    def timed(fn, reps=3):
        fn()
        ctx.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(ext)
        for _ in range(reps):
            fn()
        e1.record(ext)
        ctx.synchronize(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    d_count = torch.zeros(1, dtype=torch.int64, device=dev)
    # rays: first a small batch to size the output (hits per ray), then the full batch
    d_rays = torch.from_numpy(rays).to(dev)
    nq = len(rays)
    probe = min(nq, 4096)
    bvh.query_dev("ray", p(d_rays), probe, None, None, None, 0, p(d_count))
    ctx.synchronize()
    cap = int(int(d_count.item()) / probe * nq * 1.15) + 65536
    d_q = torch.empty(cap, dtype=torch.int32, device=dev)
    d_v = torch.empty(cap, dtype=torch.int32, device=dev)
    d_pt = torch.empty(cap * 3, dtype=torch.float32, device=dev)
    ms = timed(lambda: bvh.query_dev("ray", p(d_rays), nq, p(d_q), p(d_v), p(d_pt), cap, p(d_count)))
    hits = int(d_count.item())
    out.update({"rays": nq, "ray_hits": hits, "ray_query_ms": ms, "ray_queries_per_s": nq / (ms * 1e-3), "ray_hits_per_s": hits / (ms * 1e-3),
                "ray_capacity_ok": hits <= cap})
    d_cv = torch.empty(nq, dtype=torch.int32, device=dev)
    d_cp = torch.empty(nq * 4, dtype=torch.float32, device=dev)
    ms = timed(lambda: bvh.query_ray_closest_dev(p(d_rays), nq, p(d_cv), p(d_cp)))
    out.update({"ray_closest_ms": ms, "ray_closest_queries_per_s": nq / (ms * 1e-3)})
    del d_q, d_v, d_pt
    # frusta
    d_pl = torch.from_numpy(np.ascontiguousarray(planes, dtype=np.float32)).to(dev)
    nf = len(planes)
    probe = min(nf, 16)
    bvh.query_dev("frustum", p(d_pl), probe, None, None, None, 0, p(d_count))
    ctx.synchronize()
    cap = int(int(d_count.item()) / probe * nf * 1.25) + 65536
    d_q = torch.empty(cap, dtype=torch.int32, device=dev)
    d_v = torch.empty(cap, dtype=torch.int32, device=dev)
    d_rel = torch.empty(cap, dtype=torch.uint8, device=dev)
    ms = timed(lambda: bvh.query_dev("frustum", p(d_pl), nf, p(d_q), p(d_v), p(d_rel), cap, p(d_count)), reps=2)
    hits = int(d_count.item())
    out.update({"frusta": nf, "frustum_hits": hits, "frustum_query_ms": ms, "frustum_queries_per_s": nf / (ms * 1e-3),
                "frustum_hits_per_s": hits / (ms * 1e-3), "frustum_capacity_ok": hits <= cap,
                "path": "bam3d_bvh_query_{ray,ray_closest,frustum}_dev: device-resident, one pass (append), CUDA events on the ctx stream; "
                        "rays: stackless traversal, one thread per ray; frusta: every (frustum, value) leaf test (few large queries)"})
    return out


def bench_pipeline(env, n=1 << 20, steps=5):
    """g1 (BASELINE configs[3] "... then narrow phase"): 2^20 shapes of the C2 mix in one volume; one step = world AABBs ->
    sweep and prune -> perm remap -> GJK+EPA on every candidate pair, through bam3d_pipeline_contacts_dev (N > 1: the sort is
    replicated, every rank searches and resolves its own shard of the pair list). CUDA events on the ctx stream."""
    from . import scenes
    from .api import ShapeSet, CollisionStrategy
    torch = env.torch
    ctx, gjk = env.ctxs[0], env.gjks[0]
    scene, L = scenes.world_scene(n)
    shapes = ShapeSet(ctx, scene["kind"], scene["params"], scene["xform"])
    cap = 6 * n
    dev = env.dev
    d_pairs = torch.empty(cap * 2, dtype=torch.int32, device=dev)
    d_contacts = torch.empty(cap * 8, dtype=torch.float32, device=dev)
    d_status = torch.empty(cap, dtype=torch.uint8, device=dev)
    p = lambda t: C.c_void_p(t.data_ptr())
    cnt = C.c_uint64(0)

    def step(k=0):
        axis = C.c_int32(0)
        ctx.check(ctx._lib.bam3d_pipeline_contacts_dev(ctx.handle, shapes.handle, C.byref(axis), C.byref(gjk.settings), int(CollisionStrategy.FullResolution),
                                                       env.rank, env.world, p(d_pairs), None, None, p(d_contacts), p(d_status), cap, C.byref(cnt)))
    step.finish = lambda: None
    ms, launches = env.timed(step, steps, 3)
    m = int(cnt.value)
    hits = int((d_status[:m] == 0).sum().item())
    m_total = int(env.sum_over_ranks(m))
    hits_total = int(env.sum_over_ranks(hits))
    shapes.close()
    if env.rank != 0:
        return None
    return {"shapes": n, "candidate_pairs": m_total, "contacts": hits_total, "ms_per_step": ms / steps,
            "shapes_per_s": n * steps / (ms * 1e-3), "candidate_pairs_per_s": m_total * steps / (ms * 1e-3),
            "path": "bam3d_pipeline_contacts_dev: bam3d_shapes_world_aabbs_dev -> sweep and prune (shard of this rank) -> perm remap -> bam3d_gjk_contact_dev, "
                    "resident; one host synchronisation (the pair count)", "gpu_launches": launches}
