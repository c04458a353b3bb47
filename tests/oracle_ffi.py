"""ctypes wrapper over the CPU oracle (oracle/oracle_capi.cpp). TEST INFRASTRUCTURE: imported only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs — never from bam3d_b200/."""
import ctypes as C
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(ROOT, "oracle", "_build", "libbam3d_oracle.so")
KAT_PATH = os.path.join(ROOT, "oracle", "_build", "kat")


class Settings(C.Structure):
    _fields_ = [("distance_tolerance", C.c_float), ("continuous_tolerance", C.c_float), ("epa_tolerance", C.c_float),
                ("max_iterations", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [("support_calls", C.c_uint64), ("gjk_iterations", C.c_uint64), ("epa_iterations", C.c_uint64),
                ("max_faces", C.c_uint32), ("max_verts", C.c_uint32), ("max_edges", C.c_uint32), ("pad", C.c_uint32)]


def _p(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(LIB_PATH)
        self.lib.orc_sap_find_pairs.restype = C.c_uint64
        self.lib.orc_sap_find_pairs_spheres.restype = C.c_uint64
        self.lib.orc_dbvt_build.restype = C.c_void_p
        for f in ("orc_dbvt_query_aabb", "orc_dbvt_query_ray", "orc_dbvt_query_frustum", "orc_dbvt_find_collider_pairs"):
            getattr(self.lib, f).restype = C.c_uint64

    # ---- narrow phase ------------------------------------------------------------------------------------
    @staticmethod
    def _scene_args(scene):
        kind = np.ascontiguousarray(scene["kind"], dtype=np.uint8)
        params = np.ascontiguousarray(scene["params"], dtype=np.float32)
        mesh_id = scene.get("mesh_id")
        mv, mo = scene.get("mesh_verts"), scene.get("mesh_offsets")
        keep = [kind, params,
                None if mesh_id is None else np.ascontiguousarray(mesh_id, dtype=np.uint32),
                None if mv is None else np.ascontiguousarray(mv, dtype=np.float32),
                None if mo is None else np.ascontiguousarray(mo, dtype=np.uint32)]
        return keep

    @staticmethod
    def _settings(settings):
        if settings is None:
            return None
        s = Settings(settings.distance_tolerance, settings.continuous_tolerance, settings.epa_tolerance, settings.max_iterations)
        return C.byref(s)

    def intersect(self, scene, settings=None, nthreads=1, counters=None):
        kind, params, mesh_id, mv, mo = self._scene_args(scene)
        xform = np.ascontiguousarray(scene["xform"], dtype=np.float32)
        pairs = np.ascontiguousarray(scene["pairs"], dtype=np.uint32)
        n = len(pairs)
        hit = np.empty(n, dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        self.lib.orc_gjk_intersect_batch(_p(kind), _p(params), _p(xform), _p(mesh_id), _p(mv), _p(mo), _p(pairs), C.c_uint64(n),
                                         self._settings(settings), _p(hit), _p(status), C.c_int(nthreads),
                                         C.byref(counters) if counters is not None else None)
        return hit, status

    def contact(self, scene, strategy=0, settings=None, nthreads=1, counters=None):
        kind, params, mesh_id, mv, mo = self._scene_args(scene)
        xform = np.ascontiguousarray(scene["xform"], dtype=np.float32)
        pairs = np.ascontiguousarray(scene["pairs"], dtype=np.uint32)
        n = len(pairs)
        hit = np.empty(n, dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        contacts = np.empty((n, 8), dtype=np.float32)
        self.lib.orc_gjk_contact_batch(_p(kind), _p(params), _p(xform), _p(mesh_id), _p(mv), _p(mo), _p(pairs), C.c_uint64(n),
                                       self._settings(settings), C.c_int(int(strategy)), _p(contacts), _p(hit), _p(status),
                                       C.c_int(nthreads), C.byref(counters) if counters is not None else None)
        return contacts, hit, status

    def contact_face_limit(self, scene, face_limit, in_phases=False, nthreads=1):
        """GJK::intersection (FullResolution) with the oracle's EPA face limit set to `face_limit` (the reference has none) and,
        optionally, Polytope::add evaluated in the device's data-parallel formulation (add_in_phases).
        -> (contacts[n, 8], status[n], max_faces[n], epa_iterations[n], removed_faces[n])"""
        kind, params, mesh_id, mv, mo = self._scene_args(scene)
        xform = np.ascontiguousarray(scene["xform"], dtype=np.float32)
        pairs = np.ascontiguousarray(scene["pairs"], dtype=np.uint32)
        n = len(pairs)
        ct, st = np.zeros((n, 8), dtype=np.float32), np.zeros(n, dtype=np.uint8)
        mf, it, rem = np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint32), np.zeros(n, dtype=np.uint64)
        self.lib.orc_gjk_contact_face_limit(_p(kind), _p(params), _p(xform), _p(mesh_id), _p(mv), _p(mo), _p(pairs), C.c_uint64(n), C.c_uint64(int(face_limit)),
                                            _p(ct), _p(st), _p(mf), _p(it), _p(rem), C.c_int(nthreads), C.c_int(1 if in_phases else 0))
        return ct, st, mf, it, rem

    def sphere_values_query(self, spheres, kind, query):
        """The leaf predicate of a DBVT over volume::Sphere bounds applied to every value: kind 0 sphere (4 floats), 1 ray (6 floats,
        -> points), 2 frustum (24 floats, -> relations). -> (values ascending, points[k, 3] | None, relation[k] | None)"""
        spheres = np.ascontiguousarray(spheres, dtype=np.float32).reshape(-1, 4)
        query = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
        n = len(spheres)
        vals, pts, rel = np.empty(n, dtype=np.uint32), np.empty((n, 3), dtype=np.float32), np.empty(n, dtype=np.uint8)
        self.lib.orc_sphere_values_query.restype = C.c_uint64
        k = self.lib.orc_sphere_values_query(_p(spheres), C.c_uint64(n), C.c_int32(kind), _p(query), _p(vals), _p(pts), _p(rel), C.c_uint64(n))
        return vals[:k], (pts[:k] if kind == 1 else None), (rel[:k] if kind == 2 else None)

    def distance(self, scene, settings=None, nthreads=1, counters=None):
        kind, params, mesh_id, mv, mo = self._scene_args(scene)
        xform = np.ascontiguousarray(scene["xform"], dtype=np.float32)
        pairs = np.ascontiguousarray(scene["pairs"], dtype=np.uint32)
        n = len(pairs)
        some = np.empty(n, dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        ref = np.empty(n, dtype=np.float32)
        geo = np.empty(n, dtype=np.float32)
        self.lib.orc_gjk_distance_batch(_p(kind), _p(params), _p(xform), _p(mesh_id), _p(mv), _p(mo), _p(pairs), C.c_uint64(n),
                                        self._settings(settings), _p(ref), _p(geo), _p(some), _p(status), C.c_int(nthreads),
                                        C.byref(counters) if counters is not None else None)
        return ref, geo, some, status

    def toi(self, scene, xform_end, settings=None, toi_cap=100, nthreads=1, counters=None):
        kind, params, mesh_id, mv, mo = self._scene_args(scene)
        x0 = np.ascontiguousarray(scene["xform"], dtype=np.float32)
        x1 = np.ascontiguousarray(xform_end, dtype=np.float32)
        pairs = np.ascontiguousarray(scene["pairs"], dtype=np.uint32)
        n = len(pairs)
        hit = np.empty(n, dtype=np.uint8)
        status = np.empty(n, dtype=np.uint8)
        contacts = np.empty((n, 8), dtype=np.float32)
        self.lib.orc_gjk_toi_batch(_p(kind), _p(params), _p(x0), _p(x1), _p(mesh_id), _p(mv), _p(mo), _p(pairs), C.c_uint64(n),
                                   self._settings(settings), C.c_uint32(toi_cap), _p(contacts), _p(hit), _p(status), C.c_int(nthreads),
                                   C.byref(counters) if counters is not None else None)
        return contacts, hit, status

    def complex(self, scene, comp_offsets, comp_xform0, comp_xform1, pairs, mode, strategy=0):
        """Compound shapes: mode 0 intersection_complex, 1 distance_complex, 2 intersection_complex_time_of_impact."""
        kind, params, mesh_id, mv, mo = self._scene_args(scene)
        xform = np.ascontiguousarray(scene["xform"], dtype=np.float32)
        co = np.ascontiguousarray(comp_offsets, dtype=np.uint32)
        x0 = np.ascontiguousarray(comp_xform0, dtype=np.float32)
        x1 = np.ascontiguousarray(comp_xform1 if comp_xform1 is not None else comp_xform0, dtype=np.float32)
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32)
        out = np.empty((len(pairs), 8), dtype=np.float32)
        st = np.empty(len(pairs), dtype=np.uint8)
        self.lib.orc_gjk_complex_batch(_p(kind), _p(params), _p(xform), _p(mesh_id), _p(mv), _p(mo), _p(co), _p(x0), _p(x1), _p(pairs),
                                       C.c_uint64(len(pairs)), C.c_int(mode), C.c_int(int(strategy)), _p(out), _p(st))
        return out, st

    def ray_cast(self, scene, rays, casts, query, nthreads=1):
        """-> (points[n, 3], status[n]); query 0 = intersects_transformed, 1 = intersection_transformed, 2 = swept particle."""
        kind, params, mesh_id, mv, mo = self._scene_args(scene)
        xform = np.ascontiguousarray(scene["xform"], dtype=np.float32)
        rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
        casts = np.ascontiguousarray(casts, dtype=np.uint32).reshape(-1, 2)
        n = len(casts)
        pts = np.zeros((n, 3), dtype=np.float32)
        st = np.zeros(n, dtype=np.uint8)
        self.lib.orc_ray_cast_batch(_p(kind), _p(params), _p(xform), _p(mesh_id), _p(mv), _p(mo), _p(rays), _p(casts), C.c_uint64(n),
                                    C.c_int32(int(query)), _p(pts), _p(st), C.c_int(nthreads))
        return pts, st

    def set_mesh_faces(self, mesh_verts, mesh_offsets, faces, face_offsets):
        """ConvexPolyhedron::new_with_faces for the meshes that have faces (HalfEdge mode in every later batch call of this
        process); set_mesh_faces(None, ...) clears. Raises where the reference would panic (degenerate face)."""
        if mesh_verts is None:
            self.lib.orc_set_mesh_faces(None, None, C.c_uint32(0), None, None)
            return
        mv = np.ascontiguousarray(mesh_verts, dtype=np.float32)
        mo = np.ascontiguousarray(mesh_offsets, dtype=np.uint32)
        f = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)
        fo = np.ascontiguousarray(face_offsets, dtype=np.uint32)
        rc = self.lib.orc_set_mesh_faces(_p(mv), _p(mo), C.c_uint32(len(mo) - 1), _p(f), _p(fo))
        if rc:
            raise ValueError(f"mesh {rc - 1}: degenerate or out-of-range face")

    def support_point(self, kind, params4, direction, xform16, verts=None):
        params4 = np.ascontiguousarray(params4, dtype=np.float32)
        direction = np.ascontiguousarray(direction, dtype=np.float32)
        xform16 = np.ascontiguousarray(xform16, dtype=np.float32)
        verts = None if verts is None else np.ascontiguousarray(verts, dtype=np.float32)
        out = np.empty(3, dtype=np.float32)
        self.lib.orc_support_point(C.c_uint8(kind), _p(params4), _p(verts), C.c_uint32(0 if verts is None else len(verts)),
                                   _p(direction), _p(xform16), _p(out))
        return out

    def mat4_from_srt(self, scale, quat, trans):
        out = np.empty(16, dtype=np.float32)
        sc, q, t = (np.ascontiguousarray(a, dtype=np.float32) for a in (scale, quat, trans))  # keep alive across the call
        self.lib.orc_mat4_from_srt(_p(sc), _p(q), _p(t), _p(out))
        return out

    def quat_from_rotation_z(self, angle):
        out = np.empty(4, dtype=np.float32)
        self.lib.orc_quat_from_rotation_z(C.c_float(angle), _p(out))
        return out

    def transform_z(self, x, y, z, angle_deg):
        """The reference tests' transform()/transform_3d(): rotation about z by angle_deg.to_radians()."""
        rad = np.float32(angle_deg) * (np.float32(np.pi) / np.float32(180.0))
        return self.mat4_from_srt([1, 1, 1], self.quat_from_rotation_z(float(rad)), [x, y, z])

    def mat4_inverse(self, m):
        out = np.empty(16, dtype=np.float32)
        m = np.ascontiguousarray(m, dtype=np.float32)
        self.lib.orc_mat4_inverse(_p(m), _p(out))
        return out

    def world_aabbs(self, scene):
        kind, params, mesh_id, mv, mo = self._scene_args(scene)
        xform = np.ascontiguousarray(scene["xform"], dtype=np.float32)
        out = np.empty((len(kind), 6), dtype=np.float32)
        self.lib.orc_world_aabb_batch(_p(kind), _p(params), _p(xform), _p(mesh_id), _p(mv), _p(mo), C.c_uint64(len(kind)), _p(out))
        return out

    def world_spheres(self, scene):
        kind, params, mesh_id, mv, mo = self._scene_args(scene)
        xform = np.ascontiguousarray(scene["xform"], dtype=np.float32)
        out = np.empty((len(kind), 4), dtype=np.float32)
        self.lib.orc_world_sphere_batch(_p(kind), _p(params), _p(xform), _p(mesh_id), _p(mv), _p(mo), C.c_uint64(len(kind)), _p(out))
        return out

    # ---- broad phase -------------------------------------------------------------------------------------
    def sap_find_pairs_spheres(self, spheres, axis=0):
        spheres = np.ascontiguousarray(spheres, dtype=np.float32).reshape(-1, 4)
        n = len(spheres)
        perm = np.empty(n, dtype=np.uint32)
        cap = max(1024, 16 * n)
        while True:
            pairs = np.empty((cap, 2), dtype=np.uint32)
            ax = C.c_int32(axis)
            cnt = self.lib.orc_sap_find_pairs_spheres(_p(spheres), C.c_uint64(n), C.byref(ax), _p(perm), _p(pairs), C.c_uint64(cap))
            if cnt <= cap:
                return pairs[:cnt].copy(), perm, int(ax.value)
            cap = int(cnt)

    def sap_find_pairs(self, aabbs, axis=0):
        aabbs = np.ascontiguousarray(aabbs, dtype=np.float32).reshape(-1, 6)
        n = len(aabbs)
        ax = C.c_int32(axis)
        perm = np.empty(n, dtype=np.uint32)
        cap = max(1024, 16 * n)
        while True:
            pairs = np.empty((cap, 2), dtype=np.uint32)
            ax = C.c_int32(axis)
            cnt = self.lib.orc_sap_find_pairs(_p(aabbs), C.c_uint64(n), C.byref(ax), _p(perm), _p(pairs), C.c_uint64(cap))
            if cnt <= cap:
                return pairs[:cnt].copy(), perm, int(ax.value)
            cap = int(cnt)

    def variance3(self, aabbs):
        """Variance3 (sweep_prune.rs:166-216) over the boxes in the given order -> (sums[6] = csum.xyz, csumsq.xyz, axis)."""
        aabbs = np.ascontiguousarray(aabbs, dtype=np.float32).reshape(-1, 6)
        out = np.zeros(6, dtype=np.float32)
        axis = self.lib.orc_sap_variance_sums(_p(aabbs), C.c_uint64(len(aabbs)), _p(out))
        return out, int(axis)

    def dbvt_build(self, aabbs, fat=None):
        aabbs = np.ascontiguousarray(aabbs, dtype=np.float32).reshape(-1, 6)
        fat = None if fat is None else np.ascontiguousarray(fat, dtype=np.float32).reshape(-1, 6)
        return OracleDbvt(self, self.lib.orc_dbvt_build(_p(aabbs), _p(fat), C.c_uint64(len(aabbs))), len(aabbs))

    def frustum_from_mat4(self, m16):
        out = np.empty(24, dtype=np.float32)
        m16 = np.ascontiguousarray(m16, dtype=np.float32)
        ok = self.lib.orc_frustum_from_mat4(_p(m16), _p(out))
        return out if ok else None

    def perspective_rh_gl(self, fov, aspect, zn, zf):
        out = np.empty(16, dtype=np.float32)
        self.lib.orc_perspective_rh_gl(C.c_float(fov), C.c_float(aspect), C.c_float(zn), C.c_float(zf), _p(out))
        return out


class OracleDbvt:
    def __init__(self, orc, handle, n):
        self.orc, self.h, self.n = orc, C.c_void_p(handle), n

    def query_aabb(self, q6):
        q6 = np.ascontiguousarray(q6, dtype=np.float32)
        out = np.empty(self.n, dtype=np.uint32)
        c = self.orc.lib.orc_dbvt_query_aabb(self.h, _p(q6), _p(out), C.c_uint64(self.n))
        return out[:c].copy()

    def query_ray(self, ray6):
        ray6 = np.ascontiguousarray(ray6, dtype=np.float32)
        vals = np.empty(self.n, dtype=np.uint32)
        pts = np.empty((self.n, 3), dtype=np.float32)
        c = self.orc.lib.orc_dbvt_query_ray(self.h, _p(ray6), _p(vals), _p(pts), C.c_uint64(self.n))
        return vals[:c].copy(), pts[:c].copy()

    def query_ray_closest(self, ray6):
        ray6 = np.ascontiguousarray(ray6, dtype=np.float32)
        v = C.c_uint32(0)
        t = C.c_float(0)
        pt = np.empty(3, dtype=np.float32)
        ok = self.orc.lib.orc_dbvt_query_ray_closest(self.h, _p(ray6), C.byref(v), _p(pt), C.byref(t))
        return (int(v.value), pt, float(t.value)) if ok else None

    def query_frustum(self, planes24):
        planes24 = np.ascontiguousarray(planes24, dtype=np.float32)
        vals = np.empty(self.n, dtype=np.uint32)
        rel = np.empty(self.n, dtype=np.uint8)
        c = self.orc.lib.orc_dbvt_query_frustum(self.h, _p(planes24), _p(vals), _p(rel), C.c_uint64(self.n))
        return vals[:c].copy(), rel[:c].copy()

    def find_collider_pairs(self, dirty):
        dirty = np.ascontiguousarray(dirty, dtype=np.uint8)
        cap = 64 * self.n + 1024
        out = np.empty((cap, 2), dtype=np.uint32)
        c = self.orc.lib.orc_dbvt_find_collider_pairs(self.h, _p(dirty), _p(out), C.c_uint64(cap))
        return out[:c].copy()

    def origin_inside_ancestor(self, value, origin):
        """True when a branch bound above `value`'s leaf contains `origin` (the reference's closest-ray prune can then miss it)."""
        o = np.ascontiguousarray(origin, dtype=np.float32)
        r = self.orc.lib.orc_dbvt_origin_inside_ancestor(self.h, C.c_uint32(int(value)), _p(o))
        assert r >= 0
        return bool(r)

    def stack_overflow(self):
        return bool(self.orc.lib.orc_dbvt_stack_overflow(self.h))

    def __del__(self):
        try:
            self.orc.lib.orc_dbvt_free(self.h)
        except Exception:
            pass
