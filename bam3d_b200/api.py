"""Host-side mirror of bam3d's narrow-phase API over the CUDA C ABI.

Names, argument meaning and None/Some behaviour follow the reference crate:
  primitives   src/primitive/{sphere,cuboid,cylinder,capsule,quad,particle,polyhedron}.rs (constructors)
  Contact, CollisionStrategy   src/contact.rs:12-69
  GJK          src/algorithm/minkowski/gjk/mod.rs:32-524 (new, new_with_settings, intersect, intersection,
               distance, intersection_time_of_impact) — the scalar methods are batch-of-1 calls into the same
               kernels (no CPU fallback), the *_batch methods are the additions that make the path data parallel.
Transforms are glam Mat4 values given as 16 f32 in column-major order (Mat4::to_cols_array).
"""
import ctypes as C
import enum
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import _ffi
from ._ffi import Bam3dError, GjkSettings


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


class Context:
    """One per host thread: owns the CUDA stream and device scratch (bam3d_ctx)."""

    def __init__(self, device=0):
        self._lib = _ffi.lib()
        h = C.c_void_p()
        rc = self._lib.bam3d_ctx_create(int(device), C.byref(h))
        if rc != _ffi.OK:
            raise Bam3dError(rc, "bam3d_ctx_create failed: no usable CUDA device (there is no CPU fallback)")
        self.handle = h
        self.device = int(device)

    def check(self, rc):
        if rc != _ffi.OK:
            raise Bam3dError(rc, self._lib.bam3d_last_error(self.handle).decode())

    def synchronize(self):
        self.check(self._lib.bam3d_ctx_synchronize(self.handle))

    @property
    def stream(self):
        return self._lib.bam3d_ctx_stream(self.handle)

    @property
    def launch_count(self):
        return int(self._lib.bam3d_ctx_launch_count(self.handle))

    def set_option(self, name, value):
        self.check(self._lib.bam3d_ctx_set_option(self.handle, name.encode(), int(value)))

    def read_timing(self):
        """-> (total_ms, calls) of the bracketed main kernels since the last read (option "timing" = 1)."""
        ms, cnt = C.c_double(0), C.c_uint64(0)
        self.check(self._lib.bam3d_ctx_read_timing(self.handle, C.byref(ms), C.byref(cnt)))
        return float(ms.value), int(cnt.value)

    def close(self):
        if self.handle:
            self._lib.bam3d_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


# ---------------------------------------------------------------------------------------------------------
# primitives (constructors mirror the reference's `new`)
# ---------------------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Particle:  # primitive/particle.rs:19-30
    kind = _ffi.KIND_PARTICLE

    def params(self):
        return (0.0, 0.0, 0.0, 0.0)


@dataclass(frozen=True)
class Sphere:  # primitive/sphere.rs:8-18
    radius: float
    kind = _ffi.KIND_SPHERE

    def params(self):
        return (self.radius, 0.0, 0.0, 0.0)


class Cuboid:  # primitive/cuboid.rs:13-35: new(dim_x, dim_y, dim_z), half_dim = dim / 2
    kind = _ffi.KIND_CUBOID

    def __init__(self, dim_x, dim_y, dim_z):
        self.dim = np.array([dim_x, dim_y, dim_z], dtype=np.float32)
        self.half_dim = self.dim / np.float32(2.0)

    def params(self):
        return (self.half_dim[0], self.half_dim[1], self.half_dim[2], 0.0)


class Cube(Cuboid):  # primitive/cuboid.rs:109-131
    def __init__(self, dim):
        super().__init__(dim, dim, dim)


class Quad:  # primitive/quad.rs:13-35
    kind = _ffi.KIND_QUAD

    def __init__(self, dim_x, dim_y):
        self.half_dim = np.array([dim_x, dim_y], dtype=np.float32) / np.float32(2.0)

    def params(self):
        return (self.half_dim[0], self.half_dim[1], 0.0, 0.0)


@dataclass(frozen=True)
class Cylinder:  # primitive/cylinder.rs:12-24 (Y-axis aligned)
    half_height: float
    radius: float
    kind = _ffi.KIND_CYLINDER

    def params(self):
        return (self.half_height, self.radius, 0.0, 0.0)


@dataclass(frozen=True)
class Capsule:  # primitive/capsule.rs:15-27 (Y-axis aligned)
    half_height: float
    radius: float
    kind = _ffi.KIND_CAPSULE

    def params(self):
        return (self.half_height, self.radius, 0.0, 0.0)


class ConvexPolyhedron:  # primitive/polyhedron.rs:69-115: new (VertexOnly mode) / new_with_faces (HalfEdge mode)
    kind = _ffi.KIND_POLYHEDRON

    def __init__(self, vertices, faces=None):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float32).reshape(-1, 3)
        self.faces = None if faces is None else np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)

    @classmethod
    def new(cls, vertices):  # polyhedron.rs:69-95
        return cls(vertices)

    @classmethod
    def new_with_faces(cls, vertices, faces):  # polyhedron.rs:98-115: hill-climbing support, ray casts
        return cls(vertices, faces)

    def params(self):
        return (0.0, 0.0, 0.0, 0.0)


class CollisionStrategy(enum.IntEnum):  # contact.rs:12-18
    FullResolution = _ffi.FULL_RESOLUTION
    CollisionOnly = _ffi.COLLISION_ONLY


@dataclass
class Contact:  # contact.rs:22-38
    strategy: CollisionStrategy
    normal: np.ndarray
    penetration_depth: float
    contact_point: np.ndarray
    time_of_impact: float = 0.0


# ---------------------------------------------------------------------------------------------------------
# device-resident packs
# ---------------------------------------------------------------------------------------------------------
class MeshLibrary:
    """Vertex lists of ConvexPolyhedron values, resident on the device (bam3d_meshes)."""

    def __init__(self, ctx, verts, offsets, faces=None, face_offsets=None):
        """faces / face_offsets (optional): triangles (vertex indices local to their mesh) of the meshes built with
        ConvexPolyhedron::new_with_faces; mesh m owns faces face_offsets[m] .. face_offsets[m + 1] (none = VertexOnly)."""
        self.ctx = ctx
        self.verts = np.ascontiguousarray(verts, dtype=np.float32).reshape(-1, 3)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        h = C.c_void_p()
        if faces is None:
            self.faces = self.face_offsets = None
            ctx.check(ctx._lib.bam3d_meshes_upload(ctx.handle, _ptr(self.verts), _ptr(self.offsets), len(self.offsets) - 1, C.byref(h)))
        else:
            self.faces = np.ascontiguousarray(faces, dtype=np.uint32).reshape(-1, 3)
            self.face_offsets = np.ascontiguousarray(face_offsets, dtype=np.uint32)
            ctx.check(ctx._lib.bam3d_meshes_upload_faces(ctx.handle, _ptr(self.verts), _ptr(self.offsets), len(self.offsets) - 1,
                                                         _ptr(self.faces), _ptr(self.face_offsets), C.byref(h)))
        self.handle = h

    def close(self):
        if self.handle:
            self.ctx._lib.bam3d_meshes_free(self.ctx.handle, self.handle)
            self.handle = None


class ShapeSet:
    """Primitives + model-to-world Mat4, packed into device buffers (bam3d_shapes)."""

    def __init__(self, ctx, kind, params, xform, mesh_id=None, meshes=None):
        self.ctx = ctx
        self.kind = np.ascontiguousarray(kind, dtype=np.uint8)
        n = len(self.kind)
        self.params = np.ascontiguousarray(params, dtype=np.float32).reshape(n, 4)
        xform = np.ascontiguousarray(xform, dtype=np.float32).reshape(n, 16)
        self.mesh_id = None if mesh_id is None else np.ascontiguousarray(mesh_id, dtype=np.uint32)
        self.meshes = meshes
        h = C.c_void_p()
        ctx.check(ctx._lib.bam3d_shapes_upload(ctx.handle, _ptr(self.kind), _ptr(self.params), _ptr(xform), _ptr(self.mesh_id), n,
                                               meshes.handle if meshes is not None else None, C.byref(h)))
        self.handle = h
        self.n = n

    @classmethod
    def from_primitives(cls, ctx, primitives, transforms):
        """primitives: sequence of primitive objects; transforms: matching sequence of 16-float Mat4 arrays."""
        kind = np.array([p.kind for p in primitives], dtype=np.uint8)
        params = np.array([p.params() for p in primitives], dtype=np.float32).reshape(-1, 4)
        mesh_id, meshes = None, None
        polys = [p for p in primitives if p.kind == _ffi.KIND_POLYHEDRON]
        if polys:
            offs, verts, mesh_id = [0], [], np.zeros(len(primitives), dtype=np.uint32)
            foffs, faces = [0], []
            m = 0
            for i, p in enumerate(primitives):
                if p.kind == _ffi.KIND_POLYHEDRON:
                    verts.append(p.vertices)
                    offs.append(offs[-1] + len(p.vertices))
                    if p.faces is not None:
                        faces.append(p.faces)
                    foffs.append(foffs[-1] + (len(p.faces) if p.faces is not None else 0))
                    mesh_id[i] = m
                    m += 1
            if faces:
                meshes = MeshLibrary(ctx, np.concatenate(verts, axis=0), np.array(offs, dtype=np.uint32), np.concatenate(faces, axis=0),
                                     np.array(foffs, dtype=np.uint32))
            else:
                meshes = MeshLibrary(ctx, np.concatenate(verts, axis=0), np.array(offs, dtype=np.uint32))
        return cls(ctx, kind, params, np.array(transforms, dtype=np.float32).reshape(-1, 16), mesh_id, meshes)

    def set_transforms(self, xform):
        xform = np.ascontiguousarray(xform, dtype=np.float32).reshape(self.n, 16)
        self.ctx.check(self.ctx._lib.bam3d_shapes_set_transforms(self.ctx.handle, self.handle, _ptr(xform)))

    def set_transforms_affine(self, xform12):
        """The same from affine3x4(xform): 12 floats per shape, a quarter less to upload."""
        xform12 = np.ascontiguousarray(xform12, dtype=np.float32).reshape(self.n, 12)
        self.ctx.check(self.ctx._lib.bam3d_shapes_set_transforms_affine(self.ctx.handle, self.handle, _ptr(xform12)))

    def close(self):
        if self.handle:
            self.ctx._lib.bam3d_shapes_free(self.ctx.handle, self.handle)
            self.handle = None


def affine3x4(xform):
    """Column-major Mat4 rows [n, 16] -> [n, 12]: the xyz of the four columns (BAM3D_XFORM_AFFINE3X4). The fourth row of a model
    transform is (0, 0, 0, 1); anything else raises (the Mat4 upload would fail with ERR_NON_AFFINE)."""
    x = np.ascontiguousarray(xform, dtype=np.float32).reshape(-1, 4, 4)
    if not (np.all(x[:, :3, 3] == 0.0) and np.all(x[:, 3, 3] == 1.0)):
        raise ValueError("affine3x4: a transform's fourth row is not (0, 0, 0, 1)")
    return np.ascontiguousarray(x[:, :, :3]).reshape(-1, 12)


def pinned_empty(ctx, shape, dtype):
    """numpy array over page-locked host memory from bam3d_host_alloc (full-rate, asynchronous PCIe copies). The
    memory is released when the array (and every view of it) is garbage-collected."""
    import weakref
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    nbytes = max(1, count * dtype.itemsize)
    ptr = C.c_void_p()
    ctx.check(ctx._lib.bam3d_host_alloc(ctx.handle, nbytes, C.byref(ptr)))
    buf = (C.c_char * nbytes).from_address(ptr.value)
    arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
    weakref.finalize(buf, ctx._lib.bam3d_host_free, ctx.handle, ptr)
    return arr


class FrameQueue:
    """bam3d_frames: the per-frame sequence (new transforms -> narrow phase -> results in host memory) with `depth`
    frames in flight, so that uploads and downloads run beside the kernels of the neighbouring frames.

        q = FrameQueue(gjk, kind, params, xform0, max_pairs=n, depth=2)
        f = q.submit_contact(strategy, xform, pairs, out_contacts, out_status)   # returns at once
        ...
        q.wait(f)                                                                # out_* now hold frame f's results

    The arrays handed to submit_* must stay alive and untouched until wait(f) returns (use pinned_empty for them)."""

    def __init__(self, gjk, kind, params, xform, mesh_id=None, meshes=None, max_pairs=0, depth=2, with_end=False):
        self.gjk, self.ctx = gjk, gjk.ctx
        kind = np.ascontiguousarray(kind, dtype=np.uint8)
        self.n_shapes = len(kind)
        params = np.ascontiguousarray(params, dtype=np.float32).reshape(self.n_shapes, 4)
        xform = np.ascontiguousarray(xform, dtype=np.float32).reshape(self.n_shapes, 16)
        mesh_id = None if mesh_id is None else np.ascontiguousarray(mesh_id, dtype=np.uint32)
        self.meshes = meshes
        self.max_pairs = int(max_pairs)
        h = C.c_void_p()
        self.ctx.check(self.ctx._lib.bam3d_frames_create(self.ctx.handle, _ptr(kind), _ptr(params), _ptr(xform), _ptr(mesh_id), self.n_shapes,
                                                         meshes.handle if meshes is not None else None, self.max_pairs, int(depth),
                                                         1 if with_end else 0, C.byref(h)))
        self.handle = h
        self.xform_width = 16
        self._keep = {}  # frame -> arrays referenced by the copies in flight

    def set_transform_layout(self, layout):
        """_ffi.XFORM_MAT4 (16 floats per shape, the default) or _ffi.XFORM_AFFINE3X4 (12: see affine3x4) for later submits."""
        self.ctx.check(self.ctx._lib.bam3d_frames_set_transform_layout(self.ctx.handle, self.handle, int(layout)))
        self.xform_width = 12 if int(layout) == _ffi.XFORM_AFFINE3X4 else 16

    def _submit(self, query, strategy, xform, xform_end, pairs, out_contacts, out_status):
        def chk(a, dtype, shape):
            if not (isinstance(a, np.ndarray) and a.dtype == dtype and a.flags.c_contiguous and a.size == int(np.prod(shape))):
                raise ValueError(f"FrameQueue: expected a C-contiguous {np.dtype(dtype).name} array of {shape} elements")
            return a
        n = pairs.size // 2
        chk(pairs, np.uint32, (n, 2)); chk(xform, np.float32, (self.n_shapes, self.xform_width)); chk(out_status, np.uint8, (n,))
        if xform_end is not None:
            chk(xform_end, np.float32, (self.n_shapes, self.xform_width))
        if out_contacts is not None:
            chk(out_contacts, np.float32, (n, 8))
        frame = C.c_uint64(0)
        self.ctx.check(self.ctx._lib.bam3d_frames_submit(self.ctx.handle, self.handle, query, _ptr(xform), _ptr(xform_end), _ptr(pairs), n,
                                                         C.byref(self.gjk.settings), int(strategy), _ptr(out_contacts), _ptr(out_status),
                                                         C.byref(frame)))
        self._keep[frame.value] = (xform, xform_end, pairs, out_contacts, out_status)
        return frame.value

    def submit_intersect(self, xform, pairs, out_status):
        return self._submit(_ffi.FRAME_INTERSECT, 0, xform, None, pairs, None, out_status)

    def submit_contact(self, strategy, xform, pairs, out_contacts, out_status):
        return self._submit(_ffi.FRAME_CONTACT, strategy, xform, None, pairs, out_contacts, out_status)

    def submit_time_of_impact(self, xform_start, xform_end, pairs, out_contacts, out_status):
        return self._submit(_ffi.FRAME_TIME_OF_IMPACT, 0, xform_start, xform_end, pairs, out_contacts, out_status)

    def submit_contact_gather(self, comm, strategy, xform, pairs, pair_index_base, out_gathered, out_status):
        """bam3d_frames_submit_gather: like submit_contact, but wait() leaves the hits of ALL ranks (bam3d_contact40 rows as
        uint32[capacity, 10]) in out_gathered instead of this rank's dense contacts; gathered(frame) gives the counts."""
        n = pairs.size // 2
        frame = C.c_uint64(0)
        self.ctx.check(self.ctx._lib.bam3d_frames_submit_gather(self.ctx.handle, self.handle, comm.handle, _ffi.FRAME_CONTACT, _ptr(xform), None,
                                                                _ptr(pairs), n, C.byref(self.gjk.settings), int(strategy), int(pair_index_base),
                                                                _ptr(out_gathered), out_gathered.size // 10, _ptr(out_status), C.byref(frame)))
        self._keep[frame.value] = (xform, pairs, out_gathered, out_status)
        self._n_ranks = comm.n_ranks
        return frame.value

    def gathered(self, frame):
        """-> (per-rank record counts, total) of a gather-mode frame, after wait(frame)."""
        counts = np.zeros(self._n_ranks, dtype=np.uint64)
        total = C.c_uint64(0)
        self.ctx.check(self.ctx._lib.bam3d_frames_gathered(self.ctx.handle, self.handle, int(frame), _ptr(counts), C.byref(total)))
        return counts, int(total.value)

    def wait(self, frame):
        try:
            self.ctx.check(self.ctx._lib.bam3d_frames_wait(self.ctx.handle, self.handle, int(frame)))
        finally:
            for f in [f for f in self._keep if f <= frame]:
                del self._keep[f]

    def close(self):
        if self.handle:
            self.ctx._lib.bam3d_frames_free(self.ctx.handle, self.handle)
            self.handle = None
            self._keep.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---------------------------------------------------------------------------------------------------------
# GJK
# ---------------------------------------------------------------------------------------------------------
class GJK:
    """Gilbert-Johnson-Keerthi narrow phase (gjk/mod.rs:24-58) running on the GPU."""

    def __init__(self, ctx=None, distance_tolerance=0.000001, continuous_tolerance=0.000001, epa_tolerance=0.00001,
                 max_iterations=100, toi_iteration_cap=100):
        self.ctx = ctx if ctx is not None else default_context()
        self.settings = GjkSettings(distance_tolerance, continuous_tolerance, epa_tolerance, max_iterations, toi_iteration_cap)

    @classmethod
    def new(cls, ctx=None):  # gjk/mod.rs:34-42
        return cls(ctx)

    @classmethod
    def new_with_settings(cls, distance_tolerance, continuous_tolerance, epa_tolerance, max_iterations, ctx=None):  # :45-58
        return cls(ctx, distance_tolerance, continuous_tolerance, epa_tolerance, max_iterations)

    # ---- batched entry points (host buffers) -------------------------------------------------------------
    @staticmethod
    def _pairs(pairs):
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        return pairs, len(pairs)

    def intersect_batch(self, shapes, pairs):
        """GJK::intersect per pair (gjk/mod.rs:74-113). Returns status[n] (0 == Some(simplex))."""
        pairs, n = self._pairs(pairs)
        status = np.empty(n, dtype=np.uint8)
        self.ctx.check(self.ctx._lib.bam3d_gjk_intersect(self.ctx.handle, shapes.handle, _ptr(pairs), n, C.byref(self.settings), _ptr(status)))
        return status

    def intersection_batch(self, strategy, shapes, pairs):
        """GJK::intersection per pair (gjk/mod.rs:317-341). Returns (contacts[n, 8], status[n]); contact rows are
        normal xyz, penetration_depth, contact_point xyz, time_of_impact."""
        pairs, n = self._pairs(pairs)
        status = np.empty(n, dtype=np.uint8)
        contacts = np.empty((n, 8), dtype=np.float32)
        self.ctx.check(self.ctx._lib.bam3d_gjk_contact(self.ctx.handle, shapes.handle, _ptr(pairs), n, C.byref(self.settings), int(strategy),
                                                       _ptr(contacts), _ptr(status)))
        return contacts, status

    def distance_batch(self, shapes, pairs):
        """GJK::distance per pair (gjk/mod.rs:232-280). Returns (ref_value[n], distance[n], status[n])."""
        pairs, n = self._pairs(pairs)
        status = np.empty(n, dtype=np.uint8)
        ref = np.empty(n, dtype=np.float32)
        dist = np.empty(n, dtype=np.float32)
        self.ctx.check(self.ctx._lib.bam3d_gjk_distance(self.ctx.handle, shapes.handle, _ptr(pairs), n, C.byref(self.settings), _ptr(ref), _ptr(dist),
                                                        _ptr(status)))
        return ref, dist, status

    def intersection_time_of_impact_batch(self, shapes_start, shapes_end, pairs):
        """GJK::intersection_time_of_impact per pair (gjk/mod.rs:130-217)."""
        pairs, n = self._pairs(pairs)
        status = np.empty(n, dtype=np.uint8)
        contacts = np.empty((n, 8), dtype=np.float32)
        self.ctx.check(self.ctx._lib.bam3d_gjk_time_of_impact(self.ctx.handle, shapes_start.handle, shapes_end.handle, _ptr(pairs), n,
                                                              C.byref(self.settings), _ptr(contacts), _ptr(status)))
        return contacts, status

    # ---- device-resident entry points (pointers are raw device addresses; asynchronous on ctx.stream) ------
    def intersect_dev(self, shapes, d_pairs, n, d_status):
        self.ctx.check(self.ctx._lib.bam3d_gjk_intersect_dev(self.ctx.handle, shapes.handle, d_pairs, n, C.byref(self.settings), d_status))

    def intersection_dev(self, strategy, shapes, d_pairs, n, d_contacts, d_status):
        self.ctx.check(self.ctx._lib.bam3d_gjk_contact_dev(self.ctx.handle, shapes.handle, d_pairs, n, C.byref(self.settings), int(strategy),
                                                           d_contacts, d_status))

    def distance_dev(self, shapes, d_pairs, n, d_ref, d_dist, d_status):
        self.ctx.check(self.ctx._lib.bam3d_gjk_distance_dev(self.ctx.handle, shapes.handle, d_pairs, n, C.byref(self.settings), d_ref, d_dist,
                                                            d_status))

    def intersection_time_of_impact_dev(self, shapes_start, shapes_end, d_pairs, n, d_contacts, d_status):
        self.ctx.check(self.ctx._lib.bam3d_gjk_time_of_impact_dev(self.ctx.handle, shapes_start.handle, shapes_end.handle, d_pairs, n,
                                                                  C.byref(self.settings), d_contacts, d_status))

    def compact_contacts_dev(self, d_contacts, d_status, n, base, d_out, capacity, d_count):
        self.ctx.check(self.ctx._lib.bam3d_compact_contacts_dev(self.ctx.handle, d_contacts, d_status, n, base, d_out, capacity, d_count))

    # ---- the reference's scalar methods, as batch-of-1 -----------------------------------------------------
    def _two(self, left, left_transform, right, right_transform):
        return ShapeSet.from_primitives(self.ctx, [left, right], [left_transform, right_transform])

    @staticmethod
    def _contact(strategy, row):
        return Contact(CollisionStrategy(strategy), row[0:3].copy(), float(row[3]), row[4:7].copy(), float(row[7]))

    def intersect(self, left, left_transform, right, right_transform):
        """-> bool (the reference returns Option<Simplex>; the simplex itself stays on the device)."""
        s = self._two(left, left_transform, right, right_transform)
        try:
            return bool(self.intersect_batch(s, [[0, 1]])[0] == _ffi.STATUS_OK)
        finally:
            s.close()

    def intersection(self, strategy, left, left_transform, right, right_transform):
        s = self._two(left, left_transform, right, right_transform)
        try:
            c, st = self.intersection_batch(strategy, s, [[0, 1]])
            return self._contact(strategy, c[0]) if st[0] == _ffi.STATUS_OK else None
        finally:
            s.close()

    def distance(self, left, left_transform, right, right_transform):
        """-> Option<f32> exactly as the reference computes it (the residual dp - d0, gjk/mod.rs:273-275)."""
        s = self._two(left, left_transform, right, right_transform)
        try:
            ref, _, st = self.distance_batch(s, [[0, 1]])
            return float(ref[0]) if st[0] == _ffi.STATUS_OK else None
        finally:
            s.close()

    def intersection_time_of_impact(self, left, left_transform_range, right, right_transform_range):
        """transform ranges are (start, end) pairs, the reference's Range<&Mat4>."""
        s0 = self._two(left, left_transform_range[0], right, right_transform_range[0])
        s1 = self._two(left, left_transform_range[1], right, right_transform_range[1])
        try:
            c, st = self.intersection_time_of_impact_batch(s0, s1, [[0, 1]])
            return self._contact(CollisionStrategy.FullResolution, c[0]) if st[0] == _ffi.STATUS_OK else None
        finally:
            s0.close()
            s1.close()


# ---------------------------------------------------------------------------------------------------------
# compound shapes (gjk/mod.rs:362-523): every primitive x primitive combination, then a per-pair reduction
# ---------------------------------------------------------------------------------------------------------
def mat4_mul(a, b):
    """glam 0.8 Mat4::mul_mat4 (a * b), vectorised over rows of 16 f32 (column-major), f32, glam's operation order:
    each result column is a.x_axis * v.x, then y_axis * v.y + r, z_axis * v.z + r, w_axis * v.w + r."""
    a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 4, 4)  # [n][column][row]
    b = np.ascontiguousarray(b, dtype=np.float32).reshape(-1, 4, 4)
    out = np.empty_like(b)
    for j in range(4):
        v = b[:, j, :]
        r = a[:, 0, :] * v[:, 0:1]
        r = a[:, 1, :] * v[:, 1:2] + r
        r = a[:, 2, :] * v[:, 2:3] + r
        r = a[:, 3, :] * v[:, 3:4] + r
        out[:, j, :] = r
    return out.reshape(-1, 16)


class CompoundSet:
    """Compound shapes: compound c owns primitives comp_offsets[c] .. comp_offsets[c+1] of (kind, params, local_xform);
    the reference's `&[(Primitive, Mat4)]` slices (gjk/mod.rs:362-371), resident on the device (bam3d_compounds)."""

    def __init__(self, ctx, kind, params, local_xform, comp_offsets, mesh_id=None, meshes=None):
        self.ctx = ctx
        self.kind = np.ascontiguousarray(kind, dtype=np.uint8)
        self.params = np.ascontiguousarray(params, dtype=np.float32).reshape(-1, 4)
        self.local = np.ascontiguousarray(local_xform, dtype=np.float32).reshape(-1, 16)
        self.offsets = np.ascontiguousarray(comp_offsets, dtype=np.uint32)
        self.mesh_id = None if mesh_id is None else np.ascontiguousarray(mesh_id, dtype=np.uint32)
        self.meshes = meshes
        if self.offsets[-1] != len(self.kind):
            raise ValueError("comp_offsets[-1] must equal the number of primitives")
        self.n_compounds = len(self.offsets) - 1
        h = C.c_void_p()
        ctx.check(ctx._lib.bam3d_compounds_upload(ctx.handle, _ptr(self.kind), _ptr(self.params), _ptr(self.local), _ptr(self.mesh_id), len(self.kind),
                                                  _ptr(self.offsets), self.n_compounds, meshes.handle if meshes is not None else None, C.byref(h)))
        self.handle = h

    def close(self):
        if self.handle:
            self.ctx._lib.bam3d_compounds_free(self.ctx.handle, self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _complex_methods():
    def _xf(compounds, x):
        return np.ascontiguousarray(x, dtype=np.float32).reshape(compounds.n_compounds, 16)

    def intersection_complex_batch(self, strategy, compounds, comp_xform, pairs):
        """GJK::intersection_complex per compound pair (gjk/mod.rs:362-407) -> (contacts[n, 8], status[n])."""
        pairs, n = self._pairs(pairs)
        out, st = np.zeros((n, 8), dtype=np.float32), np.empty(n, dtype=np.uint8)
        self.ctx.check(self.ctx._lib.bam3d_gjk_contact_complex(self.ctx.handle, compounds.handle, _ptr(_xf(compounds, comp_xform)), _ptr(pairs), n,
                                                               C.byref(self.settings), int(strategy), _ptr(out), _ptr(st)))
        return out, st

    def distance_complex_batch(self, compounds, comp_xform, pairs):
        """GJK::distance_complex per compound pair (gjk/mod.rs:422-456) -> (value[n], status[n]); value = the reference's
        return (min of the per-primitive residuals)."""
        pairs, n = self._pairs(pairs)
        out, st = np.zeros(n, dtype=np.float32), np.empty(n, dtype=np.uint8)
        self.ctx.check(self.ctx._lib.bam3d_gjk_distance_complex(self.ctx.handle, compounds.handle, _ptr(_xf(compounds, comp_xform)), _ptr(pairs), n,
                                                                C.byref(self.settings), _ptr(out), _ptr(st)))
        return out, st

    def intersection_complex_time_of_impact_batch(self, strategy, compounds, comp_xform_start, comp_xform_end, pairs):
        """GJK::intersection_complex_time_of_impact per compound pair (gjk/mod.rs:478-523)."""
        pairs, n = self._pairs(pairs)
        out, st = np.zeros((n, 8), dtype=np.float32), np.empty(n, dtype=np.uint8)
        self.ctx.check(self.ctx._lib.bam3d_gjk_time_of_impact_complex(self.ctx.handle, compounds.handle, _ptr(_xf(compounds, comp_xform_start)),
                                                                      _ptr(_xf(compounds, comp_xform_end)), _ptr(pairs), n, C.byref(self.settings),
                                                                      int(strategy), _ptr(out), _ptr(st)))
        return out, st

    GJK.intersection_complex_batch = intersection_complex_batch
    GJK.distance_complex_batch = distance_complex_batch
    GJK.intersection_complex_time_of_impact_batch = intersection_complex_time_of_impact_batch


_complex_methods()


# ---------------------------------------------------------------------------------------------------------
# ray casts against primitives
# ---------------------------------------------------------------------------------------------------------
@dataclass
class Ray:  # ray.rs:7-24: origin + direction as given (not normalised by the constructor)
    origin: Sequence[float]
    direction: Sequence[float]

    def row(self):
        return np.concatenate([np.asarray(self.origin, dtype=np.float32), np.asarray(self.direction, dtype=np.float32)])


def ray_cast_batch(shapes, rays, casts, query=_ffi.RAY_INTERSECTION):
    """Batched Discrete/ContinuousTransformed<Ray> for Primitive3 (ray.rs:58-78, primitive3.rs:133-166) and the swept
    particle (particle.rs:39-128). rays: [m, 6] (origin, direction; PARTICLE_SWEEP: start, end); casts: [n, 2] =
    (ray index, shape index). -> (points[n, 3], status[n]); status 0 = true / Some(point), 1 = false / None,
    STATUS_UNSUPPORTED for a ConvexPolyhedron."""
    ctx = shapes.ctx
    rays = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 6)
    casts = np.ascontiguousarray(casts, dtype=np.uint32).reshape(-1, 2)
    n = len(casts)
    points = np.zeros((n, 3), dtype=np.float32)
    status = np.ones(n, dtype=np.uint8)
    ctx.check(ctx._lib.bam3d_ray_cast(ctx.handle, shapes.handle, _ptr(rays), len(rays), _ptr(casts), n, int(query), _ptr(points), _ptr(status)))
    return points, status


def intersects_transformed(primitive, ray, transform, ctx=None):
    """DiscreteTransformed<Ray>::intersects_transformed (ray.rs:58-66) as a batch of one."""
    ctx = ctx if ctx is not None else default_context()
    s = ShapeSet.from_primitives(ctx, [primitive], [transform])
    _, st = ray_cast_batch(s, [ray.row()], [[0, 0]], _ffi.RAY_INTERSECTS)
    if st[0] == _ffi.STATUS_UNSUPPORTED:
        raise NotImplementedError("ray casts against ConvexPolyhedron are not built (polyhedron.rs:373-522)")
    return bool(st[0] == _ffi.STATUS_OK)


def intersection_transformed(primitive, ray, transform, ctx=None):
    """ContinuousTransformed<Ray>::intersection_transformed (ray.rs:68-78) as a batch of one: hit point or None."""
    ctx = ctx if ctx is not None else default_context()
    s = ShapeSet.from_primitives(ctx, [primitive], [transform])
    pts, st = ray_cast_batch(s, [ray.row()], [[0, 0]], _ffi.RAY_INTERSECTION)
    if st[0] == _ffi.STATUS_UNSUPPORTED:
        raise NotImplementedError("ray casts against ConvexPolyhedron are not built (polyhedron.rs:373-522)")
    return pts[0] if st[0] == _ffi.STATUS_OK else None


def particle_intersection_transformed(primitive, start, end, transform, ctx=None):
    """ContinuousTransformed<(Particle, Range<Vec3>)> (particle.rs:108-128): where the particle moving start -> end hits."""
    ctx = ctx if ctx is not None else default_context()
    s = ShapeSet.from_primitives(ctx, [primitive], [transform])
    seg = np.concatenate([np.asarray(start, dtype=np.float32), np.asarray(end, dtype=np.float32)])
    pts, st = ray_cast_batch(s, [seg], [[0, 0]], _ffi.PARTICLE_SWEEP)
    return pts[0] if st[0] == _ffi.STATUS_OK else None
