"""Seeded synthetic scenes for the configs BASELINE.json names.

Counter-based RNG: SplitMix64 output function of (seed, stream, index), so any slice of a scene can be generated
independently (each rank of a multi-GPU run generates exactly its own slice). Transforms are built the way glam's
Mat4::from_scale_rotation_translation(1, q, t) builds them, in f32 with glam's operation order.
"""
import numpy as np

from . import _ffi

SEED_C1, SEED_C2, SEED_C3, SEED_C4, SEED_C5 = 0xB3D00001, 0xB3D00002, 0xB3D00003, 0xB3D00004, 0xB3D00005
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def _bits(seed, stream, idx):
    with np.errstate(over="ignore"):
        key = splitmix64(np.uint64(seed) ^ (np.uint64(stream) * np.uint64(0xD1B54A32D192ED03)))
        return splitmix64(key + idx.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15))


def uniform(seed, stream, idx, lo=0.0, hi=1.0):
    """U[lo, hi) as f32; idx is an integer array of counters."""
    u = (_bits(seed, stream, idx) >> np.uint64(40)).astype(np.float64) * (1.0 / (1 << 24))
    return (lo + (hi - lo) * u).astype(np.float32)


def randint(seed, stream, idx, n):
    return (_bits(seed, stream, idx) % np.uint64(n)).astype(np.uint32)


def gaussian(seed, stream, idx):
    u1 = ((_bits(seed, stream * 2 + 1000, idx) >> np.uint64(11)).astype(np.float64) + 1.0) * (1.0 / (1 << 53))
    u2 = (_bits(seed, stream * 2 + 1001, idx) >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def random_quat(seed, stream, idx):
    """Uniform rotations: normalised 4-Gaussian quaternions, f32 (x, y, z, w)."""
    q = np.stack([gaussian(seed, stream + k, idx) for k in range(4)], axis=1)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q.astype(np.float32)


def mat4_from_rotation_translation(q, t):
    """glam 0.8 Mat4::from_scale_rotation_translation(Vec3::splat(1), q, t), vectorised in f32 (n x 16, column-major)."""
    q = q.astype(np.float32)
    t = t.astype(np.float32)
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    one = np.float32(1.0)
    x2, y2, z2 = x + x, y + y, z + z
    xx, xy, xz = x * x2, x * y2, x * z2
    yy, yz, zz = y * y2, y * z2, z * z2
    wx, wy, wz = w * x2, w * y2, w * z2
    m = np.zeros((len(q), 16), dtype=np.float32)
    m[:, 0] = (one - (yy + zz)) * one
    m[:, 1] = (xy + wz) * one
    m[:, 2] = (xz - wy) * one
    m[:, 4] = (xy - wz) * one
    m[:, 5] = (one - (xx + zz)) * one
    m[:, 6] = (yz + wx) * one
    m[:, 8] = (xz + wy) * one
    m[:, 9] = (yz - wx) * one
    m[:, 10] = (one - (xx + yy)) * one
    m[:, 12:15] = t
    m[:, 15] = one
    return m


def _pair_transforms(seed, idx, s):
    """left: uniform rotation, translation U[-10,10]^3; right: uniform rotation, left translation + U[-s,s]^3."""
    tl = np.stack([uniform(seed, 10 + k, idx, -10.0, 10.0) for k in range(3)], axis=1)
    off = np.stack([uniform(seed, 13 + k, idx, -s, s) for k in range(3)], axis=1)
    tr = (tl + off).astype(np.float32)
    ql = random_quat(seed, 20, idx)
    qr = random_quat(seed, 30, idx)
    return mat4_from_rotation_translation(ql, tl), mat4_from_rotation_translation(qr, tr)


def _interleave(a, b):
    out = np.empty((2 * len(a),) + a.shape[1:], dtype=a.dtype)
    out[0::2] = a
    out[1::2] = b
    return out


def _analytic_side(seed, stream, idx, kinds):
    """Random analytic primitive per index drawn uniformly from `kinds`."""
    pick = randint(seed, stream, idx, len(kinds))
    kind = np.asarray(kinds, dtype=np.uint8)[pick]
    a = uniform(seed, stream + 1, idx)
    b = uniform(seed, stream + 2, idx)
    c = uniform(seed, stream + 3, idx)
    params = np.zeros((len(idx), 4), dtype=np.float32)
    sph = kind == _ffi.KIND_SPHERE
    cub = kind == _ffi.KIND_CUBOID
    cyl = (kind == _ffi.KIND_CYLINDER) | (kind == _ffi.KIND_CAPSULE)
    quad = kind == _ffi.KIND_QUAD
    lo, hi = np.float32(0.25), np.float32(1.0)
    params[sph, 0] = (lo + (hi - lo) * a[sph])                       # radius U[0.25, 1]
    dims = np.stack([a, b, c], axis=1) * np.float32(1.5) + np.float32(0.5)  # dims U[0.5, 2]
    params[cub, 0:3] = dims[cub] / np.float32(2.0)                   # half_dim = dim / 2 (cuboid.rs:29)
    params[cyl, 0] = (lo + (hi - lo) * a[cyl])                       # half_height U[0.25, 1]
    params[cyl, 1] = (lo + (hi - lo) * b[cyl])                       # radius U[0.25, 1]
    params[quad, 0:2] = dims[quad, 0:2] / np.float32(2.0)
    return kind, params


def cuboid_pairs(n, s=2.0, seed=SEED_C1, start=0):
    """C1: n Cuboid-Cuboid pairs, dims U[0.5,2]^3, uniform rotations, relative offset U[-s,s]^3."""
    return mixed_pairs(n, s=s, seed=seed, start=start, kinds=(_ffi.KIND_CUBOID,))


def mixed_pairs(n, s=1.5, seed=SEED_C2, start=0, kinds=(_ffi.KIND_SPHERE, _ffi.KIND_CUBOID, _ffi.KIND_CYLINDER)):
    """C2: each side uniformly from `kinds` (Sphere r U[.25,1], Cuboid dims U[.5,2]^3, Cylinder hh,r U[.25,1])."""
    idx = np.arange(start, start + n, dtype=np.uint64)
    kl, pl = _analytic_side(seed, 40, idx, kinds)
    kr, pr = _analytic_side(seed, 50, idx, kinds)
    ml, mr = _pair_transforms(seed, idx, s)
    pairs = np.arange(2 * n, dtype=np.uint32).reshape(n, 2)
    return dict(kind=_interleave(kl, kr), params=_interleave(pl, pr), xform=_interleave(ml, mr), mesh_id=None,
                pairs=pairs, mesh_verts=None, mesh_offsets=None)


def mesh_library(n_meshes, vmin=64, vmax=256, seed=SEED_C3):
    """n_meshes convex polyhedra: V ~ U{vmin..vmax} points on an ellipsoid with radii U[0.5,1.5]^3 (all extreme)."""
    m = np.arange(n_meshes, dtype=np.uint64)
    counts = (vmin + randint(seed, 60, m, vmax - vmin + 1)).astype(np.uint32)
    offsets = np.zeros(n_meshes + 1, dtype=np.uint32)
    offsets[1:] = np.cumsum(counts)
    total = int(offsets[-1])
    vi = np.arange(total, dtype=np.uint64)
    g = np.stack([gaussian(seed, 70 + k, vi) for k in range(3)], axis=1)
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    radii = np.stack([uniform(seed, 61 + k, m, 0.5, 1.5) for k in range(3)], axis=1)
    owner = np.repeat(np.arange(n_meshes), counts)
    verts = (g * radii[owner]).astype(np.float32)
    return verts, offsets


def polyhedron_pairs(n, n_meshes=4096, s=2.0, seed=SEED_C3, start=0, library=None, vmin=64, vmax=256):
    """C3: n ConvexPolyhedron pairs instancing a library of n_meshes meshes (VertexOnly semantics)."""
    verts, offsets = library if library is not None else mesh_library(n_meshes, vmin, vmax, seed)
    n_meshes = len(offsets) - 1
    idx = np.arange(start, start + n, dtype=np.uint64)
    ml, mr = _pair_transforms(seed, idx, s)
    mesh_id = _interleave(randint(seed, 80, idx, n_meshes), randint(seed, 81, idx, n_meshes))
    kind = np.full(2 * n, _ffi.KIND_POLYHEDRON, dtype=np.uint8)
    params = np.zeros((2 * n, 4), dtype=np.float32)
    pairs = np.arange(2 * n, dtype=np.uint32).reshape(n, 2)
    return dict(kind=kind, params=params, xform=_interleave(ml, mr), mesh_id=mesh_id, pairs=pairs,
                mesh_verts=verts, mesh_offsets=offsets)


def end_transforms(scene, seed=SEED_C5, start=0, span=3.0):
    """TOI: end transform = start transform translated by U[-span, span]^3 (same rotation)."""
    n = len(scene["kind"])
    idx = np.arange(start, start + n, dtype=np.uint64)
    d = np.stack([uniform(seed, 90 + k, idx, -span, span) for k in range(3)], axis=1)
    x1 = scene["xform"].copy()
    x1[:, 12:15] = (x1[:, 12:15] + d).astype(np.float32)
    return x1


def all_kinds_pairs(n, s=1.5, seed=0xB3D000AA, start=0):
    """Parity-test scene touching every analytic primitive kind, including Particle, Quad and Capsule."""
    kinds = (_ffi.KIND_PARTICLE, _ffi.KIND_QUAD, _ffi.KIND_SPHERE, _ffi.KIND_CUBOID, _ffi.KIND_CYLINDER, _ffi.KIND_CAPSULE)
    return mixed_pairs(n, s=s, seed=seed, start=start, kinds=kinds)


# ---------------------------------------------------------------------------------------------------------
# broad phase (C4)
# ---------------------------------------------------------------------------------------------------------
def aabb_scene(n, seed=SEED_C4, start=0, density=4.0, extent=None):
    """n Aabb3: centres U[0, L]^3, half-extents U[0.1, 0.5]^3. C4 uses L = 100 for n = 2^22 (density 4 boxes per unit
    volume, ~7 overlaps per box); smaller n keep that density unless `extent` is given."""
    L = float(extent) if extent is not None else (n / density) ** (1.0 / 3.0)
    idx = np.arange(start, start + n, dtype=np.uint64)
    c = np.stack([uniform(seed, 100 + k, idx, 0.0, L) for k in range(3)], axis=1)
    h = np.stack([uniform(seed, 103 + k, idx, 0.1, 0.5) for k in range(3)], axis=1)
    return np.concatenate([(c - h).astype(np.float32), (c + h).astype(np.float32)], axis=1), L


def world_scene(n, seed=SEED_C4, density=0.1, kinds=(_ffi.KIND_SPHERE, _ffi.KIND_CUBOID, _ffi.KIND_CYLINDER), extent=None):
    """n shapes of the C2 mix scattered in one volume (positions U[0, L]^3, uniform rotations) — the input of the broad -> narrow
    pipeline: no pair list, sweep and prune finds the candidates. density = shapes per unit volume (0.1: ~2.7 candidate pairs per
    shape, about a quarter of them in contact)."""
    L = float(extent) if extent is not None else (n / density) ** (1.0 / 3.0)
    idx = np.arange(n, dtype=np.uint64)
    kind, params = _analytic_side(seed, 140, idx, kinds)
    t = np.stack([uniform(seed, 150 + k, idx, 0.0, L) for k in range(3)], axis=1)
    q = random_quat(seed, 160, idx)
    return dict(kind=kind, params=params, xform=mat4_from_rotation_translation(q, t), mesh_id=None, mesh_verts=None, mesh_offsets=None,
                pairs=np.zeros((0, 2), dtype=np.uint32)), L


def rays(nq, L, seed=SEED_C4, start=0):
    """Rays with origin U in the volume and direction uniform on the sphere: nq x (origin.xyz, direction.xyz)."""
    idx = np.arange(start, start + nq, dtype=np.uint64)
    o = np.stack([uniform(seed, 110 + k, idx, 0.0, L) for k in range(3)], axis=1)
    d = np.stack([gaussian(seed, 113 + k, idx) for k in range(3)], axis=1)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([o, d.astype(np.float32)], axis=1).astype(np.float32)


def perspective_rh_gl(fov_y, aspect, z_near, z_far):
    """glam Mat4::perspective_rh_gl as a column-major 16-float array (f32 arithmetic)."""
    f32 = np.float32
    inv_length = f32(1.0) / (f32(z_near) - f32(z_far))
    f = f32(1.0) / np.tan(f32(0.5) * f32(fov_y), dtype=np.float32)
    a = f / f32(aspect)
    b = (f32(z_near) + f32(z_far)) * inv_length
    c = (f32(2.0) * f32(z_near) * f32(z_far)) * inv_length
    return np.array([a, 0, 0, 0, 0, f, 0, 0, 0, 0, b, -1, 0, 0, c, 0], dtype=np.float32)


def view_projections(nq, L, seed=SEED_C4, start=0):
    """nq projection * view matrices: perspective_rh_gl(60 deg, 16/9, 0.1, 50) times a random rigid view."""
    idx = np.arange(start, start + nq, dtype=np.uint64)
    q = random_quat(seed, 120, idx)
    t = np.stack([uniform(seed, 125 + k, idx, 0.0, L) for k in range(3)], axis=1)
    view = mat4_from_rotation_translation(q, -t).reshape(nq, 4, 4)  # column-major: [col][row]
    proj = perspective_rh_gl(np.float32(np.pi / 3), 16.0 / 9.0, 0.1, 50.0).reshape(4, 4)
    # column-major product P * V: result column j = P * V[:, j]
    P = proj.T.astype(np.float32)  # row-major matrix
    out = np.empty((nq, 16), dtype=np.float32)
    for k in range(nq):
        V = view[k].T
        out[k] = (P @ V).T.reshape(16)
    return out


def primitive_rays(scene, n_rays, seed=0xB3D000F1, start=0):
    """Rays / particle segments aimed at the shapes of `scene` for the ray-cast parity tests: cast i pairs ray i with
    shape (i mod n_shapes). Origins sit U[-3, 3]^3 around the shape's position; a third of the directions point at the
    shape (plus noise), a third are random unit vectors, the rest are axis-parallel (exact zeros: the inf/NaN slab
    arithmetic and the cylinder's vertical-ray branch) or not normalised. -> (rays[n, 6], segments[n, 6], casts[n, 2])."""
    n_shapes = len(scene["kind"])
    idx = np.arange(start, start + n_rays, dtype=np.uint64)
    shape = (idx % np.uint64(n_shapes)).astype(np.uint32)
    centre = scene["xform"][shape, 12:15]
    off = np.stack([uniform(seed, 1 + k, idx, -3.0, 3.0) for k in range(3)], axis=1)
    origin = (centre + off).astype(np.float32)
    noise = np.stack([uniform(seed, 4 + k, idx, -0.6, 0.6) for k in range(3)], axis=1)
    aimed = (centre - origin + noise).astype(np.float32)
    rnd = np.stack([gaussian(seed, 7 + k, idx) for k in range(3)], axis=1).astype(np.float32)
    mode = randint(seed, 10, idx, 6)
    d = np.where((mode < 2)[:, None], aimed, rnd)
    ln = np.linalg.norm(d.astype(np.float64), axis=1, keepdims=True)
    d = (d / np.maximum(ln, 1e-30)).astype(np.float32)
    axis = randint(seed, 11, idx, 3)
    sign = np.where(randint(seed, 12, idx, 2) == 0, np.float32(1.0), np.float32(-1.0))
    axial = np.zeros((n_rays, 3), dtype=np.float32)
    axial[np.arange(n_rays), axis] = sign
    # axis-parallel rays start on the shape's axis line half of the time so that they actually hit
    on_axis = origin.copy()
    keep = np.arange(n_rays), axis
    on_axis[:] = centre
    on_axis[keep] = origin[keep]
    d = np.where((mode == 4)[:, None], axial, d)
    origin = np.where(((mode == 4) & (randint(seed, 13, idx, 2) == 0))[:, None], on_axis, origin).astype(np.float32)
    d = np.where((mode == 5)[:, None], aimed, d).astype(np.float32)  # not normalised
    rays = np.concatenate([origin, d], axis=1).astype(np.float32)
    reach = uniform(seed, 14, idx, 0.5, 6.0)[:, None]
    segments = np.concatenate([origin, (origin + d * reach).astype(np.float32)], axis=1).astype(np.float32)
    casts = np.stack([np.arange(n_rays, dtype=np.uint32), shape], axis=1)
    return rays, segments, casts


def mesh_faces(verts, offsets):
    """Triangulated convex hulls of the meshes of a library (scipy), counter-clockwise seen from outside, as
    (faces[total, 3] with vertex indices local to each mesh, face_offsets[n_meshes + 1]) — what
    ConvexPolyhedron::new_with_faces takes. Every vertex of a mesh_library() mesh lies on its hull."""
    from scipy.spatial import ConvexHull
    faces, foffs = [], [0]
    for m in range(len(offsets) - 1):
        v = np.asarray(verts[offsets[m]:offsets[m + 1]], dtype=np.float64)
        hull = ConvexHull(v)
        tri = hull.simplices.astype(np.int64).copy()
        n = np.cross(v[tri[:, 1]] - v[tri[:, 0]], v[tri[:, 2]] - v[tri[:, 0]])
        flip = np.einsum("ij,ij->i", n, hull.equations[:, :3]) < 0
        tri[flip] = tri[flip][:, [0, 2, 1]]
        faces.append(tri.astype(np.uint32))
        foffs.append(foffs[-1] + len(tri))
    return np.concatenate(faces, axis=0), np.array(foffs, dtype=np.uint32)
