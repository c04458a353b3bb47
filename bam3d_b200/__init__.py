"""bam3d_b200 — B200-native batched collision pipeline behind bam3d's narrow/broad-phase API.

Layout: csrc/ (CUDA kernels + the C ABI of include/bam3d_gpu.h), _ffi.py (ctypes binding), api.py / broad.py
(host-side mirror of the reference's GJK / SweepAndPrune3 / DBVT-query interface), scenes.py (seeded synthetic
scenes for the BASELINE.json configs). No CPU fallback exists: everything computes on the GPU through
libbam3d_gpu.so.
"""
from ._ffi import (Bam3dError, GjkSettings, ERR_ARG, ERR_CAPACITY, ERR_COMM, ERR_CUDA, ERR_NON_AFFINE, ERR_OOM, KIND_CAPSULE, KIND_CUBOID, KIND_CYLINDER, KIND_PARTICLE, KIND_POLYHEDRON,  # noqa: F401
                   KIND_QUAD, KIND_SPHERE, STATUS_CAPACITY, STATUS_MAX_ITER, STATUS_MISS, STATUS_NAN, STATUS_OK,
                   STATUS_REF_PANIC, STATUS_UNSUPPORTED, RAY_INTERSECTS, RAY_INTERSECTION, PARTICLE_SWEEP, FRAME_INTERSECT,
                   FRAME_CONTACT, FRAME_TIME_OF_IMPACT)
from .api import (Capsule, CollisionStrategy, Contact, Context, ConvexPolyhedron, Cube, Cuboid, Cylinder, FrameQueue, GJK,  # noqa: F401
                  MeshLibrary, Particle, Quad, Ray, ShapeSet, Sphere, default_context, intersection_transformed,
                  intersects_transformed, particle_intersection_transformed, pinned_empty, ray_cast_batch)
