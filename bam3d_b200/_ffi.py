"""ctypes binding of the C ABI declared in include/bam3d_gpu.h (libbam3d_gpu.so, built by bam3d_b200/build.py).

There is no CPU fallback: if the CUDA library is missing or no device is present, loading / context creation
raises. Nothing under oracle/ is ever imported from here.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BAM3D_LIB: development knob to load a differently-built copy of the same library (kernel A/B timing)
LIB_PATH = os.environ.get("BAM3D_LIB") or os.path.join(_HERE, "libbam3d_gpu.so")

OK = 0
ERR_CUDA, ERR_ARG, ERR_NON_AFFINE, ERR_CAPACITY, ERR_OOM, ERR_COMM = -1, -2, -3, -4, -5, -6
XFORM_MAT4, XFORM_AFFINE3X4 = 0, 1   # transform layouts of bam3d_frames_set_transform_layout / *_set_transforms_affine
UNIQUE_ID_BYTES = 256

KIND_PARTICLE, KIND_QUAD, KIND_SPHERE, KIND_CUBOID, KIND_CYLINDER, KIND_CAPSULE, KIND_POLYHEDRON = range(7)
STATUS_OK, STATUS_MISS, STATUS_MAX_ITER, STATUS_REF_PANIC, STATUS_NAN, STATUS_CAPACITY, STATUS_UNSUPPORTED = range(7)
RAY_INTERSECTS, RAY_INTERSECTION, PARTICLE_SWEEP = 0, 1, 2
FULL_RESOLUTION, COLLISION_ONLY = 0, 1
FRAME_INTERSECT, FRAME_CONTACT, FRAME_TIME_OF_IMPACT = 0, 1, 2


class GjkSettings(C.Structure):
    _fields_ = [("distance_tolerance", C.c_float), ("continuous_tolerance", C.c_float), ("epa_tolerance", C.c_float),
                ("max_iterations", C.c_uint32), ("toi_iteration_cap", C.c_uint32)]


class Bam3dError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"bam3d error {code}: {msg}")
        self.code = code


_vp, _u64, _i32, _u32 = C.c_void_p, C.c_uint64, C.c_int32, C.c_uint32
_PS = C.POINTER(GjkSettings)

# name -> (restype, argtypes); every symbol include/bam3d_gpu.h declares
SIGNATURES = {
    "bam3d_ctx_create": (_i32, [_i32, C.POINTER(_vp)]),
    "bam3d_ctx_destroy": (None, [_vp]),
    "bam3d_last_error": (C.c_char_p, [_vp]),
    "bam3d_ctx_synchronize": (_i32, [_vp]),
    "bam3d_ctx_stream": (_vp, [_vp]),
    "bam3d_default_settings": (None, [_PS]),
    "bam3d_ctx_set_option": (_i32, [_vp, C.c_char_p, _i32]),
    "bam3d_ctx_read_timing": (_i32, [_vp, C.POINTER(C.c_double), C.POINTER(_u64)]),
    "bam3d_host_alloc": (_i32, [_vp, _u64, C.POINTER(_vp)]),
    "bam3d_host_free": (None, [_vp, _vp]),
    "bam3d_meshes_upload": (_i32, [_vp, _vp, _vp, _u32, C.POINTER(_vp)]),
    "bam3d_meshes_upload_faces": (_i32, [_vp, _vp, _vp, _u32, _vp, _vp, C.POINTER(_vp)]),
    "bam3d_meshes_free": (None, [_vp, _vp]),
    "bam3d_shapes_upload": (_i32, [_vp, _vp, _vp, _vp, _vp, _u32, _vp, C.POINTER(_vp)]),
    "bam3d_shapes_set_transforms": (_i32, [_vp, _vp, _vp]),
    "bam3d_shapes_set_transforms_affine": (_i32, [_vp, _vp, _vp]),
    "bam3d_frames_set_transform_layout": (_i32, [_vp, _vp, _i32]),
    "bam3d_shapes_free": (None, [_vp, _vp]),
    "bam3d_shapes_count": (_u32, [_vp]),
    "bam3d_gjk_intersect": (_i32, [_vp, _vp, _vp, _u64, _PS, _vp]),
    "bam3d_gjk_contact": (_i32, [_vp, _vp, _vp, _u64, _PS, _i32, _vp, _vp]),
    "bam3d_gjk_distance": (_i32, [_vp, _vp, _vp, _u64, _PS, _vp, _vp, _vp]),
    "bam3d_gjk_time_of_impact": (_i32, [_vp, _vp, _vp, _vp, _u64, _PS, _vp, _vp]),
    "bam3d_gjk_intersect_dev": (_i32, [_vp, _vp, _vp, _u64, _PS, _vp]),
    "bam3d_gjk_contact_dev": (_i32, [_vp, _vp, _vp, _u64, _PS, _i32, _vp, _vp]),
    "bam3d_gjk_distance_dev": (_i32, [_vp, _vp, _vp, _u64, _PS, _vp, _vp, _vp]),
    "bam3d_gjk_time_of_impact_dev": (_i32, [_vp, _vp, _vp, _vp, _u64, _PS, _vp, _vp]),
    "bam3d_compounds_upload": (_i32, [_vp, _vp, _vp, _vp, _vp, _u32, _vp, _u32, _vp, C.POINTER(_vp)]),
    "bam3d_compounds_free": (None, [_vp, _vp]),
    "bam3d_gjk_contact_complex": (_i32, [_vp, _vp, _vp, _vp, _u64, _PS, _i32, _vp, _vp]),
    "bam3d_gjk_distance_complex": (_i32, [_vp, _vp, _vp, _vp, _u64, _PS, _vp, _vp]),
    "bam3d_gjk_time_of_impact_complex": (_i32, [_vp, _vp, _vp, _vp, _vp, _u64, _PS, _i32, _vp, _vp]),
    "bam3d_ray_cast": (_i32, [_vp, _vp, _vp, _u64, _vp, _u64, _i32, _vp, _vp]),
    "bam3d_ray_cast_dev": (_i32, [_vp, _vp, _vp, _u64, _vp, _u64, _i32, _vp, _vp]),
    "bam3d_compact_contacts_dev": (_i32, [_vp, _vp, _vp, _u64, _u64, _vp, _u64, _vp]),
    "bam3d_ctx_launch_count": (_u64, [_vp]),
    "bam3d_sap_find_pairs": (_i32, [_vp, _vp, _u64, C.POINTER(_i32), _vp, _vp, _u64, C.POINTER(_u64)]),
    "bam3d_sap_find_pairs_dev": (_i32, [_vp, _vp, _u64, C.POINTER(_i32), _vp, _vp, _u64, C.POINTER(_u64)]),
    "bam3d_sap_find_pairs_shard": (_i32, [_vp, _vp, _u64, C.POINTER(_i32), _vp, _vp, _u64, C.POINTER(_u64), _u32, _u32]),
    "bam3d_sap_find_pairs_shard_dev": (_i32, [_vp, _vp, _u64, C.POINTER(_i32), _vp, _vp, _u64, C.POINTER(_u64), _u32, _u32]),
    "bam3d_sap_variance": (_i32, [_vp, _vp, _u64, _vp, C.POINTER(_i32)]),
    "bam3d_shapes_world_aabbs": (_i32, [_vp, _vp, _vp]),
    "bam3d_shapes_world_aabbs_dev": (_i32, [_vp, _vp, _vp]),
    "bam3d_shapes_world_spheres": (_i32, [_vp, _vp, _vp]),
    "bam3d_shapes_world_spheres_dev": (_i32, [_vp, _vp, _vp]),
    "bam3d_sap_find_pairs_spheres": (_i32, [_vp, _vp, _u64, C.POINTER(_i32), _vp, _vp, _u64, C.POINTER(_u64)]),
    "bam3d_sap_find_pairs_spheres_dev": (_i32, [_vp, _vp, _u64, C.POINTER(_i32), _vp, _vp, _u64, C.POINTER(_u64)]),
    "bam3d_pipeline_contacts": (_i32, [_vp, _vp, C.POINTER(_i32), _PS, _i32, _vp, _vp, _vp, _vp, _vp, _u64, C.POINTER(_u64)]),
    "bam3d_pipeline_contacts_dev": (_i32, [_vp, _vp, C.POINTER(_i32), _PS, _i32, _u32, _u32, _vp, _vp, _vp, _vp, _vp, _u64, C.POINTER(_u64)]),
    "bam3d_bvh_build": (_i32, [_vp, _vp, _u64, C.POINTER(_vp)]),
    "bam3d_bvh_build_dev": (_i32, [_vp, _vp, _u64, C.POINTER(_vp)]),
    "bam3d_bvh_refit": (_i32, [_vp, _vp, _vp]),
    "bam3d_bvh_refit_dev": (_i32, [_vp, _vp, _vp]),
    "bam3d_bvh_free": (None, [_vp, _vp]),
    "bam3d_bvh_count": (_u32, [_vp]),
    "bam3d_bvh_query_aabb": (_i32, [_vp, _vp, _vp, _u64, _vp, _vp, _u64, C.POINTER(_u64)]),
    "bam3d_bvh_build_spheres": (_i32, [_vp, _vp, _u64, C.POINTER(_vp)]),
    "bam3d_bvh_build_spheres_dev": (_i32, [_vp, _vp, _u64, C.POINTER(_vp)]),
    "bam3d_bvh_query_sphere": (_i32, [_vp, _vp, _vp, _u64, _vp, _vp, _u64, C.POINTER(_u64)]),
    "bam3d_bvh_query_sphere_dev": (_i32, [_vp, _vp, _vp, _u64, _vp, _vp, _u64, _vp]),
    "bam3d_bvh_query_ray": (_i32, [_vp, _vp, _vp, _u64, _vp, _vp, _vp, _u64, C.POINTER(_u64)]),
    "bam3d_bvh_query_frustum": (_i32, [_vp, _vp, _vp, _u64, _vp, _vp, _vp, _u64, C.POINTER(_u64)]),
    "bam3d_bvh_query_ray_closest": (_i32, [_vp, _vp, _vp, _u64, _vp, _vp]),
    "bam3d_bvh_find_collider_pairs": (_i32, [_vp, _vp, _vp, _vp, _u64, C.POINTER(_u64)]),
    "bam3d_bvh_query_aabb_dev": (_i32, [_vp, _vp, _vp, _u64, _vp, _vp, _u64, _vp]),
    "bam3d_bvh_query_ray_dev": (_i32, [_vp, _vp, _vp, _u64, _vp, _vp, _vp, _u64, _vp]),
    "bam3d_bvh_query_frustum_dev": (_i32, [_vp, _vp, _vp, _u64, _vp, _vp, _vp, _u64, _vp]),
    "bam3d_bvh_query_ray_closest_dev": (_i32, [_vp, _vp, _vp, _u64, _vp, _vp]),
    "bam3d_frustum_from_mat4": (_i32, [_vp, _vp]),
    "bam3d_frames_create": (_i32, [_vp, _vp, _vp, _vp, _vp, _u32, _vp, _u64, _u32, _i32, C.POINTER(_vp)]),
    "bam3d_frames_submit": (_i32, [_vp, _vp, _i32, _vp, _vp, _vp, _u64, _PS, _i32, _vp, _vp, C.POINTER(_u64)]),
    "bam3d_frames_wait": (_i32, [_vp, _vp, _u64]),
    "bam3d_frames_free": (None, [_vp, _vp]),
    "bam3d_frames_submit_gather": (_i32, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _u64, _PS, _i32, _u64, _vp, _u64, _vp, C.POINTER(_u64)]),
    "bam3d_frames_gathered": (_i32, [_vp, _vp, _u64, _vp, C.POINTER(_u64)]),
    "bam3d_comm_unique_id": (_i32, [_vp]),
    "bam3d_comm_create": (_i32, [_i32, _vp, _i32, _i32, C.POINTER(_vp)]),
    "bam3d_comm_destroy": (None, [_vp]),
    "bam3d_comm_rank": (_i32, [_vp]),
    "bam3d_comm_size": (_i32, [_vp]),
    "bam3d_comm_stream": (_vp, [_vp]),
    "bam3d_comm_unavailable_reason": (C.c_char_p, []),
    "bam3d_gather_contacts_begin": (_i32, [_vp, _vp, _vp, C.POINTER(_u64)]),
    "bam3d_gather_contacts_finish": (_i32, [_vp, _vp, _u64, _vp, _vp, _u64, _vp, C.POINTER(_u64)]),
    "bam3d_comm_join": (_i32, [_vp, _vp]),
    "bam3d_gather_contacts_wait": (_i32, [_vp, _vp, _u64]),
    "bam3d_gather_contacts_dev": (_i32, [_vp, _vp, _vp, _vp, _vp, _u64, _vp, C.POINTER(_u64)]),
    "bam3d_gather_contacts": (_i32, [_vp, _vp, _vp, _vp, _u64, _u64, _vp, _u64, _vp, C.POINTER(_u64)]),
}

_lib = None


def lib():
    """Load libbam3d_gpu.so (once). Raises if it has not been built — the product path has no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python bam3d_b200/build.py` "
                              "(nvcc, sm_100a). bam3d_b200 has no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def register(signatures):
    """Add more entry points (broad phase) to the table before the library is loaded."""
    SIGNATURES.update(signatures)
    if _lib is not None:
        for name, (res, args) in signatures.items():
            fn = getattr(_lib, name)
            fn.restype, fn.argtypes = res, args
