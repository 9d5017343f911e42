"""ctypes view of include/robast_b200.h (structs, prototypes) for pointer-level callers."""
import ctypes as C

from . import lib

RBG_RUN, RBG_STOP, RBG_EXIT, RBG_FOCUSED, RBG_SUSPEND, RBG_ABSORB = range(6)
# shape types of the flat scene ABI (include/robast_b200.h)
RBG_SHAPE_BBOX, RBG_SHAPE_TUBE, RBG_SHAPE_SPHERE, RBG_SHAPE_PARABOLOID, RBG_SHAPE_PGON, RBG_SHAPE_PCON, RBG_SHAPE_ASPHERE = range(7)
RBG_SHAPE_WINSTON2D, RBG_SHAPE_WINSTONPOLY, RBG_SHAPE_UNION, RBG_SHAPE_INTERSECTION, RBG_SHAPE_SUBTRACTION, RBG_SHAPE_ARB8, RBG_SHAPE_XTRU = range(7, 14)
RBG_QUIRK_STEPBACK, RBG_QUIRK_BOUNDARY_PUSH = 1, 2
RBG_QUIRKS_DEFAULT = 3


class rbg_trace_opts(C.Structure):
    _fields_ = [("limit", C.c_int32), ("disable_fresnel", C.c_int32), ("quirks", C.c_uint32),
                ("steps_per_launch", C.c_int32), ("seed", C.c_uint64), ("ray_id_offset", C.c_uint64)]


_dp = C.c_void_p


class rbg_rays(C.Structure):
    _fields_ = [("n", C.c_int64), ("on_device", C.c_int32), ("pad", C.c_int32)] + \
        [(k, _dp) for k in ("x", "y", "z", "t", "dx", "dy", "dz", "lambda_",
                            "ox", "oy", "oz", "ot", "odx", "ody", "odz", "status", "last_node", "npoints")]


class rbg_history(C.Structure):
    _fields_ = [("max_points", C.c_int32), ("pad", C.c_int32)] + [(k, _dp) for k in ("hx", "hy", "hz", "ht", "hnode")]


class rbg_shoot_desc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("pad", C.c_int32),
                ("dx", C.c_double), ("dy", C.c_double), ("lambda_min", C.c_double), ("lambda_max", C.c_double),
                ("rot", C.c_double * 9), ("tr", C.c_double * 3), ("dir", C.c_double * 3), ("seed", C.c_uint64)]


def _proto(name, res, args):
    f = getattr(lib, name)
    f.restype = res
    f.argtypes = args
    return f


rbg_abi_version = _proto("rbg_abi_version", C.c_int, [])
rbg_last_error = _proto("rbg_last_error", C.c_char_p, [])
rbg_device_count = _proto("rbg_device_count", C.c_int, [])
rbg_scene_create = _proto("rbg_scene_create", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)])
rbg_scene_destroy = _proto("rbg_scene_destroy", C.c_int, [C.c_void_p])
rbg_scene_num_nodes = _proto("rbg_scene_num_nodes", C.c_int, [C.c_void_p])
rbg_scene_node_name = _proto("rbg_scene_node_name", C.c_char_p, [C.c_void_p, C.c_int])
rbg_scene_kernel_variant = _proto("rbg_scene_kernel_variant", C.c_char_p, [C.c_void_p])
rbg_trace = _proto("rbg_trace", C.c_int, [C.c_void_p, C.POINTER(rbg_trace_opts), C.POINTER(rbg_rays), C.c_void_p])
rbg_trace_history = _proto("rbg_trace_history", C.c_int, [C.c_void_p, C.POINTER(rbg_trace_opts), C.POINTER(rbg_rays), C.POINTER(rbg_history), C.c_void_p])
rbg_launch_count = _proto("rbg_launch_count", C.c_int64, [])
rbg_profile_enable = _proto("rbg_profile_enable", C.c_int, [C.c_int])
rbg_profile_read = _proto("rbg_profile_read", C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_int64), C.POINTER(C.c_double), C.POINTER(C.c_int64)])
rbg_shoot = _proto("rbg_shoot", C.c_int, [C.POINTER(rbg_shoot_desc), C.c_int64, C.c_int64] + [_dp] * 8 + [C.c_int, C.c_void_p])
class rbg_bunches(C.Structure):
    _fields_ = [("nbunches", C.c_int64)] + [(k, C.POINTER(C.c_float)) for k in ("x", "y", "time", "cx", "cy", "cz", "lambda_", "photons")] + \
        [(k, C.c_double) for k in ("z", "telescope_z", "refractive_index", "lambda_min_nm", "lambda_max_nm")] + [("seed", C.c_uint64)]


rbg_bunch_rays = _proto("rbg_bunch_rays", C.c_int, [C.POINTER(rbg_bunches), C.POINTER(C.c_int64)])
rbg_shoot_bunches = _proto("rbg_shoot_bunches", C.c_int, [C.POINTER(rbg_bunches), C.c_int64, C.c_int64] + [_dp] * 8 + [C.c_int, C.c_void_p])
rbg_hist2d = _proto("rbg_hist2d", C.c_int, [C.c_int64, _dp, _dp, _dp, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_double, C.c_double, _dp, C.c_int, C.c_void_p])
rbg_hist2d_stats = _proto("rbg_hist2d_stats", C.c_int, [C.c_int64, _dp, _dp, _dp, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_double, C.c_double, _dp, _dp, C.c_int, C.c_void_p])
rbg_containment_radius = _proto("rbg_containment_radius", C.c_int, [C.c_int32, _dp, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_double, C.c_double, _dp, C.c_double, _dp, C.c_int, C.c_void_p])
rbg_containment_radius_host = _proto("rbg_containment_radius_host", C.c_int, [_dp, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_double, C.c_double, _dp, C.c_double, _dp, C.c_int])
rbg_moments = _proto("rbg_moments", C.c_int, [C.c_int64, _dp, _dp, _dp, _dp, C.c_int32, _dp, _dp, C.c_int, C.c_void_p])
rbg_tmm = _proto("rbg_tmm", C.c_int, [C.c_void_p, C.c_int, C.c_int64, _dp, _dp, _dp, _dp, C.c_void_p])
rbg_tmm_general_host = _proto("rbg_tmm_general_host", C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64, _dp, _dp, _dp, _dp, _dp])
rbg_tmm_host = _proto("rbg_tmm_host", C.c_int, [C.c_void_p, C.c_int, C.c_int64, _dp, _dp, _dp, _dp])
rbg_multi_create = _proto("rbg_multi_create", C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)])
rbg_multi_destroy = _proto("rbg_multi_destroy", C.c_int, [C.c_void_p])
rbg_multi_num_devices = _proto("rbg_multi_num_devices", C.c_int, [C.c_void_p])
rbg_multi_trace = _proto("rbg_multi_trace", C.c_int, [C.c_void_p, C.POINTER(rbg_trace_opts), C.POINTER(rbg_rays)])
rbg_multi_shoot_trace_reduce = _proto("rbg_multi_shoot_trace_reduce", C.c_int, [C.c_void_p, C.POINTER(rbg_trace_opts), C.POINTER(rbg_shoot_desc), C.c_int64, C.c_int64, C.c_int32,
                                                                              C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_double, C.c_double, _dp, _dp, _dp])

ABI_SYMBOLS = ["rbg_abi_version", "rbg_last_error", "rbg_device_count", "rbg_scene_create", "rbg_scene_destroy",
               "rbg_scene_num_nodes", "rbg_scene_node_name", "rbg_scene_kernel_variant", "rbg_trace", "rbg_trace_history", "rbg_launch_count", "rbg_profile_enable",
               "rbg_profile_read", "rbg_shoot", "rbg_bunch_rays", "rbg_shoot_bunches", "rbg_hist2d", "rbg_hist2d_stats", "rbg_containment_radius", "rbg_containment_radius_host", "rbg_moments", "rbg_tmm", "rbg_tmm_host", "rbg_tmm_general_host",
               "rbg_multi_create", "rbg_multi_destroy", "rbg_multi_num_devices", "rbg_multi_trace", "rbg_multi_shoot_trace_reduce"]


class RbgError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise RbgError("rbg error %d: %s" % (rc, rbg_last_error().decode()))
