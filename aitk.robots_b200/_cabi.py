# -*- coding: utf-8 -*-
"""ctypes binding of libbrs.so (include/brs.h).

The library is loaded lazily and loudly: if it is missing or a symbol from
``include/brs.h`` is absent the import of any compute path raises.  There is
no CPU fallback anywhere in this package (BASELINE.json north_star).
"""

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libbrs.so")
if os.environ.get("BRS_LIB"):  # tuning builds: another libbrs.so of the same ABI
    LIB_PATH = os.path.abspath(os.environ["BRS_LIB"])

BRS_ABI_VERSION = 3
BRS_MAX_ROBOTS = 8

# brs_buffer
(
    BUF_SEG,
    BUF_SEG_COUNT,
    BUF_WORLD_SIZE,
    BUF_BULBS,
    BUF_GRID,
    BUF_POSE,
    BUF_VEL,
    BUF_TARGET_VEL,
    BUF_STALLED,
    BUF_RANGE,
    BUF_CAMERA,
    BUF_LIGHT,
    BUF_SMELL,
    BUF_COUNT,
) = range(14)

STEP_MOVE, STEP_SENSE, STEP_CAMERA, STEP_ALL = 1, 2, 4, 7


class RobotDesc(C.Structure):
    _fields_ = [
        ("bbox", C.c_double * 4),
        ("height", C.c_double),
        ("vx_ramp", C.c_double),
        ("vy_ramp", C.c_double),
        ("va_ramp", C.c_double),
        ("rgba", C.c_uint8 * 4),
        ("_pad", C.c_int32),
    ]


class RangeDesc(C.Structure):
    _fields_ = [
        ("robot", C.c_int32),
        ("n_rays", C.c_int32),
        ("dist_from_center", C.c_double),
        ("dir_from_center", C.c_double),
        ("a", C.c_double),
        ("max", C.c_double),
        ("width", C.c_double),
    ]


class CameraDesc(C.Structure):
    _fields_ = [
        ("robot", C.c_int32),
        ("width", C.c_int32),
        ("height", C.c_int32),
        ("_pad", C.c_int32),
        ("dist_from_center", C.c_double),
        ("dir_from_center", C.c_double),
        ("fov", C.c_double),
        ("max_range", C.c_double),
        ("color_fade", C.c_double),
        ("size_fade", C.c_double),
    ]


class PointDesc(C.Structure):
    _fields_ = [
        ("robot", C.c_int32),
        ("_pad", C.c_int32),
        ("dist_from_center", C.c_double),
        ("dir_from_center", C.c_double),
    ]


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32),
        ("device", C.c_int32),
        ("n_worlds", C.c_int32),
        ("n_robots", C.c_int32),
        ("seg_cap", C.c_int32),
        ("n_bulbs", C.c_int32),
        ("grid_w", C.c_int32),
        ("grid_h", C.c_int32),
        ("n_range", C.c_int32),
        ("n_camera", C.c_int32),
        ("n_light", C.c_int32),
        ("n_smell", C.c_int32),
        ("smell_cell_size", C.c_double),
        ("light_multiplier", C.c_double),
        ("sky_rgb", C.c_double * 3),
        ("sky_grad", C.c_double * 3),
        ("ground_rgb", C.c_double * 3),
        ("ground_grad", C.c_double * 3),
        ("robots", C.POINTER(RobotDesc)),
        ("ranges", C.POINTER(RangeDesc)),
        ("cameras", C.POINTER(CameraDesc)),
        ("lights", C.POINTER(PointDesc)),
        ("smells", C.POINTER(PointDesc)),
    ]


# every symbol include/brs.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("brs_abi_version", C.c_int, []),
    ("brs_last_error", C.c_char_p, []),
    ("brs_device_count", C.c_int, []),
    ("brs_create", C.c_int, [C.POINTER(Config), C.POINTER(_P)]),
    ("brs_destroy", C.c_int, [_P]),
    ("brs_device_ptr", C.c_int, [_P, C.c_int, C.POINTER(_P), C.POINTER(C.c_size_t)]),
    ("brs_bind_output", C.c_int, [_P, C.c_int, _P]),
    ("brs_upload", C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, _P]),
    ("brs_download", C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, _P]),
    ("brs_step", C.c_int, [_P, C.c_double, C.c_uint, _P]),
    ("brs_bind_peer_outputs", C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    ("brs_peer_alloc", C.c_int, [_P, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_ubyte)]),
    ("brs_peer_free", C.c_int, [_P, _P]),
    ("brs_peer_open", C.c_int, [_P, C.POINTER(C.c_ubyte), C.POINTER(C.c_void_p)]),
    ("brs_peer_close", C.c_int, [_P, _P]),
    ("brs_set_noise", C.c_int, [_P, C.c_ulonglong, C.c_uint, C.c_double, C.c_double]),
    ("brs_step_counted", C.c_int, [_P, C.c_double, C.c_uint, _P, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]),
    ("brs_steps", C.c_int, [_P, C.c_double, C.c_uint, C.c_int, _P]),
    ("brs_steps_launches", C.c_int, [_P, C.c_uint, C.c_int, C.POINTER(C.c_int)]),
    ("brs_steps_info", C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("brs_step_host", C.c_int, [_P, C.c_double, C.c_uint, _P, _P, _P, _P, _P, _P, _P, _P]),
    ("brs_host_alloc", C.c_int, [_P, C.c_size_t, C.POINTER(C.c_void_p)]),
    ("brs_host_free", C.c_int, [_P, _P]),
    ("brs_render_world", C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, _P, _P]),
    ("brs_synchronize", C.c_int, [_P, _P]),
    ("brs_count_tests", C.c_int, [_P, C.c_uint, _P, C.POINTER(C.c_ulonglong)]),
    ("brs_launch_info", C.c_int, [_P, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    ("brs_tile_info", C.c_int, [_P] + [C.POINTER(C.c_int)] * 5),
    ("brs_path_info", C.c_int, [_P] + [C.POINTER(C.c_int)] * 2),
    ("brs_measure_write_bw", C.c_int, [C.c_int, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    ("brs_measure_fp32_peak", C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
]

_lib = None


class BrsError(RuntimeError):
    pass


def load():
    """Load libbrs.so and bind every declared symbol; raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BrsError(
            "libbrs.so not built at %s -- run `python -c 'import __graft_entry__ as g; g.build()'`. "
            "There is no CPU fallback." % LIB_PATH
        )
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the export is missing
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.brs_abi_version() != BRS_ABI_VERSION:
        raise BrsError("libbrs.so ABI version %d != %d" % (lib.brs_abi_version(), BRS_ABI_VERSION))
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().brs_last_error()
        raise BrsError("libbrs error %d: %s" % (rc, msg.decode() if msg else "?"))
