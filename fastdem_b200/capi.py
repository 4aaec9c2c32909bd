"""ctypes declarations for include/fastdem_b200.h (libfastdem_b200.so).

This is the whole Python<->native seam: plain pointers and sizes.  The library is the
product; if it is missing it is built from source (nvcc), and if that fails importing the
compute path fails loudly — there is no Python/CPU fallback."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libfastdem_b200.so"

FDEM_OK = 0
SENSOR_CONSTANT, SENSOR_LIDAR, SENSOR_RGBD = 0, 1, 2
MODE_LOCAL, MODE_GLOBAL = 0, 1
EST_KALMAN, EST_P2QUANTILE = 0, 1
MOVE_CLEAR_ALL_LAYERS, MOVE_CLEAR_BASIC_LAYERS = 0, 1
VOXEL_SORT_MSD, VOXEL_SORT_LIBRARY = 0, 1
CELL_SORT_TILE, CELL_SORT_GLOBAL = 0, 1
STAGE_NAMES = ["h2d", "preprocess_bin", "commit_move_clear", "sort_by_cell", "segreduce_estimate",
               "voxel_raycast"]


class FdemConfig(C.Structure):
    """struct fdem_config == fastdem::Config flattened (include/fastdem_b200.h)."""
    _fields_ = [
        ("z_min", C.c_float), ("z_max", C.c_float), ("range_min", C.c_float), ("range_max", C.c_float),
        ("sensor_type", C.c_int32),
        ("lidar_range_noise", C.c_float), ("lidar_angular_noise", C.c_float),
        ("rgbd_normal_a", C.c_float), ("rgbd_normal_b", C.c_float), ("rgbd_normal_c", C.c_float),
        ("rgbd_lateral_factor", C.c_float),
        ("constant_uncertainty", C.c_float),
        ("mode", C.c_int32), ("estimation_type", C.c_int32),
        ("kalman_min_variance", C.c_float), ("kalman_max_variance", C.c_float),
        ("kalman_process_noise", C.c_float),
        ("p2_dn", C.c_float * 5), ("p2_elevation_marker", C.c_int32), ("p2_max_sample_count", C.c_float),
        ("raycasting_enabled", C.c_int32),
        ("rc_height_conflict_threshold", C.c_float), ("rc_log_odds_observed", C.c_float),
        ("rc_log_odds_ghost", C.c_float), ("rc_log_odds_max", C.c_float), ("rc_clear_threshold", C.c_float),
        ("move_clear_policy", C.c_int32),
    ]

    def copy(self) -> "FdemConfig":
        c = FdemConfig()
        C.memmove(C.byref(c), C.byref(self), C.sizeof(FdemConfig))
        return c


class FdemScanStats(C.Structure):
    _fields_ = [("n_input", C.c_int64), ("n_kept", C.c_int64), ("n_cells", C.c_int64),
                ("n_voxels", C.c_int64), ("integrated", C.c_int32), ("voxel_box_violations", C.c_int32)]


class FdemPointCloud2Layout(C.Structure):
    _fields_ = [("point_step", C.c_uint32), ("off_x", C.c_int32), ("off_y", C.c_int32),
                ("off_z", C.c_int32), ("off_intensity", C.c_int32), ("intensity_type", C.c_int32),
                ("off_rgb", C.c_int32)]


class FdemGeometry(C.Structure):
    _fields_ = [("rows", C.c_int32), ("cols", C.c_int32), ("resolution", C.c_double),
                ("length", C.c_double * 2), ("position", C.c_double * 2),
                ("start_index", C.c_int32 * 2), ("row_begin", C.c_int32), ("row_end", C.c_int32)]


_P = C.c_void_p
_ST = C.c_int32
class FdemIpcHandle(C.Structure):
    _fields_ = [("bytes", C.c_uint8 * 64), ("size", C.c_uint64)]


_f32p, _u8p, _f64p = C.c_void_p, C.c_void_p, C.c_void_p  # addresses are passed as plain ints

# name -> (restype, argtypes); every symbol include/fastdem_b200.h declares
SIGNATURES = {
    "fdem_abi_version": (C.c_int32, []),
    "fdem_last_error": (C.c_char_p, []),
    "fdem_status_string": (C.c_char_p, [_ST]),
    "fdem_config_default": (None, [C.POINTER(FdemConfig)]),
    "fdem_map_create": (_ST, [C.c_float, C.c_float, C.c_float, C.c_int32, _P, C.POINTER(_P)]),
    "fdem_map_create_stripe": (_ST, [C.c_float, C.c_float, C.c_float, C.c_int32, C.c_int32,
                                     C.c_int32, _P, C.POINTER(_P)]),
    "fdem_map_set_geometry": (_ST, [_P, C.c_float, C.c_float, C.c_float]),
    "fdem_config_validate": (_ST, [C.POINTER(FdemConfig), C.POINTER(C.c_int32)]),
    "fdem_map_destroy": (_ST, [_P]),
    "fdem_map_get_geometry": (_ST, [_P, C.POINTER(FdemGeometry)]),
    "fdem_map_set_position": (_ST, [_P, C.c_double, C.c_double]),
    "fdem_map_set_start_index": (_ST, [_P, C.c_int32, C.c_int32]),
    "fdem_map_move": (_ST, [_P, C.c_double, C.c_double, C.c_int32, C.POINTER(C.c_int32)]),
    "fdem_map_is_inside": (_ST, [_P, C.c_double, C.c_double, C.POINTER(C.c_int32)]),
    "fdem_map_get_index": (_ST, [_P, C.c_double, C.c_double, C.POINTER(C.c_int32),
                                 C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "fdem_map_get_cell_position": (_ST, [_P, C.c_int32, C.c_int32, _f64p, _f64p]),
    "fdem_map_layer_exists": (_ST, [_P, C.c_char_p, C.POINTER(C.c_int32)]),
    "fdem_map_layer_add": (_ST, [_P, C.c_char_p, C.c_float]),
    "fdem_map_layer_count": (_ST, [_P, C.POINTER(C.c_int32)]),
    "fdem_map_layer_name": (_ST, [_P, C.c_int32, C.c_char_p, C.c_int32]),
    "fdem_map_layer_download": (_ST, [_P, C.c_char_p, _f32p]),
    "fdem_map_layer_upload": (_ST, [_P, C.c_char_p, _f32p]),
    "fdem_map_layer_device_ptr": (_ST, [_P, C.c_char_p, C.POINTER(_P)]),
    "fdem_map_cell_get": (_ST, [_P, C.c_char_p, C.c_int32, C.c_int32, C.POINTER(C.c_float)]),
    "fdem_map_cell_set": (_ST, [_P, C.c_char_p, C.c_int32, C.c_int32, C.c_float]),
    "fdem_map_clear": (_ST, [_P, C.c_char_p]),
    "fdem_map_clear_all": (_ST, [_P]),
    "fdem_map_clear_at": (_ST, [_P, C.c_int32, C.c_int32]),
    "fdem_map_is_empty": (_ST, [_P, C.POINTER(C.c_int32)]),
    "fdem_map_sync": (_ST, [_P]),
    "fdem_map_stream": (_P, [_P]),
    "fdem_mapper_create": (_ST, [_P, C.POINTER(FdemConfig), C.POINTER(_P)]),
    "fdem_mapper_destroy": (_ST, [_P]),
    "fdem_mapper_set_config": (_ST, [_P, C.POINTER(FdemConfig)]),
    "fdem_mapper_get_config": (_ST, [_P, C.POINTER(FdemConfig)]),
    "fdem_mapper_integrate": (_ST, [_P, _f32p, _f32p, _u8p, C.c_size_t, _f64p, _f64p,
                                    C.POINTER(FdemScanStats)]),
    "fdem_mapper_integrate_async": (_ST, [_P, _f32p, _f32p, _u8p, C.c_size_t, _f64p, _f64p]),
    "fdem_mapper_integrate_pointcloud2": (_ST, [_P, _u8p, C.c_size_t, C.POINTER(FdemPointCloud2Layout),
                                                _f64p, _f64p, C.POINTER(FdemScanStats)]),
    "fdem_mapper_submit_pointcloud2": (_ST, [_P, _u8p, C.c_size_t, C.POINTER(FdemPointCloud2Layout), _f64p, _f64p,
                                             C.POINTER(C.c_uint64)]),
    "fdem_mapper_wait": (_ST, [_P, C.POINTER(FdemScanStats)]),
    "fdem_mapper_integrate_batch": (_ST, [_P, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p]),
    "fdem_mapper_last_batch_stats": (_ST, [_P, C.c_void_p, C.c_int32]),
    "fdem_mapper_submit": (_ST, [_P, _f32p, _f32p, _u8p, C.c_size_t, _f64p, _f64p, C.POINTER(C.c_uint64)]),
    "fdem_mapper_collect": (_ST, [_P, C.c_uint64, C.POINTER(FdemScanStats)]),
    "fdem_mapper_update": (_ST, [_P, _f32p, _f32p, _f32p, _u8p, C.c_size_t, C.c_double, C.c_double,
                                 C.POINTER(FdemScanStats)]),
    "fdem_mapper_integrate_with_cov": (_ST, [_P, _f32p, _f32p, _f32p, _u8p, C.c_size_t, _f64p, _f64p,
                                             C.POINTER(FdemScanStats)]),
    "fdem_mapper_last_preprocessed": (_ST, [_P, _f32p, _f32p, C.c_void_p, C.POINTER(C.c_int64)]),
    "fdem_mapper_last_rasterized": (_ST, [_P, _f32p, C.POINTER(C.c_int64)]),
    "fdem_raycast": (_ST, [_P, _f32p, C.c_size_t, C.POINTER(C.c_float), C.POINTER(FdemConfig)]),
    "fdem_voxel_grid_any": (_ST, [_P, _f32p, C.c_size_t, C.c_float, C.c_void_p, C.POINTER(C.c_int64)]),
    "fdem_inpaint": (_ST, [_P, C.c_int32, C.c_int32, C.c_int32]),
    "fdem_inpaint_stripe_sweep": (_ST, [_P, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "fdem_spatial_smoothing": (_ST, [_P, C.c_char_p, C.c_int32, C.c_int32]),
    "fdem_map_pack_pointcloud2": (_ST, [_P, C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]),
    "fdem_map_pointcloud2_field": (_ST, [_P, C.c_int32, C.c_char_p, C.c_int32, C.POINTER(C.c_uint32)]),
    "fdem_map_pointcloud2_data": (_ST, [_P, C.c_void_p, C.POINTER(C.c_void_p)]),
    "fdem_bind_thread_to_device": (_ST, [C.c_int32, C.POINTER(C.c_int32)]),
    "fdem_device_alloc": (_ST, [C.c_int32, C.c_size_t, C.POINTER(C.c_void_p)]),
    "fdem_device_free": (_ST, [C.c_int32, C.c_void_p]),
    "fdem_ipc_export": (_ST, [C.c_int32, C.c_void_p, C.c_size_t, C.POINTER(FdemIpcHandle)]),
    "fdem_ipc_import": (_ST, [C.c_int32, C.POINTER(FdemIpcHandle), C.POINTER(C.c_void_p)]),
    "fdem_ipc_close": (_ST, [C.c_int32, C.c_void_p]),
    "fdem_shard_create": (_ST, [_P, C.c_int32, C.c_int32, C.c_size_t, C.POINTER(C.c_void_p)]),
    "fdem_shard_destroy": (_ST, [_P]),
    "fdem_shard_export": (_ST, [_P, C.POINTER(FdemIpcHandle)]),
    "fdem_shard_connect": (_ST, [_P, C.c_void_p]),
    "fdem_shard_integrate": (_ST, [_P, _f32p, _f32p, _u8p, C.c_size_t, _f64p, _f64p]),
    "fdem_shard_wait": (_ST, [_P, C.POINTER(FdemScanStats)]),
    "fdem_shard_slice_plan": (_ST, [C.POINTER(C.c_uint32), C.c_int32, C.c_uint32, C.c_uint32,
                                    C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "fdem_uncertainty_fusion": (_ST, [_P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int32]),
    "fdem_feature_extraction": (_ST, [_P, C.c_float, C.c_int32, C.c_float, C.c_float]),
    "fdem_mapper_launch_count": (_ST, [_P, C.POINTER(C.c_int64)]),
    "fdem_mapper_library_launch_count": (_ST, [_P, C.POINTER(C.c_int64)]),
    "fdem_mapper_set_stage_timing": (_ST, [_P, C.c_int32]),
    "fdem_mapper_set_cell_sort": (_ST, [_P, C.c_int32]),
    "fdem_mapper_set_voxel_sort": (_ST, [_P, C.c_int32]),
    "fdem_mapper_stage_times": (_ST, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
}

_lib = None


class FdemError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"fdem status {status}: {msg}")
        self.status = status


def load_library(build_if_missing: bool = True) -> C.CDLL:
    """dlopen libfastdem_b200.so (building it with nvcc first if it is not there).  Needs no
    GPU: only calls that touch a device do."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        if not build_if_missing:
            raise FileNotFoundError(f"{LIB_PATH} is missing; run `python -m fastdem_b200._build`")
        from . import _build
        _build.build_library()
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == ABI drift; fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != FDEM_OK:
        msg = load_library().fdem_last_error()
        raise FdemError(status, msg.decode(errors="replace") if msg else "")


def default_config() -> FdemConfig:
    c = FdemConfig()
    load_library().fdem_config_default(C.byref(c))
    return c
