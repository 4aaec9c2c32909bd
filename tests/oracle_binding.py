"""ctypes binding of oracle/libfdem_oracle.so — the CPU restatement of the reference path.
TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

REPO = Path(__file__).resolve().parent.parent
ORACLE_DIR = REPO / "oracle"
LIB_PATH = ORACLE_DIR / "libfdem_oracle.so"

from fastdem_b200.capi import FdemConfig, FdemScanStats  # same field layout (tests/test_abi_layout.py)

_lib = None
_f64p = C.POINTER(C.c_double)


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        srcs = [ORACLE_DIR / "fdem_oracle.hpp", ORACLE_DIR / "fdem_oracle_capi.cpp",
                ORACLE_DIR / "fdem_oracle_io.cpp"]
        if not LIB_PATH.exists() or any(s.stat().st_mtime > LIB_PATH.stat().st_mtime for s in srcs):
            subprocess.run(["make", "-C", str(ORACLE_DIR)], check=True, stdout=subprocess.DEVNULL)
        L = C.CDLL(str(LIB_PATH))
        L.orc_map_create.restype = C.c_void_p
        L.orc_map_create.argtypes = [C.c_float, C.c_float, C.c_float]
        L.orc_map_destroy.argtypes = [C.c_void_p]
        L.orc_map_geometry.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), _f64p,
                                       _f64p, _f64p, C.POINTER(C.c_int32)]
        L.orc_map_layer_exists.argtypes = [C.c_void_p, C.c_char_p]
        L.orc_map_layer_count.argtypes = [C.c_void_p]
        L.orc_map_layer_name.argtypes = [C.c_void_p, C.c_int, C.c_char_p, C.c_int]
        L.orc_map_layer_get.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.orc_map_layer_set.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
        L.orc_map_layer_add.argtypes = [C.c_void_p, C.c_char_p, C.c_float]
        L.orc_map_clear_all.argtypes = [C.c_void_p]
        L.orc_map_is_empty.argtypes = [C.c_void_p]
        L.orc_map_is_inside.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_map_get_index.argtypes = [C.c_void_p, C.c_double, C.c_double, C.POINTER(C.c_int32),
                                        C.POINTER(C.c_int32)]
        L.orc_map_get_position.argtypes = [C.c_void_p, C.c_int32, C.c_int32, _f64p, _f64p]
        L.orc_map_move.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_int32]
        L.orc_map_set_position.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.orc_map_set_start_index.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.orc_mapper_create.restype = C.c_void_p
        L.orc_mapper_create.argtypes = [C.c_void_p, C.POINTER(FdemConfig)]
        L.orc_mapper_destroy.argtypes = [C.c_void_p]
        L.orc_mapper_integrate.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                           _f64p, _f64p, C.POINTER(FdemScanStats), _f64p]
        L.orc_mapper_update.restype = C.c_int64
        L.orc_mapper_update.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_size_t, C.c_double, C.c_double]
        L.orc_preprocess.restype = C.c_int64
        L.orc_preprocess.argtypes = [C.POINTER(FdemConfig), C.c_void_p, C.c_size_t, _f64p, _f64p,
                                     C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_sensor_cov.argtypes = [C.POINTER(FdemConfig), C.c_void_p, C.c_void_p]
        L.orc_transform.argtypes = [_f64p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_kalman_step.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float,
                                      C.c_float, C.c_void_p]
        L.orc_p2_step.restype = C.c_float
        L.orc_p2_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_int,
                                  C.c_float]
        L.orc_voxel_any.restype = C.c_int64
        L.orc_voxel_any.argtypes = [C.c_void_p, C.c_size_t, C.c_float, C.c_void_p]
        L.orc_raycast.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(FdemConfig)]
        L.orc_inpaint.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_pack_color.restype = C.c_float
        L.orc_pack_color.argtypes = [C.c_uint8, C.c_uint8, C.c_uint8]
        L.orc_config_default.argtypes = [C.POINTER(FdemConfig)]
        _lib = L
    return _lib


def default_config() -> FdemConfig:
    c = FdemConfig()
    lib().orc_config_default(C.byref(c))
    return c


def _colmajor16(T) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(T, np.float64).T).reshape(16)


def _xyzw(points) -> np.ndarray:
    a = np.asarray(points, np.float32)
    if a.ndim == 1:
        a = a.reshape(-1, 3)
    if a.shape[1] == 3:
        a = np.concatenate([a, np.ones((a.shape[0], 1), np.float32)], axis=1)
    return np.ascontiguousarray(a)


class OracleMap:
    def __init__(self, width, height, resolution):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_map_create(width, height, resolution))

    def __del__(self):
        try:
            self.L.orc_map_destroy(self.h)
        except Exception:
            pass

    def geometry(self):
        r, c = C.c_int32(), C.c_int32()
        res = C.c_double()
        ln, pos = (C.c_double * 2)(), (C.c_double * 2)()
        st = (C.c_int32 * 2)()
        self.L.orc_map_geometry(self.h, C.byref(r), C.byref(c), C.byref(res), ln, pos, st)
        return dict(rows=r.value, cols=c.value, resolution=res.value, length=tuple(ln),
                    position=tuple(pos), start_index=tuple(st))

    def exists(self, name):
        return bool(self.L.orc_map_layer_exists(self.h, name.encode()))

    def layers(self):
        n = self.L.orc_map_layer_count(self.h)
        out = []
        buf = C.create_string_buffer(128)
        for i in range(n):
            self.L.orc_map_layer_name(self.h, i, buf, 128)
            out.append(buf.value.decode())
        return out

    def get(self, name):
        g = self.geometry()
        a = np.empty((g["rows"], g["cols"]), np.float32, order="F")
        if self.L.orc_map_layer_get(self.h, name.encode(), a.ctypes.data) != 0:
            raise KeyError(name)
        return a

    def set(self, name, values):
        a = np.asfortranarray(np.asarray(values, np.float32))
        self.L.orc_map_layer_set(self.h, name.encode(), a.ctypes.data)

    def add(self, name, fill=float("nan")):
        self.L.orc_map_layer_add(self.h, name.encode(), fill)

    def clearAll(self):
        self.L.orc_map_clear_all(self.h)

    def isEmpty(self):
        return bool(self.L.orc_map_is_empty(self.h))

    def isInside(self, pos):
        return bool(self.L.orc_map_is_inside(self.h, pos[0], pos[1]))

    def getIndex(self, pos):
        r, c = C.c_int32(), C.c_int32()
        ok = self.L.orc_map_get_index(self.h, pos[0], pos[1], C.byref(r), C.byref(c))
        return bool(ok), (r.value, c.value)

    def getPosition(self, idx):
        x, y = C.c_double(), C.c_double()
        self.L.orc_map_get_position(self.h, idx[0], idx[1], C.byref(x), C.byref(y))
        return (x.value, y.value)

    def move(self, pos, policy=0):
        return bool(self.L.orc_map_move(self.h, pos[0], pos[1], policy))

    def at(self, name, idx):
        return float(self.get(name)[idx[0], idx[1]])

    def setAt(self, name, idx, v):
        a = self.get(name)
        a[idx[0], idx[1]] = v
        self.set(name, a)

    def setPosition(self, pos):
        self.L.orc_map_set_position(self.h, pos[0], pos[1])

    def setStartIndex(self, idx):
        self.L.orc_map_set_start_index(self.h, idx[0], idx[1])

    def raycast(self, points, origin, cfg):
        p = _xyzw(points)
        o = np.asarray(origin, np.float32)
        self.L.orc_raycast(self.h, p.ctypes.data, p.shape[0], o.ctypes.data, C.byref(cfg))

    def inpaint(self, max_iterations=3, min_valid=2, inplace=False):
        self.L.orc_inpaint(self.h, max_iterations, min_valid, 1 if inplace else 0)


class OracleFastDEM:
    def __init__(self, omap: OracleMap, cfg: FdemConfig):
        self.L = lib()
        self.map = omap
        self.cfg = cfg
        self.h = C.c_void_p(self.L.orc_mapper_create(omap.h, C.byref(cfg)))

    def __del__(self):
        try:
            self.L.orc_mapper_destroy(self.h)
        except Exception:
            pass

    def integrate(self, points, Tbs, Twb, intensity=None, rgb=None):
        p = _xyzw(points)
        i = None if intensity is None else np.ascontiguousarray(intensity, np.float32)
        c = None if rgb is None else np.ascontiguousarray(rgb, np.uint8)
        st = FdemScanStats()
        el = C.c_double()
        a, b = _colmajor16(Tbs), _colmajor16(Twb)
        ok = self.L.orc_mapper_integrate(
            self.h, p.ctypes.data, None if i is None else i.ctypes.data,
            None if c is None else c.ctypes.data, p.shape[0], a.ctypes.data_as(_f64p),
            b.ctypes.data_as(_f64p), C.byref(st), C.byref(el))
        return bool(ok), st, el.value

    def update(self, points, robot_xy, var_z=None, intensity=None, rgb=None):
        p = _xyzw(points)
        v = None if var_z is None else np.ascontiguousarray(var_z, np.float32)
        i = None if intensity is None else np.ascontiguousarray(intensity, np.float32)
        c = None if rgb is None else np.ascontiguousarray(rgb, np.uint8)
        return int(self.L.orc_mapper_update(
            self.h, p.ctypes.data, None if v is None else v.ctypes.data,
            None if i is None else i.ctypes.data, None if c is None else c.ctypes.data, p.shape[0],
            float(robot_xy[0]), float(robot_xy[1])))


def preprocess(cfg, points, Tbs, Twb):
    p = _xyzw(points)
    n = p.shape[0]
    out = np.empty((n, 4), np.float32)
    cov = np.empty((n, 9), np.float32)
    src = np.empty(n, np.int32)
    a, b = _colmajor16(Tbs), _colmajor16(Twb)
    k = lib().orc_preprocess(C.byref(cfg), p.ctypes.data, n, a.ctypes.data_as(_f64p),
                             b.ctypes.data_as(_f64p), out.ctypes.data, cov.ctypes.data, src.ctypes.data)
    return out[:k], cov[:k], src[:k]


def sensor_cov(cfg, p3):
    p = np.asarray(p3, np.float32)
    out = np.empty(9, np.float32)
    lib().orc_sensor_cov(C.byref(cfg), p.ctypes.data, out.ctypes.data)
    return out.reshape(3, 3).T  # column-major -> [r, c]


def transform(T, points):
    p = _xyzw(points)
    out = np.empty_like(p)
    a = _colmajor16(T)
    lib().orc_transform(a.ctypes.data_as(_f64p), p.ctypes.data, p.shape[0], out.ctypes.data)
    return out


def kalman_step(state6, z, var, min_v=1e-4, max_v=1e-2, q=0.0):
    s = np.asarray(state6, np.float32).copy()
    out = np.empty(2, np.float32)
    lib().orc_kalman_step(s.ctypes.data, z, var, min_v, max_v, q, out.ctypes.data)
    return s, out


def p2_step(q5, n5, count, x, dn=(0.01, 0.16, 0.5, 0.84, 0.99), marker=3, max_count=0.0):
    q = np.asarray(q5, np.float32).copy()
    n = np.asarray(n5, np.float32).copy()
    c = C.c_float(count)
    d = np.asarray(dn, np.float32)
    e = lib().orc_p2_step(q.ctypes.data, n.ctypes.data, C.byref(c), x, d.ctypes.data, marker, max_count)
    return q, n, c.value, e


def voxel_any(points, voxel):
    p = _xyzw(points)
    out = np.empty(max(p.shape[0], 1), np.uint32)
    k = lib().orc_voxel_any(p.ctypes.data, p.shape[0], voxel, out.ctypes.data)
    if k < 0:
        raise ValueError("voxel_size must be in [0.001, 100]")
    return out[:k].copy()


def from_pointcloud2(data, n, layout):
    """nanopcl::from(msg) on the oracle: -> (xyzw [k,4], intensity [k] | None, rgb [k,3] | None)."""
    L = lib()
    L.orc_from_pointcloud2.restype = C.c_int64
    L.orc_from_pointcloud2.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    d = np.ascontiguousarray(data, np.uint8)
    xyzw = np.empty((max(n, 1), 4), np.float32)
    inten = np.empty(max(n, 1), np.float32) if layout.off_intensity >= 0 else None
    rgb = np.empty((max(n, 1), 3), np.uint8) if layout.off_rgb >= 0 else None
    k = L.orc_from_pointcloud2(d.ctypes.data, n, C.byref(layout), xyzw.ctypes.data,
                               None if inten is None else inten.ctypes.data,
                               None if rgb is None else rgb.ctypes.data)
    return xyzw[:k], None if inten is None else inten[:k], None if rgb is None else rgb[:k]


def spatial_smoothing(omap, layer, kernel_size=3, min_valid=5):
    L = lib()
    L.orc_spatial_smoothing.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int]
    L.orc_spatial_smoothing(omap.h, layer.encode(), kernel_size, min_valid)


def save_npz(omap, filename, frame_id="", layers=None):
    L = lib()
    L.orc_save_npz.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p]
    return bool(L.orc_save_npz(omap.h, str(filename).encode(), frame_id.encode(),
                               None if layers is None else "\n".join(layers).encode()))


def load_npz(omap, filename):
    """-> (ok, frame_id); re-creates the oracle map's geometry and layers from the file."""
    L = lib()
    L.orc_load_npz.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.c_int]
    buf = C.create_string_buffer(256)
    ok = bool(L.orc_load_npz(omap.h, str(filename).encode(), buf, 256))
    return ok, buf.value.decode()


def uncertainty_fusion(omap, search_radius=0.15, spatial_sigma=0.05, quantile_lower=0.01,
                       quantile_upper=0.99, min_valid=3):
    L = lib()
    L.orc_uncertainty_fusion.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int]
    L.orc_uncertainty_fusion.restype = None
    L.orc_uncertainty_fusion(omap.h, search_radius, spatial_sigma, quantile_lower, quantile_upper, min_valid)


def feature_extraction(omap, analysis_radius=0.3, min_valid=4, step_lower=0.05, step_upper=0.95):
    L = lib()
    L.orc_feature_extraction.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_float, C.c_float]
    L.orc_feature_extraction.restype = None
    L.orc_feature_extraction(omap.h, analysis_radius, min_valid, step_lower, step_upper)


def eig3(cov):
    """Eigen-style direct 3x3 solver of the oracle: (ascending eigenvalues, eigenvectors as rows)."""
    L = lib()
    L.orc_eig3.argtypes = [C.POINTER(C.c_float)] * 3
    L.orc_eig3.restype = None
    a = np.ascontiguousarray(cov, dtype=np.float32).reshape(9)
    val = np.zeros(3, np.float32)
    vec = np.zeros(9, np.float32)
    L.orc_eig3(a.ctypes.data_as(C.POINTER(C.c_float)), val.ctypes.data_as(C.POINTER(C.c_float)),
               vec.ctypes.data_as(C.POINTER(C.c_float)))
    return val, vec.reshape(3, 3)


def to_pointcloud2(omap, elevation_layer="elevation", sub_start=None, sub_size=None):
    """toPointCloud2Impl of the oracle: (fields, point_step, width, data bytes as uint8 array)."""
    L = lib()
    L.orc_to_pointcloud2.restype = C.c_int64
    L.orc_to_pointcloud2.argtypes = [C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                     C.c_int64, C.c_char_p, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    g = omap.geometry()
    r0, c0 = sub_start if sub_start is not None else g["start_index"]
    nr, nc = sub_size if sub_size is not None else (g["rows"], g["cols"])
    ps, w = C.c_uint32(0), C.c_uint32(0)
    names = C.create_string_buffer(4096)
    nbytes = L.orc_to_pointcloud2(omap.h, elevation_layer.encode(), r0, c0, nr, nc, None, 0, names, 4096,
                                  C.byref(ps), C.byref(w))
    data = np.zeros(max(nbytes, 1), np.uint8)
    L.orc_to_pointcloud2(omap.h, elevation_layer.encode(), r0, c0, nr, nc, data.ctypes.data, nbytes, names, 4096,
                         C.byref(ps), C.byref(w))
    return names.value.decode().split("\n"), ps.value, w.value, data[:nbytes]
