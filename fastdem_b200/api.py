"""Python mirror of the reference's C++ interface for the integrate() path.

Same names, argument meaning and return/error behaviour as
  fastdem::ElevationMap      fastdem/include/fastdem/elevation_map.hpp:65-177
  fastdem::FastDEM           fastdem/include/fastdem/fastdem.hpp:55-156
  fastdem::ElevationMapping  fastdem/include/fastdem/mapping/elevation_mapping.hpp:20-58
  fastdem::applyRaycasting   fastdem/include/fastdem/postprocess/raycasting.hpp:49-51
  nanopcl::PointCloud        fastdem/lib/nanoPCL/include/nanopcl/core/point_cloud.hpp:14-184
so the parity tests read like the reference's own gtest files.  Every method is a thin
call into libfastdem_b200.so (capi.py); no map arithmetic happens in Python.

Arrays: numpy (host) or torch CUDA tensors (device; zero-copy) are accepted wherever a
cloud channel is passed."""
from __future__ import annotations

import ctypes as C
from typing import Callable, Iterable, Optional, Sequence

import numpy as np

from . import capi
from .capi import (EST_KALMAN, EST_P2QUANTILE, MODE_GLOBAL, MODE_LOCAL, SENSOR_CONSTANT,
                   SENSOR_LIDAR, SENSOR_RGBD, FdemConfig, FdemGeometry, FdemScanStats, check)


class layer:  # namespace fastdem::layer (elevation_map.hpp:29-46 + estimator / raycast headers)
    elevation = "elevation"
    elevation_min = "elevation_min"
    elevation_max = "elevation_max"
    variance = "variance"
    n_points = "n_points"
    upper_bound = "upper_bound"
    lower_bound = "lower_bound"
    obstacle = "obstacle"
    intensity = "intensity"
    color = "color"
    kalman_p = "_kalman_p"
    sample_mean = "_sample_mean"
    sample_m2 = "_sample_m2"
    p2_q = ["_p2_q%d" % i for i in range(5)]
    p2_n = ["_p2_n%d" % i for i in range(5)]
    ghost_removal = "ghost_removal"
    raycasting = "raycasting"
    visibility_logodds = "_visibility_logodds"
    elevation_inpainted = "elevation_inpainted"

    @staticmethod
    def isInternal(name: str) -> bool:
        return bool(name) and name[0] == "_"


class MappingMode:
    LOCAL = MODE_LOCAL
    GLOBAL = MODE_GLOBAL


class EstimationType:
    Kalman = EST_KALMAN
    P2Quantile = EST_P2QUANTILE


class SensorType:
    Constant = SENSOR_CONSTANT
    LiDAR = SENSOR_LIDAR
    RGBD = SENSOR_RGBD


def Config() -> FdemConfig:
    """fastdem::Config{} with the reference's defaults (config/fastdem.hpp:21-37)."""
    return capi.default_config()


def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class DeviceArray:
    """A raw device buffer (this GPU's or a peer's mapped through CUDA IPC) holding `n` rows of
    a cloud channel: what the sharded driver passes when the scan lives in another rank's HBM.
    Assign instances directly to PointCloud.xyzw / .intensity / .color."""
    __slots__ = ("addr", "shape", "owner")

    def __init__(self, addr: int, n_rows: int, n_cols: int = 1, owner=None):
        self.addr = int(addr)
        self.shape = (int(n_rows), int(n_cols))
        self.owner = owner


def _ptr(x, dtype, shape_last: Optional[int] = None):
    """(address, keepalive, n_rows) of a numpy array or torch tensor as contiguous `dtype`."""
    if x is None:
        return None, None, 0
    if type(x) is DeviceArray:
        return x.addr, x, x.shape[0]
    if _is_torch(x):
        import torch
        tdt = {np.float32: torch.float32, np.uint8: torch.uint8}[dtype]
        t = x if (x.dtype == tdt and x.is_contiguous()) else x.to(tdt).contiguous()
        n = t.shape[0] if t.dim() > 0 else 0
        return t.data_ptr(), t, n
    a = np.ascontiguousarray(x, dtype=dtype)
    n = a.shape[0] if a.ndim > 0 else 0
    return a.ctypes.data, a, n


class PointCloud:
    """nanopcl::PointCloud restricted to the channels the path reads: points (xyzw, w = 1),
    optional intensity, optional color."""

    def __init__(self, xyz=None, intensity=None, color=None, frame_id: str = "", timestamp: int = 0):
        self._frame_id = frame_id
        self._timestamp = timestamp
        self._pending: list = []
        self.xyzw = None
        self.intensity = intensity
        self.color = color
        if xyz is not None:
            self.set_points(xyz)

    def set_points(self, xyz) -> None:
        if _is_torch(xyz):
            import torch
            if xyz.shape[-1] == 4:
                self.xyzw = xyz.to(torch.float32).contiguous()
            else:
                w = torch.ones((xyz.shape[0], 1), dtype=torch.float32, device=xyz.device)
                self.xyzw = torch.cat([xyz.to(torch.float32), w], dim=1).contiguous()
            return
        a = np.asarray(xyz, dtype=np.float32)
        if a.ndim != 2 or a.shape[1] not in (3, 4):
            raise ValueError("points must be N x 3 or N x 4")
        if a.shape[1] == 3:
            a = np.concatenate([a, np.ones((a.shape[0], 1), np.float32)], axis=1)  # add(): w = 1
        self.xyzw = np.ascontiguousarray(a)

    def add(self, x: float, y: float, z: float) -> None:  # point_cloud_impl.hpp:116-119
        self._pending.append((x, y, z, 1.0))

    def _flush(self) -> None:
        if self._pending:
            new = np.asarray(self._pending, dtype=np.float32)
            self.xyzw = new if self.xyzw is None else np.concatenate([np.asarray(self.xyzw), new])
            self._pending = []

    def size(self) -> int:
        self._flush()
        return 0 if self.xyzw is None else int(self.xyzw.shape[0])

    __len__ = size

    def empty(self) -> bool:
        return self.size() == 0

    def hasIntensity(self) -> bool:
        return self.intensity is not None

    def hasColor(self) -> bool:
        return self.color is not None

    def frameId(self) -> str:
        return self._frame_id

    def setFrameId(self, f: str) -> None:
        self._frame_id = f

    def timestamp(self) -> int:
        return self._timestamp

    def setTimestamp(self, ns: int) -> None:
        self._timestamp = ns


# sensor_msgs/PointField datatypes (nanopcl/bridge/ros/impl.hpp:22-31)
PF_INT8, PF_UINT8, PF_INT16, PF_UINT16, PF_INT32, PF_UINT32, PF_FLOAT32, PF_FLOAT64 = 1, 2, 3, 4, 5, 6, 7, 8


class PointCloud2:
    """A sensor_msgs/PointCloud2 message body: `data` (bytes / uint8 numpy / uint8 torch CUDA
    tensor), width x height points of `point_step` bytes, `fields` = [(name, offset, datatype)].
    FastDEM.integrate(msg, ...) == integrate(nanopcl::from(msg), ...) with the unpacking done
    on the device."""

    def __init__(self, data, width: int, height: int, point_step: int, fields, frame_id: str = "",
                 timestamp: int = 0):
        self.data = np.frombuffer(data, dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else data
        self.width, self.height, self.point_step = int(width), int(height), int(point_step)
        self.fields = list(fields)
        self._frame_id, self._timestamp = frame_id, timestamp

    def size(self) -> int:
        return self.width * self.height

    def empty(self) -> bool:
        return self.size() == 0

    def frameId(self) -> str:
        return self._frame_id

    def timestamp(self) -> int:
        return self._timestamp

    def layout(self) -> capi.FdemPointCloud2Layout:
        """FieldOffsets::parse (nanopcl/bridge/ros/impl.hpp:66-104)."""
        lo = capi.FdemPointCloud2Layout(self.point_step, -1, -1, -1, -1, 0, -1)
        for name, offset, datatype in self.fields:
            if name == "x":
                lo.off_x = offset
            elif name == "y":
                lo.off_y = offset
            elif name == "z":
                lo.off_z = offset
            elif name == "intensity":
                lo.off_intensity, lo.intensity_type = offset, datatype
            elif name in ("rgb", "rgba"):
                lo.off_rgb = offset
        return lo

    @staticmethod
    def from_arrays(xyz, intensity=None, rgb=None, intensity_type=PF_FLOAT32, pad: int = 0):
        """Pack arrays the way ROS drivers do (x, y, z, [intensity], [rgb], padding)."""
        xyz = np.asarray(xyz, np.float32)
        n = xyz.shape[0]
        fields = [("x", 0, PF_FLOAT32), ("y", 4, PF_FLOAT32), ("z", 8, PF_FLOAT32)]
        dt = [("x", "<f4"), ("y", "<f4"), ("z", "<f4")]
        off = 12
        if intensity is not None:
            np_t = {PF_UINT8: "u1", PF_UINT16: "<u2", PF_FLOAT32: "<f4", PF_FLOAT64: "<f8"}[intensity_type]
            align = 4 if intensity_type in (PF_FLOAT32, PF_FLOAT64) else np.dtype(np_t).itemsize
            while off % align:
                dt.append((f"_p{off}", "u1"))
                off += 1
            fields.append(("intensity", off, intensity_type))
            dt.append(("intensity", np_t))
            off += np.dtype(np_t).itemsize
        if rgb is not None:
            while off % 4:
                dt.append((f"_p{off}", "u1"))
                off += 1
            fields.append(("rgb", off, PF_FLOAT32))
            dt.append(("rgb", "<u4"))
            off += 4
        step = -(-(off + pad) // 4) * 4
        while off < step:
            dt.append((f"_p{off}", "u1"))
            off += 1
        a = np.zeros(n, dtype=np.dtype(dt))
        a["x"], a["y"], a["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
        if intensity is not None:
            a["intensity"] = np.asarray(intensity).astype(a.dtype["intensity"])
        if rgb is not None:
            c = np.asarray(rgb, np.uint32)
            a["rgb"] = (c[:, 0] << 16) | (c[:, 1] << 8) | c[:, 2]
        return PointCloud2(a.view(np.uint8).reshape(-1), n, 1, step, fields)


class _Iso16:
    """A transform ready for the C-ABI: column-major double[16] + its address.  Built once per
    pose; per-call pointer extraction (ndarray.ctypes) costs microseconds, which matters when a
    scan is 20 us of GPU time."""
    __slots__ = ("a", "p")

    def __init__(self, a: np.ndarray):
        self.a = a
        self.p = a.ctypes.data


def _iso(T) -> "_Iso16":
    """Eigen::Isometry3d -> column-major double[16]."""
    if type(T) is _Iso16:
        return T  # hot-loop fast path: a pose converted earlier
    if isinstance(T, np.ndarray) and T.shape == (16,) and T.dtype == np.float64 and T.flags.c_contiguous:
        return _Iso16(T)  # already column-major double[16]
    M = np.asarray(T, dtype=np.float64)
    if M.shape != (4, 4):
        raise ValueError("transform must be 4x4")
    return _Iso16(np.ascontiguousarray(M.T).reshape(16))  # row-major of the transpose == column-major


class ElevationMap:
    """fastdem::ElevationMap on the device.  Layers are float32 rows x cols, column-major."""

    def __init__(self, width: float = 0.0, height: float = 0.0, resolution: float = 0.0,
                 frame_id: str = "", device: int = 0, stream: int = 0,
                 row_stripe: Optional[Sequence[int]] = None):
        self._lib = capi.load_library()
        self._h = C.c_void_p()
        self._frame_id = frame_id
        self._device = device
        self._stream = stream
        if width > 0 and height > 0 and resolution > 0:
            self.setGeometry(width, height, resolution, row_stripe)

    # ── lifetime ──
    def setGeometry(self, width: float, height: float, resolution: float,
                    row_stripe: Optional[Sequence[int]] = None) -> None:
        if self._h and row_stripe is None:
            # in place, like nanoGrid setGeometry + clearAll (elevation_map.hpp:112-116): the handle
            # and every FastDEM bound to it stay valid (io.loadNpz restores into a live map)
            check(self._lib.fdem_map_set_geometry(self._h, width, height, resolution))
            return
        if self._h:
            check(self._lib.fdem_map_destroy(self._h))
            self._h = C.c_void_p()
        h = C.c_void_p()
        if row_stripe is None:
            check(self._lib.fdem_map_create(width, height, resolution, self._device,
                                            C.c_void_p(self._stream), C.byref(h)))
        else:
            check(self._lib.fdem_map_create_stripe(width, height, resolution, int(row_stripe[0]),
                                                   int(row_stripe[1]), self._device,
                                                   C.c_void_p(self._stream), C.byref(h)))
        self._h = h

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.fdem_map_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> C.c_void_p:
        if not self._h:
            raise RuntimeError("ElevationMap has no geometry (call setGeometry)")
        return self._h

    # ── geometry (nanogrid::GridMap accessors) ──
    def isInitialized(self) -> bool:
        return bool(self._h)

    def geometry(self) -> FdemGeometry:
        g = FdemGeometry()
        check(self._lib.fdem_map_get_geometry(self.handle, C.byref(g)))
        return g

    def getSize(self):
        g = self.geometry()
        return (g.rows, g.cols)

    def getResolution(self) -> float:
        return self.geometry().resolution

    def getLength(self):
        g = self.geometry()
        return (g.length[0], g.length[1])

    def getPosition(self, index=None):
        if index is None:
            g = self.geometry()
            return (g.position[0], g.position[1])
        x, y = C.c_double(), C.c_double()
        check(self._lib.fdem_map_get_cell_position(self.handle, int(index[0]), int(index[1]),
                                                   C.byref(x), C.byref(y)))
        return (x.value, y.value)

    def getStartIndex(self):
        g = self.geometry()
        return (g.start_index[0], g.start_index[1])

    def rowStripe(self):
        g = self.geometry()
        return (g.row_begin, g.row_end)

    def setPosition(self, pos) -> None:
        check(self._lib.fdem_map_set_position(self.handle, float(pos[0]), float(pos[1])))

    def setStartIndex(self, idx) -> None:
        check(self._lib.fdem_map_set_start_index(self.handle, int(idx[0]), int(idx[1])))

    def getFrameId(self) -> str:
        return self._frame_id

    def setFrameId(self, f: str) -> None:
        self._frame_id = f

    def isInside(self, pos) -> bool:
        v = C.c_int32()
        check(self._lib.fdem_map_is_inside(self.handle, float(pos[0]), float(pos[1]), C.byref(v)))
        return bool(v.value)

    def getIndex(self, pos):
        """-> (inside, (row, col)) like `bool getIndex(position, index&)`."""
        r, c, ins = C.c_int32(), C.c_int32(), C.c_int32()
        check(self._lib.fdem_map_get_index(self.handle, float(pos[0]), float(pos[1]), C.byref(r),
                                           C.byref(c), C.byref(ins)))
        return bool(ins.value), (r.value, c.value)

    def move(self, pos, clear_policy: int = capi.MOVE_CLEAR_ALL_LAYERS) -> bool:
        moved = C.c_int32()
        check(self._lib.fdem_map_move(self.handle, float(pos[0]), float(pos[1]), clear_policy,
                                      C.byref(moved)))
        return bool(moved.value)

    # ── layers ──
    def exists(self, name: str) -> bool:
        v = C.c_int32()
        check(self._lib.fdem_map_layer_exists(self.handle, name.encode(), C.byref(v)))
        return bool(v.value)

    def add(self, name: str, fill=float("nan")) -> None:
        if isinstance(fill, (int, float)):
            check(self._lib.fdem_map_layer_add(self.handle, name.encode(), float(fill)))
        else:
            check(self._lib.fdem_map_layer_add(self.handle, name.encode(), float("nan")))
            self.set(name, fill)

    def getLayers(self) -> list:
        n = C.c_int32()
        check(self._lib.fdem_map_layer_count(self.handle, C.byref(n)))
        out = []
        buf = C.create_string_buffer(128)
        for i in range(n.value):
            check(self._lib.fdem_map_layer_name(self.handle, i, buf, 128))
            out.append(buf.value.decode())
        return out

    def _local_shape(self):
        g = self.geometry()
        return (g.row_end - g.row_begin, g.cols)

    def get(self, name: str) -> np.ndarray:
        """Copy of the layer as a (rows, cols) Fortran-order float32 array (== Eigen::MatrixXf)."""
        rows, cols = self._local_shape()
        a = np.empty((rows, cols), dtype=np.float32, order="F")
        check(self._lib.fdem_map_layer_download(self.handle, name.encode(), a.ctypes.data))
        return a

    def set(self, name: str, values) -> None:
        rows, cols = self._local_shape()
        a = np.asfortranarray(np.asarray(values, dtype=np.float32))
        if a.shape != (rows, cols):
            raise ValueError(f"layer must be {rows}x{cols}")
        check(self._lib.fdem_map_layer_upload(self.handle, name.encode(), a.ctypes.data))

    def tensor(self, name: str):
        """Zero-copy torch view (cols, rows) of the device slab: tensor[c, r] == layer(r, c)."""
        import torch
        p = C.c_void_p()
        check(self._lib.fdem_map_layer_device_ptr(self.handle, name.encode(), C.byref(p)))
        rows, cols = self._local_shape()

        class _Holder:  # __cuda_array_interface__ provider
            pass

        h = _Holder()
        h.__cuda_array_interface__ = {"shape": (cols, rows), "typestr": "<f4",
                                      "data": (p.value, False), "version": 3, "strides": None}
        h._owner = self
        return torch.as_tensor(h, device=f"cuda:{self._device}")

    def at(self, name: str, index) -> float:
        v = C.c_float()
        check(self._lib.fdem_map_cell_get(self.handle, name.encode(), int(index[0]), int(index[1]),
                                          C.byref(v)))
        return v.value

    def setAt(self, name: str, index, value: float) -> None:
        check(self._lib.fdem_map_cell_set(self.handle, name.encode(), int(index[0]), int(index[1]),
                                          float(value)))

    def clear(self, name: str) -> None:
        check(self._lib.fdem_map_clear(self.handle, name.encode()))

    def clearAll(self) -> None:
        check(self._lib.fdem_map_clear_all(self.handle))

    def clearAt(self, index) -> None:
        check(self._lib.fdem_map_clear_at(self.handle, int(index[0]), int(index[1])))

    # ── ElevationMap conveniences (elevation_map.hpp:118-177) ──
    def isEmpty(self) -> bool:
        v = C.c_int32()
        check(self._lib.fdem_map_is_empty(self.handle, C.byref(v)))
        return bool(v.value)

    def isEmptyAt(self, index) -> bool:
        return bool(np.isnan(self.at(layer.elevation, index)))

    def elevationAt(self, where) -> float:
        """Position (floats) or Index (ints), like the two C++ overloads."""
        if all(isinstance(v, (int, np.integer)) for v in where):
            return self.at(layer.elevation, where)
        inside, idx = self.getIndex(where)
        if not inside:
            return float("nan")
        return self.at(layer.elevation, idx)

    def hasElevationAt(self, where) -> bool:
        return bool(np.isfinite(self.elevationAt(where)))

    def sync(self) -> None:
        check(self._lib.fdem_map_sync(self.handle))

    def stream(self) -> int:
        return int(self._lib.fdem_map_stream(self.handle) or 0)


CloudCallback = Callable[[PointCloud], None]


class FastDEM:
    """fastdem::FastDEM — scan-sequential elevation mapping API (fastdem.hpp:55-156)."""

    def __init__(self, map: ElevationMap, cfg: Optional[FdemConfig] = None):
        self._lib = capi.load_library()
        self._map = map
        self._cfg = cfg.copy() if cfg is not None else Config()
        self._h = C.c_void_p()
        check(self._lib.fdem_mapper_create(map.handle, C.byref(self._cfg), C.byref(self._h)))
        self._calibration = None
        self._odometry = None
        self._on_preprocessed: Optional[CloudCallback] = None
        self._on_rasterized: Optional[CloudCallback] = None
        self._keep = None  # inputs of in-flight async scans

    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.fdem_mapper_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _push_config(self) -> "FastDEM":
        check(self._lib.fdem_mapper_set_config(self._h, C.byref(self._cfg)))
        return self

    # fluent setters (fastdem.cpp:28-66)
    def setMappingMode(self, mode: int) -> "FastDEM":
        self._cfg.mode = mode
        return self._push_config()

    def setEstimatorType(self, t: int) -> "FastDEM":
        self._cfg.estimation_type = t
        return self._push_config()

    def setSensorModel(self, t: int) -> "FastDEM":
        self._cfg.sensor_type = t
        return self._push_config()

    def setHeightFilter(self, z_min: float, z_max: float) -> "FastDEM":
        self._cfg.z_min, self._cfg.z_max = z_min, z_max
        return self._push_config()

    def setRangeFilter(self, range_min: float, range_max: float) -> "FastDEM":
        self._cfg.range_min, self._cfg.range_max = range_min, range_max
        return self._push_config()

    def enableRaycasting(self, enabled: bool = True) -> "FastDEM":
        self._cfg.raycasting_enabled = 1 if enabled else 0
        return self._push_config()

    def setCalibrationProvider(self, calibration) -> "FastDEM":
        self._calibration = calibration
        return self

    def setOdometryProvider(self, odometry) -> "FastDEM":
        self._odometry = odometry
        return self

    def setTransformProvider(self, system) -> "FastDEM":
        return self.setCalibrationProvider(system).setOdometryProvider(system)

    def hasTransformProvider(self) -> bool:
        return self._calibration is not None and self._odometry is not None

    def reset(self) -> None:  # fastdem.cpp:26
        self._map.clearAll()

    def config(self) -> FdemConfig:
        return self._cfg

    def onScanPreprocessed(self, cb: CloudCallback) -> None:
        self._on_preprocessed = cb

    def onScanRasterized(self, cb: CloudCallback) -> None:
        self._on_rasterized = cb

    # ── integrate ──
    def integrate(self, cloud: PointCloud, T_base_sensor=None, T_world_base=None) -> bool:
        """Both overloads: integrate(cloud) resolves transforms through the providers
        (fastdem.cpp:83-120); integrate(cloud, T_base_sensor, T_world_base) uses them as given
        (fastdem.cpp:122-131)."""
        if T_base_sensor is None and T_world_base is None:
            if self._calibration is None or self._odometry is None:
                return False  # "Transform providers not set"
            if cloud is None or cloud.empty():
                return False
            if not cloud.frameId():
                return False  # "Input cloud has no frameId"
            T_base_sensor = self._calibration.getExtrinsic(cloud.frameId())
            if T_base_sensor is None:
                return False
            T_world_base = self._odometry.getPoseAt(cloud.timestamp())
            if T_world_base is None:
                return False
        if isinstance(cloud, PointCloud2):
            return bool(self.integrate_pointcloud2(cloud, T_base_sensor, T_world_base).integrated)
        stats = self.integrate_stats(cloud, T_base_sensor, T_world_base)
        return bool(stats.integrated)

    def integrate_pointcloud2(self, msg: "PointCloud2", T_base_sensor, T_world_base) -> FdemScanStats:
        """integrate(nanopcl::from(msg), ...): the message body goes to the device as it is."""
        p, keep, _ = _ptr(msg.data, np.uint8)
        lo = msg.layout()
        Tbs, Twb = _iso(T_base_sensor), _iso(T_world_base)
        stats = FdemScanStats()
        check(self._lib.fdem_mapper_integrate_pointcloud2(
            self._h, p, msg.size(), C.byref(lo), Tbs.p,
            Twb.p, C.byref(stats)))
        return stats

    def submit_pointcloud2(self, msg: "PointCloud2", T_base_sensor, T_world_base) -> int:
        """Streaming form of integrate_pointcloud2: returns a ticket for collect()."""
        c = getattr(msg, "_abi_cache", None)
        if c is None or c[0] is not msg.data:
            p, keep, _ = _ptr(msg.data, np.uint8)
            c = (msg.data, p, keep, msg.layout(), msg.size())
            msg._abi_cache = c
        Tbs, Twb = _iso(T_base_sensor), _iso(T_world_base)
        t = C.c_uint64()
        check(self._lib.fdem_mapper_submit_pointcloud2(self._h, c[1], c[4], C.byref(c[3]), Tbs.p, Twb.p, C.byref(t)))
        if not hasattr(self, "_inflight"):
            self._inflight = {}
        self._inflight[t.value] = c
        return t.value

    def _channels(self, cloud: PointCloud):
        # pointers of a cloud's channels are cached on the cloud (keyed by the identity of the
        # channel objects): extracting them costs ~3 us per channel
        c = getattr(cloud, "_abi_cache", None)
        if c is not None and c[0] is cloud.xyzw and c[1] is cloud.intensity and c[2] is cloud.color:
            return c[3]
        n = cloud.size()
        pxyzw, k0, _ = _ptr(cloud.xyzw, np.float32)
        pint, k1, ni = _ptr(cloud.intensity, np.float32)
        prgb, k2, nc = _ptr(cloud.color, np.uint8)
        if cloud.intensity is not None and ni != n:
            raise ValueError("intensity length mismatch")
        if cloud.color is not None and nc != n:
            raise ValueError("color length mismatch")
        out = (n, pxyzw, pint, prgb, (k0, k1, k2))
        try:
            cloud._abi_cache = (cloud.xyzw, cloud.intensity, cloud.color, out)
        except AttributeError:
            pass
        return out

    def integrate_stats(self, cloud: PointCloud, T_base_sensor, T_world_base) -> FdemScanStats:
        n, pxyzw, pint, prgb, keep = self._channels(cloud)
        Tbs, Twb = _iso(T_base_sensor), _iso(T_world_base)
        stats = FdemScanStats()
        check(self._lib.fdem_mapper_integrate(
            self._h, pxyzw, pint, prgb, n, Tbs.p,
            Twb.p, C.byref(stats)))
        del keep
        if stats.integrated:
            if self._on_preprocessed is not None:
                self._on_preprocessed(self._last_preprocessed())
            if self._on_rasterized is not None and stats.n_cells > 0:
                self._on_rasterized(self._last_rasterized())
        return stats

    def integrate_async(self, cloud: PointCloud, T_base_sensor, T_world_base) -> None:
        n, pxyzw, pint, prgb, keep = self._channels(cloud)
        Tbs, Twb = _iso(T_base_sensor), _iso(T_world_base)
        check(self._lib.fdem_mapper_integrate_async(
            self._h, pxyzw, pint, prgb, n, Tbs.p,
            Twb.p))
        self._keep = keep

    def integrate_batch(self, clouds, poses, wait: bool = True):
        """Up to 16 consecutive integrate() calls as one graph in which scan k+1's front half runs
        beside scan k's estimator (fdem_mapper_integrate_batch).  clouds: list of PointCloud;
        poses: list of (T_base_sensor, T_world_base).  Returns the per-scan stats (wait=True) or
        None (queued; see wait())."""
        S = len(clouds)
        key = tuple(id(c) for c in clouds)
        caches = getattr(self, "_batch_caches", None)
        if caches is None:
            caches = self._batch_caches = {}
        cache = caches.get(key)
        if cache is not None and any(a is not b for a, b in zip(cache[11], clouds)):
            cache = None   # an id() was recycled for a different object
        if cache is None:
            ch = [self._channels(c) for c in clouds]
            has_i = ch[0][2] is not None
            has_c = ch[0][3] is not None
            vp = C.c_void_p * S
            px = vp(*[c[1] for c in ch])
            pi = vp(*[c[2] for c in ch]) if has_i else None
            pc = vp(*[c[3] for c in ch]) if has_c else None
            pn = (C.c_size_t * S)(*[c[0] for c in ch])
            cache = (key, px, pi, pc, pn, ch, np.empty(S * 16, np.float64), np.empty(S * 16, np.float64),
                     (FdemScanStats * S)())
            cache = cache + (cache[6].ctypes.data, cache[7].ctypes.data, list(clouds))
            if len(caches) >= 256:
                caches.clear()
            caches[key] = cache
        _, px, pi, pc, pn, ch, tbs, twb, stats, ptbs, ptwb, _ = cache
        for i, (a, b) in enumerate(poses):
            tbs[16 * i:16 * i + 16] = _iso(a).a
            twb[16 * i:16 * i + 16] = _iso(b).a
        check(self._lib.fdem_mapper_integrate_batch(self._h, S, px, pi, pc, pn, ptbs, ptwb,
                                                    stats if wait else None))
        if not wait:
            self._keep = ch
            return None
        return [FdemScanStats.from_buffer_copy(st) for st in stats]

    def last_batch_stats(self, n_scans: int):
        """Per-scan stats of the most recent integrate_batch(..., wait=False) (waits for it)."""
        arr = (FdemScanStats * n_scans)()
        check(self._lib.fdem_mapper_last_batch_stats(self._h, arr, n_scans))
        return [FdemScanStats.from_buffer_copy(st) for st in arr]

    def submit(self, cloud: PointCloud, T_base_sensor, T_world_base) -> int:
        """Queue one scan; returns its ticket.  `submit(k+1); collect(k)` overlaps the
        host->device copy of scan k+1 with the kernels of scan k."""
        n, pxyzw, pint, prgb, keep = self._channels(cloud)
        Tbs, Twb = _iso(T_base_sensor), _iso(T_world_base)
        t = C.c_uint64()
        check(self._lib.fdem_mapper_submit(
            self._h, pxyzw, pint, prgb, n, Tbs.p,
            Twb.p, C.byref(t)))
        if not hasattr(self, "_inflight"):
            self._inflight = {}
        self._inflight[t.value] = keep
        return t.value

    def collect(self, ticket: int) -> FdemScanStats:
        stats = FdemScanStats()
        check(self._lib.fdem_mapper_collect(self._h, ticket, C.byref(stats)))
        getattr(self, "_inflight", {}).pop(ticket, None)
        return stats

    def wait(self) -> FdemScanStats:
        stats = FdemScanStats()
        check(self._lib.fdem_mapper_wait(self._h, C.byref(stats)))
        self._keep = None
        return stats

    def integrate_with_covariances(self, cloud: PointCloud, cov9, T_base_sensor, T_world_base) -> FdemScanStats:
        """setSensorModel(std::unique_ptr<SensorModel>) path: the caller's computeCovariances()
        output (N x 9, column-major 3x3 each) replaces the built-in sensor models."""
        n, pxyzw, pint, prgb, keep = self._channels(cloud)
        pcov, kc, nc = _ptr(cov9, np.float32)
        Tbs, Twb = _iso(T_base_sensor), _iso(T_world_base)
        stats = FdemScanStats()
        check(self._lib.fdem_mapper_integrate_with_cov(
            self._h, pxyzw, pcov, pint, prgb, n, Tbs.p,
            Twb.p, C.byref(stats)))
        return stats

    def _last_preprocessed(self) -> PointCloud:
        nk = C.c_int64()
        check(self._lib.fdem_mapper_last_preprocessed(self._h, None, None, None, C.byref(nk)))
        xyzw = np.empty((max(nk.value, 1), 4), np.float32)
        src = np.empty(max(nk.value, 1), np.int32)
        check(self._lib.fdem_mapper_last_preprocessed(self._h, xyzw.ctypes.data, None,
                                                      src.ctypes.data, C.byref(nk)))
        xyzw = xyzw[:nk.value]
        pc = PointCloud(frame_id=self._map.getFrameId())
        pc.var_z = xyzw[:, 3].copy()
        pc.src_index = src[:nk.value]
        xyzw = xyzw.copy()
        xyzw[:, 3] = 1.0
        pc.xyzw = xyzw
        return pc

    def _last_rasterized(self) -> PointCloud:
        nc = C.c_int64()
        check(self._lib.fdem_mapper_last_rasterized(self._h, None, C.byref(nc)))
        xyz = np.empty((max(nc.value, 1), 3), np.float32)
        check(self._lib.fdem_mapper_last_rasterized(self._h, xyz.ctypes.data, C.byref(nc)))
        return PointCloud(xyz[:nc.value], frame_id=self._map.getFrameId())

    # tuning probes: only a -DFDEM_PROBES build exports these (tools/phase_probe.py)
    def debug_cta_times(self):
        fn = self._lib.fdem_mapper_debug_cta_times  # AttributeError: production build
        fn.restype, fn.argtypes = C.c_int32, [C.c_void_p, C.POINTER(C.c_uint64)]
        out = (C.c_uint64 * 1024)()
        check(fn(self._h, out))
        return np.array(out, dtype=np.uint64).reshape(512, 2)

    def debug_phase_clocks(self):
        fn = self._lib.fdem_mapper_debug_phase_clocks
        fn.restype, fn.argtypes = C.c_int32, [C.c_void_p, C.POINTER(C.c_int64)]
        out = (C.c_int64 * 16)()
        check(fn(self._h, out))
        return list(out)

    def set_cell_sort(self, mode: int) -> None:
        """capi.CELL_SORT_TILE (default) or capi.CELL_SORT_GLOBAL; results are identical."""
        check(self._lib.fdem_mapper_set_cell_sort(self._h, mode))

    def set_voxel_sort(self, mode: int) -> None:
        """capi.VOXEL_SORT_LIBRARY (default, faster) or capi.VOXEL_SORT_MSD (no library launches); results are identical."""
        check(self._lib.fdem_mapper_set_voxel_sort(self._h, mode))

    def set_stage_timing(self, enabled: bool) -> None:
        check(self._lib.fdem_mapper_set_stage_timing(self._h, 1 if enabled else 0))

    def stage_times(self):
        """-> ({stage: total device ms}, n_scans) accumulated since the last call."""
        ms = (C.c_double * len(capi.STAGE_NAMES))()
        n = C.c_int64()
        check(self._lib.fdem_mapper_stage_times(self._h, ms, C.byref(n)))
        return dict(zip(capi.STAGE_NAMES, list(ms))), n.value

    def library_launch_count(self) -> int:
        v = C.c_int64()
        check(self._lib.fdem_mapper_library_launch_count(self._h, C.byref(v)))
        return v.value

    def launch_count(self) -> int:
        v = C.c_int64()
        check(self._lib.fdem_mapper_launch_count(self._h, C.byref(v)))
        return v.value

    # the lower seam, fastdem::ElevationMapping::update(cloud, robot_position)
    def update(self, cloud: PointCloud, robot_position, var_z=None) -> FdemScanStats:
        n, pxyzw, pint, prgb, keep = self._channels(cloud)
        pvar, kv, nv = _ptr(var_z, np.float32)
        stats = FdemScanStats()
        check(self._lib.fdem_mapper_update(self._h, pxyzw, pvar, pint, prgb, n,
                                           float(robot_position[0]), float(robot_position[1]),
                                           C.byref(stats)))
        return stats


class ElevationMapping:
    """fastdem::ElevationMapping(map, config::Mapping) — the seam test_dual_layer.cpp drives."""

    def __init__(self, map: ElevationMap, cfg: Optional[FdemConfig] = None):
        self._dem = FastDEM(map, cfg)

    def update(self, cloud: PointCloud, robot_position, var_z=None) -> FdemScanStats:
        return self._dem.update(cloud, robot_position, var_z)


def applyRaycasting(map: ElevationMap, scan: PointCloud, sensor_origin, cfg: FdemConfig) -> None:
    lib = capi.load_library()
    n = scan.size()
    p, keep, _ = _ptr(scan.xyzw, np.float32)
    o = (C.c_float * 3)(*[float(v) for v in sensor_origin])
    check(lib.fdem_raycast(map.handle, p, n, o, C.byref(cfg)))


def voxelGridAny(map: ElevationMap, cloud: PointCloud, voxel_size: float) -> np.ndarray:
    """nanopcl::filters::voxelGrid(cloud, voxel_size, VoxelMode::ANY): selected source indices."""
    lib = capi.load_library()
    n = cloud.size()
    p, keep, _ = _ptr(cloud.xyzw, np.float32)
    out = np.empty(max(n, 1), np.uint32)
    nv = C.c_int64()
    check(lib.fdem_voxel_grid_any(map.handle, p, n, float(voxel_size), out.ctypes.data, C.byref(nv)))
    return out[:nv.value].copy()


def applyInpainting(map: ElevationMap, max_iterations: int = 3, min_valid_neighbors: int = 2,
                    inplace: bool = False) -> None:
    lib = capi.load_library()
    check(lib.fdem_inpaint(map.handle, max_iterations, min_valid_neighbors, 1 if inplace else 0))


def applySpatialSmoothing(map: ElevationMap, layer_name: str, kernel_size: int = 3,
                          min_valid_neighbors: int = 5) -> None:
    """fastdem::applySpatialSmoothing (postprocess/spatial_smoothing.hpp:38-67)."""
    lib = capi.load_library()
    check(lib.fdem_spatial_smoothing(map.handle, layer_name.encode(), kernel_size, min_valid_neighbors))


def applyUncertaintyFusion(map: ElevationMap, search_radius: float = 0.15, spatial_sigma: float = 0.05,
                           quantile_lower: float = 0.01, quantile_upper: float = 0.99,
                           min_valid_neighbors: int = 3, enabled: bool = True) -> None:
    """fastdem::applyUncertaintyFusion(map, config::UncertaintyFusion)
    (src/uncertainty_fusion.cpp:103-186); keyword defaults = config/postprocess.hpp:33-40
    except `enabled` (the reference's config defaults to disabled = no-op)."""
    if not enabled:
        return
    lib = capi.load_library()
    check(lib.fdem_uncertainty_fusion(map.handle, search_radius, spatial_sigma, quantile_lower,
                                      quantile_upper, min_valid_neighbors))


def applyFeatureExtraction(map: ElevationMap, analysis_radius: float = 0.3, min_valid_neighbors: int = 4,
                           step_lower_percentile: float = 0.05, step_upper_percentile: float = 0.95) -> None:
    """fastdem::applyFeatureExtraction (src/feature_extraction.cpp:28-118)."""
    lib = capi.load_library()
    check(lib.fdem_feature_extraction(map.handle, analysis_radius, min_valid_neighbors,
                                      step_lower_percentile, step_upper_percentile))


def toPointCloud2(map: ElevationMap, elevation_layer: str = "elevation", sub_start=None, sub_size=None,
                  stamp: int = 0) -> PointCloud2:
    """fastdem::ros::toPointCloud2Impl (include/fastdem/bridge/ros/impl.hpp:29-174): the map as a
    PointCloud2 message (packed on the device).  sub_start/sub_size = buffer start index / size of
    a sub-region; default = the full map."""
    lib = capi.load_library()
    w, ps, nf = C.c_uint32(0), C.c_uint32(0), C.c_int32(0)
    r0, c0 = sub_start if sub_start is not None else (0, 0)
    nr, nc = sub_size if sub_size is not None else (-1, -1)
    check(lib.fdem_map_pack_pointcloud2(map.handle, elevation_layer.encode(), r0, c0, nr, nc,
                                        C.byref(w), C.byref(ps), C.byref(nf)))
    fields = []
    buf = C.create_string_buffer(128)
    off = C.c_uint32(0)
    for i in range(nf.value):
        check(lib.fdem_map_pointcloud2_field(map.handle, i, buf, 128, C.byref(off)))
        fields.append((buf.value.decode(), off.value, PF_FLOAT32))
    data = np.zeros(max(w.value * ps.value, 1), np.uint8)
    check(lib.fdem_map_pointcloud2_data(map.handle, data.ctypes.data, None))
    return PointCloud2(data[:w.value * ps.value], w.value, 1, ps.value, fields, map.getFrameId(), stamp)
