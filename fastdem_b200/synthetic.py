"""Seeded synthetic scans + poses for the BASELINE.json configurations.

There is no network and the reference ships no recorded data (its benchmark's KITTI file is
not in the tree, fastdem/benchmarks/benchmark_height_update.cpp:668), so the workloads are
generated: a LiDAR / RGB-D sensor moving over the terrain
    z = 0.3 sin(0.5 x) cos(0.5 y) - 0.5
(the reference's own example terrain, fastdem/examples/common/data_loader.hpp:32-53),
with a surrounding cylinder wall for the non-ground beams.  Everything is a pure function
of (config name, scan index) — numpy RandomState(42 + scan_idx)."""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from . import capi


def terrain(x, y):
    return 0.3 * np.sin(0.5 * x) * np.cos(0.5 * y) - 0.5


@dataclass
class Workload:
    name: str
    description: str
    map_width: float
    map_height: float
    resolution: float
    beams: int            # LiDAR rows / image rows
    azimuths: int         # LiDAR columns / image cols
    sensor: str           # "lidar" | "rgbd"
    elev_min_deg: float = 0.0
    elev_max_deg: float = 0.0
    wall_radius: float = 12.0
    has_intensity: bool = True
    has_color: bool = False
    ghost_fraction: float = 0.0
    loop_radius: float = 0.0     # > 0: robot drives a circle (global mapping)
    cfg_overrides: dict = field(default_factory=dict)

    @property
    def points_per_scan(self) -> int:
        return self.beams * self.azimuths

    def config(self, default=None) -> capi.FdemConfig:
        """fastdem::Config for this workload.  `default` = the function that yields Config{}
        (the library's by default; the CPU-only reference arm passes the oracle's so that it
        never maps the CUDA library)."""
        c = (default or capi.default_config)()
        for k, v in self.cfg_overrides.items():
            setattr(c, k, v)
        return c


_LIDAR_FILTER = dict(z_min=-1.0, z_max=2.0, range_min=0.5, range_max=20.0)  # config/default.yaml:19-23

WORKLOADS = {
    # BASELINE.json configs[0]: the reference's own published case (README.md:59)
    "c1_vlp16_local": Workload(
        "c1_vlp16_local", "VLP-16 28.8K pts, 15x15 m @ 0.1 m, Kalman, LiDAR model, LOCAL",
        15.0, 15.0, 0.1, 16, 1800, "lidar", -15.0, 15.0, 12.0,
        cfg_overrides=dict(_LIDAR_FILTER)),
    # configs[1]: the configuration the metric is quoted on (bench default)
    "c2_lidar64_local": Workload(
        "c2_lidar64_local", "64-beam LiDAR 131K pts, 30x30 m @ 0.05 m, Kalman, LiDAR model, LOCAL",
        30.0, 30.0, 0.05, 64, 2048, "lidar", -24.8, 2.0, 25.0,
        cfg_overrides=dict(_LIDAR_FILTER, range_max=30.0)),
    # configs[2]: RGB-D + P2.  ("stereo" does not exist in the reference: SensorType::RGBD)
    "c3_rgbd_p2": Workload(
        "c3_rgbd_p2", "RGB-D 640x480 307K pts + RGB, 20x20 m @ 0.05 m, P2 quantile, RGBD model, LOCAL",
        20.0, 20.0, 0.05, 480, 640, "rgbd", has_intensity=False, has_color=True,
        cfg_overrides=dict(range_min=0.3, range_max=10.0, sensor_type=capi.SENSOR_RGBD,
                           estimation_type=capi.EST_P2QUANTILE)),
    # configs[3]: dense LiDAR + raycasting
    "c4_dense_raycast": Workload(
        "c4_dense_raycast", "dense LiDAR 1.05M pts, 50x50 m @ 0.05 m, Kalman, raycasting on, LOCAL",
        50.0, 50.0, 0.05, 128, 8192, "lidar", -25.0, 15.0, 30.0, ghost_fraction=0.02,
        cfg_overrides=dict(_LIDAR_FILTER, range_max=40.0, raycasting_enabled=1)),
    # configs[4]: global map, row-striped over GPUs
    "c5_global": Workload(
        "c5_global", "global map 400x400 m @ 0.05 m (64M cells), 1.05M pts/scan, Kalman, GLOBAL",
        400.0, 400.0, 0.05, 128, 8192, "lidar", -25.0, 15.0, 30.0, loop_radius=48.0,
        cfg_overrides=dict(_LIDAR_FILTER, range_max=40.0, mode=capi.MODE_GLOBAL)),
    # small case for CPU-side and quick GPU tests
    "tiny": Workload(
        "tiny", "tiny LiDAR 2K pts, 10x10 m @ 0.1 m", 10.0, 10.0, 0.1, 8, 256, "lidar",
        -20.0, 5.0, 6.0, cfg_overrides=dict(_LIDAR_FILTER)),
}


def _rot_z(a: float) -> np.ndarray:
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def _rot_y(a: float) -> np.ndarray:
    c, s = math.cos(a), math.sin(a)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


def _iso(R: np.ndarray, t) -> np.ndarray:
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = t
    return T


def pose(wl: Workload, k: int):
    """(T_base_sensor, T_world_base) for scan k: 1 m/s at 10 Hz, 0.05 rad/s yaw (§8d)."""
    yaw = 0.005 * k
    if wl.loop_radius > 0:
        ang = (0.1 * k) / wl.loop_radius
        px, py = wl.loop_radius * math.cos(ang), wl.loop_radius * math.sin(ang)
        yaw = ang + math.pi / 2
    else:
        px, py = 0.1 * k, 0.02 * k
    pz = float(terrain(px, py)) + 0.5
    T_world_base = _iso(_rot_z(yaw), [px, py, pz])
    if wl.sensor == "rgbd":
        # optical frame (z fwd, x right, y down) -> base (x fwd, y left, z up), pitched 30 deg down
        R_opt = np.array([[0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [0.0, -1.0, 0.0]])
        T_base_sensor = _iso(_rot_y(math.radians(30.0)) @ R_opt, [0.1, 0.0, 1.0])
    else:
        T_base_sensor = _iso(np.eye(3), [0.0, 0.0, 0.5])
    return T_base_sensor, T_world_base


def _ray_ranges(origin, dirs_world, max_range, rng, wl: Workload, k: int):
    """range along each world-frame unit ray to the terrain or the cylinder wall."""
    dz = dirs_world[:, 2]
    ox, oy, oz = origin
    horiz = np.maximum(np.hypot(dirs_world[:, 0], dirs_world[:, 1]), 1e-6)
    t_wall = wl.wall_radius / horiz
    t = np.full(dirs_world.shape[0], np.inf)
    down = dz < -1e-3
    td = (oz + 0.5) / (-dz[down])
    for _ in range(4):  # fixed-point iterations on the height field
        gx = ox + td * dirs_world[down, 0]
        gy = oy + td * dirs_world[down, 1]
        td = (oz - terrain(gx, gy)) / (-dz[down])
    t[down] = td
    t = np.minimum(t, t_wall)
    if wl.ghost_fraction > 0:
        # a moving box: a narrow azimuth sector whose downward rays stop short (ghost source)
        az = np.arctan2(dirs_world[:, 1], dirs_world[:, 0])
        centre = ((0.37 * k) % (2 * math.pi)) - math.pi
        width = wl.ghost_fraction * 2 * math.pi
        d = np.abs(((az - centre + math.pi) % (2 * math.pi)) - math.pi)
        box = (d < width / 2) & down
        t_box = 4.0 / horiz
        hit_h = oz + t_box * dz
        box &= (hit_h < terrain(ox, oy) + 1.2) & (t_box < t)
        t = np.where(box, t_box, t)
    t = np.minimum(t, max_range)
    return t


def make_scan(wl: Workload, k: int):
    """-> dict(xyzw float32 [N,4], intensity float32 [N] | None, rgb uint8 [N,3] | None,
    T_base_sensor, T_world_base) for scan k of workload `wl`."""
    rng = np.random.RandomState(42 + k)
    Tbs, Twb = pose(wl, k)
    Tws = Twb @ Tbs
    origin = Tws[:3, 3]
    Rws = Tws[:3, :3]
    if wl.sensor == "lidar":
        elev = np.radians(np.linspace(wl.elev_min_deg, wl.elev_max_deg, wl.beams))
        az = np.linspace(-math.pi, math.pi, wl.azimuths, endpoint=False)
        E, A = np.meshgrid(elev, az, indexing="ij")
        d_s = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], axis=-1).reshape(-1, 3)
        d_w = d_s @ Rws.T
        t = _ray_ranges(origin, d_w, 60.0, rng, wl, k)
        t = t + rng.normal(0.0, 0.02, size=t.shape)  # range noise
        pts = d_s * t[:, None]
        # clip wall points to z <= 2 in the world frame (§8d) by shortening nothing: the height
        # filter drops them; keep geometry simple and deterministic
        intensity = rng.uniform(0.0, 1.0, size=t.shape).astype(np.float32) if wl.has_intensity else None
        rgb = None
    else:
        fx = fy = 525.0
        cx, cy = (wl.azimuths - 1) / 2.0, (wl.beams - 1) / 2.0
        v, u = np.meshgrid(np.arange(wl.beams), np.arange(wl.azimuths), indexing="ij")
        x = (u - cx) / fx
        y = (v - cy) / fy
        d_s = np.stack([x, y, np.ones_like(x)], axis=-1).reshape(-1, 3)
        norm = np.linalg.norm(d_s, axis=1)
        d_unit = d_s / norm[:, None]
        d_w = d_unit @ Rws.T
        t = _ray_ranges(origin, d_w, 60.0, rng, wl, k)
        depth = np.clip(t / norm, 0.3, 8.0)  # z-depth along the optical axis
        sigma = 0.001 + 0.002 * (depth - 0.4) ** 2  # RGBD model noise (rgbd_model.hpp:92-93)
        depth = depth + rng.normal(0.0, 1.0, size=depth.shape) * sigma
        pts = d_s * depth[:, None]
        intensity = None
        h = (u.astype(np.uint32) * 73856093) ^ (v.astype(np.uint32) * 19349663)
        h = h.reshape(-1)
        rgb = np.stack([h & 0xFF, (h >> 8) & 0xFF, (h >> 16) & 0xFF], axis=-1).astype(np.uint8) \
            if wl.has_color else None
    xyzw = np.empty((pts.shape[0], 4), np.float32)
    xyzw[:, :3] = pts.astype(np.float32)
    xyzw[:, 3] = 1.0
    return dict(xyzw=xyzw, intensity=intensity, rgb=rgb, T_base_sensor=Tbs, T_world_base=Twb)
