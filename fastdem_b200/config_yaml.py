"""YAML -> fastdem::Config, with the reference's parse / validate / clamp rules
(fastdem/src/config_fastdem.cpp:57-277) — a "next" row of SURVEY.md §8f.  Unknown enum
strings fall back with a warning, fatal inconsistencies raise (std::invalid_argument ->
ValueError, load failure std::runtime_error -> RuntimeError), everything else is clamped
with a warning.  Host-side code: no GPU involved."""
from __future__ import annotations

import logging
from pathlib import Path

import numpy as np
import yaml

from . import capi

log = logging.getLogger("fastdem_b200.config")
DEFAULT_YAML = Path(__file__).resolve().parent / "config" / "default.yaml"


def _load(node, key, cfg, attr, cast=float):
    if isinstance(node, dict) and key in node and node[key] is not None:
        setattr(cfg, attr, cast(node[key]))


def parseConfig(root) -> capi.FdemConfig:
    """detail::parse + detail::validate (config_fastdem.cpp:57-126, 128-260)."""
    cfg = capi.default_config()
    root = root or {}
    m = root.get("mapping")
    if isinstance(m, dict):
        mode = m.get("mode")
        if mode:
            if mode == "local":
                cfg.mode = capi.MODE_LOCAL
            elif mode == "global":
                cfg.mode = capi.MODE_GLOBAL
            else:
                log.warning("[Config] Unknown mapping mode '%s', defaulting to local", mode)
                cfg.mode = capi.MODE_LOCAL
        typ = m.get("type")
        if typ:
            if typ == "kalman_filter":
                cfg.estimation_type = capi.EST_KALMAN
            elif typ == "p2_quantile":
                cfg.estimation_type = capi.EST_P2QUANTILE
            else:
                log.warning("[Config] Unknown estimation type '%s', defaulting to kalman_filter", typ)
                cfg.estimation_type = capi.EST_KALMAN
        k = m.get("kalman")
        _load(k, "min_variance", cfg, "kalman_min_variance")
        _load(k, "max_variance", cfg, "kalman_max_variance")
        _load(k, "process_noise", cfg, "kalman_process_noise")
        p = m.get("p2")
        if isinstance(p, dict):
            for i in range(5):
                if p.get(f"dn{i}") is not None:
                    cfg.p2_dn[i] = float(p[f"dn{i}"])
            _load(p, "elevation_marker", cfg, "p2_elevation_marker", int)
            _load(p, "max_sample_count", cfg, "p2_max_sample_count")
    f = root.get("point_filter")
    for key in ("z_min", "z_max", "range_min", "range_max"):
        _load(f, key, cfg, key)
    r = root.get("raycasting")
    _load(r, "enabled", cfg, "raycasting_enabled", lambda v: 1 if bool(v) else 0)
    _load(r, "height_conflict_threshold", cfg, "rc_height_conflict_threshold")
    _load(r, "log_odds_observed", cfg, "rc_log_odds_observed")
    _load(r, "log_odds_ghost", cfg, "rc_log_odds_ghost")
    _load(r, "log_odds_max", cfg, "rc_log_odds_max")
    _load(r, "clear_threshold", cfg, "rc_clear_threshold")
    s = root.get("sensor_model")
    if isinstance(s, dict):
        typ = s.get("type")
        if typ:
            if typ in ("lidar", "laser"):
                cfg.sensor_type = capi.SENSOR_LIDAR
            elif typ == "rgbd":
                cfg.sensor_type = capi.SENSOR_RGBD
            elif typ in ("constant", "none"):
                cfg.sensor_type = capi.SENSOR_CONSTANT
            else:
                log.warning("[Config] Unknown sensor_model.type '%s', defaulting to LiDAR", typ)
                cfg.sensor_type = capi.SENSOR_LIDAR
        l = s.get("lidar")
        _load(l, "range_noise", cfg, "lidar_range_noise")
        _load(l, "angular_noise", cfg, "lidar_angular_noise")
        g = s.get("rgbd")
        _load(g, "normal_a", cfg, "rgbd_normal_a")
        _load(g, "normal_b", cfg, "rgbd_normal_b")
        _load(g, "normal_c", cfg, "rgbd_normal_c")
        _load(g, "lateral_factor", cfg, "rgbd_lateral_factor")
        c = s.get("constant")
        _load(c, "uncertainty", cfg, "constant_uncertainty")
    validate(cfg)
    return cfg


def validate(cfg: capi.FdemConfig) -> None:
    """detail::validate (config_fastdem.cpp:128-260): raise on fatal, warn + clamp otherwise."""
    if cfg.kalman_min_variance >= cfg.kalman_max_variance:
        raise ValueError(f"mapping.kalman: min_variance ({cfg.kalman_min_variance}) >= max_variance "
                         f"({cfg.kalman_max_variance})")

    def reset(attr, bad, value, msg):
        if bad(getattr(cfg, attr)):
            log.warning("[Config] %s (%s) %s, clamping to %s", attr, getattr(cfg, attr), msg, value)
            setattr(cfg, attr, value)

    if cfg.raycasting_enabled:
        reset("rc_height_conflict_threshold", lambda v: v <= 0.0, 0.05, "must be > 0")
        reset("rc_log_odds_observed", lambda v: v <= 0.0, 0.4, "must be > 0")
        reset("rc_log_odds_ghost", lambda v: v <= 0.0, 0.2, "must be > 0")
        reset("rc_log_odds_max", lambda v: v <= 0.0, 2.0, "must be > 0")
        reset("rc_clear_threshold", lambda v: v >= 0.0, -1.0, "must be < 0")
    reset("kalman_min_variance", lambda v: v <= 0.0, 0.0001, "must be > 0")
    reset("kalman_process_noise", lambda v: v < 0.0, 0.0, "must be >= 0")
    if cfg.p2_elevation_marker < 0 or cfg.p2_elevation_marker > 4:
        log.warning("[Config] mapping.p2.elevation_marker (%d) out of range [0, 4], clamping",
                    cfg.p2_elevation_marker)
        cfg.p2_elevation_marker = min(max(cfg.p2_elevation_marker, 0), 4)
    for i in range(5):
        if cfg.p2_dn[i] < 0.0 or cfg.p2_dn[i] > 1.0:
            log.warning("[Config] mapping.p2.dn%d (%s) out of [0, 1], clamping", i, cfg.p2_dn[i])
            cfg.p2_dn[i] = min(max(cfg.p2_dn[i], 0.0), 1.0)
    dn = list(cfg.p2_dn)
    if any(dn[i] > dn[i + 1] for i in range(4)):
        raise ValueError(f"mapping.p2: markers must be sorted (dn0 <= dn1 <= dn2 <= dn3 <= dn4), got {dn}")
    reset("lidar_range_noise", lambda v: v <= 0.0, 0.02, "must be > 0")
    reset("lidar_angular_noise", lambda v: v < 0.0, 0.0, "must be >= 0")
    reset("constant_uncertainty", lambda v: v <= 0.0, 0.1, "must be > 0")
    reset("rgbd_normal_a", lambda v: v < 0.0, 0.0, "must be >= 0")
    reset("rgbd_normal_b", lambda v: v < 0.0, 0.0, "must be >= 0")
    reset("rgbd_normal_c", lambda v: v < 0.0, 0.0, "must be >= 0")
    reset("rgbd_lateral_factor", lambda v: v < 0.0, 0.0, "must be >= 0")


def loadConfig(path) -> capi.FdemConfig:
    """fastdem::loadConfig (config_fastdem.cpp:270-277)."""
    try:
        with open(path, "r") as f:
            root = yaml.safe_load(f)
    except (OSError, yaml.YAMLError) as e:
        raise RuntimeError(f"Failed to load config: {path} - {e}") from e
    return parseConfig(root)
