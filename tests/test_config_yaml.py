"""fastdem/tests/test_config.cpp (ConfigLoadTest / ConfigValidationTest) against
fastdem_b200.config_yaml — same YAML snippets, same expectations.  CPU only."""
import numpy as np
import pytest

from fastdem_b200 import capi, config_yaml as cy

F = np.float32


def write(tmp_path, text):
    p = tmp_path / "cfg.yaml"
    p.write_text(text)
    return str(p)


def test_load_default_yaml():  # :36-43
    cfg = cy.loadConfig(cy.DEFAULT_YAML)
    assert cfg.estimation_type == capi.EST_KALMAN and cfg.sensor_type == capi.SENSOR_LIDAR
    assert cfg.raycasting_enabled == 1
    assert (cfg.z_min, cfg.z_max, cfg.range_min, cfg.range_max) == (F(-1.0), F(2.0), F(0.5), F(20.0))


def test_nonexistent_file_throws():  # :45-47
    with pytest.raises(RuntimeError):
        cy.loadConfig("/nonexistent/path.yaml")


def test_empty_and_partial_yaml_keep_defaults(tmp_path):  # :49-74
    d = capi.default_config()
    cfg = cy.loadConfig(write(tmp_path, "{}\n"))
    assert bytes(cfg) == bytes(d)
    cfg = cy.loadConfig(write(tmp_path, "mapping:\n  type: p2_quantile\n"))
    assert cfg.estimation_type == capi.EST_P2QUANTILE and cfg.mode == d.mode
    assert cfg.lidar_range_noise == d.lidar_range_noise


def test_all_enum_spellings(tmp_path):  # :76-113
    for s, want in (("kalman_filter", capi.EST_KALMAN), ("p2_quantile", capi.EST_P2QUANTILE), ("bogus", capi.EST_KALMAN)):
        assert cy.loadConfig(write(tmp_path, f"mapping:\n  type: {s}\n")).estimation_type == want
    for s, want in (("lidar", 1), ("laser", 1), ("rgbd", 2), ("constant", 0), ("none", 0), ("bogus", 1)):
        assert cy.loadConfig(write(tmp_path, f"sensor_model:\n  type: {s}\n")).sensor_type == want
    assert cy.loadConfig(write(tmp_path, "mapping:\n  mode: global\n")).mode == capi.MODE_GLOBAL
    assert cy.loadConfig(write(tmp_path, "mapping:\n  type: kalman_filter\n")).mode == capi.MODE_LOCAL


def test_kalman_and_point_filter_parsed(tmp_path):  # :115-158
    cfg = cy.loadConfig(write(tmp_path, "mapping:\n  type: kalman_filter\n  kalman:\n    min_variance: 0.001\n"
                                        "    max_variance: 0.05\n    process_noise: 0.001\n"))
    assert (cfg.kalman_min_variance, cfg.kalman_max_variance, cfg.kalman_process_noise) == (F(0.001), F(0.05), F(0.001))
    cfg = cy.loadConfig(write(tmp_path, "point_filter:\n  z_min: -0.5\n  z_max: 2.0\n  range_min: 0.5\n  range_max: 20.0\n"))
    assert (cfg.z_min, cfg.z_max, cfg.range_min, cfg.range_max) == (F(-0.5), F(2.0), F(0.5), F(20.0))
    d = capi.default_config()
    cfg = cy.loadConfig(write(tmp_path, "mapping:\n  type: kalman_filter\n"))
    assert (cfg.z_min, cfg.z_max, cfg.range_min, cfg.range_max) == (d.z_min, d.z_max, d.range_min, d.range_max)


def test_fatal_validation_throws(tmp_path):  # :160-181
    with pytest.raises(ValueError):
        cy.loadConfig(write(tmp_path, "mapping:\n  kalman:\n    min_variance: 0.1\n    max_variance: 0.01\n"))
    with pytest.raises(ValueError):
        cy.loadConfig(write(tmp_path, "mapping:\n  p2:\n    dn0: 0.5\n    dn1: 0.16\n"))


def test_clamping(tmp_path):  # :185-224, :335-344
    assert cy.loadConfig(write(tmp_path, "sensor_model:\n  lidar:\n    range_noise: -0.5\n")).lidar_range_noise == F(0.02)
    assert cy.loadConfig(write(tmp_path, "sensor_model:\n  lidar:\n    angular_noise: -0.5\n")).lidar_angular_noise == 0.0
    assert cy.loadConfig(write(tmp_path, "sensor_model:\n  constant:\n    uncertainty: -1\n")).constant_uncertainty == F(0.1)
    assert cy.loadConfig(write(tmp_path, "mapping:\n  kalman:\n    process_noise: -0.01\n")).kalman_process_noise == 0.0
    assert cy.loadConfig(write(tmp_path, "mapping:\n  p2:\n    elevation_marker: 9\n")).p2_elevation_marker == 4
    assert cy.loadConfig(write(tmp_path, "mapping:\n  p2:\n    elevation_marker: -3\n")).p2_elevation_marker == 0
    cfg = cy.loadConfig(write(tmp_path, "raycasting:\n  enabled: true\n  log_odds_ghost: -1\n  clear_threshold: 0.5\n"))
    assert cfg.rc_log_odds_ghost == F(0.2) and cfg.rc_clear_threshold == -1.0
    cfg = cy.loadConfig(write(tmp_path, "raycasting:\n  enabled: false\n  log_odds_ghost: -1\n"))
    assert cfg.rc_log_odds_ghost == -1.0     # raycasting checks only run when it is enabled
