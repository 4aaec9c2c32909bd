"""Pins the CPU oracle (oracle/fdem_oracle.hpp) against every known-answer value the
REFERENCE's own tests assert for the integrate() path.  Each test names the
reference test it ports (paths relative to /root/reference/fastdem/tests/ unless noted).
No GPU needed."""
import math

import numpy as np
import pytest

import oracle_binding as ob

NAN = float("nan")


def kalman_cfg(**kw):
    c = ob.default_config()
    c.mode = 1  # GLOBAL
    c.estimation_type = 0
    for k, v in kw.items():
        setattr(c, k, v)
    return c


# ───────────────────────── Kalman (test_kalman_estimation.cpp) ─────────────────────────

def fresh_state():
    # elevation NaN, P 0, count 0, sample_mean NaN, variance 0, m2 0 (ensureLayers fills)
    return np.array([NAN, 0, 0, NAN, 0, 0], np.float32)


def test_kalman_first_measurement_initializes():  # :18-28
    s, b = ob.kalman_step(fresh_state(), 5.0, 0.04, 1e-4, 1e-2, 0.0)
    assert s[0] == np.float32(5.0)
    assert s[1] == np.float32(0.04)
    assert s[2] == 1.0


def test_kalman_p_reduced_and_clamped():  # :30-62
    s, _ = ob.kalman_step(fresh_state(), 5.0, 0.5, 1e-4, 1.0, 0.0)
    p0 = s[1]
    for _ in range(20):
        s, _ = ob.kalman_step(s, 5.0, 0.01, 1e-4, 1.0, 0.0)
    assert s[1] < p0
    s, _ = ob.kalman_step(fresh_state(), 5.0, 0.05, 1e-3, 0.1, 0.0)
    for _ in range(100):
        s, _ = ob.kalman_step(s, 5.0, 1e-4, 1e-3, 0.1, 0.0)
    assert np.float32(1e-3) <= s[1] <= np.float32(0.1)


def test_kalman_bounds_are_two_sigma():  # :64-80
    s = fresh_state()
    rng = np.random.RandomState(0)
    for z in 5.0 + 0.1 * rng.randn(20):
        s, b = ob.kalman_step(s, float(z), 0.01)
    sigma = math.sqrt(max(0.0, float(s[4])))
    assert abs(b[0] - (s[0] + 2 * sigma)) < 1e-5
    assert abs(b[1] - (s[0] - 2 * sigma)) < 1e-5


def test_kalman_zero_variance_uses_max_variance():  # :82-90
    s, _ = ob.kalman_step(fresh_state(), 2.0, 0.0, 1e-4, 1e-2, 0.0)
    assert s[1] == np.float32(1e-2)


def test_kalman_tracks_after_outlier():  # :92-104
    s, _ = ob.kalman_step(fresh_state(), 10.0, 0.01)
    for _ in range(50):
        s, _ = ob.kalman_step(s, 5.0, 0.01)
    assert abs(s[0] - 5.0) < 0.1


def test_kalman_sample_variance_of_3_and_7_is_8():  # :106-119
    s, _ = ob.kalman_step(fresh_state(), 3.0, 0.01)
    s, _ = ob.kalman_step(s, 7.0, 0.01)
    assert s[4] == np.float32(8.0)
    assert s[2] == 2.0


def test_kalman_identical_measurements():  # :121-139 (Kalman(1e-4, 1.0, 0))
    s = fresh_state()
    for _ in range(50):
        s, _ = ob.kalman_step(s, 5.0, 0.01, 1e-4, 1.0, 0.0)
    assert abs(s[1] - 1e-4) < 1e-3  # EXPECT_NEAR(kalman_p, 0.0001f, 0.001f)
    assert abs(s[4]) < 1e-6         # sample variance ~ 0 (NOT the Kalman P)


# ───────────────────────── P2 (test_quantile_estimation.cpp) ───────────────────────────

def p2_fresh():
    return np.full(5, NAN, np.float32), np.arange(5, dtype=np.float32), 0.0


def test_p2_count_after_three():  # :40-46
    q, n, c = p2_fresh()
    for x in (1.0, 2.0, 3.0):
        q, n, c, _ = ob.p2_step(q, n, c, x)
    assert c == 3.0


def test_p2_five_samples_sorted():  # :48-67
    q, n, c = p2_fresh()
    for x in (5.0, 3.0, 1.0, 4.0, 2.0):
        q, n, c, _ = ob.p2_step(q, n, c, x)
    assert c == 5.0
    assert list(q) == [1.0, 2.0, 3.0, 4.0, 5.0]
    assert list(n) == [0.0, 1.0, 2.0, 3.0, 4.0]


def test_p2_markers_monotone_uniform():  # :69-87 (mt19937(42) uniform[0,10) x100)
    rng = np.random.RandomState(42)
    q, n, c = p2_fresh()
    for x in rng.uniform(0, 10, 100):
        q, n, c, _ = ob.p2_step(q, n, c, float(x))
    assert all(q[i] <= q[i + 1] for i in range(4))
    assert c == 100.0


def test_p2_median_of_gaussian():  # :89-103
    rng = np.random.RandomState(42)
    q, n, c = p2_fresh()
    for x in rng.normal(5.0, 1.0, 1000):
        q, n, c, _ = ob.p2_step(q, n, c, float(x))
    assert abs(q[2] - 5.0) < 0.2


def test_p2_update_elevation_before_and_after_five():  # :121-148
    q, n, c = p2_fresh()
    for i, x in enumerate((1.0, 2.0, 3.0, 4.0)):
        q, n, c, e = ob.p2_step(q, n, c, x)
        assert e == x  # update() alone writes the latest sample while count < 5
    q, n, c, e = ob.p2_step(q, n, c, 5.0)
    assert e == q[3]   # marker 3 once initialised


def test_p2_count_nan_guard():  # quantile_estimation.hpp:183 (clearAll() leaves NaN counters)
    q, n = np.full(5, NAN, np.float32), np.full(5, NAN, np.float32)
    q, n, c, e = ob.p2_step(q, n, NAN, 2.5)
    assert c == 1.0 and q[0] == 2.5


# ───────────────────────── sensor models (test_sensor_models.cpp) ──────────────────────

def test_lidar_point_on_x_axis():  # :113-129
    c = ob.default_config()
    c.sensor_type = 1
    cov = ob.sensor_cov(c, [10.0, 0.0, 0.0])
    assert abs(cov[0, 0] - 4e-4) < 1e-6   # sigma_r^2
    assert abs(cov[1, 1] - 1e-4) < 1e-6   # (10 * 0.001)^2
    assert abs(cov[2, 2] - 1e-4) < 1e-6
    assert abs(cov[0, 1]) < 1e-9 and abs(cov[0, 2]) < 1e-9


def test_lidar_zero_point_fallback():  # :104-111
    c = ob.default_config()
    cov = ob.sensor_cov(c, [0.0, 0.0, 0.0])
    assert np.allclose(cov, 0.01 * np.eye(3))


def test_lidar_eigenvalues_on_diagonal_ray():  # :148-159
    c = ob.default_config()
    p = np.array([3.0, 3.0, 3.0], np.float32)
    cov = ob.sensor_cov(c, p).astype(np.float64)
    w = np.sort(np.linalg.eigvalsh((cov + cov.T) / 2))
    d = float(np.linalg.norm(p))
    v_l = max((d * 0.001) ** 2, 1e-6)
    v_r = max(0.02 ** 2, 1e-6)
    assert np.allclose(w, sorted([v_l, v_l, v_r]), rtol=1e-4)


def test_lidar_negative_params_abs():  # lidar_model.hpp:59-62
    c = ob.default_config()
    a = ob.sensor_cov(c, [4.0, 1.0, -2.0])
    c.lidar_range_noise, c.lidar_angular_noise = -0.02, -0.001
    b = ob.sensor_cov(c, [4.0, 1.0, -2.0])
    assert np.array_equal(a, b)


def test_rgbd_at_optimal_depth():  # :250-254
    c = ob.default_config()
    c.sensor_type = 2
    cov = ob.sensor_cov(c, [0.0, 0.0, c.rgbd_normal_c])
    assert abs(cov[2, 2] - c.rgbd_normal_a ** 2) < 1e-10


def test_rgbd_nonpositive_depth_fallback():  # :222-228, :242-248
    c = ob.default_config()
    c.sensor_type = 2
    for z in (0.0, -1.0):
        assert np.allclose(ob.sensor_cov(c, [0.1, 0.2, z]), 0.01 * np.eye(3))


def test_rgbd_diagonal_and_lateral_scaling():  # :199-207, :230-240
    c = ob.default_config()
    c.sensor_type = 2
    a = ob.sensor_cov(c, [0.3, -0.2, 1.0])
    b = ob.sensor_cov(c, [0.3, -0.2, 2.0])
    assert a[0, 1] == 0 and a[0, 2] == 0 and a[1, 2] == 0
    assert abs(b[0, 0] / a[0, 0] - 4.0) < 1e-4   # var_lat ∝ depth^2
    assert a[0, 0] == a[1, 1]


def test_constant_model():  # sensor_model.hpp:87-93
    c = ob.default_config()
    c.sensor_type = 0
    c.constant_uncertainty = 0.1
    assert np.allclose(ob.sensor_cov(c, [1, 2, 3]), np.float32(0.1) * np.float32(0.1) * np.eye(3))


# ───────────────────────── transform / crop (nanoPCL tests) ────────────────────────────

def test_transform_identity_translation_rotation():  # lib/nanoPCL/tests/test_transform.cpp:25-165
    pts = np.array([[1, 2, 3], [-1, 0.5, 2]], np.float32)
    T = np.eye(4)
    assert np.allclose(ob.transform(T, pts)[:, :3], pts, atol=1e-5)
    T[:3, 3] = [1, -2, 0.5]
    assert np.allclose(ob.transform(T, pts)[:, :3], pts + [1, -2, 0.5], atol=1e-5)
    Rz = np.eye(4)
    Rz[:2, :2] = [[0, -1], [1, 0]]  # +90 deg about z: (1,0,0) -> (0,1,0)
    out = ob.transform(Rz, np.array([[1, 0, 0]], np.float32))
    assert np.allclose(out[0, :3], [0, 1, 0], atol=1e-5)
    assert out[0, 3] == 1.0


def test_preprocess_filters_in_base_frame_and_stable_order():  # fastdem.cpp:164-190
    c = ob.default_config()
    c.sensor_type = 0
    c.range_min, c.range_max = 0.5, 5.0
    c.z_min, c.z_max = -10.0, 10.0
    pts = np.array([[0.1, 0, 0], [1, 0, 0], [3, 0, 0], [6, 0, 0], [2, 2, 0]], np.float32)
    Tbs = np.eye(4)
    Twb = np.eye(4)
    Twb[:3, 3] = [10, 0, 0]  # world offset must NOT affect the range test
    out, cov, src = ob.preprocess(c, pts, Tbs, Twb)
    assert list(src) == [1, 2, 4]            # cropRange keeps [0.5, 5], order preserved
    assert np.allclose(out[:, 0], [11, 13, 12])
    # NaN points fail both compares and are dropped
    pts2 = np.array([[NAN, 0, 0], [1, 0, 0]], np.float32)
    out, cov, src = ob.preprocess(c, pts2, Tbs, Twb)
    assert list(src) == [1]


def test_preprocess_cov_rotation():  # fastdem.cpp:181-187
    c = ob.default_config()
    c.sensor_type = 1
    p = np.array([[5.0, 0.0, 0.0]], np.float32)
    Twb = np.eye(4)
    Twb[:3, :3] = [[0, 0, 1], [0, 1, 0], [-1, 0, 0]]  # x -> -z : radial axis becomes vertical
    out, cov, src = ob.preprocess(c, p, np.eye(4), Twb)
    C = cov[0].reshape(3, 3).T
    assert abs(C[2, 2] - 4e-4) < 1e-7   # sigma_r^2 now on z
    assert abs(C[0, 0] - 2.5e-5) < 1e-8  # (5 * 0.001)^2


# ───────────────────────── grid (test_elevation_map.cpp + Appendix A) ─────────────────

def test_grid_fixture_20x20_and_round_trips():  # test_postprocess.cpp:28, test_elevation_map.cpp:40-49
    m = ob.OracleMap(10.0, 10.0, 0.5)
    g = m.geometry()
    assert (g["rows"], g["cols"]) == (20, 20)
    assert m.isEmpty()
    for pos in [(1.0, 1.0), (0.0, 0.0), (-2.0, -2.0), (4.9, -4.9)]:
        ok, idx = m.getIndex(pos)
        assert ok
        x, y = m.getPosition(idx)
        assert abs(x - pos[0]) <= 0.25 + 1e-9 and abs(y - pos[1]) <= 0.25 + 1e-9
        assert m.getIndex((x, y)) == (True, idx)
    assert not m.getIndex((100.0, 100.0))[0]   # :30-33 out of bounds


def test_grid_row_axis_is_minus_x_and_column_major():  # feature_extraction.cpp:73-77, io_npz.cpp:142-144
    m = ob.OracleMap(10.0, 10.0, 0.5)
    _, a = m.getIndex((4.9, 4.9))
    _, b = m.getIndex((-4.9, -4.9))
    assert a == (0, 0) and b == (19, 19)
    m.setAt("elevation", (3, 7), 42.0)
    flat = m.get("elevation").reshape(-1, order="F")
    assert flat[7 * 20 + 3] == 42.0


def test_grid_move_wraps_and_clears():  # Appendix A; test_fastdem_integration.cpp:198-215
    m = ob.OracleMap(10.0, 10.0, 0.5)
    m.add("n_points", 0.0)
    e = np.arange(400, dtype=np.float32).reshape(20, 20, order="F")
    m.set("elevation", e)
    assert m.move((1.0, 0.0))          # +2 cells in x  => buffer shift -2 rows
    g = m.geometry()
    assert g["position"] == (1.0, 0.0) and g["start_index"] == (18, 0)
    cur = m.get("elevation")
    assert np.isnan(cur[18:20, :]).all() and not np.isnan(cur[:18, :]).any()
    assert np.isnan(m.get("n_points")[18:20, :]).all()   # ALL layers (default policy)
    # same physical location keeps its value: (0.25, 0.25) was logical (9,9)
    ok, idx = m.getIndex((0.25, 0.25))
    assert ok and cur[idx] == e[9, 9]
    # a 100 m jump empties the map and puts the old origin outside
    assert m.move((100.0, 0.0))
    assert m.isEmpty() and not m.isInside((0.0, 0.0))


def test_grid_move_basic_policy_keeps_internal_layers():
    m = ob.OracleMap(10.0, 10.0, 0.5)
    m.add("n_points", 3.0)
    m.set("elevation", np.ones((20, 20), np.float32))
    m.move((0.0, -0.5), policy=1)
    assert np.isnan(m.get("elevation")).sum() == 20
    assert not np.isnan(m.get("n_points")).any()


def test_color_packing():  # test_rasterization.cpp:107-127
    import ctypes
    v = ob.lib().orc_pack_color(0x12, 0x34, 0x56)
    assert np.float32(v).view(np.uint32) == 0x123456


# ───────────────────────── dual layer (test_dual_layer.cpp) ───────────────────────────

def make_mapping(est=0):
    m = ob.OracleMap(10.0, 10.0, 0.5)
    c = kalman_cfg(kalman_min_variance=1e-4, kalman_max_variance=1.0, estimation_type=est)
    return m, ob.OracleFastDEM(m, c)


def test_dual_ground_obstacle_separation():  # :66-83
    m, d = make_mapping()
    d.update([[0, 0, 0.0], [0, 0, 3.0]], (0, 0))
    _, idx = m.getIndex((0, 0))
    assert abs(m.at("elevation", idx)) < 0.1
    assert abs(m.at("obstacle", idx) - 3.0) < 0.1


def test_dual_single_point_only_ground():  # :107-119
    m, d = make_mapping()
    d.update([[0, 0, 2.0]], (0, 0))
    _, idx = m.getIndex((0, 0))
    assert abs(m.at("elevation", idx) - 2.0) < 0.1
    assert math.isnan(m.at("obstacle", idx))


def test_dual_second_frame_obstacle_exact():  # :121-143
    m, d = make_mapping()
    d.update([[0, 0, 0.0], [0, 0, 3.0]], (0, 0))
    d.update([[0, 0, 0.1], [0, 0, 3.1]], (0, 0))
    _, idx = m.getIndex((0, 0))
    assert -0.05 < m.at("elevation", idx) < 0.15
    assert np.float32(m.at("obstacle", idx)) == np.float32(3.1)


def test_dual_quantile():  # :145-165
    m, d = make_mapping(est=1)
    for i in range(10):
        noise = 0.05 if i % 2 == 0 else -0.05
        d.update([[0, 0, 0.0 + noise], [0, 0, 5.0 + noise]], (0, 0))
    _, idx = m.getIndex((0, 0))
    assert abs(m.at("elevation", idx)) < 0.5
    assert abs(m.at("obstacle", idx) - 5.0) < 0.1


def test_dual_elevation_max_true_max():  # :167-186
    m, d = make_mapping()
    _, idx = m.getIndex((0, 0))
    d.update([[0, 0, 0.0], [0, 0, 3.0]], (0, 0))
    assert m.at("elevation_max", idx) == 3.0
    d.update([[0, 0, 0.0], [0, 0, 5.0]], (0, 0))
    assert m.at("elevation_max", idx) == 5.0


def test_dual_obstacle_clears_when_flat():  # :188-203
    m, d = make_mapping()
    _, idx = m.getIndex((0, 0))
    d.update([[0, 0, 0.0], [0, 0, 2.0]], (0, 0))
    assert m.at("obstacle", idx) == 2.0
    d.update([[0, 0, 0.0]], (0, 0))
    assert math.isnan(m.at("obstacle", idx))


def test_rasterize_three_nearby_points_share_a_cell():  # test_rasterization.cpp:33-43,159-174
    m = ob.OracleMap(10.0, 10.0, 1.0)
    d = ob.OracleFastDEM(m, kalman_cfg())
    n = d.update([[0.1, 0.1, 1.0], [0.2, 0.3, 2.0], [0.4, 0.2, 3.0]], (0, 0))
    assert n == 1
    _, idx = m.getIndex((0.25, 0.25))
    assert m.at("elevation_min", idx) == 1.0 and m.at("elevation_max", idx) == 3.0


# ───────────────────────── pipeline (test_fastdem_integration.cpp) ─────────────────────

def ground_cloud(height, half=3, spacing=0.3):  # :32-41
    g = np.arange(-half, half + 1, dtype=np.float32) * np.float32(spacing)
    xx, yy = np.meshgrid(g, g, indexing="ij")
    return np.stack([xx.ravel(), yy.ravel(), np.full(xx.size, height, np.float32)], axis=1)


def pipeline(**kw):
    m = ob.OracleMap(10.0, 10.0, 0.5)
    c = ob.default_config()
    for k, v in kw.items():
        setattr(c, k, v)
    return m, ob.OracleFastDEM(m, c)


def test_integrate_updates_elevation():  # :45-59
    m, d = pipeline(z_min=-2, z_max=5, range_min=0, range_max=20, sensor_type=0)
    ok, st, _ = d.integrate(ground_cloud(1.0), np.eye(4), np.eye(4))
    assert ok
    _, idx = m.getIndex((0, 0))
    assert abs(m.at("elevation", idx) - 1.0) < 0.1


def test_empty_and_all_filtered_return_false():  # :61-69, :357-378
    m, d = pipeline()
    ok, st, _ = d.integrate(np.zeros((0, 3), np.float32), np.eye(4), np.eye(4))
    assert not ok and m.isEmpty()
    m, d = pipeline(z_min=100.0, z_max=200.0)
    ok, st, _ = d.integrate(ground_cloud(1.0), np.eye(4), np.eye(4))
    assert not ok and st.n_kept == 0 and m.isEmpty()


def test_height_and_range_filters():  # :71-80, :287-316
    m, d = pipeline(z_min=0.0, z_max=2.0)
    d.integrate(ground_cloud(10.0), np.eye(4), np.eye(4))
    assert m.isEmpty()
    m, d = pipeline(range_min=5.0, range_max=20.0)
    d.integrate(ground_cloud(1.0, half=2), np.eye(4), np.eye(4))
    assert m.isEmpty()


def test_multiple_integrations_blend():  # :82-104
    m, d = pipeline(z_min=-5, z_max=15, range_max=20, sensor_type=0)
    d.integrate(ground_cloud(1.0), np.eye(4), np.eye(4))
    d.integrate(ground_cloud(1.5), np.eye(4), np.eye(4))
    _, idx = m.getIndex((0, 0))
    assert 0.9 < m.at("elevation", idx) < 1.6


def test_p2_needs_five_scans():  # :159-175 + Appendix C.6
    m, d = pipeline(z_min=-5, z_max=15, range_max=20, sensor_type=0, estimation_type=1)
    _, idx = m.getIndex((0, 0))
    for i in range(6):
        d.integrate(ground_cloud(1.0 + i * 0.01), np.eye(4), np.eye(4))
        e = m.at("elevation", idx)
        if i < 3:
            assert math.isnan(e)      # q[3] not written yet (computeBounds overwrite quirk)
        else:
            assert abs(e - 1.0) < 0.2


def test_global_fixed_local_follows():  # :179-215
    m, d = pipeline(mode=1, z_min=-5, z_max=15, sensor_type=0)
    d.integrate(ground_cloud(1.0), np.eye(4), np.eye(4))
    T = np.eye(4)
    T[0, 3] = 3.0
    d.integrate(ground_cloud(2.0), np.eye(4), T)
    _, idx = m.getIndex((0, 0))
    assert np.isfinite(m.at("elevation", idx))
    m, d = pipeline(mode=0, z_min=-5, z_max=15, sensor_type=0)
    d.integrate(ground_cloud(1.0), np.eye(4), np.eye(4))
    T[0, 3] = 100.0
    d.integrate(ground_cloud(2.0), np.eye(4), T)
    assert not m.isInside((0.0, 0.0))


def test_sensor_offset_applied():  # :253-268
    m, d = pipeline(z_min=-5, z_max=15, sensor_type=0)
    Tbs = np.eye(4)
    Tbs[2, 3] = 1.0
    d.integrate(ground_cloud(0.0), Tbs, np.eye(4))
    _, idx = m.getIndex((0, 0))
    assert abs(m.at("elevation", idx) - 1.0) < 0.2


def test_all_filtered_does_not_move_local_map():  # Appendix C.2 (fastdem.cpp:137-138)
    m, d = pipeline(z_min=100.0, z_max=200.0)
    T = np.eye(4)
    T[0, 3] = 3.0
    ok, _, _ = d.integrate(ground_cloud(1.0), np.eye(4), T)
    assert not ok and m.geometry()["position"] == (0.0, 0.0)


def test_points_outside_map_move_but_do_not_reset_obstacle():  # Appendix C.3
    m, d = pipeline(z_min=-5, z_max=15, sensor_type=0)
    d.integrate(np.array([[0, 0, 0.0], [0, 0, 2.0]], np.float32), np.eye(4), np.eye(4))
    _, idx = m.getIndex((0, 0))
    assert m.at("obstacle", idx) == 2.0
    far = np.array([[50.0, 50.0, 0.0]], np.float32)   # survives filters, outside the map
    ok, st, _ = d.integrate(far, np.eye(4), np.eye(4))
    assert ok and st.n_cells == 0
    assert m.at("obstacle", idx) == 2.0              # no whole-layer clear happened


# ───────────────────────── raycasting (test_postprocess.cpp) ───────────────────────────

def rc_cfg(**kw):
    c = ob.default_config()
    c.raycasting_enabled = 1
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def test_raycast_layers_created_and_disabled_noop():  # :71-89, :177-189
    m = ob.OracleMap(10.0, 10.0, 0.5)
    c = ob.default_config()
    m.raycast([[1.0, 0.0, 0.5]], (0, 0, 5), c)  # disabled
    assert not m.exists("raycasting")
    m.raycast([[1.0, 0.0, 0.5]], (0, 0, 5), rc_cfg())
    for name in ("ghost_removal", "raycasting", "_visibility_logodds"):
        assert m.exists(name)


def test_raycast_clears_ghost_cell():  # :92-115
    m = ob.OracleMap(10.0, 10.0, 0.5)
    _, g = m.getIndex((2.0, 0.0))
    m.setAt("elevation", g, 10.0)
    m.raycast([[4.0, 0.0, 0.0]], (0, 0, 5), rc_cfg(rc_log_odds_ghost=0.5, rc_clear_threshold=-0.4))
    assert math.isnan(m.at("elevation", g))
    assert m.at("ghost_removal", g) == 1.0


def test_raycast_observed_cell_protected():  # :117-144
    m = ob.OracleMap(10.0, 10.0, 0.5)
    _, g = m.getIndex((2.0, 0.0))
    m.setAt("elevation", g, 2.0)
    m.raycast([[4.0, 0.0, 0.0], [2.0, 0.0, 0.3]], (0, 0, 5),
              rc_cfg(rc_log_odds_observed=0.8, rc_log_odds_ghost=0.5, rc_clear_threshold=-0.4))
    assert not math.isnan(m.at("elevation", g))
    assert abs(m.at("_visibility_logodds", g) - 0.3) < 1e-6


def test_raycast_ghost_requires_accumulation():  # :146-175
    m = ob.OracleMap(10.0, 10.0, 0.5)
    _, g = m.getIndex((2.0, 0.0))
    m.setAt("elevation", g, 10.0)
    cfg = rc_cfg(rc_log_odds_ghost=0.2, rc_clear_threshold=-0.9)
    for _ in range(4):
        m.raycast([[4.0, 0.0, 0.0]], (0, 0, 5), cfg)
    assert not math.isnan(m.at("elevation", g))
    m.raycast([[4.0, 0.0, 0.0]], (0, 0, 5), cfg)
    assert math.isnan(m.at("elevation", g))


def test_raycast_sensor_outside_map_is_noop():  # raycasting.cpp:230-234
    m = ob.OracleMap(10.0, 10.0, 0.5)
    m.raycast([[1.0, 0.0, 0.0]], (50.0, 0.0, 5.0), rc_cfg())
    assert not m.exists("raycasting")


def test_voxel_any_bad_size_and_selection_rule():  # voxel_grid_impl.hpp:31-33, :171-172
    pts = np.array([[0.01, 0.01, 0.01], [0.02, 0.02, 0.02], [0.03, 0.01, 0.02], [5.0, 5.0, 5.0]], np.float32)
    with pytest.raises(ValueError):
        ob.voxel_any(pts, 0.0001)
    sel = ob.voxel_any(pts, 0.1)
    # voxel A = {0,1,2} starts at 0 -> (3*7 + 0*13) % 3 = 0 -> idx 0; voxel B = {3} at start 3
    assert list(sel) == [0, 3]


# ───────────────────────── inpainting (test_postprocess.cpp:41-69) ─────────────────────

def test_inpaint_fills_hole_from_eight_neighbours():
    m = ob.OracleMap(10.0, 10.0, 0.5)
    e = np.full((20, 20), NAN, np.float32)
    e[9:12, 9:12] = 1.0
    e[10, 10] = NAN
    m.set("elevation", e)
    m.inpaint(3, 2, False)
    assert abs(m.get("elevation_inpainted")[10, 10] - 1.0) < 0.01
    assert math.isnan(m.get("elevation")[10, 10])  # not in place


def test_spatial_smoothing_removes_spike():  # test_postprocess.cpp:249-266
    m = ob.OracleMap(10.0, 10.0, 0.5)
    e = np.full((20, 20), NAN, np.float32)
    e[8:13, 8:13] = 1.0
    e[10, 10] = 100.0
    m.set("elevation", e)
    ob.spatial_smoothing(m, "elevation", 3, 5)
    assert abs(m.get("elevation")[10, 10] - 1.0) < 0.01
    ob.spatial_smoothing(m, "nonexistent_layer")  # :268-271 must not crash


# ───────────── uncertainty fusion / feature extraction (test_postprocess.cpp:191-345) ─────────────
def _bounds_block(m):
    up = np.full((20, 20), NAN, np.float32)
    lo = np.full((20, 20), NAN, np.float32)
    el = np.full((20, 20), NAN, np.float32)
    for dr in (-1, 0, 1):
        for dc in (-1, 0, 1):
            h = np.float32(1.0) + np.float32(0.1) * dr
            el[10 + dr, 10 + dc] = h
            up[10 + dr, 10 + dc] = h + np.float32(0.2)
            lo[10 + dr, 10 + dc] = h - np.float32(0.2)
    m.set("elevation", el)
    m.add("upper_bound")
    m.add("lower_bound")
    m.set("upper_bound", up)
    m.set("lower_bound", lo)
    return up, lo


def test_uncertainty_fusion_computes_bounds():  # :193-225
    m = ob.OracleMap(10.0, 10.0, 0.5)
    up0, lo0 = _bounds_block(m)
    ob.uncertainty_fusion(m, 0.6, 0.3, 0.01, 0.99, 1)
    up, lo = m.get("upper_bound"), m.get("lower_bound")
    assert np.isfinite(up[10, 10]) and np.isfinite(lo[10, 10]) and up[10, 10] > lo[10, 10]
    # radius 0.6 m on a 0.5 m grid reaches the four edge neighbours: quantile 0.99 of the
    # upper bounds {1.1, 1.2, 1.2, 1.2, 1.3} is the largest, quantile 0.01 of the lower the smallest
    assert abs(up[10, 10] - 1.3) < 1e-6
    assert abs(lo[10, 10] - 0.7) < 1e-6
    # cells outside the block stay NaN
    assert np.isnan(up[5, 5]) and np.isnan(lo[5, 5])


def test_uncertainty_fusion_skips_missing_bounds():  # :227-237
    m = ob.OracleMap(10.0, 10.0, 0.5)
    ob.uncertainty_fusion(m)  # no bound layers: early return, nothing added
    assert not m.exists("upper_bound")


def test_feature_extraction_flat_tilted_step():  # :273-345
    m = ob.OracleMap(10.0, 10.0, 0.5)
    m.set("elevation", np.ones((20, 20), np.float32))
    ob.feature_extraction(m, 0.6, 4)
    for name in ("step", "slope", "roughness", "curvature", "_normal_x", "_normal_y", "_normal_z"):
        assert m.exists(name)
    assert abs(m.get("slope")[10, 10]) < 1.0
    assert abs(m.get("roughness")[10, 10]) < 1e-3
    assert abs(m.get("step")[10, 10]) < 1e-3
    assert abs(m.get("_normal_z")[10, 10] - 1.0) < 0.01
    # tilted plane: z = logical row * res * 0.5 -> slope = atan(0.5) = 26.57 deg
    e = np.fromfunction(lambda r, c: r * 0.5 * 0.5, (20, 20)).astype(np.float32)
    m.set("elevation", e)
    ob.feature_extraction(m, 0.6, 4)
    s = m.get("slope")[10, 10]
    assert 10.0 < s < 45.0 and abs(s - math.degrees(math.atan(0.5))) < 0.01
    # step edge between the column halves
    e = np.zeros((20, 20), np.float32)
    e[:, 10:] = 1.0
    m.set("elevation", e)
    ob.feature_extraction(m, 0.6, 4)
    assert m.get("step")[10, 10] > 0.5


def test_feature_extraction_skips_nan_cells():  # :342-350
    m = ob.OracleMap(10.0, 10.0, 0.5)
    ob.feature_extraction(m, 0.6, 4)
    assert m.exists("slope") and not np.isfinite(m.get("slope")[10, 10])


def test_direct_eigen_solver_against_lapack():
    """The oracle restates Eigen's closed-form 3x3 solver (Eigen itself is not available here):
    its eigenvalues / eigenvectors must agree with LAPACK on random covariance matrices."""
    rng = np.random.RandomState(0)
    for t in range(500):
        A = (rng.randn(3, 3) * rng.choice([1e-3, 1.0, 10.0])).astype(np.float32)
        S = (A @ A.T).astype(np.float32)
        if t % 5 == 0:
            S[2, :] *= 1e-2
            S[:, 2] *= 1e-2
        val, vec = ob.eig3(S)
        w, v = np.linalg.eigh(S.astype(np.float64))
        assert np.abs(val - w).max() <= 5e-5 * max(np.abs(w).max(), 1e-30)
        gaps = np.diff(w) / max(np.abs(w).max(), 1e-30)
        if gaps.min() > 1e-2:  # well separated: eigenvectors are defined up to sign
            for k in range(3):
                assert abs(np.dot(vec[k], v[:, k])) > 0.999
