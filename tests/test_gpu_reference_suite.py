"""-m gpu: the reference's own gtest cases for the integrate() path, re-expressed against the
CUDA implementation through the mirrored API (fastdem_b200.api), plus the behavioural edge
cases of SURVEY.md Appendix C.  Where the reference asserts a tolerance the same tolerance is
used; in addition every scenario is replayed on the CPU oracle and compared layer by layer.
File:line citations are relative to /root/reference/fastdem/tests/."""
import math

import numpy as np
import pytest

import oracle_binding as ob
from parity_utils import compare_layer, compare_maps, run_pair

pytestmark = pytest.mark.gpu

I4 = np.eye(4)


def ground_cloud(height, half=3, spacing=0.3):  # test_fastdem_integration.cpp:32-41
    g = np.arange(-half, half + 1, dtype=np.float32) * np.float32(spacing)
    xx, yy = np.meshgrid(g, g, indexing="ij")
    return np.stack([xx.ravel(), yy.ravel(), np.full(xx.size, height, np.float32)], axis=1)


class Pair:
    """The same map + mapper on the GPU and on the oracle, driven in lock step."""

    def __init__(self, fdem, size=(10.0, 10.0, 0.5), **cfg_kw):
        self.fd = fdem
        self.cfg = fdem.Config()
        for k, v in cfg_kw.items():
            setattr(self.cfg, k, v)
        self.gmap = fdem.ElevationMap(*size, "map")
        self.gdem = fdem.FastDEM(self.gmap, self.cfg)
        self.omap = ob.OracleMap(*size)
        self.odem = ob.OracleFastDEM(self.omap, self.cfg)

    def integrate(self, pts, Tbs=I4, Twb=I4, intensity=None, rgb=None):
        cloud = self.fd.PointCloud(pts, intensity, rgb)
        got = self.gdem.integrate(cloud, Tbs, Twb)
        want, ost, _ = self.odem.integrate(pts, Tbs, Twb, intensity, rgb)
        assert got == want
        return got

    def update(self, pts, robot=(0.0, 0.0), var_z=None, intensity=None, rgb=None):
        cloud = self.fd.PointCloud(pts, intensity, rgb)
        st = self.gdem.update(cloud, robot, var_z)
        n = self.odem.update(pts, robot, var_z, intensity, rgb)
        assert st.n_cells == n
        return st

    def check(self):
        return compare_maps(self.gmap, self.omap)


# ───────────────────────── test_elevation_map.cpp ─────────────────────────

def test_elevation_map_basics(fdem):
    m = fdem.ElevationMap(10.0, 10.0, 0.5, "world")
    assert m.isInitialized() and m.isEmpty() and m.getFrameId() == "world"   # :14-28, :71-77
    assert not fdem.ElevationMap().isInitialized()                          # :18-21
    assert math.isnan(m.elevationAt((100.0, 100.0)))                         # :30-33
    assert not m.hasElevationAt((0.0, 0.0))                                  # :35-38
    ok, idx = m.getIndex((1.0, 1.0))
    assert ok
    m.setAt("elevation", idx, 1.5)                                           # :40-49
    assert m.hasElevationAt((1.0, 1.0)) and m.elevationAt((1.0, 1.0)) == 1.5
    assert m.hasElevationAt(idx) and not m.isEmptyAt(idx) and not m.isEmpty()
    m.clearAt(idx)                                                           # :51-61
    assert not m.hasElevationAt((1.0, 1.0)) and math.isnan(m.elevationAt(idx))
    assert m.getSize() == (20, 20) and m.getLayers()[:3] == ["elevation", "elevation_min", "elevation_max"]


def test_geometry_matches_oracle(fdem):
    m = fdem.ElevationMap(15.0, 15.0, 0.1)
    o = ob.OracleMap(15.0, 15.0, 0.1)
    rng = np.random.RandomState(1)
    for step in range(6):
        pos = (rng.uniform(-3, 3), rng.uniform(-3, 3))
        assert m.move(pos) == o.move(pos)
        g, og = m.geometry(), o.geometry()
        assert (g.position[0], g.position[1]) == og["position"]
        assert (g.start_index[0], g.start_index[1]) == og["start_index"]
        for p in rng.uniform(-9, 9, size=(200, 2)):
            assert m.getIndex(p) == o.getIndex(p)
            assert m.isInside(p) == o.isInside(p)
        for idx in [(0, 0), (149, 149), (17, 93)]:
            assert m.getPosition(idx) == o.getPosition(idx)


def test_move_clears_like_oracle(fdem):
    for policy in (0, 1):
        m = fdem.ElevationMap(10.0, 10.0, 0.5)
        o = ob.OracleMap(10.0, 10.0, 0.5)
        m.add("n_points", 3.0)
        o.add("n_points", 3.0)
        e = np.arange(400, dtype=np.float32).reshape(20, 20, order="F")
        m.set("elevation", e)
        o.set("elevation", e)
        for pos in [(1.0, 0.0), (1.0, -1.6), (-3.2, 2.7), (0.1, 0.1)]:
            assert m.move(pos, policy) == o.move(pos, policy)
            for name in ("elevation", "n_points", "elevation_min"):
                compare_layer(name, m.get(name), o.get(name))
        assert m.move((100.0, 0.0), policy) == o.move((100.0, 0.0), policy)   # whole map dropped
        assert m.isEmpty() and not m.isInside((0.0, 0.0))


# ───────────────────────── test_dual_layer.cpp (ElevationMapping::update, GLOBAL) ─────────

def dual(fdem, est=0):
    return Pair(fdem, mode=1, estimation_type=est, kalman_min_variance=1e-4, kalman_max_variance=1.0)


def test_dual_ground_obstacle_separation(fdem):  # :66-83
    p = dual(fdem)
    p.update([[0, 0, 0.0], [0, 0, 3.0]])
    _, idx = p.gmap.getIndex((0, 0))
    assert abs(p.gmap.at("elevation", idx)) < 0.1 and abs(p.gmap.at("obstacle", idx) - 3.0) < 0.1
    p.check()


def test_dual_single_point_and_flat(fdem):  # :85-119
    p = dual(fdem)
    p.update([[0, 0, 2.0]])
    _, idx = p.gmap.getIndex((0, 0))
    assert abs(p.gmap.at("elevation", idx) - 2.0) < 0.1 and math.isnan(p.gmap.at("obstacle", idx))
    p.update([[0, 0, 1.0], [0, 0, 1.02]])
    p.check()


def test_dual_kalman_second_frame(fdem):  # :121-143
    p = dual(fdem)
    p.update([[0, 0, 0.0], [0, 0, 3.0]])
    p.update([[0, 0, 0.1], [0, 0, 3.1]])
    _, idx = p.gmap.getIndex((0, 0))
    assert -0.05 < p.gmap.at("elevation", idx) < 0.15
    assert np.float32(p.gmap.at("obstacle", idx)) == np.float32(3.1)
    p.check()


def test_dual_quantile(fdem):  # :145-165
    p = dual(fdem, est=1)
    for i in range(10):
        noise = 0.05 if i % 2 == 0 else -0.05
        p.update([[0, 0, 0.0 + noise], [0, 0, 5.0 + noise]])
    _, idx = p.gmap.getIndex((0, 0))
    assert abs(p.gmap.at("elevation", idx)) < 0.5 and abs(p.gmap.at("obstacle", idx) - 5.0) < 0.1
    p.check()


def test_dual_elevation_max_and_obstacle_clear(fdem):  # :167-203
    p = dual(fdem)
    _, idx = p.gmap.getIndex((0, 0))
    p.update([[0, 0, 0.0], [0, 0, 3.0]])
    assert p.gmap.at("elevation_max", idx) == 3.0
    p.update([[0, 0, 0.0], [0, 0, 5.0]])
    assert p.gmap.at("elevation_max", idx) == 5.0
    p.update([[0, 0, 0.0], [0, 0, 2.0]])
    assert p.gmap.at("obstacle", idx) == 2.0
    p.update([[0, 0, 0.0]])
    assert math.isnan(p.gmap.at("obstacle", idx))
    p.check()


def test_update_with_variance_channel(fdem):
    """cloud.covariance(i)(2,2) supplied by the caller (elevation_mapping.cpp:58-60)."""
    p = dual(fdem)
    rng = np.random.RandomState(3)
    pts = rng.uniform(-4, 4, size=(500, 3)).astype(np.float32)
    var = rng.uniform(1e-4, 1e-2, size=500).astype(np.float32)
    for _ in range(3):
        p.update(pts, var_z=var)
    p.check()


# ───────────────────────── test_fastdem_integration.cpp ─────────────────────────

def test_integrate_updates_elevation(fdem):  # :45-59
    p = Pair(fdem, z_min=-2, z_max=5, range_min=0, range_max=20, sensor_type=0)
    assert p.integrate(ground_cloud(1.0))
    assert p.gmap.hasElevationAt((0.0, 0.0)) and abs(p.gmap.elevationAt((0.0, 0.0)) - 1.0) < 0.1
    p.check()


def test_empty_cloud_is_noop_and_false(fdem):  # :61-69, :365-370
    p = Pair(fdem)
    assert not p.integrate(np.zeros((0, 3), np.float32))
    assert p.gmap.isEmpty()


def test_filters_reject(fdem):  # :71-80, :287-316, :372-378
    p = Pair(fdem, z_min=0.0, z_max=2.0)
    assert not p.integrate(ground_cloud(10.0))
    assert p.gmap.isEmpty()
    p = Pair(fdem, range_min=5.0, range_max=20.0)
    assert not p.integrate(ground_cloud(1.0, half=2))
    assert p.gmap.isEmpty()
    p = Pair(fdem, z_min=0.0, z_max=3.0, range_max=20.0, sensor_type=0)
    assert p.integrate(ground_cloud(1.0)) and not p.gmap.isEmpty()
    p.gmap.clearAll()
    p.omap.clearAll()
    assert not p.integrate(ground_cloud(5.0)) and p.gmap.isEmpty()


def test_multiple_integrations_and_sensor_models(fdem):  # :82-104, :128-157
    for st in (0, 1, 2):
        p = Pair(fdem, z_min=-5, z_max=15, range_max=20, sensor_type=st)
        assert p.integrate(ground_cloud(1.0))
        assert p.integrate(ground_cloud(1.5))
        assert 0.9 < p.gmap.elevationAt((0.0, 0.0)) < 1.6
        p.check()


def test_p2_quantile_estimator(fdem):  # :159-175 + Appendix C.6
    p = Pair(fdem, z_min=-5, z_max=15, range_max=20, sensor_type=0, estimation_type=1)
    for i in range(6):
        p.integrate(ground_cloud(1.0 + i * 0.01))
        e = p.gmap.elevationAt((0.0, 0.0))
        assert math.isnan(e) if i < 3 else abs(e - 1.0) < 0.2
        p.check()


def test_global_fixed_local_follows(fdem):  # :179-215
    p = Pair(fdem, mode=1, z_min=-5, z_max=15, sensor_type=0)
    p.integrate(ground_cloud(1.0))
    T = np.eye(4)
    T[0, 3] = 3.0
    p.integrate(ground_cloud(2.0), I4, T)
    assert p.gmap.hasElevationAt((0.0, 0.0))
    p.check()
    p = Pair(fdem, mode=0, z_min=-5, z_max=15, sensor_type=0)
    p.integrate(ground_cloud(1.0))
    T[0, 3] = 100.0
    p.integrate(ground_cloud(2.0), I4, T)
    assert not p.gmap.isInside((0.0, 0.0))
    p.check()


def test_transforms(fdem):  # :253-283
    p = Pair(fdem, z_min=-5, z_max=15, sensor_type=0)
    Tbs = np.eye(4)
    Tbs[2, 3] = 1.0
    p.integrate(ground_cloud(0.0), Tbs, I4)
    assert abs(p.gmap.elevationAt((0.0, 0.0)) - 1.0) < 0.2
    Twb = np.eye(4)
    c, s = math.cos(math.pi / 2), math.sin(math.pi / 2)
    Twb[:2, :2] = [[c, -s], [s, c]]
    assert p.integrate(ground_cloud(1.0), I4, Twb) and not p.gmap.isEmpty()
    p.check()


def test_fluent_setters_and_config(fdem):  # :45-52, :219-249
    gmap = fdem.ElevationMap(10.0, 10.0, 0.5)
    dem = fdem.FastDEM(gmap)
    dem.setHeightFilter(-2.0, 5.0).setRangeFilter(0.0, 20.0).setSensorModel(fdem.SensorType.Constant) \
        .setEstimatorType(fdem.EstimationType.Kalman)
    assert dem.integrate(fdem.PointCloud(ground_cloud(1.0)), I4, I4)
    assert gmap.exists("_kalman_p") and not gmap.exists("_p2_q0")
    dem.setEstimatorType(fdem.EstimationType.P2Quantile)   # re-creates the mapping; adds P2 layers
    assert gmap.exists("_p2_q0") and gmap.exists("_kalman_p")
    cfg = fdem.Config()
    cfg.z_min, cfg.z_max = 0.0, 2.0
    gmap2 = fdem.ElevationMap(10.0, 10.0, 0.5)
    assert not fdem.FastDEM(gmap2, cfg).integrate(fdem.PointCloud(ground_cloud(5.0)), I4, I4)
    assert gmap2.isEmpty()


def test_callbacks(fdem):  # :320-353
    gmap = fdem.ElevationMap(10.0, 10.0, 0.5)
    dem = fdem.FastDEM(gmap)
    dem.setHeightFilter(-5.0, 15.0).setSensorModel(fdem.SensorType.Constant)
    seen = {}
    dem.onScanPreprocessed(lambda c: seen.__setitem__("pre", c.size()))
    dem.onScanRasterized(lambda c: seen.__setitem__("ras", np.asarray(c.xyzw).copy()))
    pts = ground_cloud(1.0)
    dem.integrate(fdem.PointCloud(pts), I4, I4)
    assert seen["pre"] == 49
    ras = seen["ras"]
    omap = ob.OracleMap(10.0, 10.0, 0.5)
    want = set()
    for q in pts:
        ok, idx = omap.getIndex((float(q[0]), float(q[1])))
        x, y = omap.getPosition(idx)
        want.add((np.float32(x), np.float32(y), np.float32(1.0)))
    assert {tuple(r[:3]) for r in ras} == want   # one point per cell at the cell centre, z = min_z


# ───────────────────────── online mode (test_online_mode.cpp) ─────────────────────────

class MockCalibration:
    def __init__(self, T=None):
        self.T, self.fail = (np.eye(4) if T is None else T), False

    def getExtrinsic(self, frame):
        return None if self.fail else self.T

    def getBaseFrame(self):
        return "base_link"


class MockOdometry:
    def __init__(self):
        self.T, self.fail = np.eye(4), False

    def getPoseAt(self, ts):
        return None if self.fail else self.T

    def getWorldFrame(self):
        return "map"


def test_online_mode_provider_failures(fdem):  # test_online_mode.cpp:70-263
    gmap = fdem.ElevationMap(10.0, 10.0, 0.5)
    dem = fdem.FastDEM(gmap)
    dem.setHeightFilter(-5.0, 15.0).setSensorModel(fdem.SensorType.Constant)
    cloud = fdem.PointCloud(ground_cloud(1.0), frame_id="lidar", timestamp=123)
    assert not dem.hasTransformProvider() and not dem.integrate(cloud)       # providers missing
    calib, odom = MockCalibration(), MockOdometry()
    dem.setCalibrationProvider(calib).setOdometryProvider(odom)
    assert dem.hasTransformProvider()
    assert not dem.integrate(fdem.PointCloud(frame_id="lidar"))              # empty
    assert not dem.integrate(fdem.PointCloud(ground_cloud(1.0)))             # no frame id
    calib.fail = True
    assert not dem.integrate(cloud) and gmap.isEmpty()
    calib.fail, odom.fail = False, True
    assert not dem.integrate(cloud) and gmap.isEmpty()
    odom.fail = False
    assert dem.integrate(cloud) and gmap.hasElevationAt((0.0, 0.0))


# ───────────────────────── Appendix C edge cases ─────────────────────────

def test_all_filtered_does_not_move(fdem):  # C.2
    p = Pair(fdem, z_min=100.0, z_max=200.0)
    T = np.eye(4)
    T[0, 3] = 3.0
    assert not p.integrate(ground_cloud(1.0), I4, T)
    assert p.gmap.getPosition() == (0.0, 0.0)
    p.check()


def test_outside_map_moves_but_keeps_obstacle(fdem):  # C.3
    p = Pair(fdem, z_min=-5, z_max=15, sensor_type=0)
    p.integrate(np.array([[0, 0, 0.0], [0, 0, 2.0]], np.float32))
    _, idx = p.gmap.getIndex((0, 0))
    assert p.gmap.at("obstacle", idx) == 2.0
    assert p.integrate(np.array([[50.0, 50.0, 0.0]], np.float32))
    assert p.gmap.at("obstacle", idx) == 2.0
    p.check()


def test_reset_then_continue(fdem):  # C.7: clearAll leaves NaN counters; estimators re-init
    for est in (0, 1):
        p = Pair(fdem, z_min=-5, z_max=15, sensor_type=1, estimation_type=est)
        for i in range(6):
            p.integrate(ground_cloud(1.0 + 0.01 * i))
        p.gdem.reset()
        p.omap.clearAll()
        assert p.gmap.isEmpty()
        for i in range(6):
            p.integrate(ground_cloud(2.0 + 0.01 * i))
        p.check()


def test_nan_and_duplicate_points(fdem):
    """NaN coordinates are dropped by the crop compares (C.4); exact z ties pick the variance of
    the lowest-index point; colour comes from the highest-index point."""
    p = Pair(fdem, z_min=-5, z_max=15, sensor_type=1)
    pts = np.array([[1.0, 1.0, 0.5], [1.1, 1.05, 0.5], [np.nan, 0, 0], [1.2, 1.1, 0.5], [0, np.nan, 1],
                    [3.0, -2.0, np.nan], [1.05, 1.0, 0.7]], np.float32)
    inten = np.array([0.2, np.nan, 0.5, 0.9, 0.1, 0.3, 0.4], np.float32)
    rgb = np.arange(21, dtype=np.uint8).reshape(7, 3)
    for _ in range(2):
        assert p.integrate(pts, I4, I4, inten, rgb)
    p.check()
    # first point of a cell carries NaN intensity -> the scan's max for that cell is NaN
    pts2 = np.array([[2.0, 2.0, 0.1], [2.05, 2.0, 0.2]], np.float32)
    p.integrate(pts2, I4, I4, np.array([np.nan, 0.8], np.float32), rgb[:2])
    p.check()


def test_crowded_cells_span_chunks_and_warps(fdem):
    """Thousands of points in a handful of cells: segments longer than a warp window, buckets
    longer than one staged chunk, ties everywhere."""
    rng = np.random.RandomState(7)
    n = 20000
    pts = np.empty((n, 3), np.float32)
    pts[:, 0] = rng.choice([0.1, 0.13, 0.61, -2.2], size=n) + rng.uniform(0, 0.01, n)
    pts[:, 1] = rng.choice([0.1, 0.33, -1.4], size=n) + rng.uniform(0, 0.01, n)
    pts[:, 2] = rng.choice(np.linspace(0.0, 1.0, 7), size=n)        # heavy z ties
    inten = rng.uniform(0, 1, n).astype(np.float32)
    rgb = rng.randint(0, 255, size=(n, 3)).astype(np.uint8)
    for est in (0, 1):
        p = Pair(fdem, z_min=-5, z_max=15, sensor_type=1, estimation_type=est)
        for k in range(3):
            perm = rng.permutation(n)                                 # destroy scanline coherence
            p.integrate(pts[perm], I4, I4, inten[perm], rgb[perm])
        p.check()


def test_random_clouds_both_sort_paths(fdem):
    rng = np.random.RandomState(11)
    for cell_sort in (0, 1):
        p = Pair(fdem, size=(12.0, 9.0, 0.1), z_min=-5, z_max=15, sensor_type=1, mode=0)
        p.gdem.set_cell_sort(cell_sort)
        for k in range(5):
            n = int(rng.randint(1, 6000))
            pts = rng.uniform(-8, 8, size=(n, 3)).astype(np.float32)
            T = np.eye(4)
            T[:2, 3] = rng.uniform(-1, 1, 2) * (k + 1)
            p.integrate(pts, I4, T, rng.uniform(0, 1, n).astype(np.float32))
        p.check()


def test_async_stream_equals_sync(fdem):
    from fastdem_b200 import synthetic as syn
    wl = syn.WORKLOADS["tiny"]
    cfg = wl.config()
    a = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution)
    b = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution)
    da, db = fdem.FastDEM(a, cfg), fdem.FastDEM(b, cfg)
    scans = [syn.make_scan(wl, k) for k in range(9)]
    for s in scans:
        da.integrate_stats(fdem.PointCloud(s["xyzw"], s["intensity"]), s["T_base_sensor"], s["T_world_base"])
    clouds = [fdem.PointCloud(s["xyzw"].copy(), s["intensity"].copy()) for s in scans]
    for c, s in zip(clouds, scans):
        db.integrate_async(c, s["T_base_sensor"], s["T_world_base"])   # no host sync in between
    last = db.wait()
    assert last.integrated == 1
    for name in a.getLayers():
        compare_layer(name, b.get(name), a.get(name), rtol=0, atol=0)
    assert a.geometry().start_index[0] == b.geometry().start_index[0]


def test_custom_sensor_model_covariances(fdem):
    """FastDEM::setSensorModel(unique_ptr<SensorModel>): caller-computed covariances."""
    rng = np.random.RandomState(5)
    pts = ground_cloud(1.0)
    n = pts.shape[0]
    gmap = fdem.ElevationMap(10.0, 10.0, 0.5)
    dem = fdem.FastDEM(gmap)
    dem.setHeightFilter(-5.0, 15.0)
    # feeding the built-in LiDAR model's covariances back in must reproduce the built-in result
    cfg = fdem.Config()
    cov = np.stack([ob.sensor_cov(cfg, q).T.reshape(9) for q in pts]).astype(np.float32)
    st = dem.integrate_with_covariances(fdem.PointCloud(pts), cov, I4, I4)
    ref = fdem.ElevationMap(10.0, 10.0, 0.5)
    rdem = fdem.FastDEM(ref)
    rdem.setHeightFilter(-5.0, 15.0)
    rdem.integrate(fdem.PointCloud(pts), I4, I4)
    assert st.integrated
    for name in ref.getLayers():
        compare_layer(name, gmap.get(name), ref.get(name), rtol=0, atol=0)


def test_device_pointer_inputs_and_tensor_views(fdem):
    import torch
    from fastdem_b200 import synthetic as syn
    wl = syn.WORKLOADS["tiny"]
    s = syn.make_scan(wl, 0)
    a = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution)
    b = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution)
    fdem.FastDEM(a, wl.config()).integrate(fdem.PointCloud(s["xyzw"], s["intensity"]), s["T_base_sensor"], s["T_world_base"])
    dev = fdem.PointCloud(torch.from_numpy(s["xyzw"]).cuda(), torch.from_numpy(s["intensity"]).cuda())
    fdem.FastDEM(b, wl.config()).integrate(dev, s["T_base_sensor"], s["T_world_base"])
    compare_layer("elevation", b.get("elevation"), a.get("elevation"), rtol=0, atol=0)
    t = b.tensor("elevation")                     # zero-copy (cols, rows) view of the device slab
    assert t.is_cuda and tuple(t.shape) == (b.getSize()[1], b.getSize()[0])
    assert np.array_equal(np.nan_to_num(t.cpu().numpy().T, nan=-9), np.nan_to_num(b.get("elevation"), nan=-9))


# ───────────────────────── raycasting / voxel / inpainting ─────────────────────────

def rc_cfg(fdem, **kw):
    c = fdem.Config()
    c.raycasting_enabled = 1
    for k, v in kw.items():
        setattr(c, k, v)
    return c


def test_raycasting_reference_cases(fdem):  # test_postprocess.cpp:71-189
    m = fdem.ElevationMap(10.0, 10.0, 0.5)
    fdem.applyRaycasting(m, fdem.PointCloud([[1.0, 0.0, 0.5]]), (0, 0, 5), fdem.Config())  # disabled
    assert not m.exists("raycasting")
    fdem.applyRaycasting(m, fdem.PointCloud([[1.0, 0.0, 0.5]]), (50.0, 0, 5), rc_cfg(fdem))  # outside
    assert not m.exists("raycasting")
    # ghost cleared after one conflict
    _, g = m.getIndex((2.0, 0.0))
    m.setAt("elevation", g, 10.0)
    fdem.applyRaycasting(m, fdem.PointCloud([[4.0, 0.0, 0.0]]), (0, 0, 5),
                         rc_cfg(fdem, rc_log_odds_ghost=0.5, rc_clear_threshold=-0.4))
    assert math.isnan(m.at("elevation", g)) and m.at("ghost_removal", g) == 1.0
    for name in ("ghost_removal", "raycasting", "_visibility_logodds"):
        assert m.exists(name)
    # observed cell protected
    m2 = fdem.ElevationMap(10.0, 10.0, 0.5)
    m2.setAt("elevation", g, 2.0)
    fdem.applyRaycasting(m2, fdem.PointCloud([[4.0, 0.0, 0.0], [2.0, 0.0, 0.3]]), (0, 0, 5),
                         rc_cfg(fdem, rc_log_odds_observed=0.8, rc_log_odds_ghost=0.5, rc_clear_threshold=-0.4))
    assert not math.isnan(m2.at("elevation", g))
    # accumulation: exactly 5 frames
    m3 = fdem.ElevationMap(10.0, 10.0, 0.5)
    m3.setAt("elevation", g, 10.0)
    cfg = rc_cfg(fdem, rc_log_odds_ghost=0.2, rc_clear_threshold=-0.9)
    for _ in range(4):
        fdem.applyRaycasting(m3, fdem.PointCloud([[4.0, 0.0, 0.0]]), (0, 0, 5), cfg)
    assert not math.isnan(m3.at("elevation", g))
    fdem.applyRaycasting(m3, fdem.PointCloud([[4.0, 0.0, 0.0]]), (0, 0, 5), cfg)
    assert math.isnan(m3.at("elevation", g))


def test_raycasting_matches_oracle_on_random_scene(fdem):
    rng = np.random.RandomState(2)
    g = fdem.ElevationMap(12.0, 12.0, 0.1)
    o = ob.OracleMap(12.0, 12.0, 0.1)
    elev = rng.uniform(-0.5, 1.5, size=(120, 120)).astype(np.float32)
    elev[rng.uniform(size=elev.shape) < 0.3] = np.nan
    g.set("elevation", np.asfortranarray(elev))
    o.set("elevation", np.asfortranarray(elev))
    g.move((0.37, -0.22))
    o.move((0.37, -0.22))
    cfg = rc_cfg(fdem, rc_log_odds_ghost=0.6, rc_clear_threshold=-1.0)
    for k in range(3):
        pts = rng.uniform(-7, 7, size=(3000, 3)).astype(np.float32)
        pts[:, 2] = rng.uniform(-0.5, 2.5, size=3000)
        origin = (0.3, -0.1, 1.8)
        fdem.applyRaycasting(g, fdem.PointCloud(pts), origin, cfg)
        o.raycast(pts, origin, cfg)
        for name in ("elevation", "raycasting", "_visibility_logodds", "ghost_removal"):
            compare_layer(name, g.get(name), o.get(name))


def test_voxel_grid_any_matches_oracle(fdem):
    rng = np.random.RandomState(4)
    m = fdem.ElevationMap(10.0, 10.0, 0.5)
    pts = rng.uniform(-3, 3, size=(5000, 3)).astype(np.float32)
    pts[::50] = np.nan
    for voxel in (0.05, 0.3, 1.0):
        got = fdem.voxelGridAny(m, fdem.PointCloud(pts), voxel)
        want = ob.voxel_any(pts, voxel)
        assert np.array_equal(got, want)
    with pytest.raises(fdem.FdemError):
        fdem.voxelGridAny(m, fdem.PointCloud(pts), 0.0001)


def test_integrate_with_raycasting_matches_oracle(fdem):
    from fastdem_b200 import synthetic as syn
    wl = syn.WORKLOADS["tiny"]
    cfg = wl.config()
    cfg.raycasting_enabled = 1
    g = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution)
    o = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
    gd, od = fdem.FastDEM(g, cfg), ob.OracleFastDEM(o, cfg)
    for k in range(6):
        s = syn.make_scan(wl, k)
        st = gd.integrate_stats(fdem.PointCloud(s["xyzw"], s["intensity"]), s["T_base_sensor"], s["T_world_base"])
        ok, ost, _ = od.integrate(s["xyzw"], s["T_base_sensor"], s["T_world_base"], s["intensity"])
        assert st.n_voxels == ost.n_voxels and st.n_cells == ost.n_cells
    compare_maps(g, o)


def test_voxel_compact_keys_match_the_63bit_keys(fdem):
    """voxelGrid(ANY) sorts 32-bit box-relative keys when the crop filters bound the kept
    points (finite range_max) and the reference's 63-bit keys otherwise; both must pick the
    same representatives, i.e. produce the same map — with a tilted, translated robot pose so
    all three box axes are exercised."""
    from fastdem_b200 import synthetic as syn
    wl = syn.WORKLOADS["tiny"]
    maps = []
    for range_max in (20.0, 3.0e6):   # 3e6 > voxel_box()'s limit -> 63-bit path; no point is that far
        cfg = wl.config()
        cfg.raycasting_enabled = 1
        cfg.range_max = range_max
        g = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution)
        o = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
        gd, od = fdem.FastDEM(g, cfg), ob.OracleFastDEM(o, cfg)
        for k in range(5):
            s = syn.make_scan(wl, k)
            Twb = np.array(s["T_world_base"], dtype=np.float64).copy()
            a = 0.03 * (k + 1)   # small roll: z of the map frame mixes with y of the base frame
            Rx = np.array([[1, 0, 0], [0, np.cos(a), -np.sin(a)], [0, np.sin(a), np.cos(a)]])
            Twb[:3, :3] = Twb[:3, :3] @ Rx
            Twb[:3, 3] += (137.0, -61.5, 0.0)   # far from the origin: box corner offsets matter
            st = gd.integrate_stats(fdem.PointCloud(s["xyzw"], s["intensity"]), s["T_base_sensor"], Twb)
            ok, ost, _ = od.integrate(s["xyzw"], s["T_base_sensor"], Twb, s["intensity"])
            assert st.n_voxels == ost.n_voxels and st.n_cells == ost.n_cells and st.n_voxels > 100
            assert st.voxel_box_violations == 0
        compare_maps(g, o)
        maps.append({n: g.get(n) for n in ("elevation", "raycasting", "_visibility_logodds")})
    for n in maps[0]:
        assert np.array_equal(maps[0][n], maps[1][n], equal_nan=True), n


def test_inpainting_matches_oracle(fdem):  # test_postprocess.cpp:41-69
    rng = np.random.RandomState(9)
    g = fdem.ElevationMap(10.0, 10.0, 0.5)
    o = ob.OracleMap(10.0, 10.0, 0.5)
    e = rng.uniform(0, 1, size=(20, 20)).astype(np.float32)
    e[rng.uniform(size=e.shape) < 0.4] = np.nan
    e[9:12, 9:12] = 1.0
    e[10, 10] = np.nan
    for m in (g, o):
        m.set("elevation", np.asfortranarray(e))
        m.move((1.0, -0.5))
    fdem.applyInpainting(g, 3, 2, False)
    o.inpaint(3, 2, False)
    compare_layer("elevation_inpainted", g.get("elevation_inpainted"), o.get("elevation_inpainted"), rtol=1e-6)
    fdem.applyInpainting(g, 2, 3, True)
    o.inpaint(2, 3, True)
    compare_layer("elevation", g.get("elevation"), o.get("elevation"), rtol=1e-6)


def test_spatial_smoothing_matches_oracle(fdem):  # test_postprocess.cpp:249-271
    rng = np.random.RandomState(13)
    g = fdem.ElevationMap(10.0, 10.0, 0.5)
    o = ob.OracleMap(10.0, 10.0, 0.5)
    e = rng.uniform(0, 1, size=(20, 20)).astype(np.float32)
    e[rng.uniform(size=e.shape) < 0.3] = np.nan
    e[8:13, 8:13] = 1.0
    e[10, 10] = 100.0
    for m in (g, o):
        m.set("elevation", np.asfortranarray(e))
        m.move((1.0, -0.5))
    _, c = g.getIndex((1.0 + 0.25, -0.5 + 0.25))
    for k, mv in ((3, 5), (5, 7), (7, 3), (1, 1)):
        fdem.applySpatialSmoothing(g, "elevation", k, mv)
        ob.spatial_smoothing(o, "elevation", k, mv)
        compare_layer("elevation", g.get("elevation"), o.get("elevation"), rtol=0, atol=0)
    fdem.applySpatialSmoothing(g, "nonexistent_layer")   # no-op like the reference
    with pytest.raises(fdem.FdemError):
        fdem.applySpatialSmoothing(g, "elevation", 4, 5)


def _random_terrain(rng, n=40, hole=0.25):
    r, c = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    e = (0.3 * np.sin(0.35 * r) * np.cos(0.27 * c) + 0.02 * rng.standard_normal((n, n))).astype(np.float32)
    e[20:, 25:] += np.float32(0.8)                     # a step edge
    e[rng.uniform(size=e.shape) < hole] = np.nan
    return e


def test_uncertainty_fusion_matches_oracle(fdem):  # test_postprocess.cpp:193-247
    rng = np.random.RandomState(21)
    g = fdem.ElevationMap(4.0, 4.0, 0.1)
    o = ob.OracleMap(4.0, 4.0, 0.1)
    fdem.applyUncertaintyFusion(g)                      # no bound layers: no-op, nothing added
    assert not g.exists("upper_bound")
    e = _random_terrain(rng)
    spread = rng.uniform(0.01, 0.3, size=e.shape).astype(np.float32)
    for m in (g, o):
        m.add("upper_bound")
        m.add("lower_bound")
        m.set("elevation", np.asfortranarray(e))
        m.set("upper_bound", np.asfortranarray(e + spread))
        m.set("lower_bound", np.asfortranarray(e - spread))
        m.move((0.7, -0.4))                             # circular buffer wraps, edge stripes are NaN
    for kw in (dict(), dict(search_radius=0.25, spatial_sigma=0.1, quantile_lower=0.1, quantile_upper=0.8,
                            min_valid_neighbors=5)):
        fdem.applyUncertaintyFusion(g, **kw)
        ob.uncertainty_fusion(o, kw.get("search_radius", 0.15), kw.get("spatial_sigma", 0.05),
                              kw.get("quantile_lower", 0.01), kw.get("quantile_upper", 0.99),
                              kw.get("min_valid_neighbors", 3))
        for name in ("upper_bound", "lower_bound"):
            # every output is one of the input samples: a quantile picks, it does not blend
            compare_layer(name, g.get(name), o.get(name), rtol=0, atol=0)
    fdem.applyUncertaintyFusion(g, enabled=False)       # disabled = no-op
    with pytest.raises(fdem.FdemError):
        fdem.applyUncertaintyFusion(g, search_radius=0.95)   # 10 cells > compiled neighbourhood


def test_feature_extraction_matches_oracle(fdem):  # test_postprocess.cpp:273-350
    rng = np.random.RandomState(22)
    g = fdem.ElevationMap(4.0, 4.0, 0.1)
    o = ob.OracleMap(4.0, 4.0, 0.1)
    fdem.applyFeatureExtraction(g, 0.3, 4)               # all NaN: layers created, nothing computed
    assert g.exists("slope") and not np.isfinite(g.get("slope")).any()
    e = _random_terrain(rng, hole=0.15)
    for m in (g, o):
        m.set("elevation", np.asfortranarray(e))
        m.move((-0.3, 0.6))
    fdem.applyFeatureExtraction(g, 0.3, 4)
    ob.feature_extraction(o, 0.3, 4)
    n_ok = int(np.isfinite(o.get("slope")).sum())
    assert n_ok > 500
    compare_layer("step", g.get("step"), o.get("step"), rtol=0, atol=0)   # order statistics: exact
    # PCA outputs go through sin/cos/atan2/acos, whose last bits differ between libm and CUDA
    compare_layer("roughness", g.get("roughness"), o.get("roughness"), rtol=1e-3, atol=1e-5)
    compare_layer("curvature", g.get("curvature"), o.get("curvature"), rtol=1e-3, atol=1e-6)
    for name in ("_normal_x", "_normal_y", "_normal_z"):
        compare_layer(name, g.get(name), o.get(name), rtol=1e-4, atol=1e-4)
    compare_layer("slope", g.get("slope"), o.get("slope"), rtol=1e-4, atol=0.05)  # acos near 1
    # percentiles away from the ends take the kernel's keep-everything-sorted path
    fdem.applyFeatureExtraction(g, 0.3, 4, 0.25, 0.75)
    ob.feature_extraction(o, 0.3, 4, 0.25, 0.75)
    compare_layer("step", g.get("step"), o.get("step"), rtol=0, atol=0)
    # reference known answers on the GPU path: flat plane, tilted plane
    g2 = fdem.ElevationMap(10.0, 10.0, 0.5)
    g2.set("elevation", np.asfortranarray(np.ones((20, 20), np.float32)))
    fdem.applyFeatureExtraction(g2, 0.6, 4)
    assert abs(g2.get("slope")[10, 10]) < 1.0 and abs(g2.get("_normal_z")[10, 10] - 1.0) < 0.01
    assert abs(g2.get("roughness")[10, 10]) < 1e-3 and abs(g2.get("step")[10, 10]) < 1e-3
    g2.set("elevation", np.asfortranarray(np.fromfunction(lambda r, c: r * 0.25, (20, 20)).astype(np.float32)))
    fdem.applyFeatureExtraction(g2, 0.6, 4)
    assert abs(g2.get("slope")[10, 10] - 26.565) < 0.01


def _decode_pc2(fields, point_step, width, data):
    a = np.frombuffer(bytes(data), dtype=np.uint32).reshape(width, point_step // 4) if width else np.zeros((0, point_step // 4), np.uint32)
    return {name: a[:, k] for k, name in enumerate(fields)}


def test_map_to_pointcloud2_matches_oracle(fdem):  # bridge/ros/impl.hpp:29-174
    from fastdem_b200 import synthetic as syn
    wl = syn.WORKLOADS["c3_rgbd_p2"]   # colour layer -> "rgb" field; LOCAL -> non-zero start index
    gmap, omap, *_ = run_pair(fdem, wl, 6)
    assert gmap.getStartIndex() != (0, 0)
    rows, cols = gmap.getSize()
    for sub in (None, ((rows - 7, cols - 40), (60, 150)), ((3, 5), (0, 10))):
        kw = {} if sub is None else dict(sub_start=sub[0], sub_size=sub[1])
        msg = fdem.toPointCloud2(gmap, "elevation", **kw)
        of, ops, ow, od = ob.to_pointcloud2(omap, "elevation", *(sub if sub else (None, None)))
        gf = [f[0] for f in msg.fields]
        assert gf[:3] == ["x", "y", "z"] and gf[-1] == "rgb" and of[-1] == "rgb"
        assert not any(n.startswith("_") for n in gf) and "elevation" not in gf and "color" not in gf
        assert sorted(gf) == sorted(of) and msg.point_step == ops == 4 * len(gf)
        assert msg.width == ow and msg.height == 1
        if sub is None:
            assert ow == int(np.isfinite(omap.get("elevation")).sum()) > 1000
        g = _decode_pc2(gf, msg.point_step, msg.width, msg.data)
        o = _decode_pc2(of, ops, ow, od)
        for name in gf:   # same points, same order, same bits (NaN payloads included)
            assert np.array_equal(g[name], o[name]), name
    with pytest.raises(fdem.FdemError):
        fdem.toPointCloud2(gmap, "no_such_layer")
