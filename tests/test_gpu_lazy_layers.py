"""-m gpu: the layer SET is part of the parity contract.

The reference creates `intensity` / `color` inside updateIntensity / updateColor, i.e. on the
first scan that carries the channel AND produces >= 1 cell (elevation_mapping.cpp:116-117,
154-156, 168-170), and the three raycasting layers once applyRaycasting passes its guards
(raycasting.cpp:221-240).  Those conditions are decided on the device here, so the storage is
allocated earlier but must stay invisible — exists() / getLayers() / save files — until then.
Also: ElevationMap.setGeometry / io.loadNpz on a map that already has a FastDEM bound to it
(the reference resizes in place; the mapper must stay valid), and the whole-layer obstacle
clear after the caller edited the obstacle layer."""
import numpy as np
import pytest

import oracle_binding as ob
from parity_utils import compare_maps

pytestmark = pytest.mark.gpu

I4 = np.eye(4)


def _pair(fdem, **kw):
    cfg = fdem.Config()
    for k, v in kw.items():
        setattr(cfg, k, v)
    gmap = fdem.ElevationMap(10.0, 10.0, 0.5, "map")
    omap = ob.OracleMap(10.0, 10.0, 0.5)
    return cfg, gmap, fdem.FastDEM(gmap, cfg), omap, ob.OracleFastDEM(omap, cfg)


def _both(fdem, gdem, odem, pts, intensity=None, rgb=None, Tbs=I4, Twb=I4):
    got = gdem.integrate(fdem.PointCloud(pts, intensity, rgb), Tbs, Twb)
    want, _, _ = odem.integrate(pts, Tbs, Twb, intensity, rgb)
    assert got == want
    return got


def test_channel_layers_appear_with_the_first_observing_scan(fdem):
    cfg, gmap, gdem, omap, odem = _pair(fdem, mode=fdem.MODE_GLOBAL)
    far = np.array([[100.0, 100.0, 0.0], [120.0, -90.0, 1.0]], np.float32)   # kept, but outside the map
    inten = np.array([0.3, 0.9], np.float32)
    rgb = np.array([[1, 2, 3], [4, 5, 6]], np.uint8)
    assert _both(fdem, gdem, odem, far, inten, rgb) is True   # integrate() is true: points survived the filters
    assert not omap.exists("intensity") and not omap.exists("color")
    assert not gmap.exists("intensity") and not gmap.exists("color")
    assert sorted(gmap.getLayers()) == sorted(omap.layers())
    with pytest.raises(fdem.FdemError):
        gmap.get("intensity")
    near = np.array([[0.1, 0.1, 0.5], [1.2, -0.7, 0.25]], np.float32)
    _both(fdem, gdem, odem, near, inten, None)
    assert gmap.exists("intensity") and not gmap.exists("color")
    compare_maps(gmap, omap)
    _both(fdem, gdem, odem, near, inten, rgb)
    assert gmap.exists("color")
    compare_maps(gmap, omap)


def test_channel_layers_through_the_async_queue(fdem):
    """The fact arrives with results the host has not looked at yet: exists() must wait for them."""
    cfg, gmap, gdem, omap, odem = _pair(fdem, mode=fdem.MODE_GLOBAL)
    near = np.array([[0.1, 0.1, 0.5], [1.2, -0.7, 0.25]], np.float32)
    far = near + np.float32(500.0)
    inten = np.array([0.3, 0.9], np.float32)
    gdem.integrate_async(fdem.PointCloud(near, inten), I4, I4)
    gdem.integrate_async(fdem.PointCloud(far, inten), I4, I4)      # newest result has no cells
    odem.integrate(near, I4, I4, inten, None)
    odem.integrate(far, I4, I4, inten, None)
    assert gmap.exists("intensity")
    gdem.wait()
    compare_maps(gmap, omap)


def test_raycasting_layers_follow_the_guards(fdem):
    cfg, gmap, gdem, omap, odem = _pair(fdem, mode=fdem.MODE_GLOBAL, raycasting_enabled=1)
    pts = np.array([[0.5, 0.5, -1.0], [1.0, 0.2, -0.8]], np.float32)
    Twb_out = np.eye(4)
    Twb_out[:3, 3] = [50.0, 0.0, 1.0]            # sensor origin outside the 10 x 10 m map
    _both(fdem, gdem, odem, pts, Twb=Twb_out)
    for name in ("raycasting", "ghost_removal", "_visibility_logodds"):
        assert not omap.exists(name) and not gmap.exists(name), name
    assert sorted(gmap.getLayers()) == sorted(omap.layers())
    Twb_in = np.eye(4)
    Twb_in[:3, 3] = [0.0, 0.0, 1.0]
    _both(fdem, gdem, odem, pts, Twb=Twb_in)
    for name in ("raycasting", "ghost_removal", "_visibility_logodds"):
        assert omap.exists(name) and gmap.exists(name), name
    compare_maps(gmap, omap)


def test_set_geometry_in_place_keeps_the_mapper_valid(fdem, tmp_path):
    """io.loadNpz calls setGeometry on the map it restores into; a FastDEM created on that map
    earlier must keep working (the reference's holds an ElevationMap&)."""
    from fastdem_b200 import io_npz
    cfg, gmap, gdem, omap, odem = _pair(fdem)
    pts = np.array([[0.1, 0.1, 0.5], [1.2, -0.7, 0.25], [-2.0, 1.0, 0.1]], np.float32)
    _both(fdem, gdem, odem, pts)
    path = str(tmp_path / "m.npz")
    assert io_npz.saveNpz(path, gmap)
    layers_before = sorted(gmap.getLayers())

    gmap.setGeometry(6.0, 8.0, 0.25)              # in place: same handle, all layers NaN
    assert gmap.getSize() == (24, 32)
    assert sorted(gmap.getLayers()) == layers_before
    assert np.isnan(gmap.get("elevation")).all() and np.isnan(gmap.get("n_points")).all()
    omap2 = ob.OracleMap(6.0, 8.0, 0.25)
    odem2 = ob.OracleFastDEM(omap2, cfg)
    omap2.clearAll()                              # setGeometry + clearAll: estimator layers are NaN too
    _both(fdem, gdem, odem2, pts)                  # the old mapper integrates into the resized map
    compare_maps(gmap, omap2, layers=["elevation", "elevation_min", "elevation_max", "obstacle"])

    assert io_npz.loadNpz(path, gmap)             # restore the checkpoint into the live map ...
    assert gmap.getSize() == (20, 20)
    compare_maps(gmap, omap)
    _both(fdem, gdem, odem, pts + np.float32(0.05))   # ... and keep mapping with the same FastDEM
    compare_maps(gmap, omap)


def test_destroying_the_map_detaches_the_mapper(fdem):
    gmap = fdem.ElevationMap(10.0, 10.0, 0.5, "map")
    gdem = fdem.FastDEM(gmap, fdem.Config())
    gmap.close()
    with pytest.raises(fdem.FdemError):
        gdem.integrate(fdem.PointCloud(np.zeros((1, 3), np.float32)), I4, I4)
    gdem.close()


def test_obstacle_edited_by_the_caller_is_cleared_by_the_next_observing_scan_only(fdem):
    """a scan WITHOUT observations must leave a restored / edited obstacle layer alone
    (update() returns before updateObstacle, elevation_mapping.cpp:116-117); the next scan WITH
    observations clears all of it (:146), not just the cells the mapper remembers."""
    cfg, gmap, gdem, omap, odem = _pair(fdem, mode=fdem.MODE_GLOBAL)
    near = np.array([[0.1, 0.1, 0.5], [0.1, 0.1, 1.5]], np.float32)
    _both(fdem, gdem, odem, near)
    ob_layer = gmap.get("obstacle")
    ob_layer[3, 4] = 7.0
    gmap.set("obstacle", ob_layer)
    omap.set("obstacle", ob_layer)
    far = near + np.float32(500.0)
    _both(fdem, gdem, odem, far)                  # no observations: nothing may be cleared
    assert gmap.at("obstacle", (3, 4)) == 7.0
    compare_maps(gmap, omap)
    _both(fdem, gdem, odem, near + np.float32(1.0))
    assert np.isnan(gmap.at("obstacle", (3, 4)))
    compare_maps(gmap, omap)
