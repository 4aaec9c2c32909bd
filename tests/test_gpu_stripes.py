"""-m gpu: the CUDA side of row-stripe sharding on ONE GPU: stripes of a GLOBAL map integrate
the same scans as the unsharded map and their rows concatenate to it bit for bit."""
import numpy as np
import pytest

from fastdem_b200 import capi, sharded
from fastdem_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2, 3])
def test_stripes_concatenate_to_the_full_map(fdem, world):
    wl = syn.WORKLOADS["tiny"]
    cfg = wl.config()
    cfg.mode = capi.MODE_GLOBAL
    full = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution)
    fdem_full = fdem.FastDEM(full, cfg)
    rows = full.getSize()[0]
    stripes = []
    for r in range(world):
        r0, r1 = sharded.stripe_bounds(rows, world, r)
        m = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution, row_stripe=(r0, r1))
        assert m.rowStripe() == (r0, r1)
        stripes.append((m, fdem.FastDEM(m, cfg)))
    for k in range(5):
        s = syn.make_scan(wl, k)
        cloud = fdem.PointCloud(s["xyzw"], s["intensity"])
        st = fdem_full.integrate_stats(cloud, s["T_base_sensor"], s["T_world_base"])
        cells = 0
        for m, d in stripes:
            cells += d.integrate_stats(cloud, s["T_base_sensor"], s["T_world_base"]).n_cells
        assert cells == st.n_cells
    for name in full.getLayers():
        whole = full.get(name)
        parts = np.concatenate([m.get(name) for m, _ in stripes], axis=0)
        assert np.array_equal(np.isnan(parts), np.isnan(whole)), name
        assert np.array_equal(np.nan_to_num(parts).view(np.uint32), np.nan_to_num(whole).view(np.uint32)), name


def test_local_mode_rejected_on_stripe(fdem):
    m = fdem.ElevationMap(10.0, 10.0, 0.5, row_stripe=(0, 10))
    d = fdem.FastDEM(m)  # default config is LOCAL
    with pytest.raises(fdem.FdemError):
        d.integrate(fdem.PointCloud([[0.0, 0.0, 1.0]]), np.eye(4), np.eye(4))


def test_sharded_map_world1(fdem):
    wl = syn.WORKLOADS["tiny"]
    cfg = wl.config()
    cfg.mode = capi.MODE_GLOBAL
    sm = sharded.ShardedGlobalMap(wl.map_width, wl.map_height, wl.resolution, cfg)
    assert sm.is_empty()
    s = syn.make_scan(wl, 0)
    st = sm.integrate(s["xyzw"], s["intensity"], None, s["T_base_sensor"], s["T_world_base"])
    assert st.integrated and not sm.is_empty()
    assert sm.gather("elevation").shape == (100, 100)
