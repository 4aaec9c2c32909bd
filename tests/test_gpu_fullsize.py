"""-m gpu: BASELINE.json's full-size configurations (C4: 1.05 M points + raycasting on a
1000x1000 map; C5: 1.05 M points on the 8000x8000 = 64 M-cell global map).

Oracle parity at full size for a couple of scans (the oracle needs ~0.1-0.4 s per scan here),
plus size-independent properties that need no oracle:
  * the two independent cell-reduction implementations (tile path: bucket partition +
    shared-memory sort; global path: CUB radix sort + warp-segmented reduce) agree bit for bit
  * counts are consistent: n_points grows by exactly one per touched cell per scan
  * order relations between layers hold cell by cell"""
import numpy as np
import pytest

from fastdem_b200 import capi
from fastdem_b200 import synthetic as syn
from parity_utils import compare_layer, compare_maps, run_pair

pytestmark = pytest.mark.gpu


def _run(fdem, wl, cfg, scans, cell_sort):
    m = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution)
    d = fdem.FastDEM(m, cfg)
    d.set_cell_sort(cell_sort)
    stats = []
    for s in scans:
        stats.append(d.integrate_stats(fdem.PointCloud(s["xyzw"], s["intensity"], s["rgb"]),
                                       s["T_base_sensor"], s["T_world_base"]))
    return m, d, stats


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("name", ["c4_dense_raycast", "c5_global"])
def test_two_sort_paths_agree_bit_for_bit(fdem, name):
    wl = syn.WORKLOADS[name]
    cfg = wl.config()
    scans = [syn.make_scan(wl, k) for k in range(3)]
    ma, da, sa = _run(fdem, wl, cfg, scans, capi.CELL_SORT_TILE)
    mb, db, sb = _run(fdem, wl, cfg, scans, capi.CELL_SORT_GLOBAL)
    for a, b in zip(sa, sb):
        assert (a.n_kept, a.n_cells, a.n_voxels, a.integrated) == (b.n_kept, b.n_cells, b.n_voxels, b.integrated)
    assert ma.getLayers() == mb.getLayers()
    for layer in ma.getLayers():
        x, y = ma.get(layer), mb.get(layer)
        assert np.array_equal(_bits(x), _bits(y)), layer
    # counts: every touched cell got exactly one estimator step per scan
    n_points = ma.get("n_points")
    assert np.nansum(n_points) == sum(s.n_cells for s in sa) or name == "c4_dense_raycast"
    # order relations (Kalman): min <= max, lower <= elevation <= upper where defined
    emin, emax = ma.get("elevation_min"), ma.get("elevation_max")
    ok = np.isfinite(emin) & np.isfinite(emax)
    assert (emin[ok] <= emax[ok]).all()
    e, lo, up = ma.get("elevation"), ma.get("lower_bound"), ma.get("upper_bound")
    ok = np.isfinite(e) & np.isfinite(lo) & np.isfinite(up)
    assert ok.sum() > 1000 and (lo[ok] <= e[ok]).all() and (e[ok] <= up[ok]).all()
    ob = ma.get("obstacle")
    ok = np.isfinite(ob) & np.isfinite(emax)
    assert (ob[ok] <= emax[ok]).all()          # this scan's max_z never exceeds the all-time max


def test_c4_oracle_parity_with_raycasting(fdem):
    """1.05 M points, 1000x1000 LOCAL map, Kalman, voxelGrid(ANY) + raycasting, 2 scans."""
    wl = syn.WORKLOADS["c4_dense_raycast"]
    gmap, omap, gdem, odem, gs, os_ = run_pair(fdem, wl, 2)
    assert gs[-1].n_voxels > 100000 and gs[-1].n_cells > 50000
    assert all(s.voxel_box_violations == 0 for s in gs)
    compare_maps(gmap, omap)


def test_c5_oracle_parity_global_64m_cells(fdem):
    """1.05 M points on the 8000 x 8000 global map, 2 scans; EVERY layer compared over all
    64 M cells (cells the scans cannot reach must still be at their initial fill)."""
    wl = syn.WORKLOADS["c5_global"]
    gmap, omap, gdem, odem, gs, os_ = run_pair(fdem, wl, 2)
    assert gmap.getSize() == (8000, 8000)
    report = compare_maps(gmap, omap)
    assert len(report) >= 12, sorted(report)


def test_long_stream_soak_matches_oracle(fdem):
    """600 VLP-16 scans of a moving robot through every entry style in turn — synchronous,
    queued (integrate_async), streaming host buffers (submit / collect, staging ring and result
    ring wrap many times), PointCloud2 bodies — with the bucket shape switching on the way.
    The map must still equal the oracle's, cell for cell, at the end."""
    import oracle_binding as ob
    from fastdem_b200.api import PointCloud2
    wl = syn.WORKLOADS["c1_vlp16_local"]
    cfg = wl.config()
    gmap = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution, "map")
    gdem = fdem.FastDEM(gmap, cfg)
    omap = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
    odem = ob.OracleFastDEM(omap, cfg)
    scans = [syn.make_scan(wl, k) for k in range(12)]
    clouds = [fdem.PointCloud(s["xyzw"], s["intensity"]) for s in scans]
    msgs = [PointCloud2.from_arrays(s["xyzw"][:, :3], s["intensity"]) for s in scans]
    pending = []
    n_cells_gpu, n_cells_cpu = 0, 0
    for k in range(600):
        j = k % len(scans)
        Tbs, Twb = syn.pose(wl, k)
        mode = (k // 50) % 4
        if mode != 2 and mode != 3 and pending:     # leaving a streaming phase: drain
            n_cells_gpu += sum(gdem.collect(t).n_cells for t in pending)
            pending = []
        if mode == 0:
            n_cells_gpu += gdem.integrate_stats(clouds[j], Tbs, Twb).n_cells
        elif mode == 1:
            gdem.integrate_async(clouds[j], Tbs, Twb)
            if k % 50 == 49:
                gdem.wait()
        elif mode == 2:
            pending.append(gdem.submit(clouds[j], Tbs, Twb))
        else:
            pending.append(gdem.submit_pointcloud2(msgs[j], Tbs, Twb))
        if len(pending) > 3:
            n_cells_gpu += gdem.collect(pending.pop(0)).n_cells
        _, ost, _ = odem.integrate(scans[j]["xyzw"], Tbs, Twb, scans[j]["intensity"], None)
        if mode != 1:
            n_cells_cpu += ost.n_cells
    n_cells_gpu += sum(gdem.collect(t).n_cells for t in pending)
    gdem.wait()
    assert n_cells_gpu == n_cells_cpu > 100000
    compare_maps(gmap, omap)


def test_voxel_sorts_agree_bit_for_bit(fdem):
    """voxelGrid(ANY) through the MSD sort written for it and through cub::DeviceRadixSort: same
    voxel counts, same representatives — hence identical raycasting layers — on the dense C4 scan
    (rows of every size: thousands of points near the sensor, single points far out)."""
    wl = syn.WORKLOADS["c4_dense_raycast"]
    cfg = wl.config()
    scans = [syn.make_scan(wl, k) for k in range(3)]
    maps, stats = [], []
    for mode in (capi.VOXEL_SORT_MSD, capi.VOXEL_SORT_LIBRARY):
        m = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution)
        d = fdem.FastDEM(m, cfg)
        d.set_voxel_sort(mode)
        st = [d.integrate_stats(fdem.PointCloud(s["xyzw"], s["intensity"], s["rgb"]),
                                s["T_base_sensor"], s["T_world_base"]) for s in scans]
        maps.append(m)
        stats.append(st)
    for a, b in zip(*stats):
        assert (a.n_kept, a.n_cells, a.n_voxels) == (b.n_kept, b.n_cells, b.n_voxels)
        assert a.voxel_box_violations == 0 and a.n_voxels > 100000
    assert maps[0].getLayers() == maps[1].getLayers()
    for layer in maps[0].getLayers():
        assert np.array_equal(_bits(maps[0].get(layer)), _bits(maps[1].get(layer))), layer
