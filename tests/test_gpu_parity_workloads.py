"""-m gpu: the CUDA path against the CPU oracle on the BASELINE.json workloads, scan by scan,
every layer.  Bit-exact NaN masks / counts / colour bits,
1e-5 relative on heights and variances (north_star tolerance)."""
import numpy as np
import pytest

from fastdem_b200 import synthetic as syn
from parity_utils import compare_maps, run_pair

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cell_sort", [0, 1], ids=["tile", "global"])
@pytest.mark.parametrize("name,n_scans", [
    ("tiny", 12),
    ("c1_vlp16_local", 8),       # BASELINE configs[0]: the reference's CPU-runnable case
    ("c2_lidar64_local", 6),     # configs[1]: the bench workload
    ("c3_rgbd_p2", 7),           # configs[2]: RGBD model + P2 (needs > 5 scans to leave phase 1)
])
def test_workload_parity(fdem, name, n_scans, cell_sort):
    wl = syn.WORKLOADS[name]
    gmap, omap, gdem, odem, gs, os_ = run_pair(fdem, wl, n_scans, cell_sort=cell_sort)
    report = compare_maps(gmap, omap)
    touched = int(np.isfinite(omap.get("elevation")).sum())
    assert touched > 0
    print(f"{name}: {touched} cells with elevation; cells differing in bits per layer: "
          f"{ {k: v for k, v in report.items() if v} }")


def test_c1_basic_layer_policy(fdem):
    """move() clearing policy switch: BASIC layers only, same check."""
    wl = syn.WORKLOADS["c1_vlp16_local"]
    cfg = wl.config()
    cfg.move_clear_policy = 1
    gmap, omap, *_ = run_pair(fdem, wl, 8, cfg)
    compare_maps(gmap, omap)


@pytest.mark.parametrize("bucket_bits", [8, 9, 10])
@pytest.mark.parametrize("name,n_scans", [("c1_vlp16_local", 6), ("c2_lidar64_local", 4), ("c3_rgbd_p2", 7)])
def test_bucket_sizes_give_the_same_map(fdem, monkeypatch, name, n_scans, bucket_bits):
    """The tile path's bucket shape (256 / 512 / 1024 cells; K3t is compiled for all three and
    by default follows the scan density) is a scheduling choice: every shape must reproduce
    the oracle.  (The default-shape runs above cover the switch itself: c1 and c3 go dense
    after their first scan.)"""
    monkeypatch.setenv("FDEM_BUCKET_BITS", str(bucket_bits))
    wl = syn.WORKLOADS[name]
    gmap, omap, *_ = run_pair(fdem, wl, n_scans)
    compare_maps(gmap, omap)


@pytest.mark.parametrize("name,n_scans,batch", [("tiny", 12, 3), ("tiny", 40, 16), ("c1_vlp16_local", 16, 8), ("c2_lidar64_local", 9, 4),
                                                ("c3_rgbd_p2", 10, 5)])
def test_batched_integration_equals_scan_by_scan(fdem, name, n_scans, batch):
    """fdem_mapper_integrate_batch runs a batch as ONE graph in which scan k+1's front half
    overlaps scan k's estimator.  Device-resident inputs (the overlapped schedule), LOCAL maps
    that move every scan, Kalman and P2, intensity and colour channels, batches mixed with
    single scans: statistics and every layer must equal the oracle's scan-by-scan result."""
    import torch
    import oracle_binding as ob
    wl = syn.WORKLOADS[name]
    cfg = wl.config()
    gmap = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution, "map")
    gdem = fdem.FastDEM(gmap, cfg)
    omap = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
    odem = ob.OracleFastDEM(omap, cfg)
    scans = [syn.make_scan(wl, k) for k in range(n_scans)]
    dev = lambda a: None if a is None else torch.from_numpy(a).cuda()
    clouds = [fdem.PointCloud(dev(s["xyzw"]), dev(s["intensity"]), dev(s["rgb"])) for s in scans]
    ostats = []
    for s in scans:
        ok, st, _ = odem.integrate(s["xyzw"], s["T_base_sensor"], s["T_world_base"], s["intensity"], s["rgb"])
        ostats.append(st)
    k = 0
    gstats = []
    while k < n_scans:
        if k == batch:   # a single scan between two batches: the two schedules share all state
            gstats.append(gdem.integrate_stats(clouds[k], scans[k]["T_base_sensor"], scans[k]["T_world_base"]))
            k += 1
            continue
        e = min(k + batch, n_scans)
        poses = [(s["T_base_sensor"], s["T_world_base"]) for s in scans[k:e]]
        if (k // batch) % 2:   # queued batch: the per-scan statistics are fetched afterwards
            assert gdem.integrate_batch(clouds[k:e], poses, wait=False) is None
            gstats += gdem.last_batch_stats(e - k)
        else:
            gstats += gdem.integrate_batch(clouds[k:e], poses)
        k = e
    assert len(gstats) == n_scans
    for g, o in zip(gstats, ostats):
        assert (g.n_kept, g.n_cells, g.integrated) == (o.n_kept, o.n_cells, o.integrated)
    compare_maps(gmap, omap)


def test_batch_with_host_inputs_falls_back(fdem):
    """Host buffers take the scan-after-scan schedule: same results."""
    import oracle_binding as ob
    wl = syn.WORKLOADS["tiny"]
    gmap = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution, "map")
    gdem = fdem.FastDEM(gmap, wl.config())
    omap = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
    odem = ob.OracleFastDEM(omap, wl.config())
    scans = [syn.make_scan(wl, k) for k in range(5)]
    st = gdem.integrate_batch([fdem.PointCloud(s["xyzw"], s["intensity"]) for s in scans],
                              [(s["T_base_sensor"], s["T_world_base"]) for s in scans])
    for s in scans:
        odem.integrate(s["xyzw"], s["T_base_sensor"], s["T_world_base"], s["intensity"], None)
    assert all(x.integrated for x in st)
    compare_maps(gmap, omap)
    with pytest.raises(fdem.FdemError):
        gdem.integrate_batch([fdem.PointCloud(scans[0]["xyzw"])] * 17, [(np.eye(4), np.eye(4))] * 17)


def test_batch_with_jumps_empty_and_filtered_scans(fdem):
    """Inside one batch: a 100 m jump (the whole LOCAL window is vacated: clear_all), a scan whose
    points are all filtered (no move, nothing touched), a scan that lands outside the window's
    previous position — the deferred commit / back prologue must treat them like the per-scan
    pipeline does."""
    import torch
    import oracle_binding as ob
    wl = syn.WORKLOADS["tiny"]
    cfg = wl.config()
    gmap = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution, "map")
    gdem = fdem.FastDEM(gmap, cfg)
    omap = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
    odem = ob.OracleFastDEM(omap, cfg)
    scans = [syn.make_scan(wl, k) for k in range(10)]
    for k, s in enumerate(scans):
        s["T_world_base"] = np.array(s["T_world_base"], dtype=np.float64).copy()
        if k >= 4:
            s["T_world_base"][0, 3] += 100.0       # scans 4.. happen 100 m away
        if k == 6:
            s["xyzw"] = s["xyzw"].copy()
            s["xyzw"][:, :3] *= 1000.0             # every point beyond range_max: filtered, no move
        if k == 8:
            s["T_world_base"][1, 3] -= 3.05        # a sideways hop of 30 cells and a half
    clouds = [fdem.PointCloud(torch.from_numpy(s["xyzw"]).cuda(), torch.from_numpy(s["intensity"]).cuda()) for s in scans]
    gst = gdem.integrate_batch(clouds, [(s["T_base_sensor"], s["T_world_base"]) for s in scans])
    for s, g in zip(scans, gst):
        ok, o, _ = odem.integrate(s["xyzw"], s["T_base_sensor"], s["T_world_base"], s["intensity"], None)
        assert (g.integrated, g.n_kept, g.n_cells) == (int(ok), o.n_kept, o.n_cells)
    assert gst[6].integrated == 0 and gst[6].n_cells == 0
    compare_maps(gmap, omap)
