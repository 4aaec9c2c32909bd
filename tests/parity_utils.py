"""Helpers shared by the parity tests: run the same scans through the CUDA path and the CPU
oracle and compare every layer."""
from __future__ import annotations

import numpy as np

import oracle_binding as ob

# north_star: bit-exact cell indices / counts, height & variance within 1e-5 relative
RTOL = 1e-5
ATOL = 1e-7
EXACT_LAYERS = {"n_points", "color", "ghost_removal"} | {"_p2_n%d" % i for i in range(5)}


def compare_layer(name, got, want, rtol=RTOL, atol=ATOL):
    """Raises AssertionError with a useful message; returns #cells that differ in bits."""
    assert got.shape == want.shape, f"{name}: shape {got.shape} vs {want.shape}"
    gn, wn = np.isnan(got), np.isnan(want)
    assert np.array_equal(gn, wn), (
        f"{name}: NaN masks differ in {np.count_nonzero(gn != wn)} cells "
        f"(first at {np.argwhere(gn != wn)[:3].tolist()})")
    g, w = got[~gn], want[~wn]
    if name in EXACT_LAYERS:
        if name == "color":
            assert np.array_equal(g.view(np.uint32), w.view(np.uint32)), f"{name}: bits differ"
        else:
            assert np.array_equal(g, w), f"{name}: values differ (must be exact)"
        return 0
    bad = ~np.isclose(g, w, rtol=rtol, atol=atol)
    assert not bad.any(), (
        f"{name}: {bad.sum()} cells outside rtol={rtol}: max abs diff "
        f"{np.abs(g - w).max():.3e}, e.g. got {g[bad][:3]} want {w[bad][:3]}")
    return int(np.count_nonzero(g.view(np.uint32) != w.view(np.uint32)))


def compare_maps(gpu_map, omap, layers=None, rtol=RTOL, atol=ATOL):
    """The two maps must hold exactly the same layer SET (the reference creates intensity /
    color / the raycasting layers lazily, and exists() / getLayers() are part of its API) and
    every layer must match.  Returns {layer: n_cells_with_different_bits}."""
    names = layers if layers is not None else omap.layers()
    if layers is None:
        assert sorted(gpu_map.getLayers()) == sorted(omap.layers()), (
            f"layer sets differ: GPU-only {sorted(set(gpu_map.getLayers()) - set(omap.layers()))}, "
            f"oracle-only {sorted(set(omap.layers()) - set(gpu_map.getLayers()))}")
    report = {}
    for name in names:
        assert gpu_map.exists(name), f"GPU map lacks layer {name}"
        report[name] = compare_layer(name, gpu_map.get(name), omap.get(name), rtol, atol)
    gg, og = gpu_map.geometry(), omap.geometry()
    assert (gg.rows, gg.cols) == (og["rows"], og["cols"])
    assert (gg.start_index[0], gg.start_index[1]) == og["start_index"], "start index differs"
    assert (gg.position[0], gg.position[1]) == og["position"], "map position differs"
    return report


def run_pair(fdem, wl, n_scans, cfg=None, scan_fn=None, cell_sort=None):
    """Integrate n_scans synthetic scans of workload `wl` on both paths.  Returns
    (gpu_map, oracle_map, [gpu stats], [oracle stats])."""
    from fastdem_b200 import synthetic as syn
    cfg = cfg if cfg is not None else wl.config()
    gmap = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution, "map")
    gdem = fdem.FastDEM(gmap, cfg)
    if cell_sort is not None:
        gdem.set_cell_sort(cell_sort)
    omap = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
    odem = ob.OracleFastDEM(omap, cfg)
    gs, os_ = [], []
    for k in range(n_scans):
        s = scan_fn(k) if scan_fn else syn.make_scan(wl, k)
        cloud = fdem.PointCloud(s["xyzw"], s["intensity"], s["rgb"])
        st = gdem.integrate_stats(cloud, s["T_base_sensor"], s["T_world_base"])
        ok, ost, _ = odem.integrate(s["xyzw"], s["T_base_sensor"], s["T_world_base"], s["intensity"], s["rgb"])
        assert bool(st.integrated) == ok, f"scan {k}: integrate() bool differs"
        assert st.n_kept == ost.n_kept, f"scan {k}: n_kept {st.n_kept} vs {ost.n_kept}"
        assert st.n_cells == ost.n_cells, f"scan {k}: n_cells {st.n_cells} vs {ost.n_cells}"
        assert st.n_voxels == ost.n_voxels, f"scan {k}: n_voxels {st.n_voxels} vs {ost.n_voxels}"
        gs.append(st)
        os_.append(ost)
    return gmap, omap, gdem, odem, gs, os_
