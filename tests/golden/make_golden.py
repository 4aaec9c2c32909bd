"""Generates the golden fixtures of tests/golden/: the CPU oracle's output (every layer, the scan
statistics, the committed geometry) for short seeded scan streams.

The reference itself cannot be built or imported here (C++ with un-vendored dependencies —
DESIGN.md §2), so these vectors come from the oracle, which is pinned against the reference's own
known-answer tests (tests/test_oracle_golden.py).  What the fixtures add: the oracle's behaviour
on whole scan streams is frozen in the repository — a later edit of oracle/ or of
fastdem_b200/synthetic.py that changes any cell shows up as a diff against committed data, on
the CPU, and the CUDA path is compared with the same committed data on the GPU.

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle_binding as ob  # noqa: E402
from fastdem_b200 import capi, synthetic as syn  # noqa: E402

# name -> (workload, scans, config overrides)
CASES = {
    "tiny_kalman_local": ("tiny", 7, {}),
    "tiny_p2_local": ("tiny", 7, {"estimation_type": capi.EST_P2QUANTILE}),
    "tiny_kalman_raycast": ("tiny", 7, {"raycasting_enabled": 1}),
    "tiny_kalman_global": ("tiny", 5, {"mode": capi.MODE_GLOBAL}),
}


def case_config(name):
    wl_name, n_scans, over = CASES[name]
    wl = syn.WORKLOADS[wl_name]
    cfg = wl.config(ob.default_config)
    for k, v in over.items():
        setattr(cfg, k, v)
    return wl, n_scans, cfg


def input_digest(wl, n_scans):
    h = hashlib.sha256()
    for k in range(n_scans):
        s = syn.make_scan(wl, k)
        for key in ("xyzw", "intensity", "rgb"):
            if s[key] is not None:
                h.update(np.ascontiguousarray(s[key]).tobytes())
        h.update(np.ascontiguousarray(s["T_base_sensor"], dtype=np.float64).tobytes())
        h.update(np.ascontiguousarray(s["T_world_base"], dtype=np.float64).tobytes())
    return h.hexdigest()


def run_oracle(name):
    wl, n_scans, cfg = case_config(name)
    omap = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
    odem = ob.OracleFastDEM(omap, cfg)
    stats = []
    for k in range(n_scans):
        s = syn.make_scan(wl, k)
        ok, st, _ = odem.integrate(s["xyzw"], s["T_base_sensor"], s["T_world_base"], s["intensity"], s["rgb"])
        stats.append([int(ok), st.n_kept, st.n_cells, st.n_voxels])
    out = {"layer:" + n: np.asarray(omap.get(n), dtype=np.float32) for n in omap.layers()}
    g = omap.geometry()
    out["stats"] = np.asarray(stats, dtype=np.int64)
    out["geometry"] = np.asarray([g["rows"], g["cols"], g["start_index"][0], g["start_index"][1]], dtype=np.int64)
    out["position"] = np.asarray(g["position"], dtype=np.float64)
    out["input_sha256"] = np.frombuffer(input_digest(wl, n_scans).encode(), dtype=np.uint8)
    return out


if __name__ == "__main__":
    for name in CASES:
        data = run_oracle(name)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        layers = sorted(k[6:] for k in data if k.startswith("layer:"))
        print(f"{name}: {len(layers)} layers {layers}, {os.path.getsize(path)} bytes")
