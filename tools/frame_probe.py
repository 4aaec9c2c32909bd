"""Per-kernel split of the whole-frame chain (run under ncu --metrics gpu__time_duration.sum)."""
import sys; sys.path.insert(0, 'tests')
import numpy as np
import fastdem_b200 as fd
from fastdem_b200 import synthetic as syn
wl = syn.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2_lidar64_local"]
cfg = wl.config(); cfg.raycasting_enabled = 1
m = fd.ElevationMap(wl.map_width, wl.map_height, wl.resolution); d = fd.FastDEM(m, cfg)
for k in range(4):
    s = syn.make_scan(wl, k)
    d.integrate_stats(fd.PointCloud(s["xyzw"], s["intensity"], s["rgb"]), *syn.pose(wl, k))
    fd.applyUncertaintyFusion(m); fd.applySpatialSmoothing(m, "elevation", 3, 5)
    fd.applyInpainting(m, 3, 2, False); fd.applyFeatureExtraction(m, 0.3, 4)
print("done")
