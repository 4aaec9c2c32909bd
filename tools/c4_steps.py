"""Run a few scans of a workload through the C-ABI with device-resident inputs (ncu target).
usage: python tools/c4_steps.py [workload] [n_scans]"""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import torch
import fastdem_b200 as fd
from fastdem_b200 import synthetic as syn

name = sys.argv[1] if len(sys.argv) > 1 else "c4_dense_raycast"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
wl = syn.WORKLOADS[name]
m = fd.ElevationMap(wl.map_width, wl.map_height, wl.resolution)
d = fd.FastDEM(m, wl.config())
clouds = []
for k in range(4):
    s = syn.make_scan(wl, k)
    clouds.append(fd.PointCloud(torch.from_numpy(s['xyzw']).cuda(),
                                None if s['intensity'] is None else torch.from_numpy(s['intensity']).cuda(),
                                None if s['rgb'] is None else torch.from_numpy(s['rgb']).cuda()))
for k in range(n):
    st = d.integrate_stats(clouds[k % 4], *syn.pose(wl, k))
print(st.n_kept, st.n_cells, st.n_voxels)
