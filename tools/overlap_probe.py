"""How much of a scan's device time could overlap with the next scan's?  Upper bound probe:
two INDEPENDENT maps on two streams, scans submitted alternately from one host thread."""
import sys, time; sys.path.insert(0, 'tests')
import torch
import fastdem_b200 as fd
from fastdem_b200 import synthetic as syn
wl = syn.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2_lidar64_local"]
dev = torch.device("cuda", 0)
def setup():
    st = torch.cuda.Stream(device=dev)
    m = fd.ElevationMap(wl.map_width, wl.map_height, wl.resolution, "map", device=0, stream=st.cuda_stream)
    return st, m, fd.FastDEM(m, wl.config())
scans = []
for k in range(16):
    s = syn.make_scan(wl, k)
    scans.append(fd.PointCloud(torch.from_numpy(s["xyzw"]).to(dev), None if s["intensity"] is None else torch.from_numpy(s["intensity"]).to(dev),
                               None if s["rgb"] is None else torch.from_numpy(s["rgb"]).to(dev)))
poses = [tuple(fd.api._iso(x) for x in syn.pose(wl, k)) for k in range(4096)]
for n_maps in (1, 2, 3):
    sets = [setup() for _ in range(n_maps)]
    for k in range(50):
        for _, _, d in sets: d.integrate_async(scans[k % 16], *poses[k])
    for _, _, d in sets: d.wait()
    torch.cuda.synchronize()
    N = 3000
    t0 = time.perf_counter()
    for k in range(N):
        for _, _, d in sets: d.integrate_async(scans[k % 16], *poses[50 + k])
    t_enq = time.perf_counter() - t0
    for _, _, d in sets: d.wait()
    torch.cuda.synchronize()
    t = time.perf_counter() - t0
    print(f"{wl.name}: {n_maps} map(s)/stream(s): {N*n_maps/t:9.0f} scans/s total, {1e6*t/(N*n_maps):6.2f} us/scan, host enqueue {1e6*t_enq/(N*n_maps):5.2f} us/scan")
