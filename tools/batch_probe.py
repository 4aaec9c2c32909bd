"""Device-resident throughput: scan-by-scan queue vs fdem_mapper_integrate_batch (S scans per graph)."""
import sys, time; sys.path.insert(0, 'tests')
import torch
import fastdem_b200 as fd
from fastdem_b200 import synthetic as syn
wl = syn.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c2_lidar64_local"]
dev = torch.device("cuda", 0)
n_ring = 64 if wl.points_per_scan < 400000 else 8
scans = []
for k in range(n_ring):
    s = syn.make_scan(wl, k)
    scans.append(fd.PointCloud(torch.from_numpy(s["xyzw"]).to(dev), None if s["intensity"] is None else torch.from_numpy(s["intensity"]).to(dev),
                               None if s["rgb"] is None else torch.from_numpy(s["rgb"]).to(dev)))
poses = [tuple(fd.api._iso(x) for x in syn.pose(wl, k)) for k in range(8192)]
N = 2048 if wl.points_per_scan < 400000 else 256
for S in (0, 2, 4, 8):
    st = torch.cuda.Stream(device=dev)
    m = fd.ElevationMap(wl.map_width, wl.map_height, wl.resolution, "map", device=0, stream=st.cuda_stream)
    d = fd.FastDEM(m, wl.config())
    batches = [[scans[(b * max(S, 1) + j) % n_ring] for j in range(max(S, 1))] for b in range(n_ring)]
    def run(k0, count):
        k = k0
        if S == 0:
            for _ in range(count):
                d.integrate_async(scans[k % n_ring], *poses[k]); k += 1
        else:
            for b in range(count // S):
                d.integrate_batch(batches[b % n_ring], poses[k:k + S], wait=False); k += S
        return k
    k = run(0, 64); d.wait(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(st); k = run(k, N); e1.record(st)
    t_enq = time.perf_counter() - t0
    d.wait(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f"{wl.name}: batch {S or 'off':>3}: {N / ms * 1e3:9.0f} scans/s  {ms / N * 1e3:6.2f} us/scan  host enqueue {t_enq / N * 1e6:5.2f} us/scan")
