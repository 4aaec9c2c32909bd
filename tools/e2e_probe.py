"""Where does the streaming (host-input) path of a workload spend its period?
    python tools/e2e_probe.py [workload]
Prints us/scan for: device inputs through submit/collect; host inputs; host inputs without
raycasting; the H2D copies alone (same sizes, same three staging-sized targets)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fastdem_b200 as fd
from fastdem_b200 import synthetic as syn

name = sys.argv[1] if len(sys.argv) > 1 else "c4_dense_raycast"
wl = syn.WORKLOADS[name]
dev = torch.device("cuda", 0)
host = [syn.make_scan(wl, k) for k in range(4)]
poses = [tuple(fd.api._iso(t) for t in syn.pose(wl, k)) for k in range(256)]


def clouds(to_dev):
    out = []
    for s in host:
        pc = fd.PointCloud()
        if to_dev:
            pc.xyzw = torch.from_numpy(s["xyzw"]).to(dev)
            pc.intensity = None if s["intensity"] is None else torch.from_numpy(s["intensity"]).to(dev)
            pc.color = None if s["rgb"] is None else torch.from_numpy(s["rgb"]).to(dev)
        else:
            keep = [torch.from_numpy(s["xyzw"]).pin_memory()]
            pc.xyzw = keep[0].numpy()
            pc.intensity = pc.color = None
            if s["intensity"] is not None:
                keep.append(torch.from_numpy(s["intensity"]).pin_memory())
                pc.intensity = keep[-1].numpy()
            if s["rgb"] is not None:
                keep.append(torch.from_numpy(s["rgb"]).pin_memory())
                pc.color = keep[-1].numpy()
            pc._pinned = keep
        out.append(pc)
    return out


def run(label, cfg, pcs, lookahead=2, n=200):
    m = fd.ElevationMap(wl.map_width, wl.map_height, wl.resolution, "map")
    d = fd.FastDEM(m, cfg)
    k = 0

    def loop(count):
        nonlocal k
        pend = []
        t_submit = 0.0
        for _ in range(count):
            t0 = time.perf_counter()
            pend.append(d.submit(pcs[k % 4], *poses[k % 256]))
            t_submit += time.perf_counter() - t0
            k += 1
            if len(pend) > lookahead:
                d.collect(pend.pop(0))
        for t in pend:
            d.collect(t)
        return t_submit

    loop(20)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ts = loop(n)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{label:54s} {1e6 * dt / n:8.1f} us/scan   (host time inside submit {1e6 * ts / n:6.1f} us/scan)", flush=True)


cfg = wl.config()
run("device inputs, submit/collect", cfg, clouds(True))
run("host inputs, submit/collect, lookahead 2", cfg, clouds(False))
run("host inputs, submit/collect, lookahead 1", cfg, clouds(False), lookahead=1)
cfg2 = wl.config()
cfg2.raycasting_enabled = 0
run("host inputs, raycasting off", cfg2, clouds(False))
run("device inputs, raycasting off", cfg2, clouds(True))

# the copies alone
pins = clouds(False)
n_pts = host[0]["xyzw"].shape[0]
targets = [(torch.empty((n_pts, 4), device=dev), torch.empty(n_pts, device=dev)) for _ in range(3)]
cs = torch.cuda.Stream()
srcs = [(torch.from_numpy(p.xyzw), None if p.intensity is None else torch.from_numpy(p.intensity)) for p in pins]
with torch.cuda.stream(cs):
    for it in range(220):
        if it == 20:
            cs.synchronize()
            t0 = time.perf_counter()
        a, b = targets[it % 3]
        x, i = srcs[it % 4]
        a.copy_(x, non_blocking=True)
        if i is not None:
            b.copy_(i, non_blocking=True)
    cs.synchronize()
print(f"{'H2D copies alone (xyzw + intensity per scan)':54s} {1e6 * (time.perf_counter() - t0) / 200:8.1f} us/scan")
