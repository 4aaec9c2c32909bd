"""Multi-GPU probe: queue `depth` sharded scans without waiting, then wait; prints per-depth timing.
    python -m torch.distributed.run --nproc-per-node N tools/shard_depth_probe.py [workload]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from fastdem_b200 import capi, sharded
from fastdem_b200 import synthetic as syn


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    wl = syn.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c5_global"]
    cfg = wl.config()
    cfg.mode = capi.MODE_GLOBAL
    n = wl.points_per_scan
    ring = sharded.PeerScanRing(4, n, wl.has_intensity, wl.has_color, device=rank, src=0, replicated=True)
    scans = [syn.make_scan(wl, k) for k in range(4)]
    for j, s in enumerate(scans):
        ring.fill(j, s["xyzw"], s["intensity"], s["rgb"])
    torch.cuda.synchronize()
    dist.barrier()
    tstream = torch.cuda.Stream(device=rank) if os.environ.get("PROBE_TORCH_STREAM") == "1" else None
    sm = sharded.ShardedMapper(wl.map_width, wl.map_height, wl.resolution, cfg, max_points=n, device=rank,
                               stream=tstream.cuda_stream if tstream is not None else 0)
    import ctypes
    import threading

    def watchdog():
        time.sleep(float(os.environ.get("PROBE_WATCHDOG", "25")))
        out = (ctypes.c_uint32 * 17)()
        try:
            fn = sm.lib.fdem_shard_debug_flags   # FDEM_PROBES=1 builds only
        except AttributeError:
            return
        fn.restype, fn.argtypes = ctypes.c_int32, [ctypes.c_void_p, ctypes.c_void_p]
        rc = fn(sm._h, out)
        print(f"[watchdog rank {rank}] rc={rc} ready={list(out[0:world])} consumed={list(out[8:8 + world])} seq={out[16]}",
              flush=True)

    threading.Thread(target=watchdog, daemon=True).start()
    k = 0
    from fastdem_b200 import api
    clouds = [ring.cloud(j, n) for j in range(4)]
    poses = [tuple(api._iso(t) for t in syn.pose(wl, kk)) for kk in range(512)]   # the host must not be the bound
    depths = [int(x) for x in os.environ.get("PROBE_DEPTHS", "1,2,3,4,8,20,40").split(",")]
    for depth in depths:
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(depth):
            sm.integrate_async(clouds[k % 4], *poses[k % 512])
            k += 1
        t_enq = time.perf_counter() - t0
        st = sm.wait()
        dt = time.perf_counter() - t0
        print(f"rank {rank} depth {depth}: {1e6 * dt / depth:.1f} us/scan  (host enqueue {1e6 * t_enq / depth:.1f} us/scan, "
              f"n_cells {st.n_cells})", flush=True)
    dist.barrier()
    sm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
