#!/usr/bin/env python
"""bench.py — scans/s and Mpoints/s of FastDEM::integrate() on B200, beside the CPU path.

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A "step" is one integrate() of one synthetic scan of the workload (default: BASELINE.json
configs[1], 64-beam LiDAR 131K pts, 30x30 m @ 0.05 m, Kalman, LOCAL).  One JSON line on
stdout (rank 0).

  value    scans/s with the scans already resident in HBM (device pointers through the C-ABI,
           queued back to back, no host sync inside the timed region), submitted 16 at a time
           through fdem_mapper_integrate_batch — identical results to 16 integrate() calls, with
           scan k+1's front half overlapping scan k's estimator (--batch 1: one call per scan;
           the JSON also carries that figure as value_scan_by_scan).  The timed steps cycle
           through distinct device-resident scans totalling more than L2 (--l2 flush: a 256 MiB
           fill before every step instead); one CUDA-event pair on the kernels' stream around
           the K steps, ms_per_step = that time / K.
  e2e      the same metric through the public API with HOST (pinned) buffers: every step
           copies the scan host->device, runs integrate() synchronously and reads the scan
           stats + committed geometry back.  Wall clock, barrier + synchronize on both sides.
  roofline the slowest pipeline stage, its algorithmic bytes over
           its mean device time, against the MEASURED copy bandwidth (MEASURED_PEAKS.json).
  cpu_baseline  the CPU oracle (a line-by-line restatement of the reference path; the
           reference itself cannot be built here — no Eigen / nanoGrid) on 1 host core.

N > 1 (torchrun): LOCAL mapping does not shard — each rank integrates its own
robot's scan stream into its own map ("replicas", weak scaling, no data-path collective);
torch.distributed (NCCL) is used only for the barrier and the max-over-ranks of the time.
`--workload c5_global` instead row-stripes ONE global map over the ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))
sys.path.insert(0, str(REPO / "tests"))

import numpy as np


def _peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.rows.append((time.perf_counter(), line.strip()))

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

    def wait_first(self, timeout_s=5.0):
        """nvidia-smi needs a few hundred ms before its first row; wait for it."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.02)

    def stop(self, region=None):
        """region = (t0, t1) perf_counter bounds of the timed steps: rows inside it are counted
        separately from rows taken while the same load kept running around it."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        in_region = 0
        for ts, r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            if region is not None and region[0] <= ts <= region[1] + 0.1:
                in_region += 1
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None,
                "samples": len(sm), "samples_in_timed_region": in_region,
                "how": "nvidia-smi -lms 100 started before warm-up and kept running while the same "
                       "step loop continues for >=1 s after the timed steps (the timed region "
                       "itself can be shorter than one sampling period)",
                "reasons": sorted(reasons)}


# pipeline stage (fdem_mapper_stage_times) -> the kernels it runs
STAGE_KERNELS = {
    "preprocess_bin": "preprocess_bin_kernel (K1)",
    "commit_move_clear": "commit_move_clear_kernel (K2)",
    "sort_by_cell": "scatter_records_kernel (L1 of the 2-level sort)",
    "segreduce_estimate": "tile_estimate_kernel<8|9|10> (K3t: per-bucket sort + segmented reduce + estimator)",
    "voxel_raycast": "voxel_keys32 + cub radix sort + voxel_select + ray_keys + cub radix sort + "
                     "raycast_scan + raycast_resolve (not HBM bound: per-ray DDA, L1/L2 resident)",
}


def algorithmic_bytes(wl, n_points, stats_list, has_i, has_c, p2):
    """B = N*b_pt + C*b_cell + C_prev*4, split per pipeline stage.
    Compulsory traffic only — no sort scratch, no intermediate copies."""
    b_pt = 16 + (4 if has_i else 0) + (3 if has_c else 0)
    b_cell = (124 if p2 else 76) + (8 if has_i else 0) + (4 if has_c else 0)
    C = statistics.mean(s.n_cells for s in stats_list) if stats_list else 0.0
    Nv = statistics.mean(s.n_kept for s in stats_list) if stats_list else 0.0
    per_stage = {
        "preprocess_bin": n_points * 16.0,                       # every input point read once
        "commit_move_clear": C * 4.0,                            # last scan's obstacle cells
        "sort_by_cell": 0.0,                                     # pure scratch traffic
        "segreduce_estimate": C * b_cell + Nv * (b_pt - 16.0),   # cell state + per-point channels
    }
    total = n_points * b_pt + C * b_cell + C * 4.0
    V = statistics.mean(s.n_voxels for s in stats_list) if stats_list else 0.0
    if V > 0:
        # raycasting (not HBM bound — an L1/L2-resident per-ray DDA; the figure is its compulsory
        # traffic only): kept points read for the voxel keys, one point per traced ray, and the
        # per-scan clear of the `raycasting` layer + read of elevation + write of the ray minimum
        # over the whole map (raycasting.cpp:242, 188-214)
        M = int(round(wl.map_width / wl.resolution)) * int(round(wl.map_height / wl.resolution))
        per_stage["voxel_raycast"] = Nv * 16.0 + V * 16.0 + M * 12.0
        total += per_stage["voxel_raycast"]
    return per_stage, total


def run_reference(args, wl, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  It cannot be
    built here (no Eigen, nanoGrid un-vendored), so this times the oracle port — a
    line-by-line restatement incl. the by-value cloud copy, 36-byte covariances and the
    unordered_map rasteriser — single-threaded, exactly like the reference path."""
    if rank != 0:
        return None
    import oracle_binding as ob
    from fastdem_b200 import synthetic as syn
    cfg = wl.config()
    omap = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
    odem = ob.OracleFastDEM(omap, cfg)
    ring = [syn.make_scan(wl, k) for k in range(min(8, args.warmup + args.steps))]
    k = 0
    for _ in range(args.warmup):
        s = ring[k % len(ring)]
        odem.integrate(s["xyzw"], *pose_for(wl, k), s["intensity"], s["rgb"])
        k += 1
    total = 0.0
    budget_s = 60.0
    done = 0
    for _ in range(args.steps):
        s = ring[k % len(ring)]
        _, _, el = odem.integrate(s["xyzw"], *pose_for(wl, k), s["intensity"], s["rgb"])
        total += el
        k += 1
        done += 1
        if total > budget_s:
            break
    n = wl.points_per_scan
    ms = 1e3 * total / max(done, 1)
    val = done / total

    # Courtesy upper bound: the reference path is single-threaded, so one map
    # cannot use more than one core; this is what the box's cores deliver on INDEPENDENT maps
    # (one robot per core), the CPU counterpart of the replicas the GPU arm runs at N > 1.
    all_cores = None
    try:
        import threading
        ncores = os.cpu_count() or 1
        counts = [0] * ncores
        stop = time.perf_counter() + 5.0

        def worker(t):
            m_ = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
            d_ = ob.OracleFastDEM(m_, cfg)
            kk = 0
            while time.perf_counter() < stop:
                s_ = ring[kk % len(ring)]
                d_.integrate(s_["xyzw"], *pose_for(wl, kk), s_["intensity"], s_["rgb"])   # ctypes releases the GIL
                kk += 1
            counts[t] = kk

        t0 = time.perf_counter()
        th = [threading.Thread(target=worker, args=(t,)) for t in range(ncores)]
        for x in th:
            x.start()
        for x in th:
            x.join()
        all_cores = {"value": sum(counts) / (time.perf_counter() - t0), "unit": "scans/s", "cores": ncores,
                     "what": "independent maps, one per host core (the single-map path cannot be threaded)"}
    except Exception as e:
        all_cores = {"error": repr(e)}
    return {
        "impl": "reference", "metric": "integrate_scans_per_sec", "value": val, "unit": "scans/s",
        "mpoints_per_s": val * n / 1e6, "n_gpus": world, "steps": done, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "description": wl.description, "points_per_scan": n},
        "cpu_baseline": {"value": val, "unit": "scans/s", "cores": 1, "kind": "port",
                         "sample": f"{done} scans of {wl.name} on 1 host core (reference path is single-threaded; "
                                   f"{os.cpu_count()} cores on the box)"},
        "e2e": {"value": val, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "all_cores_independent_maps": all_cores,
    }


def frame_pipeline(fd, wl, host_scans, cpu_seconds):
    """The reference's whole per-frame chain (its published 48.3 ms/frame on a Jetson Orin =
    mapping + uncertainty fusion + raycasting + spike removal + inpainting, README.md:59 /
    assets/fastdem_jetson_benchmark.svg:1999-2098), plus feature extraction: integrate() with
    raycasting on, then the post-process functions, synchronously, host (pinned) input — on
    the GPU and on the CPU oracle (bounded sample).  Reported beside the headline metric."""
    import oracle_binding as ob
    from fastdem_b200 import synthetic as syn
    cfg = wl.config()
    cfg.raycasting_enabled = 1

    def chain_gpu(m):
        fd.applyUncertaintyFusion(m)
        fd.applySpatialSmoothing(m, "elevation", 3, 5)
        fd.applyInpainting(m, 3, 2, False)
        fd.applyFeatureExtraction(m, 0.3, 4)

    def chain_cpu(m):
        ob.uncertainty_fusion(m)
        ob.spatial_smoothing(m, "elevation", 3, 5)
        m.inpaint(3, 2, False)
        ob.feature_extraction(m, 0.3, 4)

    gmap = fd.ElevationMap(wl.map_width, wl.map_height, wl.resolution, "map")
    gdem = fd.FastDEM(gmap, cfg)
    k = 0
    for _ in range(5):
        gdem.integrate_stats(host_scans[k % len(host_scans)], *syn.pose(wl, k))
        chain_gpu(gmap)
        k += 1
    n_gpu = 40
    t0 = time.perf_counter()
    t_int = 0.0
    for _ in range(n_gpu):
        t1 = time.perf_counter()
        gdem.integrate_stats(host_scans[k % len(host_scans)], *syn.pose(wl, k))
        t_int += time.perf_counter() - t1
        chain_gpu(gmap)
        k += 1
    gpu_ms = 1e3 * (time.perf_counter() - t0) / n_gpu
    gpu_int_ms = 1e3 * t_int / n_gpu

    omap = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
    odem = ob.OracleFastDEM(omap, cfg)
    kk, tot, n_cpu, tot_int = 0, 0.0, 0, 0.0
    while (tot < cpu_seconds and n_cpu < 200) or n_cpu < 2:
        s = host_scans[kk % len(host_scans)]
        t1 = time.perf_counter()
        odem.integrate(s.xyzw, *syn.pose(wl, kk), s.intensity, s.color)
        t2 = time.perf_counter()
        chain_cpu(omap)
        t3 = time.perf_counter()
        if kk >= 1:  # first frame warms the allocator
            tot += t3 - t1
            tot_int += t2 - t1
            n_cpu += 1
        kk += 1
    return {"what": "integrate(raycasting on) + applyUncertaintyFusion + applySpatialSmoothing(elevation,3,5) + "
                    "applyInpainting(3,2) + applyFeatureExtraction(0.3,4), synchronous, host input",
            "gpu_ms_per_frame": gpu_ms, "gpu_frames_per_s": 1e3 / gpu_ms, "gpu_integrate_ms": gpu_int_ms,
            "cpu_ms_per_frame": 1e3 * tot / n_cpu, "cpu_integrate_ms": 1e3 * tot_int / n_cpu,
            "cpu_frames": n_cpu, "cpu_cores": 1, "gpu_frames": n_gpu}


def pose_for(wl, k):
    from fastdem_b200 import synthetic as syn
    return syn.pose(wl, k)


def main():
    # stdout carries exactly ONE JSON line: libraries (NCCL's version banner, torchrun notices)
    # write to fd 1 as well, so park fd 1 on stderr until the result is ready
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_lidar64_local")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget")
    ap.add_argument("--transport", default="peer", choices=["peer", "nccl"],
                    help="c5_global at N>1: how the stripes get the scan — 'peer': every rank's K1 reads it in "
                         "place from the ingest rank's HBM over NVLink (CUDA IPC); 'nccl': dist.broadcast")
    ap.add_argument("--batch", type=int, default=16,
                    help="scans per fdem_mapper_integrate_batch call in the device-resident `value` loop "
                         "(1 = one fdem_mapper_integrate_async per scan)")
    ap.add_argument("--no-frame", action="store_true", help="skip the whole-frame (mapping + post-process) block")
    ap.add_argument("--l2", default="ring", choices=["ring", "flush", "none"],
                    help="ring: the timed steps cycle through distinct device-resident scans whose total "
                         "size exceeds L2 (inputs always cold, the persistent map stays warm, steps back "
                         "to back); flush: a 256 MiB fill before every timed step (everything cold)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from fastdem_b200 import synthetic as syn
    wl = syn.WORKLOADS[args.workload]

    if args.impl == "reference":
        out = run_reference(args, wl, rank, world)
        if out is not None:
            emit(out)
        return 0

    args.warmup = max(args.warmup, 3)
    import torch
    import torch.distributed as dist
    import fastdem_b200 as fd

    if not torch.cuda.is_available():
        emit({"error": "no CUDA device: the hot path has no CPU fallback"})
        return 2
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    sharded = wl.name == "c5_global" and world > 1
    cfg = wl.config()
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        if sharded:
            from fastdem_b200.sharded import stripe_bounds
            rows = int(round(wl.map_width / wl.resolution))
            r0, r1 = stripe_bounds(rows, world, rank)
            gmap = fd.ElevationMap(wl.map_width, wl.map_height, wl.resolution, "map", device=local_rank,
                                   stream=stream.cuda_stream, row_stripe=(r0, r1))
        else:
            gmap = fd.ElevationMap(wl.map_width, wl.map_height, wl.resolution, "map", device=local_rank,
                                   stream=stream.cuda_stream)
        dem = fd.FastDEM(gmap, cfg)

        # distinct scans, generated once; replicas shift the scan index so ranks differ
        n_ring = 8 if wl.points_per_scan <= 400_000 else 4
        base = 0 if sharded else 1000 * rank
        n = wl.points_per_scan
        probe = syn.make_scan(wl, base)
        has_i, has_c = probe["intensity"] is not None, probe["rgb"] is not None
        scan_bytes = n * (16 + (4 if has_i else 0) + (3 if has_c else 0))
        L2_BYTES = 126 << 20
        # device ring: distinct scans totalling > L2 (ring mode) so every timed step reads inputs
        # that are not in L2; the first n_ring of them double as the host-side ring
        n_dev = max(n_ring, -(-(160 << 20) // scan_bytes)) if args.l2 == "ring" else n_ring
        n_dev = min(n_dev, 512)
        all_scans = [probe] + [syn.make_scan(wl, base + k) for k in range(1, n_dev)]
        host = all_scans[:n_ring]
        dev_scans, pin_scans = [], []
        peer = sharded and args.transport == "peer"
        ring = None
        if peer:
            # the scans live ONCE, in the ingest rank's HBM; every other rank maps them (CUDA IPC)
            # and its kernels read them across NVLink while binning: no per-scan collective
            from fastdem_b200.sharded import PeerScanRing
            ring = PeerScanRing(n_dev, n, has_i, has_c, device=local_rank, src=0)
            for j, s in enumerate(all_scans):
                ring.fill(j, s["xyzw"], s["intensity"], s["rgb"])
            torch.cuda.synchronize(dev)
            dist.barrier()
            dev_scans = [ring.cloud(j, n) for j in range(n_dev)]
        else:
            for s in all_scans:
                d = dict(xyzw=torch.from_numpy(s["xyzw"]).to(dev),
                         intensity=None if not has_i else torch.from_numpy(s["intensity"]).to(dev),
                         rgb=None if not has_c else torch.from_numpy(s["rgb"]).to(dev))
                dev_scans.append(fd.PointCloud(d["xyzw"], d["intensity"], d["rgb"]))
        all_scans = None
        for s in host:
            p = dict(xyzw=torch.from_numpy(s["xyzw"]).pin_memory(),
                     intensity=None if not has_i else torch.from_numpy(s["intensity"]).pin_memory(),
                     rgb=None if not has_c else torch.from_numpy(s["rgb"]).pin_memory())
            pc = fd.PointCloud()
            pc.xyzw = p["xyzw"].numpy()
            pc.intensity = None if not has_i else p["intensity"].numpy()
            pc.color = None if not has_c else p["rgb"].numpy()
            pc._pinned = p
            pin_scans.append(pc)
        flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

        _pose_cache = {}

        def pose_of(kk):  # building the 4x4s with numpy costs more CPU time than enqueueing a scan
            kk = kk % 4096
            if kk not in _pose_cache:
                a, b = syn.pose(wl, base + kk)
                _pose_cache[kk] = (fd.api._iso(a), fd.api._iso(b))
            return _pose_cache[kk]

        for kk in range(min(4096, args.warmup + 4 * args.steps + 200)):
            pose_of(kk)

        if peer:
            def submit_dev(kk):
                dem.integrate_async(dev_scans[kk % n_dev], *pose_of(kk))
        elif sharded:
            # every rank needs the scan: rank 0 (the ingest rank) broadcasts it over NCCL/NVLink
            # inside the timed step; the other ranks integrate out of their receive buffers
            rx = dev_scans[0]

            def submit_dev(kk):
                c = dev_scans[kk % n_dev] if rank == 0 else rx
                dist.broadcast(c.xyzw, src=0)
                if has_i:
                    dist.broadcast(c.intensity, src=0)
                if has_c:
                    dist.broadcast(c.color, src=0)
                dem.integrate_async(c, *pose_of(kk))
        else:
            def submit_dev(kk):
                dem.integrate_async(dev_scans[kk % n_dev], *pose_of(kk))

        # Batched submission for the timed `value` loop: S consecutive scans per
        # fdem_mapper_integrate_batch call (same results as S integrate() calls; scan k+1's front
        # half overlaps scan k's estimator inside one graph).  Raycasting and the sharded global map
        # keep the scan-by-scan queue.
        S = max(1, min(16, args.batch))
        if (sharded and not peer) or cfg.raycasting_enabled:
            S = 1   # the NCCL transport broadcasts inside every step; raycasting is scan by scan

        def submit_many(k0, count):
            kk = k0
            while count - (kk - k0) >= S and S > 1:
                dem.integrate_batch([dev_scans[(kk + j) % n_dev] for j in range(S)],
                                    [pose_of(kk + j) for j in range(S)], wait=False)
                kk += S
            while kk - k0 < count:
                submit_dev(kk)
                kk += 1
            return kk

        def barrier():
            torch.cuda.synchronize(dev)
            if distributed:
                dist.barrier()
            torch.cuda.synchronize(dev)

        # ── warm-up (also sizes every scratch buffer) ──
        sampler = ClockSampler(local_rank)
        sampler.start()
        k = 0
        for _ in range(args.warmup):
            submit_dev(k)
            k += 1
        dem.wait()
        if S > 1:   # build the batch graph (and let the bucket shape settle) outside the timed region
            for _ in range(3):
                k = submit_many(k, S)
                dem.wait()

        # ── timed region: device-resident inputs, CUDA events on the kernels' stream ──
        sampler.wait_first()
        l0, lib0 = dem.launch_count(), dem.library_launch_count()
        stats_ring = []
        if args.l2 == "flush":
            # everything cold: 256 MiB fill before each step, one event pair per step
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
                  for _ in range(args.steps)]
            barrier()
            t_cpu0 = time.perf_counter()
            for i in range(args.steps):
                flush_buf.fill_(i & 0xFF)
                ev[i][0].record(stream)
                submit_dev(k)
                ev[i][1].record(stream)
                k += 1
            cpu_enqueue_s = time.perf_counter() - t_cpu0
            last = dem.wait()
            barrier()
            total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
        else:
            # steps back to back; in ring mode each step's scan has been pushed out of L2 by the
            # >126 MB of other scans read since its last use
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            t_cpu0 = time.perf_counter()
            e0.record(stream)
            k = submit_many(k, args.steps)
            e1.record(stream)
            cpu_enqueue_s = time.perf_counter() - t_cpu0
            last = dem.wait()
            barrier()
            total_ms = float(e0.elapsed_time(e1))
        t_region = (t_cpu0, time.perf_counter())
        launches = dem.launch_count() - l0
        # the same loop through the per-scan call (what `value` was before batching), for reference
        value_scan_by_scan = None
        if S > 1 and args.l2 != "flush":
            n2 = max(args.steps // 2, 1)
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            f0.record(stream)
            for _ in range(n2):
                submit_dev(k)
                k += 1
            f1.record(stream)
            dem.wait()
            barrier()
            value_scan_by_scan = n2 * (1 if sharded else world) / (float(f0.elapsed_time(f1)) * 1e-3)
        lib_launches = dem.library_launch_count() - lib0
        # keep the identical load running until nvidia-smi has seen >= 1 s of it (same step
        # count on every rank: the sharded mode broadcasts inside each step)
        per_step_s = max(total_ms * 1e-3 / max(args.steps, 1), 1e-6)
        extra = int(min(200000, max(0.0, 1.2 - (t_region[1] - t_region[0])) / per_step_s))
        if distributed:
            t_extra = torch.tensor([extra], device=dev, dtype=torch.int64)
            dist.all_reduce(t_extra, op=dist.ReduceOp.MAX)
            extra = int(t_extra.item())
        for i in range(extra):
            if args.l2 == "flush":
                flush_buf.fill_(i & 0xFF)
            submit_dev(k)
            k += 1
            if (i & 255) == 255:
                dem.wait()
        dem.wait()
        clocks = sampler.stop(t_region)

        # ── stage attribution pass (separate from the timed region: bracketing every stage
        #    with events costs ~2.7 us per event and forces one launch per kernel) ──
        dem.set_stage_timing(True)
        for i in range(min(args.steps, 64)):
            if args.l2 == "flush":
                flush_buf.fill_(i & 0xFF)
            submit_dev(k)
            k += 1
        dem.wait()
        stage_ms, stage_scans = dem.stage_times()
        dem.set_stage_timing(False)

        # per-scan statistics for the roofline's algorithmic bytes (a few synchronous scans)
        for j in range(min(8, args.steps)):
            stats_ring.append(dem.integrate_stats(dev_scans[k % n_dev], *pose_of(k)))
            k += 1

        # ── e2e: public API, HOST (pinned) buffers.  Every step copies that step's scan
        #    host->device and reads that step's stats + committed geometry back; the streaming
        #    form submit(k+1); collect(k) lets the copy of the next scan overlap the kernels of
        #    the current one (double-buffered staging on a copy stream inside the library) ──
        e2e_steps = args.steps
        if peer:
            # host scan on the ingest rank -> a ring slot in its HBM -> barrier -> every stripe
            # integrates straight out of that slot (peer reads); the next barrier also keeps the
            # ingest rank from rewriting a slot a stripe may still be reading
            def e2e_step(kk):
                if rank == 0:
                    pp_ = pin_scans[kk % n_ring]._pinned
                    ring.fill(kk % n_dev, pp_["xyzw"], pp_["intensity"] if has_i else None,
                              pp_["rgb"] if has_c else None)
                    torch.cuda.synchronize(dev)
                dist.barrier()
                return dem.integrate_stats(dev_scans[kk % n_dev], *pose_of(kk))
        elif sharded:
            # host scan on the ingest rank -> its GPU -> NCCL broadcast -> every stripe integrates;
            # the per-step result read is the stats of this rank's stripe
            def e2e_step(kk):
                if rank == 0:
                    p = pin_scans[kk % n_ring]._pinned
                    rx.xyzw.copy_(p["xyzw"], non_blocking=True)
                    if has_i:
                        rx.intensity.copy_(p["intensity"], non_blocking=True)
                    if has_c:
                        rx.color.copy_(p["rgb"], non_blocking=True)
                dist.broadcast(rx.xyzw, src=0)
                if has_i:
                    dist.broadcast(rx.intensity, src=0)
                if has_c:
                    dist.broadcast(rx.color, src=0)
                return dem.integrate_stats(rx, *pose_of(kk))
        if sharded:
            if not peer:
                rx = fd.PointCloud(torch.empty_like(dev_scans[0].xyzw),
                                   None if not has_i else torch.empty_like(dev_scans[0].intensity),
                                   None if not has_c else torch.empty_like(dev_scans[0].color))
            for _ in range(3):
                e2e_step(k)
                k += 1
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step(k)
                k += 1
            barrier()
            e2e_s = e2e_sync_s = time.perf_counter() - t0
        for _ in range(0 if sharded else 3):
            dem.integrate_stats(pin_scans[k % n_ring], *pose_of(k))
            k += 1
        barrier()
        t0 = time.perf_counter()
        prev = None
        for _ in range(0 if sharded else e2e_steps):
            t = dem.submit(pin_scans[k % n_ring], *pose_of(k))
            k += 1
            if prev is not None:
                dem.collect(prev)
            prev = t
        if prev is not None:
            dem.collect(prev)
        barrier()
        if not sharded:
            e2e_s = time.perf_counter() - t0
        # the same scans as sensor_msgs/PointCloud2 bodies (what a ROS driver hands the node:
        # x, y, z, intensity[, rgb] packed per point — 16 B instead of 20 B for a LiDAR point),
        # parsed on the device inside K1: fewer bytes over PCIe per scan
        pc2_value = pc2_bytes = None
        if not sharded:
            msgs = []
            for s_ in host:
                mm = fd.PointCloud2.from_arrays(s_["xyzw"][:, :3], s_["intensity"], s_["rgb"])
                t_ = torch.from_numpy(np.ascontiguousarray(mm.data)).pin_memory()
                msgs.append((fd.PointCloud2(t_.numpy(), mm.width, mm.height, mm.point_step, mm.fields), t_))
            pc2_bytes = int(msgs[0][0].point_step) * n
            for _ in range(3):
                dem.collect(dem.submit_pointcloud2(msgs[k % n_ring][0], *pose_of(k)))
                k += 1
            barrier()
            t0 = time.perf_counter()
            prev = None
            for _ in range(e2e_steps):
                t = dem.submit_pointcloud2(msgs[k % n_ring][0], *pose_of(k))
                k += 1
                if prev is not None:
                    dem.collect(prev)
                prev = t
            dem.collect(prev)
            barrier()
            pc2_value = e2e_steps * world / (time.perf_counter() - t0)
        # same thing fully synchronous (one scan in flight): integrate() per step
        barrier()
        t0 = time.perf_counter()
        for _ in range(0 if sharded else e2e_steps):
            dem.integrate_stats(pin_scans[k % n_ring], *pose_of(k))
            k += 1
        barrier()
        if not sharded:
            e2e_sync_s = time.perf_counter() - t0
        # context: what the PCIe link does for this scan size (pinned H2D, CUDA events)
        hb = pin_scans[0]._pinned["xyzw"]
        db_ = torch.empty((n, 4), dtype=torch.float32, device=dev)
        for _ in range(3):
            db_.copy_(hb, non_blocking=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(20):
            db_.copy_(hb, non_blocking=True)
        e1.record(stream)
        torch.cuda.synchronize(dev)
        h2d_gbs = 20 * hb.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9

    # ── reduce over ranks: max time ──
    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_s = float(t[0]), float(t[1])
    units = args.steps * (1 if sharded else world)   # scans all ranks processed
    value = units / (total_ms / 1e3)
    e2e_value = (e2e_steps * (1 if sharded else world)) / e2e_s

    out = None
    if rank == 0:
        peak, peak_src = _peaks()
        per_stage_bytes, total_bytes = algorithmic_bytes(
            wl, n, stats_ring, has_i, has_c, cfg.estimation_type == fd.EST_P2QUANTILE)
        stage_avg = {s: (ms / max(stage_scans, 1)) for s, ms in stage_ms.items()}
        timed = {s: v for s, v in stage_avg.items() if s != "h2d"}
        dominant = max(timed, key=timed.get)
        dom_ms = timed[dominant]
        dom_bytes = per_stage_bytes.get(dominant, 0.0)
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        traffic = None
        try:  # measured DRAM bytes per launch of that kernel (one ncu --set full capture)
            traffic = json.loads((REPO / "profiles" / "r1_traffic.json").read_text()).get(wl.name, {}).get(dominant)
        except Exception:
            pass
        roofline = {
            "bound": "hbm", "kernel": dominant, "kernel_names": STAGE_KERNELS.get(dominant), "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": dom_bytes,
            "kernel_ms": dom_ms,
            "stage_ms": stage_avg,
            "stage_ms_note": "separate pass, one launch per kernel, ~2.7 us event overhead inside each stage",
            "stage_share": {s: v / max(sum(timed.values()), 1e-12) for s, v in timed.items()},
            "pipeline": {"algorithmic_bytes_per_scan": total_bytes,
                         "achieved": total_bytes / ((total_ms / args.steps) * 1e-3) / 1e9,
                         "frac": total_bytes / ((total_ms / args.steps) * 1e-3) / 1e9 / peak},
        }

        # ── CPU baseline: the oracle on one host core, bounded sample ──
        import oracle_binding as ob
        omap = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
        odem = ob.OracleFastDEM(omap, cfg)
        cpu_total, cpu_n, kk = 0.0, 0, 0
        cpu_budget = args.cpu_seconds if world == 1 else 1.0   # full sample at N=1; a token one beside N>1
        for _ in range(2):
            s = host[kk % n_ring]
            odem.integrate(s["xyzw"], *syn.pose(wl, base + kk), s["intensity"], s["rgb"])
            kk += 1
        while cpu_total < cpu_budget and cpu_n < 2000:
            s = host[kk % n_ring]
            _, _, el = odem.integrate(s["xyzw"], *syn.pose(wl, base + kk), s["intensity"], s["rgb"])
            cpu_total += el
            cpu_n += 1
            kk += 1
        cpu_val = cpu_n / cpu_total

        frame = None
        if not args.no_frame and not sharded and world == 1 and wl.name in ("c1_vlp16_local", "c2_lidar64_local"):
            try:
                frame = frame_pipeline(fd, wl, pin_scans, min(args.cpu_seconds, 6.0))
            except Exception as e:  # the block is additive: never lose the headline line over it
                frame = {"error": repr(e)}

        h2d = n * (16 + (4 if has_i else 0) + (3 if has_c else 0))
        out = {
            "metric": "integrate_scans_per_sec", "value": value, "unit": "scans/s",
            "mpoints_per_s": value * n / 1e6,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "description": wl.description, "points_per_scan": n,
                       "arithmetic": "float32 cell state and point math, float64 grid geometry (as the reference)",
                       "map_cells": int(round(wl.map_width / wl.resolution)) * int(round(wl.map_height / wl.resolution)),
                       "parallelism": ("row-stripes x%d, scan read in place from the ingest GPU over NVLink (CUDA IPC)" % world
                                       if peer else "row-stripes x%d, scan broadcast with NCCL" % world) if sharded
                       else ("replicas x%d" % world),
                       "l2": {"ring": f"inputs larger than L2: {n_dev} distinct device-resident scans = "
                                      f"{n_dev * scan_bytes >> 20} MiB cycled (> 126 MiB L2); map state stays warm; "
                                      "steps back to back, one CUDA-event pair",
                              "flush": "256 MiB L2 flush before every timed step, one CUDA-event pair per step",
                              "none": "no flush, small ring"}[args.l2],
                       "scan_ring": n_dev,
                       "submission": (f"fdem_mapper_integrate_batch, {S} scans per call (results identical to {S} "
                                      "integrate() calls; scan k+1's front half overlaps scan k's estimator)")
                       if S > 1 else "one fdem_mapper_integrate_async per scan"},
            "cpu_enqueue_us_per_step": 1e6 * cpu_enqueue_s / args.steps,
            "value_scan_by_scan": value_scan_by_scan,
            "e2e": {"value": e2e_value, "unit": "scans/s", "mpoints_per_s": e2e_value * n / 1e6,
                    "ms_per_step": 1e3 * e2e_s / e2e_steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32 + 88,
                    "how": "fdem_mapper_submit(k+1)/collect(k) on pinned host buffers, wall clock",
                    "sync_value": e2e_steps * (1 if sharded else world) / e2e_sync_s,
                    "sync_how": "fdem_mapper_integrate() per step, one scan in flight",
                    "pointcloud2_value": pc2_value, "pointcloud2_h2d_bytes_per_step": pc2_bytes,
                    "pointcloud2_how": "fdem_mapper_submit_pointcloud2(k+1)/collect(k): the same scans as packed "
                                       "PointCloud2 bodies (x, y, z, intensity[, rgb]) parsed on the device",
                    "pinned_h2d_gbs": h2d_gbs},
            "gpu_launches": int(launches), "library_launches": int(lib_launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_val, "unit": "scans/s", "cores": 1, "kind": "port",
                             "ms_per_scan": 1e3 / cpu_val,
                             "sample": f"{cpu_n} scans of {wl.name}, oracle port (-O3), 1 of {os.cpu_count()} host cores; "
                                       "the reference path is single-threaded"},
            "last_scan": {"n_kept": int(last.n_kept), "n_cells": int(last.n_cells)},
            "frame_pipeline": frame,
        }
        emit(out)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
