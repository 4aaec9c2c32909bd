#!/usr/bin/env python
"""bench.py — scans/s and Mpoints/s of FastDEM::integrate() on B200, beside the CPU path.

    python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A "step" is one integrate() of one synthetic scan.  One JSON line on stdout (rank 0).

Workload.  N = 1: BASELINE.json's largest single-GPU configuration, configs[3] = C4
(`c4_dense_raycast`: 1.05 M points/scan, 50x50 m @ 0.05 m, Kalman, LOCAL, raycasting on) is the
top-level line; C1, C2, C3 and C5 (one GPU) follow as `configs.{name}` sub-blocks with the same
keys.  N > 1 (torchrun): the only configuration that shards, configs[4] = C5 (`c5_global`, 64 M
cells) row-striped over the N ranks — STRONG scaling (one scan stream, one map); LOCAL maps do
not shard.  The line then also carries the same-box one-GPU C5 figure
(`strong_scaling.n1_value`) so the speed-up is read off one run.

  value    scans/s with the scans already resident in HBM, ONE integrate call per scan (the
           reference's API granularity), queued back to back on the map's stream.  The timed
           region is exactly K steps between two CUDA events on that stream, bracketed by barrier
           + synchronize; the region is REPEATED until >= 0.5 s of device time has been measured
           and the MEDIAN region is reported, so the figure does not depend on K (`regions`).
           Inputs cycle through > L2 worth of distinct device buffers (`config.l2`).
           `value_batched`: the same scans through fdem_mapper_integrate_batch (16 per call).
  e2e      the same metric through the public API with HOST (pinned) buffers: submit(k+1) /
           collect(k) — every step copies that step's scan host->device and reads its stats +
           committed geometry back.  Wall clock, same repetition rule.  `sync_value`: one
           blocking integrate() per step.
  roofline the slowest pipeline stage (CUDA events around every stage, separate pass), its
           algorithmic bytes over its mean device time, against the
           MEASURED copy bandwidth (MEASURED_PEAKS.json); `traffic` = that stage's DRAM bytes per
           launch from the ncu capture of the same workload (profiles/r2_traffic.json).
  cpu_baseline  the CPU oracle (a line-by-line restatement of the reference path; the reference
           itself cannot be built here — no Eigen / nanoGrid) on 1 host core, bounded sample.

`--impl reference`: the same workload on the CPU oracle, nothing from libfastdem_b200.so.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
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

L2_BYTES = 126 << 20
RING_BYTES = 160 << 20     # distinct device-resident scan buffers cycled by the timed steps (> L2)


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
        t0 = time.perf_counter()
        while self.proc is not None and not self.rows and time.perf_counter() - t0 < timeout_s:
            time.sleep(0.02)

    def stop(self, region=None):
        """region = (t0, t1) perf_counter bounds of the timed regions."""
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
            inside = region is None or region[0] <= ts <= region[1] + 0.1
            if not inside:
                continue
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
                "samples_in_timed_regions": in_region,
                "how": "nvidia-smi -lms 100; only rows taken between the first and the last timed region are used",
                "reasons": sorted(reasons)}


# pipeline stage (fdem_mapper_stage_times) -> the kernels it runs
STAGE_KERNELS = {
    "preprocess_bin": "preprocess_bin_kernel (K1)",
    "commit_move_clear": "commit_move_clear_kernel (K2)",
    "sort_by_cell": "scatter_records_kernel (L1 of the 2-level sort)",
    "segreduce_estimate": "tile_estimate_kernel<8|9|10> (K3t: per-bucket sort + segmented reduce + estimator)",
    "voxel_raycast": "voxel_keys32 + cub radix sort + voxel_select_rays + ray_bin_scan + ray_bin_scatter + "
                     "raycast_dda + raycast_resolve (not HBM bound: per-ray DDA on an L2-resident scratch)",
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
        # raycasting (not HBM bound; the figure is its compulsory traffic only): kept points read
        # for the voxel keys, one point per traced ray, and the per-scan clear of the `raycasting`
        # layer + read of elevation + write of the ray minimum over the map (raycasting.cpp:242, 188-214)
        M = int(round(wl.map_width / wl.resolution)) * int(round(wl.map_height / wl.resolution))
        per_stage["voxel_raycast"] = Nv * 16.0 + V * 16.0 + M * 12.0
        total += per_stage["voxel_raycast"]
    return per_stage, total, C


def default_workload(world: int) -> str:
    return "c4_dense_raycast" if world == 1 else "c5_global"


# ───────────────────────────── reference arm ─────────────────────────────────────────────

def run_reference(args, wl, rank, world):
    """--impl reference: the reference's own CPU implementation of the path.  It cannot be built
    here (no Eigen, nanoGrid un-vendored), so this times the oracle port — a line-by-line
    restatement incl. the by-value cloud copy, 36-byte covariances and the unordered_map
    rasteriser — single-threaded, exactly like the reference path.  Nothing of the CUDA library
    is loaded: the config defaults come from the oracle too."""
    if rank != 0:
        return None
    import oracle_binding as ob
    from fastdem_b200 import synthetic as syn
    cfg = wl.config(ob.default_config)
    omap = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
    odem = ob.OracleFastDEM(omap, cfg)
    ring = [syn.make_scan(wl, k) for k in range(min(8, args.warmup + args.steps))]
    k = 0
    for _ in range(args.warmup):
        s = ring[k % len(ring)]
        odem.integrate(s["xyzw"], *syn.pose(wl, k), s["intensity"], s["rgb"])
        k += 1
    total, done, budget_s = 0.0, 0, 60.0
    for _ in range(args.steps):
        s = ring[k % len(ring)]
        _, _, el = odem.integrate(s["xyzw"], *syn.pose(wl, k), s["intensity"], s["rgb"])
        total += el
        k += 1
        done += 1
        if total > budget_s:
            break
    n = wl.points_per_scan
    val = done / total

    # Courtesy upper bound: the reference path is single-threaded, so one map
    # cannot use more than one core; this is what the box's cores deliver on INDEPENDENT maps.
    all_cores = None
    if wl.map_width * wl.map_height / wl.resolution ** 2 <= 4e6:   # a 64 M-cell map per core does not fit
        try:
            ncores = os.cpu_count() or 1
            counts = [0] * ncores
            stop = time.perf_counter() + 5.0

            def worker(t):
                m_ = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
                d_ = ob.OracleFastDEM(m_, cfg)
                kk = 0
                while time.perf_counter() < stop:
                    s_ = ring[kk % len(ring)]
                    d_.integrate(s_["xyzw"], *syn.pose(wl, kk), s_["intensity"], s_["rgb"])   # ctypes releases the GIL
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
    loaded = [ln.split()[-1] for ln in open("/proc/self/maps") if "libfastdem_b200" in ln]
    return {
        "impl": "reference", "metric": "integrate_scans_per_sec", "value": val, "unit": "scans/s",
        "mpoints_per_s": val * n / 1e6, "n_gpus": world, "steps": done, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / max(done, 1), "higher_is_better": True,
        "scaling": "strong" if wl.name == "c5_global" and world > 1 else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "description": wl.description, "points_per_scan": n},
        "cpu_baseline": {"value": val, "unit": "scans/s", "cores": 1, "kind": "port",
                         "sample": f"{done} scans of {wl.name} on 1 host core (the reference path is "
                                   f"single-threaded; {os.cpu_count()} cores on the box)"},
        "e2e": {"value": val, "unit": "scans/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "all_cores_independent_maps": all_cores,
        "cuda_library_mapped": bool(loaded),
    }


# ───────────────────────────── helpers of the repo arm ───────────────────────────────────

class Ctx:
    """torch / dist / device plumbing shared by the measurements."""

    def __init__(self, torch, dist, fd, dev, stream, rank, world, args):
        self.torch, self.dist, self.fd = torch, dist, fd
        self.dev, self.stream = dev, stream
        self.rank, self.world, self.args = rank, world, args
        self.distributed = world > 1

    def barrier(self):
        self.torch.cuda.synchronize(self.dev)
        if self.distributed:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, values):
        t = self.torch.tensor(list(values), dtype=self.torch.float64, device=self.dev)
        if self.distributed:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def regions_device(self, run_steps, wait, steps, min_total_s, max_regions=1500):
        """Repeat the K-step timed region (CUDA events on the kernels' stream, barrier +
        synchronize on both sides) until >= min_total_s of device time; per-region max over
        ranks; returns (median ms, all ms)."""
        torch = self.torch

        def one():
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            self.barrier()
            e0.record(self.stream)
            run_steps(steps)
            e1.record(self.stream)
            wait()
            self.barrier()
            return float(e0.elapsed_time(e1))

        pilot = self.max_over_ranks([one()])[0]
        R = int(min(max_regions, max(3, math.ceil(min_total_s * 1e3 / max(pilot, 1e-3)))))
        ms = self.max_over_ranks([one() for _ in range(R)])
        return statistics.median(ms), ms

    def regions_wall(self, run_steps, steps, min_total_s, max_regions=400):
        def one():
            self.barrier()
            t0 = time.perf_counter()
            run_steps(steps)
            self.barrier()
            return 1e3 * (time.perf_counter() - t0)

        pilot = self.max_over_ranks([one()])[0]
        R = int(min(max_regions, max(3, math.ceil(min_total_s * 1e3 / max(pilot, 1e-3)))))
        ms = self.max_over_ranks([one() for _ in range(R)])
        return statistics.median(ms), ms


def make_scans(syn, wl, base, n_host):
    return [syn.make_scan(wl, base + k) for k in range(n_host)]


def ring_size(scan_bytes, cap=512):
    return int(min(cap, max(4, -(-RING_BYTES // scan_bytes))))


def cpu_baseline(ob, syn, wl, cfg, host, base, budget_s, max_scans=2000):
    omap = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
    odem = ob.OracleFastDEM(omap, cfg)
    kk = 0
    for _ in range(2):
        s = host[kk % len(host)]
        odem.integrate(s["xyzw"], *syn.pose(wl, base + kk), s["intensity"], s["rgb"])
        kk += 1
    tot, n = 0.0, 0
    while (tot < budget_s and n < max_scans) or n < 2:
        s = host[kk % len(host)]
        _, _, el = odem.integrate(s["xyzw"], *syn.pose(wl, base + kk), s["intensity"], s["rgb"])
        tot += el
        n += 1
        kk += 1
    val = n / tot
    return {"value": val, "unit": "scans/s", "cores": 1, "kind": "port", "ms_per_scan": 1e3 / val,
            "sample": f"{n} scans of {wl.name}, oracle port (-O3), 1 of {os.cpu_count()} host cores; "
                      "the reference path is single-threaded"}


def traffic_for(wl_name, stage):
    """measured DRAM bytes per launch of the stage's kernels (ncu --set full capture of the same
    workload, summarised by tools/traffic_from_ncu.py into profiles/r2_traffic.json)"""
    try:
        d = json.loads((REPO / "profiles" / "r2_traffic.json").read_text())
        return d.get(wl_name, {}).get(stage)
    except Exception:
        return None


def frame_pipeline(fd, wl, host_scans, cpu_seconds):
    """The reference's whole per-frame chain (its published 48.3 ms/frame on a Jetson Orin =
    mapping + uncertainty fusion + raycasting + spike removal + inpainting, README.md:59 /
    assets/fastdem_jetson_benchmark.svg:1999-2098), plus feature extraction: integrate() with
    raycasting on, then the post-process functions, synchronously, host (pinned) input — on
    the GPU and on the CPU oracle (bounded sample)."""
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


# ───────────────────────────── one GPU, one map ──────────────────────────────────────────

def measure_single(cx: Ctx, wl, *, full: bool, base: int = 0, cpu_seconds: float = 12.0,
                   sampler: ClockSampler | None = None):
    """One workload on ONE GPU through fdem_mapper_*: value / e2e / roofline / cpu_baseline.
    full = the headline treatment (all e2e variants, long CPU sample); otherwise a sub-block."""
    torch, fd, dev, stream, args = cx.torch, cx.fd, cx.dev, cx.stream, cx.args
    from fastdem_b200 import synthetic as syn
    import oracle_binding as ob
    cfg = wl.config()
    steps = args.steps
    min_s = args.min_region_s if full else min(args.min_region_s, 0.25)
    n = wl.points_per_scan
    with torch.cuda.stream(stream):
        gmap = fd.ElevationMap(wl.map_width, wl.map_height, wl.resolution, "map", device=dev.index,
                               stream=stream.cuda_stream)
        dem = fd.FastDEM(gmap, cfg)
        n_host = 8 if n <= 400_000 else 4
        host = make_scans(syn, wl, base, n_host)
        has_i, has_c = host[0]["intensity"] is not None, host[0]["rgb"] is not None
        scan_bytes = n * (16 + (4 if has_i else 0) + (3 if has_c else 0))
        n_dev = ring_size(scan_bytes)
        # > L2 worth of distinct device buffers; their CONTENT cycles through the n_host distinct
        # scans (what defeats L2 is the address, not the value)
        dev_scans = []
        for j in range(n_dev):
            s = host[j % n_host]
            dev_scans.append(fd.PointCloud(torch.from_numpy(s["xyzw"]).to(dev),
                                           None if not has_i else torch.from_numpy(s["intensity"]).to(dev),
                                           None if not has_c else torch.from_numpy(s["rgb"]).to(dev)))
        pin_scans = []
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

        poses = {}

        def pose_of(kk):  # building the 4x4s with numpy costs more CPU time than enqueueing a scan
            kk = kk % 4096
            if kk not in poses:
                a, b = syn.pose(wl, base + kk)
                poses[kk] = (fd.api._iso(a), fd.api._iso(b))
            return poses[kk]

        for kk in range(4096):
            pose_of(kk)
        K = [0]   # running scan index

        def submit_dev(count):
            for _ in range(count):
                dem.integrate_async(dev_scans[K[0] % n_dev], *pose_of(K[0]))
                K[0] += 1

        S = 1 if cfg.raycasting_enabled else 16

        def submit_batched(count):
            done = 0
            while count - done >= S:
                dem.integrate_batch([dev_scans[(K[0] + j) % n_dev] for j in range(S)],
                                    [pose_of(K[0] + j) for j in range(S)], wait=False)
                K[0] += S
                done += S
            submit_dev(count - done)

        # ── warm-up (sizes every scratch buffer, lets the bucket shape settle) ──
        submit_dev(max(args.warmup, 3))
        dem.wait()
        for _ in range(3):
            submit_dev(4)
            dem.wait()

        # ── value: device-resident inputs, one integrate call per scan ──
        if sampler is not None:
            sampler.wait_first()
        t_first = time.perf_counter()
        l0, lib0, k0 = dem.launch_count(), dem.library_launch_count(), K[0]
        t_cpu0 = time.perf_counter()
        submit_dev(steps)
        cpu_enqueue_us = 1e6 * (time.perf_counter() - t_cpu0) / steps
        dem.wait()
        launches = (dem.launch_count() - l0) / (K[0] - k0) * steps
        lib_launches = (dem.library_launch_count() - lib0) / (K[0] - k0) * steps
        med_ms, all_ms = cx.regions_device(submit_dev, dem.wait, steps, min_s)
        value = steps * cx.world / (med_ms * 1e-3)
        out = {
            "value": value, "unit": "scans/s", "mpoints_per_s": value * n / 1e6,
            "ms_per_step": med_ms / steps,
            "regions": {"count": len(all_ms), "steps_per_region": steps, "median_ms": med_ms,
                        "min_ms": min(all_ms), "max_ms": max(all_ms), "total_s": sum(all_ms) * 1e-3,
                        "how": "each region = K integrate calls between two CUDA events on the map's stream, "
                               "barrier + synchronize on both sides; repeated until >= %.2f s; median region" % min_s},
            "cpu_enqueue_us_per_step": cpu_enqueue_us,
            "gpu_launches": int(round(launches)), "library_launches": int(round(lib_launches)),
        }
        if S > 1:
            for _ in range(3):   # builds the batch graph outside the timed regions
                submit_batched(S)
                dem.wait()
            bsteps = max(steps, S)
            bmed, ball = cx.regions_device(submit_batched, dem.wait, bsteps, min_s)
            out["value_batched"] = bsteps * cx.world / (bmed * 1e-3)
            out["value_batched_how"] = (f"fdem_mapper_integrate_batch, {S} scans per call (results identical to {S} "
                                        "integrate() calls; scan k+1's front half overlaps scan k's estimator) — a "
                                        "replay / throughput API, not the per-scan call a robot makes")
        t_last = time.perf_counter()

        # ── stage attribution pass (separate: bracketing every stage with events costs ~2.7 us per
        #    event and serialises the raycasting branch, which otherwise runs beside K3t) ──
        dem.set_stage_timing(True)
        submit_dev(min(max(steps, 16), 64))
        dem.wait()
        stage_ms, stage_scans = dem.stage_times()
        dem.set_stage_timing(False)
        stats_ring = []
        for _ in range(min(8, max(steps, 2))):
            stats_ring.append(dem.integrate_stats(dev_scans[K[0] % n_dev], *pose_of(K[0])))
            K[0] += 1

        # ── e2e: public API, HOST (pinned) buffers; submit(k+1) / collect(k) ──
        LOOKAHEAD = 2   # scans submitted ahead of the one being collected (the library stages 3 deep)

        def e2e_stream(count):
            pending = []
            for _ in range(count):
                pending.append(dem.submit(pin_scans[K[0] % n_host], *pose_of(K[0])))
                K[0] += 1
                if len(pending) > LOOKAHEAD:
                    dem.collect(pending.pop(0))      # every scan's result is read back, in order
            for t in pending:
                dem.collect(t)

        def e2e_sync(count):
            for _ in range(count):
                dem.integrate_stats(pin_scans[K[0] % n_host], *pose_of(K[0]))
                K[0] += 1

        e2e_stream(3)
        e_med, e_all = cx.regions_wall(e2e_stream, steps, min_s)
        e2e_value = steps * cx.world / (e_med * 1e-3)
        h2d = scan_bytes
        out["e2e"] = {"value": e2e_value, "unit": "scans/s", "mpoints_per_s": e2e_value * n / 1e6,
                      "ms_per_step": e_med / steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 32 + 88,
                      "regions": len(e_all),
                      "how": "fdem_mapper_submit(k+2)/collect(k) on pinned host buffers (two scans' copies queued "
                             "ahead, every scan's statistics read back), wall clock, median region"}
        s_med, s_all = cx.regions_wall(e2e_sync, steps, min_s)
        out["e2e"]["sync_value"] = steps * cx.world / (s_med * 1e-3)
        out["e2e"]["sync_how"] = "fdem_mapper_integrate() per step, one scan in flight"
        if full:
            # the same scans as sensor_msgs/PointCloud2 bodies (what a ROS driver hands the node:
            # x, y, z, intensity[, rgb] packed per point — 16 B instead of 20 B for a LiDAR point),
            # parsed on the device inside K1: fewer bytes over PCIe per scan
            msgs = []
            for s_ in host:
                mm = fd.PointCloud2.from_arrays(s_["xyzw"][:, :3], s_["intensity"], s_["rgb"])
                t_ = torch.from_numpy(np.ascontiguousarray(mm.data)).pin_memory()
                msgs.append((fd.PointCloud2(t_.numpy(), mm.width, mm.height, mm.point_step, mm.fields), t_))

            def e2e_pc2(count):
                pending = []
                for _ in range(count):
                    pending.append(dem.submit_pointcloud2(msgs[K[0] % n_host][0], *pose_of(K[0])))
                    K[0] += 1
                    if len(pending) > LOOKAHEAD:
                        dem.collect(pending.pop(0))
                for t in pending:
                    dem.collect(t)

            e2e_pc2(3)
            p_med, _ = cx.regions_wall(e2e_pc2, steps, min_s)
            out["e2e"]["pointcloud2_value"] = steps * cx.world / (p_med * 1e-3)
            out["e2e"]["pointcloud2_h2d_bytes_per_step"] = int(msgs[0][0].point_step) * n
            out["e2e"]["pointcloud2_how"] = ("fdem_mapper_submit_pointcloud2(k+2)/collect(k): the same scans as "
                                             "packed PointCloud2 bodies parsed on the device")
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
            out["e2e"]["pinned_h2d_gbs"] = 20 * hb.numel() * 4 / (e0.elapsed_time(e1) * 1e-3) / 1e9

    # ── roofline of the dominant stage ──
    peak, peak_src = _peaks()
    per_stage_bytes, total_bytes, C = algorithmic_bytes(wl, n, stats_ring, has_i, has_c,
                                                        cfg.estimation_type == fd.EST_P2QUANTILE)
    stage_avg = {s: (ms / max(stage_scans, 1)) for s, ms in stage_ms.items()}
    timed = {s: v for s, v in stage_avg.items() if s != "h2d"}
    dominant = max(timed, key=timed.get)
    dom_ms, dom_bytes = timed[dominant], per_stage_bytes.get(dominant, 0.0)
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    step_ms = out["ms_per_step"]
    out["roofline"] = {
        "bound": "hbm", "kernel": dominant, "kernel_names": STAGE_KERNELS.get(dominant),
        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic_for(wl.name, dominant), "peak_source": peak_src,
        "algorithmic_bytes_per_launch": dom_bytes, "kernel_ms": dom_ms, "stage_ms": stage_avg,
        "stage_ms_note": "separate pass, stages serialised on one stream, ~2.7 us event overhead inside each "
                         "stage (in the timed regions the raycasting branch runs beside scatter + K3t)",
        "stage_share": {s: v / max(sum(timed.values()), 1e-12) for s, v in timed.items()},
        "pipeline": {"algorithmic_bytes_per_scan": total_bytes,
                     "achieved": total_bytes / (step_ms * 1e-3) / 1e9,
                     "frac": total_bytes / (step_ms * 1e-3) / 1e9 / peak},
    }
    out["last_scan"] = {"n_kept": int(stats_ring[-1].n_kept), "n_cells": int(stats_ring[-1].n_cells),
                        "n_voxels": int(stats_ring[-1].n_voxels)}
    out["config"] = {
        "workload": wl.name, "description": wl.description, "points_per_scan": n,
        "arithmetic": "float32 cell state and point math, float64 grid geometry (as the reference)",
        "map_cells": int(round(wl.map_width / wl.resolution)) * int(round(wl.map_height / wl.resolution)),
        "parallelism": "one map on one GPU" if cx.world == 1 else "replicas x%d (one map per GPU)" % cx.world,
        "l2": f"inputs larger than L2: {n_dev} distinct device-resident scan buffers = "
              f"{n_dev * scan_bytes >> 20} MiB cycled (> 126 MiB L2); map state stays warm; steps back to back",
        "scan_ring": n_dev,
        "submission": "one fdem_mapper_integrate_async per scan",
    }
    out["_timed_wall"] = (t_first, t_last)
    if cx.rank == 0:
        out["cpu_baseline"] = cpu_baseline(ob, syn, wl, cfg, host, base, cpu_seconds)
        if full and not args.no_frame:
            try:
                fwl = syn.WORKLOADS["c1_vlp16_local"]
                fhost = make_scans(syn, fwl, 0, 8)
                fpin = []
                for s in fhost:
                    pc = fd.PointCloud()
                    t_ = torch.from_numpy(s["xyzw"]).pin_memory()
                    ti = torch.from_numpy(s["intensity"]).pin_memory()
                    pc.xyzw, pc.intensity, pc.color, pc._pinned = t_.numpy(), ti.numpy(), None, (t_, ti)
                    fpin.append(pc)
                out["frame_pipeline"] = dict(frame_pipeline(fd, fwl, fpin, min(cpu_seconds, 5.0)),
                                             workload="c1_vlp16_local (the configuration the reference publishes its "
                                                      "48.3 ms whole-frame figure on, Jetson Orin)")
            except Exception as e:  # the block is additive: never lose the headline line over it
                out["frame_pipeline"] = {"error": repr(e)}
    del dem, gmap
    return out


# ───────────────────────────── N GPUs, one striped map ───────────────────────────────────

def measure_sharded(cx: Ctx, wl, sampler):
    """C5 row-striped over the ranks (fastdem_b200.sharded.ShardedMapper): strong scaling."""
    torch, dist, fd, dev, stream, args = cx.torch, cx.dist, cx.fd, cx.dev, cx.stream, cx.args
    from fastdem_b200 import synthetic as syn
    from fastdem_b200.sharded import PeerScanRing, ShardedMapper
    import oracle_binding as ob
    cfg = wl.config()
    steps, n = args.steps, wl.points_per_scan
    rank, world = cx.rank, cx.world
    with torch.cuda.stream(stream):
        n_host = 4
        host = make_scans(syn, wl, 0, n_host)   # every rank generates the same scans (pure function of the index)
        has_i, has_c = host[0]["intensity"] is not None, host[0]["rgb"] is not None
        scan_bytes = n * (16 + (4 if has_i else 0) + (3 if has_c else 0))
        n_dev = ring_size(scan_bytes)
        # every rank holds the scans in its own HBM (a deployment uploads the scan to every GPU over
        # that GPU's own PCIe link, in parallel): which slice a rank bins is decided on the device
        # from the stripes' loads, so any rank may be handed most of a scan
        ring = PeerScanRing(n_dev, n, has_i, has_c, device=dev.index, src=0, replicated=True)
        for j in range(n_dev):
            s = host[j % n_host]
            ring.fill(j, s["xyzw"], s["intensity"], s["rgb"])
        cx.barrier()
        clouds = [ring.cloud(j, n) for j in range(n_dev)]
        sm = ShardedMapper(wl.map_width, wl.map_height, wl.resolution, cfg, max_points=n, device=dev.index,
                           stream=stream.cuda_stream)
        poses = {}

        def pose_of(kk):
            kk = kk % 4096
            if kk not in poses:
                a, b = syn.pose(wl, kk)
                poses[kk] = (fd.api._iso(a), fd.api._iso(b))
            return poses[kk]

        for kk in range(4096):
            pose_of(kk)
        K = [0]

        def submit(count):
            for _ in range(count):
                sm.integrate_async(clouds[K[0] % n_dev], *pose_of(K[0]))
                K[0] += 1

        if os.environ.get("FDEM_BENCH_WATCHDOG"):   # debugging aid: the handshake flags if the queue stalls
            def flags_watchdog():
                time.sleep(float(os.environ["FDEM_BENCH_WATCHDOG"]) * 0.5)
                out = (ctypes.c_uint32 * 17)()
                try:
                    fn = sm.lib.fdem_shard_debug_flags   # FDEM_PROBES=1 builds only
                except AttributeError:
                    return
                fn.restype, fn.argtypes = ctypes.c_int32, [ctypes.c_void_p, ctypes.c_void_p]
                rc = fn(sm._h, out)
                print(f"[flags rank {rank}] rc={rc} ready={list(out[0:world])} consumed={list(out[8:8 + world])} "
                      f"seq={out[16]}", file=sys.stderr, flush=True)
            threading.Thread(target=flags_watchdog, daemon=True).start()
        submit(max(args.warmup, 3))
        sm.wait()
        sampler.wait_first()
        t_first = time.perf_counter()
        l0, k0 = sm.dem.launch_count(), K[0]
        t_cpu0 = time.perf_counter()
        submit(steps)
        cpu_enqueue_us = 1e6 * (time.perf_counter() - t_cpu0) / steps
        sm.wait()
        launches = (sm.dem.launch_count() - l0) / (K[0] - k0) * steps
        med_ms, all_ms = cx.regions_device(submit, sm.wait, steps, args.min_region_s)
        t_last = time.perf_counter()
        value = steps / (med_ms * 1e-3)            # ONE scan stream over all ranks: strong scaling
        # per-scan statistics (summed over ranks) for the algorithmic bytes
        stats = []
        for _ in range(4):
            st = sm.integrate(clouds[K[0] % n_dev], *pose_of(K[0]))
            K[0] += 1
            t = torch.tensor([st.n_kept, st.n_cells], dtype=torch.float64, device=dev)
            dist.all_reduce(t)
            stats.append(type("S", (), {"n_kept": float(t[0]), "n_cells": float(t[1]), "n_voxels": 0})())

        # ── e2e: every rank uploads the host scan (pinned) into a slot of its own ring over its own
        #    PCIe link -> integrates out of that slot -> per-rank stats read back ──
        pin = None
        if True:
            pin = [dict(xyzw=torch.from_numpy(s["xyzw"]).pin_memory(),
                        intensity=None if not has_i else torch.from_numpy(s["intensity"]).pin_memory(),
                        rgb=None if not has_c else torch.from_numpy(s["rgb"]).pin_memory()) for s in host]

        copy_stream = torch.cuda.Stream(device=dev)
        uploaded = {}

        def upload(kk):
            # scan kk -> slot kk % n_dev of this rank's ring, on the copy stream (beside the kernels)
            p = pin[kk % n_host]
            with torch.cuda.stream(copy_stream):
                ring.fill(kk % n_dev, p["xyzw"], p["intensity"], p["rgb"])
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            uploaded[kk] = ev

        def e2e_steps(count):
            for _ in range(count):
                kk = K[0]
                if kk not in uploaded:
                    upload(kk)
                uploaded.pop(kk).synchronize()   # the slot is complete before this rank's front half reads it
                sm.integrate_async(clouds[kk % n_dev], *pose_of(kk))
                upload(kk + 1)                    # the next scan's upload runs beside this scan's kernels
                sm.wait()                         # this scan's statistics are on the host
                K[0] += 1

        e2e_steps(3)
        e_med, e_all = cx.regions_wall(e2e_steps, steps, args.min_region_s)
        e2e_value = steps / (e_med * 1e-3)

    peak, peak_src = _peaks()
    per_stage_bytes, total_bytes, C = algorithmic_bytes(wl, n, stats, has_i, has_c, False)
    step_ms = med_ms / steps
    out = {
        "value": value, "unit": "scans/s", "mpoints_per_s": value * n / 1e6, "ms_per_step": step_ms,
        "regions": {"count": len(all_ms), "steps_per_region": steps, "median_ms": med_ms, "min_ms": min(all_ms),
                    "max_ms": max(all_ms), "total_s": sum(all_ms) * 1e-3,
                    "how": "each region = K scans between two CUDA events on every rank's stream, barrier + "
                           "synchronize on both sides, max over ranks; repeated until >= %.2f s; median region"
                           % args.min_region_s},
        "cpu_enqueue_us_per_step": cpu_enqueue_us, "gpu_launches": int(round(launches)), "library_launches": 0,
        "e2e": {"value": e2e_value, "unit": "scans/s", "mpoints_per_s": e2e_value * n / 1e6,
                "ms_per_step": e_med / steps, "h2d_bytes_per_step": scan_bytes * world,
                "h2d_bytes_per_step_per_gpu": scan_bytes, "d2h_bytes_per_step": (32 + 88) * world,
                "regions": len(e_all),
                "how": "every rank: pinned host scan -> slot of its own ring over its own PCIe link (in parallel, "
                       "the next scan's upload beside this scan's kernels), integrates the slice the device hands "
                       "it, reads its stats back every scan; wall clock, median region, max over ranks"},
        "roofline": {
            "bound": "hbm", "kernel": "whole scan",
            "kernel_names": "per rank: shard_begin + preprocess_bin (slice) + shard_alloc + scatter_records (push) + "
                            "shard_publish_front | shard_gather + tile_estimate_light_shard + tile_estimate_shard "
                            "(its stripe)",
            "achieved": total_bytes / (step_ms * 1e-3) / 1e9, "peak": peak * world, "unit": "GB/s",
            "frac": total_bytes / (step_ms * 1e-3) / 1e9 / (peak * world), "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": total_bytes, "kernel_ms": step_ms,
            "note": "algorithmic bytes of one scan over the whole job's step time, against N x the measured "
                    "copy bandwidth; the path is latency bound at this scan size (DESIGN.md)"},
        "last_scan": {"n_kept": int(stats[-1].n_kept), "n_cells": int(stats[-1].n_cells)},
        "config": {
            "workload": wl.name, "description": wl.description, "points_per_scan": n,
            "arithmetic": "float32 cell state and point math, float64 grid geometry (as the reference)",
            "map_cells": int(round(wl.map_width / wl.resolution)) * int(round(wl.map_height / wl.resolution)),
            "parallelism": f"row-stripes x{world}: every rank bins a slice of the scan sized on the device from "
                           f"the stripes' loads (busy owners bin less), pushes each record into its owner's arena "
                           f"(NVLink stores, CUDA IPC), owners estimate out of local memory; device-side "
                           f"ready/consumed flags; no collective on the data path",
            "l2": f"inputs larger than L2: {n_dev} distinct scan buffers = {n_dev * scan_bytes >> 20} MiB in "
                  "every rank's HBM, cycled",
            "scan_ring": n_dev, "submission": "one fdem_shard_integrate per scan per rank"},
        "_timed_wall": (t_first, t_last),
    }
    sm.close()
    ring.close()
    del sm
    cx.barrier()
    # ── same box, ONE GPU, same workload: the denominator of the strong-scaling figure ──
    n1 = torch.zeros(1, dtype=torch.float64, device=dev)
    if rank == 0:
        solo = Ctx(torch, dist, fd, dev, stream, 0, 1, args)
        solo.distributed = False
        r1 = measure_single(solo, wl, full=False, cpu_seconds=min(args.cpu_seconds, 6.0))
        n1[0] = r1["value"]
        out["cpu_baseline"] = r1["cpu_baseline"]
        out["strong_scaling"] = {"n1_value": r1["value"], "n1_ms_per_step": r1["ms_per_step"],
                                 "n1_value_batched": r1.get("value_batched"),
                                 "speedup": value / r1["value"], "efficiency": value / r1["value"] / world,
                                 "what": "same box, same run: c5_global on ONE GPU (rank 0, unsharded map, one "
                                         "integrate call per scan) vs the striped map on all ranks"}
    dist.all_reduce(n1)
    return out


# ───────────────────────────── main ──────────────────────────────────────────────────────

def main():
    # stdout carries exactly ONE JSON line: libraries (NCCL's version banner, torchrun notices)
    # write to fd 1 as well, so park fd 1 on stderr until the result is ready
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(real_stdout, (json.dumps(obj) + "\n").encode())

    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None,
                    help="default: c4_dense_raycast on one GPU, c5_global (row stripes) under torchrun")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU baseline sample budget (headline)")
    ap.add_argument("--min-region-s", type=float, default=0.5,
                    help="the K-step timed region is repeated until this much time has been measured")
    ap.add_argument("--no-subconfigs", action="store_true", help="skip the configs.{c1,c2,c3,c5} sub-blocks")
    ap.add_argument("--no-frame", action="store_true", help="skip the whole-frame (mapping + post-process) block")
    args = ap.parse_args()

    if os.environ.get("FDEM_BENCH_WATCHDOG"):   # debugging aid: dump every thread's stack if the run stalls
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ["FDEM_BENCH_WATCHDOG"]), repeat=True, file=sys.stderr)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    from fastdem_b200 import synthetic as syn
    wl = syn.WORKLOADS[args.workload or default_workload(world)]

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
    # host thread (and with it the pinned buffers it first-touches) onto the GPU's NUMA node
    ncpu = ctypes.c_int32(0)
    numa_bound = fd.load_library().fdem_bind_thread_to_device(local_rank, ctypes.byref(ncpu)) == 0
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    cx = Ctx(torch, dist, fd, dev, stream, rank, world, args)
    sampler = ClockSampler(local_rank)
    sampler.start()

    sharded = world > 1 and wl.name == "c5_global"
    if sharded:
        res = measure_sharded(cx, wl, sampler)
    else:
        res = measure_single(cx, wl, full=True, base=1000 * rank, cpu_seconds=args.cpu_seconds, sampler=sampler)
    clocks = sampler.stop(res.pop("_timed_wall"))

    configs = None
    if world == 1 and not args.no_subconfigs and args.workload is None:
        configs = {}
        for name in ("c1_vlp16_local", "c2_lidar64_local", "c3_rgbd_p2", "c5_global"):
            try:
                r = measure_single(cx, syn.WORKLOADS[name], full=False, cpu_seconds=2.5)
                r.pop("_timed_wall", None)
                configs[name] = r
            except Exception as e:   # additive: never lose the headline line over a sub-block
                configs[name] = {"error": repr(e)}

    if rank == 0:
        out = {
            "metric": "integrate_scans_per_sec", "value": res["value"], "unit": "scans/s",
            "mpoints_per_s": res["mpoints_per_s"], "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True,
            "scaling": "strong" if sharded else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        }
        for k in ("config", "regions", "cpu_enqueue_us_per_step", "value_batched", "value_batched_how", "e2e",
                  "gpu_launches", "library_launches", "roofline", "cpu_baseline", "strong_scaling", "last_scan",
                  "frame_pipeline"):
            if k in res:
                out[k] = res[k]
        out["clocks"] = clocks
        out["host"] = {"cpus": os.cpu_count(), "numa_bound_to_gpu_node": bool(numa_bound), "cpus_in_affinity": int(ncpu.value)}
        if configs is not None:
            out["configs"] = configs
        emit(out)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
