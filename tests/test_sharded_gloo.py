"""CPU, world_size 2, gloo: the collective plumbing of the row-stripe sharded global map
(scan broadcast, layer gather, isEmpty all-reduce, halo-row exchange for the inpainting
stencil) with an oracle-backed local engine.  The CUDA stripe engine itself is covered by
tests/test_gpu_stripes.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_binding as ob
from fastdem_b200 import capi, sharded
from fastdem_b200 import synthetic as syn


def test_stripe_bounds_partition():
    for rows in (1, 7, 20, 8000, 8001):
        for world in (1, 2, 3, 4, 8):
            if world > rows:
                continue
            spans = [sharded.stripe_bounds(rows, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == rows
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    assert sharded.grid_rows(400.0, 0.05) == 8000 and sharded.grid_rows(10.0, 0.5) == 20


class OracleStripe:
    """Local engine for the CPU tests: the full map on the oracle, exposing only this rank's rows."""

    def __init__(self, width, height, resolution, cfg, r0, r1):
        self.map = ob.OracleMap(width, height, resolution)
        self.dem = ob.OracleFastDEM(self.map, cfg)
        self.r0, self.r1 = r0, r1

    def integrate(self, xyzw, intensity, rgb, Tbs, Twb):
        ok, st, _ = self.dem.integrate(xyzw, Tbs, Twb, intensity, rgb)
        return st

    def get(self, name):
        return self.map.get(name)[self.r0:self.r1]

    def exists(self, name):
        return self.map.exists(name)

    def is_empty(self):
        return bool(np.isnan(self.get("elevation")).all())

    # stencil interface of a stripe engine (the CUDA engine runs fdem_inpaint_stripe_sweep)
    def inpaint_begin(self):
        self._inp = np.asarray(self.get("elevation"), np.float32).copy()

    def border_rows(self, name):
        return torch.from_numpy(self._inp[0].copy()), torch.from_numpy(self._inp[-1].copy())

    def inpaint_sweep(self, name, above, below, min_valid):
        cur = self._inp
        R, cols = cur.shape
        padded = np.full((R + 2, cols + 2), np.nan, np.float32)
        padded[1:-1, 1:-1] = cur
        if above is not None:
            padded[0, 1:-1] = above.numpy()
        if below is not None:
            padded[-1, 1:-1] = below.numpy()
        total = np.zeros((R, cols), np.float32)
        count = np.zeros((R, cols), np.int32)
        for dr in (-1, 0, 1):          # the oracle's neighbour order: dr outer, dc inner
            for dc in (-1, 0, 1):
                if dr == 0 and dc == 0:
                    continue
                nb = padded[1 + dr: 1 + dr + R, 1 + dc: 1 + dc + cols]
                ok = np.isfinite(nb)
                total = np.where(ok, total + np.where(ok, nb, np.float32(0)), total).astype(np.float32)
                count += ok
        fill = np.isnan(cur) & (count >= min_valid)
        nxt = cur.copy()
        nxt[fill] = (total[fill] / count[fill].astype(np.float32)).astype(np.float32)
        self._inp = nxt

    def result(self, name):
        return self._inp


def _global_cfg():
    wl = syn.WORKLOADS["tiny"]
    cfg = wl.config()
    cfg.mode = capi.MODE_GLOBAL
    return wl, cfg


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wl, cfg = _global_cfg()
        sm = sharded.ShardedGlobalMap(wl.map_width, wl.map_height, wl.resolution, cfg,
                                      engine_factory=OracleStripe)
        assert (sm.r0, sm.r1) == sharded.stripe_bounds(100, world, rank)
        assert sm.is_empty()
        for k in range(4):
            if rank == 0:   # the ingest rank has the scan; the others receive it
                s = syn.make_scan(wl, k)
                st = sm.integrate(s["xyzw"], s["intensity"], s["rgb"], s["T_base_sensor"], s["T_world_base"], src=0)
            else:
                st = sm.integrate(src=0)
            assert st.integrated == 1
        assert not sm.is_empty()
        full = {name: sm.gather(name, dst=0) for name in ("elevation", "n_points", "variance", "intensity")}
        inp = sm.inpaint(3, 2)
        parts = [None] * world if rank == 0 else None
        dist.gather_object(inp, parts, dst=0)
        if rank == 0:
            np.savez(os.path.join(out_dir, "result.npz"), inpainted=np.concatenate(parts, axis=0), **full)
        else:
            assert all(v is None for v in full.values())
    finally:
        dist.barrier()
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_sharded_map_two_ranks_gloo(tmp_path):
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "result.npz")
    # single-process truth
    wl, cfg = _global_cfg()
    m = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
    d = ob.OracleFastDEM(m, cfg)
    for k in range(4):
        s = syn.make_scan(wl, k)
        d.integrate(s["xyzw"], s["T_base_sensor"], s["T_world_base"], s["intensity"], s["rgb"])
    for name in ("elevation", "n_points", "variance", "intensity"):
        want = m.get(name)
        assert np.array_equal(np.isnan(got[name]), np.isnan(want)), name
        assert np.array_equal(np.nan_to_num(got[name]), np.nan_to_num(want)), name
    m.inpaint(3, 2, False)
    want = m.get("elevation_inpainted")
    assert np.array_equal(np.isnan(got["inpainted"]), np.isnan(want))
    assert np.allclose(np.nan_to_num(got["inpainted"]), np.nan_to_num(want), rtol=1e-6, atol=0)
    assert np.isfinite(want).sum() > np.isfinite(m.get("elevation")).sum()  # holes were filled


# ── the load-aware slice split of the multi-GPU front half (fdem_shard_slice_plan: the same
#    integer arithmetic shard_begin_kernel runs on the device) ──

def _check_tiling(plan, n):
    pos = 0
    for b, c in plan:
        assert b == pos, plan
        pos += c
    assert pos == n, plan
    for b, c in plan[:-1]:
        assert (b + c) % 32 == 0 or b + c == n, plan   # warp-aligned boundaries


def test_slice_plan_equal_split_without_loads():
    from fastdem_b200.sharded import slice_plan
    n = 1048576
    for world in (1, 2, 3, 4, 8):
        plan = slice_plan([0] * world, n)
        _check_tiling(plan, n)
        counts = [c for _, c in plan]
        assert max(counts) - min(counts) <= 64 + n % world, plan


def test_slice_plan_busy_owner_bins_less():
    from fastdem_b200.sharded import slice_plan
    n = 1048576
    # one stripe owns every touched cell: with a back half as dear as a front half it bins nothing
    plan = slice_plan([102000, 0], n, back_weight_q8=256)
    _check_tiling(plan, n)
    assert plan[0][1] == 0 and plan[1][1] == n
    # back half = half a front half: work levels at 0.75 each -> the owner bins a quarter
    plan = slice_plan([102000, 0], n, back_weight_q8=128)
    _check_tiling(plan, n)
    assert abs(plan[0][1] / n - 0.25) < 1e-3
    # eight ranks, three owners: owners bin least, the idle ranks share the rest equally
    loads = [0, 0, 40000, 45000, 17000, 0, 0, 0]
    plan = slice_plan(loads, n, back_weight_q8=256)
    _check_tiling(plan, n)
    counts = [c for _, c in plan]
    idle = [counts[r] for r in range(8) if loads[r] == 0]
    assert max(idle) - min(idle) <= 64 + 32 * 8          # rounding to warp boundaries only
    assert counts[3] <= counts[2] <= counts[4] <= min(idle)   # more cells, fewer points
    # level: points share + cells share is the same for every rank that bins anything
    tot = sum(loads)
    work = [counts[r] / n + loads[r] / tot for r in range(8) if counts[r] > 0]
    assert max(work) - min(work) < 2e-3


def test_slice_plan_is_deterministic_and_total_for_ragged_sizes():
    from fastdem_b200.sharded import slice_plan
    import random
    rng = random.Random(7)
    for _ in range(200):
        world = rng.randint(1, 8)
        n = rng.choice([1, 31, 32, 33, 1000, 28800, 131072, 1048576, 4000001])
        loads = [rng.choice([0, 0, rng.randint(1, 200000)]) for _ in range(world)]
        w = rng.choice([0, 1, 64, 128, 256, 1024, 4096])
        a = slice_plan(loads, n, w)
        assert a == slice_plan(loads, n, w)
        _check_tiling(a, n)
