"""-m gpu: the compute-scaling multi-GPU path (fdem_shard_*, fastdem_b200.sharded.ShardedMapper).

Every rank bins its 1/world slice of the scan for every stripe, owners pull their buckets'
record pieces from all ranks' arenas, device-side flags order the halves.  Here: `world`
PROCESSES on the one GPU of the test box (CUDA IPC mappings instead of NVLink peer mappings —
the same code path); on a multi-GPU box the same worker runs one rank per GPU
(tools/shard_parity.py under `gpurun --gpus N`).  The stripes must concatenate to the oracle's
single map: every layer, the same layer set, the same cell counts."""
import os
import socket

import numpy as np
import pytest

import oracle_binding as ob
from fastdem_b200 import capi, sharded
from fastdem_b200 import synthetic as syn

pytestmark = pytest.mark.gpu


def run_rank(rank, world, port, out_dir, wl_name, n_scans, one_gpu=True, backend="gloo"):
    """One rank of the sharded mapper: integrates n_scans scans of workload `wl_name` and saves
    its stripe of every layer."""
    import torch
    import torch.distributed as dist
    import fastdem_b200 as fd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = 0 if one_gpu else rank
    torch.cuda.set_device(dev)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        wl = syn.WORKLOADS[wl_name]
        cfg = wl.config()
        cfg.mode = capi.MODE_GLOBAL
        n = wl.points_per_scan
        has_i, has_c = wl.has_intensity, wl.has_color
        ring = sharded.PeerScanRing(3, n, has_i, has_c, device=dev, src=0)
        sm = sharded.ShardedMapper(wl.map_width, wl.map_height, wl.resolution, cfg, max_points=n, device=dev)
        cells = kept = 0
        for k in range(n_scans):
            s = syn.make_scan(wl, k)
            if rank == 0:
                ring.fill(k, s["xyzw"], s["intensity"], s["rgb"])
                torch.cuda.synchronize()
            dist.barrier()          # the slot is complete before any rank reads it
            if k % 2 == 0:
                st = sm.integrate(ring.cloud(k, n), s["T_base_sensor"], s["T_world_base"])
            else:                   # queued: the next scan's front half follows without a host sync
                sm.integrate_async(ring.cloud(k, n), s["T_base_sensor"], s["T_world_base"])
                st = sm.wait()
            cells += st.n_cells
            kept += st.n_kept
            dist.barrier()          # every rank is done with the slot before it is reused
        out = {name: sm.map.get(name) for name in sm.map.getLayers()}
        out["__cells"] = np.array([cells])
        out["__kept"] = np.array([kept])
        np.savez(os.path.join(out_dir, f"stripe{rank}.npz"), **out)
        sm.close()
        ring.close()
    finally:
        dist.barrier()
        dist.destroy_process_group()


def check_against_oracle(out_dir, world, wl_name, n_scans):
    parts = [np.load(os.path.join(out_dir, f"stripe{r}.npz")) for r in range(world)]
    wl = syn.WORKLOADS[wl_name]
    cfg = wl.config()
    cfg.mode = capi.MODE_GLOBAL
    om = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
    od = ob.OracleFastDEM(om, cfg)
    cells = kept = 0
    for k in range(n_scans):
        s = syn.make_scan(wl, k)
        _, st, _ = od.integrate(s["xyzw"], s["T_base_sensor"], s["T_world_base"], s["intensity"], s["rgb"])
        cells += st.n_cells
        kept += st.n_kept
    assert sum(int(p["__cells"][0]) for p in parts) == cells
    assert sum(int(p["__kept"][0]) for p in parts) == kept
    names = sorted(k for k in parts[0].files if not k.startswith("__"))
    for p in parts:
        assert sorted(k for k in p.files if not k.startswith("__")) == names
    assert names == sorted(om.layers()), (names, sorted(om.layers()))
    bits_differ = 0
    for name in names:
        got = np.concatenate([p[name] for p in parts], axis=0)
        want = om.get(name)
        assert got.shape == want.shape, name
        assert np.array_equal(np.isnan(got), np.isnan(want)), name
        ok = ~np.isnan(want)
        if name == "color":
            assert np.array_equal(got[ok].view(np.uint32), want[ok].view(np.uint32)), name
        else:
            assert np.allclose(got[ok], want[ok], rtol=1e-5, atol=1e-7), name
        bits_differ += int(np.count_nonzero(got[ok].view(np.uint32) != want[ok].view(np.uint32)))
    return bits_differ


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_mapper_matches_the_oracle(fdem, tmp_path, world):
    import torch.multiprocessing as mp
    mp.spawn(run_rank, args=(world, _free_port(), str(tmp_path), "tiny", 6), nprocs=world, join=True)
    check_against_oracle(str(tmp_path), world, "tiny", 6)


def test_sharded_mapper_one_rank_per_gpu(fdem, tmp_path):
    """The real thing where the box has it: one rank per GPU, NCCL for the handle exchange, records
    pushed over NVLink (tools/shard_parity.py is the same run as a script; on 2 x B200 it reports
    0 bit-different cells for tiny, C2-sized and C5 scans).  Skipped on a one-GPU box — the
    one-GPU worlds above run the same kernels and flags."""
    import torch
    n_gpus = torch.cuda.device_count()
    if n_gpus < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(run_rank, args=(world, _free_port(), str(tmp_path), "tiny", 6, False, "nccl"),
             nprocs=world, join=True)
    assert check_against_oracle(str(tmp_path), world, "tiny", 6) == 0   # bit-identical cells


def test_sharded_mapper_world_1_equals_plain_mapper(fdem, tmp_path):
    """Degenerate case in one process: the two-half pipeline against the one-GPU pipeline."""
    wl = syn.WORKLOADS["c1_vlp16_local"]
    cfg = wl.config()
    cfg.mode = capi.MODE_GLOBAL
    sm = sharded.ShardedMapper(wl.map_width, wl.map_height, wl.resolution, cfg, max_points=wl.points_per_scan)
    m = fdem.ElevationMap(wl.map_width, wl.map_height, wl.resolution, "map")
    d = fdem.FastDEM(m, cfg)
    for k in range(5):
        s = syn.make_scan(wl, k)
        import torch
        cloud = fdem.PointCloud(torch.from_numpy(s["xyzw"]).cuda(), torch.from_numpy(s["intensity"]).cuda())
        a = sm.integrate(cloud, s["T_base_sensor"], s["T_world_base"])
        b = d.integrate_stats(cloud, s["T_base_sensor"], s["T_world_base"])
        assert (a.n_kept, a.n_cells) == (b.n_kept, b.n_cells)
    assert sm.map.getLayers() == m.getLayers()
    for name in m.getLayers():
        x, y = sm.map.get(name), m.get(name)
        assert np.array_equal(x.view(np.uint32), y.view(np.uint32)), name


def _inpaint_worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        wl = syn.WORKLOADS["tiny"]
        cfg = wl.config()
        cfg.mode = capi.MODE_GLOBAL
        sm = sharded.ShardedGlobalMap(wl.map_width, wl.map_height, wl.resolution, cfg, device=0)
        for k in range(4):
            s = syn.make_scan(wl, k)
            if rank == 0:
                sm.integrate(s["xyzw"], s["intensity"], s["rgb"], s["T_base_sensor"], s["T_world_base"], src=0)
            else:
                sm.integrate(src=0)
        inp = sm.inpaint(3, 2)          # CUDA stencil on the stripe, halo rows exchanged between sweeps
        np.save(os.path.join(out_dir, f"inp{rank}.npy"), inp)
    finally:
        dist.barrier()
        dist.destroy_process_group()


def test_sharded_inpainting_runs_the_cuda_stencil_with_halo_rows(fdem, tmp_path):
    """applyInpainting over row stripes: fdem_inpaint_stripe_sweep per stripe, one halo row to /
    from each neighbour between sweeps (device tensors under NCCL; here two processes on one GPU
    with a gloo group).  The stripes must concatenate to the oracle's whole-map result."""
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_inpaint_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = np.concatenate([np.load(tmp_path / f"inp{r}.npy") for r in range(world)], axis=0)
    wl = syn.WORKLOADS["tiny"]
    cfg = wl.config()
    cfg.mode = capi.MODE_GLOBAL
    m = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
    d = ob.OracleFastDEM(m, cfg)
    for k in range(4):
        s = syn.make_scan(wl, k)
        d.integrate(s["xyzw"], s["T_base_sensor"], s["T_world_base"], s["intensity"], s["rgb"])
    m.inpaint(3, 2, False)
    want = m.get("elevation_inpainted")
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.allclose(np.nan_to_num(got), np.nan_to_num(want), rtol=1e-6, atol=0)
    assert np.isfinite(want).sum() > np.isfinite(m.get("elevation")).sum()
