"""-m gpu: the scan ring that row stripes read in place from the ingest rank's HBM (CUDA IPC,
fastdem_b200.sharded.PeerScanRing).  Two PROCESSES on the one GPU of the test box: the ingest
process fills ring slots, the other maps them and its kernels read them through the IPC mapping
(on a multi-GPU node the same code path goes over NVLink); each process integrates its row
stripe and the stripes must concatenate to the oracle's map, layer by layer."""
import os
import socket

import numpy as np
import pytest

import oracle_binding as ob
from fastdem_b200 import capi, sharded
from fastdem_b200 import synthetic as syn

pytestmark = pytest.mark.gpu
LAYERS = ("elevation", "n_points", "variance", "intensity", "elevation_max", "obstacle")


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    import fastdem_b200 as fd
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        wl = syn.WORKLOADS["tiny"]
        cfg = wl.config()
        cfg.mode = capi.MODE_GLOBAL
        n = wl.points_per_scan
        ring = sharded.PeerScanRing(3, n, True, False, device=0, src=0)
        rows = sharded.grid_rows(wl.map_width, wl.resolution)
        r0, r1 = sharded.stripe_bounds(rows, world, rank)
        m = fd.ElevationMap(wl.map_width, wl.map_height, wl.resolution, "map", device=0, row_stripe=(r0, r1))
        dem = fd.FastDEM(m, cfg)
        cells = 0
        for k in range(6):
            s = syn.make_scan(wl, k)
            if rank == 0:
                ring.fill(k, s["xyzw"], s["intensity"], None)
                torch.cuda.synchronize()
            dist.barrier()          # the slot is complete before any stripe reads it
            st = dem.integrate_stats(ring.cloud(k, n), s["T_base_sensor"], s["T_world_base"])
            assert st.integrated == 1
            cells += st.n_cells
            dist.barrier()          # every stripe is done with the slot before it is reused
        out = {name: m.get(name) for name in LAYERS}
        out["cells"] = np.array([cells])
        np.savez(os.path.join(out_dir, f"stripe{rank}.npz"), **out)
        ring.close()
    finally:
        dist.barrier()
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_stripes_read_the_scan_through_cuda_ipc(fdem, tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    parts = [np.load(tmp_path / f"stripe{r}.npz") for r in range(2)]
    wl = syn.WORKLOADS["tiny"]
    cfg = wl.config()
    cfg.mode = capi.MODE_GLOBAL
    om = ob.OracleMap(wl.map_width, wl.map_height, wl.resolution)
    od = ob.OracleFastDEM(om, cfg)
    cells = 0
    for k in range(6):
        s = syn.make_scan(wl, k)
        _, st, _ = od.integrate(s["xyzw"], s["T_base_sensor"], s["T_world_base"], s["intensity"], None)
        cells += st.n_cells
    assert int(parts[0]["cells"][0] + parts[1]["cells"][0]) == cells
    for name in LAYERS:
        got = np.concatenate([p[name] for p in parts], axis=0)
        want = om.get(name)
        assert np.array_equal(np.isnan(got), np.isnan(want)), name
        assert np.allclose(np.nan_to_num(got), np.nan_to_num(want), rtol=1e-5, atol=1e-7), name


def test_device_alloc_export_roundtrip(fdem):
    """Same-process sanity of the plumbing: alloc, export a handle, free; bad device is an error."""
    import ctypes as C
    lib = capi.load_library()
    p = C.c_void_p()
    capi.check(lib.fdem_device_alloc(0, 4096, C.byref(p)))
    h = capi.FdemIpcHandle()
    capi.check(lib.fdem_ipc_export(0, p, 4096, C.byref(h)))
    assert h.size == 4096 and any(h.bytes)
    capi.check(lib.fdem_device_free(0, p))
    assert lib.fdem_device_alloc(99, 4096, C.byref(p)) != 0
