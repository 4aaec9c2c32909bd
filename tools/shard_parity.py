"""Real multi-GPU parity of the sharded mapper: one rank per GPU (NVLink peer reads), stripes
concatenated == the CPU oracle's single map.

    gpurun --gpus 2 -- python tools/shard_parity.py 2 [workload] [n_scans]"""
import os
import sys
import tempfile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))


def main():
    import torch.multiprocessing as mp
    import test_gpu_shard as t
    world = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    wl = sys.argv[2] if len(sys.argv) > 2 else "c5_global"
    n_scans = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    out = tempfile.mkdtemp()
    mp.spawn(t.run_rank, args=(world, t._free_port(), out, wl, n_scans, False, "nccl"), nprocs=world, join=True)
    diff = t.check_against_oracle(out, world, wl, n_scans)
    print(f"shard parity OK: world={world} workload={wl} scans={n_scans} bit-different cells={diff}")


if __name__ == "__main__":
    main()
