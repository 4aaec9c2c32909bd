"""Row-stripe sharding of ONE global map over the ranks of a torch.distributed job
(SURVEY.md §8e): rank g owns logical rows [g*R/G, (g+1)*R/G) of every layer.  integrate()
needs no collective on the data path — after binning, a cell's update depends only on that
cell's state and the scan's points in it — so every rank receives the full scan (broadcast
from the ingest rank) and keeps only the keys whose row falls in its stripe."""
from __future__ import annotations


def stripe_bounds(rows: int, world: int, rank: int):
    """Contiguous, balanced row stripes: the first rows % world ranks get one extra row."""
    base, extra = divmod(rows, world)
    r0 = rank * base + min(rank, extra)
    r1 = r0 + base + (1 if rank < extra else 0)
    return r0, r1
