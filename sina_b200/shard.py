"""Query sharding across ranks (one process per GPU, `torch.distributed`): the index is replicated, every rank
aligns a contiguous block of the queries, results are gathered on rank 0 in input order -- the multi-process
equivalent of the reference's sequencer_node (src/sina.cpp:529-538). There is no collective on the data path
itself; the gather moves finished results only. Backend: nccl on GPUs, gloo in the CPU tests."""
import numpy as np


def shard_bounds(n, world, rank):
    """contiguous block [lo, hi) of n queries owned by `rank` (sizes differ by at most one)"""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_queries(qmasks, qoff, world, rank):
    """(masks, offsets rebased to 0, lo, hi) of this rank's block"""
    nq = len(qoff) - 1
    lo, hi = shard_bounds(nq, world, rank)
    a, b = int(qoff[lo]), int(qoff[hi])
    return np.ascontiguousarray(qmasks[a:b]), (qoff[lo:hi + 1] - qoff[lo]).astype(np.uint64), lo, hi


def max_over_ranks(dist, seconds, device="cpu"):
    """a timing is the max over ranks (the slowest GPU defines the step)"""
    import torch
    t = torch.tensor([float(seconds)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


def gather_ordered(dist, out_cols, out_masks, results, qoff_local, dst=0):
    """Gather per-rank outputs on `dst` in rank (= input) order. Returns (cols, masks, results, qoff) on dst,
    None elsewhere. Shapes differ per rank, so objects are gathered (host memory; results are small next to the
    DP that produced them)."""
    payload = (np.asarray(out_cols), np.asarray(out_masks), np.asarray(results), np.asarray(qoff_local, np.uint64))
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return payload
    world, rank = dist.get_world_size(), dist.get_rank()
    bucket = [None] * world if rank == dst else None
    dist.gather_object(payload, bucket, dst=dst)
    if rank != dst:
        return None
    cols = np.concatenate([b[0][:int(b[3][-1])] for b in bucket])
    masks = np.concatenate([b[1][:int(b[3][-1])] for b in bucket])
    res = np.concatenate([b[2] for b in bucket])
    offs, base = [np.zeros(1, np.uint64)], 0
    for b in bucket:
        offs.append(b[3][1:] + np.uint64(base))
        base += int(b[3][-1])
    return cols, masks, res, np.concatenate(offs)
