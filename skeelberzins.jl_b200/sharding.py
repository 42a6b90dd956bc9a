"""Multi-GPU execution: the batch is sharded by instance index, one process per GPU.

The instances of a batch are independent PDE problems, so there is no collective on the data path
(SURVEY.md 8(e)): rank r solves the contiguous block of instances ``shard_range(B, world, r)`` on its
own GPU with its own Newton loop (no lock-step across GPUs), and only the results are gathered to
rank 0 at the end (one transfer per rank, ``torch.distributed.gather``).  ``solve_sharded`` is
backend-agnostic (NCCL on the GPU box, gloo in the CPU tests).
"""
import os

import numpy as np


def shard_range(B, world_size, rank):
    """Contiguous block [lo, hi) of instances of rank `rank`: ceil(B/world) per rank, the last ranks may get fewer."""
    per = -(-int(B) // int(world_size))
    lo = min(B, rank * per)
    return lo, min(B, lo + per)


def shard_indices(B, world_size, rank, mode="strided"):
    """Instance indices of rank `rank`.  "strided": rank, rank + world, rank + 2*world, ... -- every rank gets the same
    mix of a sweep, so ranks finish together (a monotone parameter sweep cut into contiguous blocks is not balanced:
    cfg2 on 8 GPUs, slowest rank 13.0 ms against a median of 12.3).  "contiguous": the block of shard_range."""
    if mode == "contiguous":
        lo, hi = shard_range(B, world_size, rank)
        return np.arange(lo, hi)
    return np.arange(int(rank), int(B), int(world_size))


def env_rank():
    """(rank, world_size, local_rank) from the torchrun environment (defaults: single process)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def solve_sharded(solver, params, *, dist=None, device=None, gather=True):
    """Solve a batch whose per-instance parameters are ``params[B, nparams]`` across the ranks of `dist`.

    solver(shard_params, lo, hi) -> np.ndarray [hi-lo, ...]  (this rank's instances, e.g. a closure over pdepe)
    Returns on rank 0 the results of all instances in instance order ([B, ...]); None on the other ranks
    (gather=False: every rank returns just its own block).  `dist` is torch.distributed (initialised) or None.
    """
    params = np.ascontiguousarray(params, dtype=np.float64)
    B = params.shape[0]
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(solver(params, 0, B))
    import torch
    rank, world = dist.get_rank(), dist.get_world_size()
    lo, hi = shard_range(B, world, rank)
    local = np.ascontiguousarray(solver(params[lo:hi], lo, hi), dtype=np.float64)
    if not gather:
        return local
    per = -(-B // world)
    tail = local.shape[1:]
    dev = device if device is not None else ("cuda" if dist.get_backend() == "nccl" else "cpu")
    buf = torch.zeros((per,) + tail, dtype=torch.float64, device=dev)   # equal-sized blocks for gather
    if hi > lo:
        buf[:hi - lo] = torch.from_numpy(local).to(dev)
    if rank == 0:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf, gather_list=parts, dst=0)
        out = np.empty((B,) + tail)
        for r, part in enumerate(parts):
            a, b = shard_range(B, world, r)
            out[a:b] = part[:b - a].cpu().numpy()
        return out
    dist.gather(buf, gather_list=None, dst=0)
    return None
