"""Multi-GPU plumbing: frames are independent, so a job shards them over ranks with no data-path collective;
only the tally vector (qkdldpc.h "Tally vector layout") is summed, once per batch.

The reference parallelises the same way over CPU threads (BS::thread_pool::detach_loop, simulation.cpp:740-746):
trial n always uses seeds[n] + combination index and writes slot n, so any partition gives the same tallies
(quirk Q16). One process per GPU; `torch.distributed` (NCCL over NVLink on the GPU box, gloo in CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_range(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous trial-index range [lo, hi) of this rank -- the static block partition detach_loop uses."""
    return n_frames * rank // world, n_frames * (rank + 1) // world


def allreduce_tally(tally, device=None) -> np.ndarray:
    """Sum of the per-rank tally vectors over the default process group (NCCL: the tensor must live on the GPU).
    Accepts a numpy uint64 vector or a torch int64 tensor; returns numpy uint64."""
    import torch
    import torch.distributed as dist
    if isinstance(tally, np.ndarray):
        t = torch.from_numpy(tally.astype(np.int64))
        if device is not None:
            t = t.to(device)
    else:
        t = tally
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy().astype(np.uint64)


def tally_from_results(iterations_num, flags, max_iterations: int, executed=None) -> np.ndarray:
    """Host-side construction of the tally vector from per-frame results (what the scheduler kernel accumulates
    on the device): used to cross-check the device tallies and in CPU tests."""
    it = np.asarray(iterations_num, np.int64)
    fl = np.asarray(flags, np.uint8)
    t = np.zeros(max_iterations + 5, np.uint64)
    ok = (fl & 1) != 0
    t[0] = it.size
    t[1] = int(ok.sum())
    t[2] = int((ok & ((fl & 2) != 0)).sum())
    t[3] = int(np.asarray(executed if executed is not None else it, np.int64).sum())
    t[4:] = np.bincount(it[ok], minlength=max_iterations + 1)[: max_iterations + 1].astype(np.uint64)
    return t
