"""Locus sharding over the GPUs of one box (one process per GPU).

Loci are independent (one RNG, one matrix, one output each; assembler.cpp:1441-1445, 1514, 1548-1603), so the only
cross-rank step is agreeing on who runs what: every rank computes the cost of the loci it generated/parsed
(bayes_locus_cost), the costs are all-gathered, and the same cost-balanced greedy partition (bayes_partition_lpt, the
multi-GPU form of the reference's largest-first work queue, assembler.cpp:1886-1916) is evaluated identically on every
rank.  No collective touches the data path; results are gathered on the host at the end.
"""
import ctypes as C

import numpy as np

from . import LocusDesc, load_library, partition_lpt


def desc_costs(descs):
    """bayes_locus_cost of every entry of a bayes_locus_desc[] (numpy structured array or ctypes array)."""
    lib = load_library()
    n = len(descs)
    out = np.zeros(n)
    if isinstance(descs, np.ndarray):
        base, step = descs.ctypes.data, descs.dtype.itemsize
        for i in range(n):
            out[i] = lib.bayes_locus_cost(C.cast(base + i * step, C.POINTER(LocusDesc)))
    else:
        for i in range(n):
            out[i] = lib.bayes_locus_cost(C.byref(descs[i]))
    return out


def gather_costs(cost_local, dist=None, device=None):
    """All-gather equally sized per-rank cost vectors (torch.distributed, nccl or gloo) -> concatenated numpy array."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(cost_local, np.float64)
    import torch
    t = torch.from_numpy(np.ascontiguousarray(cost_local, np.float64))
    if device is not None:
        t = t.to(device)
    parts = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, t)
    return torch.cat(parts).cpu().numpy()


def shard_loci(cost_all, world, rank):
    """-> (global indices owned by `rank`, per-rank total cost).  Deterministic: identical on every rank."""
    part = partition_lpt(cost_all, world)
    mine = np.nonzero(part == rank)[0]
    load = np.bincount(part, weights=cost_all, minlength=world)
    return mine, load
