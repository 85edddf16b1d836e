"""Sharding of NGPTOT over ranks and the two reductions the dwarf needs.

Columns are independent (doc/documentation.tex:1169-1171), so the only "collectives" are the ones the
reference's MPI mode performs after the kernel: the step time is the MAX over ranks
(src/common/module/timer_mod.F90:183-186) and the validation norms combine as MIN / MAX / SUM
(src/common/module/validate_mod.F90:197-199).  There is no exchange on the data path.

Weak scaling (bench.py): rank r owns global columns [r*n, (r+1)*n) -- the 100-column source pattern
continues across ranks (first_column = r*n in state.blocked_fields / load_state_set_first_column).
Strong scaling (C driver, dwarf_cloudsc.F90:74-77): nblocks // world blocks per rank, the last rank
takes the remainder.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def weak_shard(rank: int, world: int, ngptot_per_rank: int) -> Tuple[int, int]:
    """-> (first global column, number of columns) of `rank`."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return rank * ngptot_per_rank, ngptot_per_rank


def strong_shard(rank: int, world: int, ngptot: int, nproma: int) -> Tuple[int, int, int]:
    """-> (first block, number of blocks, number of valid columns) of `rank`; block-contiguous."""
    nblocks = ngptot // nproma + min(ngptot % nproma, 1)
    world = min(world, nblocks)
    if rank >= world:
        return nblocks, 0, 0
    per = nblocks // world
    b0 = rank * per
    nb = nblocks - b0 if rank == world - 1 else per
    ncols = min((b0 + nb) * nproma, ngptot) - b0 * nproma
    return b0, nb, ncols


def reduce_max(value: float, dist=None, device=None) -> float:
    """MAX over ranks of a host scalar (the step time); identity without a process group."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value: float, dist=None, device=None) -> float:
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def reduce_field_errors(errs: Sequence, ngptot_global: int, dist=None, device=None) -> List:
    """Combine per-rank validate.FieldError lists into the global table: min/max by MIN/MAX, the sums
    by SUM, derived columns recomputed (validate_mod.F90:197-199 does the same three all-reduces)."""
    from .validate import EPS, FieldError
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return list(errs)
    import torch
    mins = torch.tensor([e.minval for e in errs], dtype=torch.float64, device=device)
    maxs = torch.tensor([[e.maxval, e.maxerr, e.refmax] for e in errs], dtype=torch.float64, device=device)
    sums = torch.tensor([[e.errsum, e.refsum_true, e.refsum_c] for e in errs], dtype=torch.float64, device=device)
    dist.all_reduce(mins, op=dist.ReduceOp.MIN)
    dist.all_reduce(maxs, op=dist.ReduceOp.MAX)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    out = []
    for i, e in enumerate(errs):
        errsum, refsum_true, refsum_c = (float(x) for x in sums[i])
        if errsum < EPS:
            rel, iopt = 0.0, 1
        elif refsum_c < EPS:
            rel, iopt = errsum / (1.0 + refsum_c), 2
        else:
            rel, iopt = errsum / refsum_c, 3
        rel_l1 = 0.0 if errsum == 0.0 else (errsum / refsum_true if refsum_true > 0 else float("inf"))
        out.append(FieldError(e.name, float(mins[i]), float(maxs[i, 0]), float(maxs[i, 1]), errsum, refsum_true, refsum_c,
                              float(maxs[i, 2]), errsum / ngptot_global, 100.0 * rel, iopt, rel_l1, rel > 10.0 * EPS))
    return out
