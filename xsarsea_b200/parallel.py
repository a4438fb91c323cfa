"""Row sharding across the GPUs of one box (one process per GPU, torch.distributed).

Pixels are independent (windspeed.py:190-281 has no cross-pixel state), so a scene is split into contiguous row
blocks, each rank inverts its block with no communication, and NCCL is used only to gather the results
(SURVEY.md section 8 row E1).  Works with the gloo backend on CPU tensors too (used by the CPU tests with a stub
compute function).
"""
from __future__ import annotations

import numpy as np


def row_shard(n_lines: int, world: int, rank: int):
    """Contiguous block [lo, hi) of `n_lines` rows owned by `rank`; the remainder goes to the first ranks."""
    base, rem = divmod(int(n_lines), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _gather_rows(local: np.ndarray, n_lines: int, dst=None, group=None):
    """all_gather (dst=None) or gather to rank `dst` of per-rank row blocks -> full [n_lines, ...] array."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    backend = dist.get_backend(group)
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    is_cplx = np.iscomplexobj(local)
    t = torch.from_numpy(np.ascontiguousarray(local))
    if is_cplx:
        t = torch.view_as_real(t)
    rows_max = max(row_shard(n_lines, world, r)[1] - row_shard(n_lines, world, r)[0] for r in range(world))
    pad = torch.zeros((rows_max,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
    pad[: t.shape[0]].copy_(t.to(device))
    if dst is None:
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
    else:
        parts = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
        dist.gather(pad, parts, dst=dst, group=group)
        if rank != dst:
            return None
    out = []
    for r, p in enumerate(parts):
        lo, hi = row_shard(n_lines, world, r)
        out.append(p[: hi - lo].cpu())
    full = torch.cat(out, dim=0)
    if is_cplx:
        full = torch.view_as_complex(full.contiguous())
    return full.numpy()


def invert_sharded(inc, sigma0, sigma0_dual=None, /, *, gather="all", group=None, _invert=None, **kwargs):
    """`invert_from_model` on a scene every rank holds in full (host arrays, first axis = line): rank r inverts rows
    row_shard(...) and the results are gathered (gather="all": every rank gets the full result; an int: only that
    rank does, the others return None; None: each rank keeps its own block).  Keyword arguments as
    `invert_from_model`; array-valued `ancillary_wind` / `dsig_cr` are sliced like the rasters."""
    import torch.distributed as dist

    if _invert is None:
        from .windspeed import invert_from_model as _invert
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_lines = np.shape(inc)[0]
    lo, hi = row_shard(n_lines, world, rank)
    cut = lambda a: a[lo:hi] if (a is not None and np.ndim(a) >= 1 and np.shape(a)[0] == n_lines) else a
    kw = dict(kwargs)
    for k in ("ancillary_wind", "dsig_cr"):
        if k in kw:
            kw[k] = cut(kw[k])
    res = _invert(cut(inc), cut(sigma0), cut(sigma0_dual), **kw) if sigma0_dual is not None else _invert(
        cut(inc), cut(sigma0), **kw)
    if gather is None:
        return res
    dst = None if gather == "all" else int(gather)
    if isinstance(res, tuple):
        parts = tuple(_gather_rows(np.asarray(r), n_lines, dst, group) for r in res)
        return None if parts[0] is None else parts
    return _gather_rows(np.asarray(res), n_lines, dst, group)
