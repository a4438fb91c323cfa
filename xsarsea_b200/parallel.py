"""Row sharding across the GPUs of one box (one process per GPU, torch.distributed).

Pixels are independent (windspeed.py:190-281 has no cross-pixel state), so a scene is split into contiguous row
blocks, each rank inverts its block with no communication, and NCCL is used only to gather the results
(the reference's analogue is the dask row-block fan-out of windspeed.py:356-364).

The gather is device-resident and moves every result byte exactly once: the destination rank inverts its own rows
straight into its slice of the full result and posts one receive per peer into that peer's row slice; the peers send
their block from where the kernel wrote it (ncclSend / ncclRecv through `batch_isend_irecv`).  No padding, no staging
through the host, no replication to ranks that did not ask for the result.  The same code runs on the gloo backend
with CPU tensors (the CPU tests use it with a stub compute function).
"""
from __future__ import annotations

import numpy as np


def row_shard(n_lines: int, world: int, rank: int):
    """Contiguous block [lo, hi) of `n_lines` rows owned by `rank`; the remainder goes to the first ranks."""
    base, rem = divmod(int(n_lines), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _as_real(t):
    import torch

    return torch.view_as_real(t) if t.is_complex() else t


def gather_rows(local, full, n_lines: int, dst: int, group=None):
    """Assemble per-rank row blocks in `full` ([n_lines, ...], only used on rank `dst`) from every rank's `local`
    ([hi - lo, ...]).  `dst`'s own block must already sit in full[lo:hi] (pass local=None there) -- it is not copied."""
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ops = []
    if rank == dst:
        for r in range(world):
            if r == dst:
                continue
            lo, hi = row_shard(n_lines, world, r)
            if hi > lo:
                ops.append(dist.P2POp(dist.irecv, _as_real(full[lo:hi]), r, group))
    else:
        if local is not None and local.shape[0] > 0:
            ops.append(dist.P2POp(dist.isend, _as_real(local), dst, group))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def invert_rows_resident(plan, blk, n_lines: int, lo: int, hi: int, *, dst=0, dsig_cr=0.1, merge_dual=False, cr_abs=False,
                         group=None, events=None, _invert=None):
    """This rank's rows [lo, hi) of one scene, device-resident: blk = (inc, sigma0_co, sigma0_cr, ancillary) tensors of
    those rows (None for an absent raster).  Inverts them with `plan` and gathers both results on the device into rank
    `dst`: returns (wind_co, wind_cr) of the whole scene there and (None, None) elsewhere; dst=None keeps every rank's
    own block.  `events` = two CUDA events recorded after the inversion and after the gather (bench.py)."""
    import torch
    import torch.distributed as dist

    world, rank = (dist.get_world_size(group), dist.get_rank(group)) if dist.is_initialized() else (1, 0)
    inc, s_co, s_cr, anc = blk
    width = tuple(inc.shape[1:])
    dev, cdt = inc.device, torch.complex128
    cr_dt = torch.float64 if cr_abs else cdt
    run = _invert if _invert is not None else (lambda i, a, b, d, c, oc, ox: plan.invert(
        i, a, b, d, c, merge_dual=merge_dual, cr_abs=cr_abs, out_co=oc, out_cr=ox))
    gather = dst is not None and world > 1
    if gather and rank == dst:
        full_co = torch.empty((n_lines,) + width, dtype=cdt, device=dev)
        full_cr = torch.empty((n_lines,) + width, dtype=cr_dt, device=dev)
        o_co, o_cr = full_co[lo:hi], full_cr[lo:hi]   # own rows are written in place
    else:
        o_co = torch.empty((hi - lo,) + width, dtype=cdt, device=dev)
        o_cr = torch.empty((hi - lo,) + width, dtype=cr_dt, device=dev)
    if hi > lo:
        cont = lambda t: None if t is None else t.contiguous()
        run(cont(inc), cont(s_co), cont(s_cr), dsig_cr if np.isscalar(dsig_cr) else cont(dsig_cr), cont(anc), o_co, o_cr)
    if events is not None:
        events[0].record()
    if gather:
        mine = rank == dst
        gather_rows(None if mine else o_co, full_co if mine else None, n_lines, dst, group)
        gather_rows(None if mine else o_cr, full_cr if mine else None, n_lines, dst, group)
    if events is not None:
        events[1].record()
    if not gather:
        return o_co, o_cr
    return (full_co, full_cr) if rank == dst else (None, None)


def invert_sharded(inc, sigma0, sigma0_dual=None, /, *, gather=0, group=None, _invert=None, **kwargs):
    """`invert_from_model` on a scene every rank holds in full (host arrays, first axis = line): rank r uploads and
    inverts rows row_shard(...), the results are gathered on the device (gather = destination rank: only that rank
    downloads and returns the full result, the others return None; "all": the destination's result is broadcast and
    every rank returns it; None: each rank returns its own block).  Keyword arguments as `invert_from_model`; array-
    valued `ancillary_wind` / `dsig_cr` are sliced like the rasters.  `_invert` (tests): a stand-in for
    `invert_from_model` working on the backend's tensors."""
    import torch
    import torch.distributed as dist

    from .windspeed import windspeed as impl

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    n_lines = np.shape(inc)[0]
    lo, hi = row_shard(n_lines, world, rank)
    on_gpu = dist.get_backend(group) == "nccl"
    invert = _invert if _invert is not None else impl.invert_from_model   # tensors in -> tensors out (device-resident path)

    def rows(a, sl):
        """Rows `sl` of a per-line array as a tensor of the backend's device; scalars and None pass through."""
        if a is None or np.isscalar(a) or np.ndim(a) == 0 or np.shape(a)[0] != n_lines:
            return a
        t = torch.as_tensor(np.ascontiguousarray(np.asarray(a)[sl]))
        return t.cuda() if on_gpu else t

    def call(sl):
        kw = {k: (rows(v, sl) if k in ("ancillary_wind", "dsig_cr") else v) for k, v in kwargs.items()}
        args = [rows(inc, sl), rows(sigma0, sl)] + ([rows(sigma0_dual, sl)] if sigma0_dual is not None else [])
        return invert(*args, **kw)

    if hi > lo:
        res = call(slice(lo, hi))
    else:  # a rank without rows still takes part in the gather: learn the result structure from one line
        probe = call(slice(0, 1))
        res = tuple(r[:0] for r in probe) if isinstance(probe, tuple) else probe[:0]
    parts = res if isinstance(res, tuple) else (res,)
    if gather is None:
        out = tuple(p.cpu().numpy() for p in parts)
        return out if isinstance(res, tuple) else out[0]
    dst = 0 if gather == "all" else int(gather)
    fulls = []
    for p in parts:
        full = None
        if rank == dst:
            full = torch.empty((n_lines,) + tuple(p.shape[1:]), dtype=p.dtype, device=p.device)
            full[lo:hi].copy_(p)
        gather_rows(None if rank == dst else p, full, n_lines, dst, group)
        if gather == "all":
            if rank != dst:
                full = torch.empty((n_lines,) + tuple(p.shape[1:]), dtype=p.dtype, device=p.device)
            dist.broadcast(_as_real(full), src=dst, group=group)
        fulls.append(full)
    if fulls[0] is None:
        return None
    out = tuple(f.cpu().numpy() for f in fulls)
    return out if isinstance(res, tuple) else out[0]
