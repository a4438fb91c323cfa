"""Row sharding across the GPUs of one box (one process per GPU, torch.distributed).

Pixels are independent (windspeed.py:190-281 has no cross-pixel state), so a scene is split into contiguous row
blocks, each rank inverts its block with no communication, and NCCL is used only to gather the results
(the reference's analogue is the dask row-block fan-out of windspeed.py:356-364).

The gather is device-resident and moves every result byte exactly once: the destination rank inverts its own rows
straight into its slice of the full result and posts receives into the peers' row slices; the peers send their rows
from where the kernel wrote them (ncclSend / ncclRecv through `batch_isend_irecv`).  No padding, no staging through
the host, no replication to ranks that did not ask for the result.  A rank's rows are inverted in a few sub-blocks and
each sub-block's results travel on a side stream while the next one is inverted, so only the last transfer is exposed.  The same code runs on the gloo backend
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


_COMM_STREAMS = {}


def _comm_stream(device):
    """One side stream per device for the gather, so that it overlaps the inversion of the following rows."""
    import torch

    key = str(device)
    if key not in _COMM_STREAMS:
        _COMM_STREAMS[key] = torch.cuda.Stream(device=device)
    return _COMM_STREAMS[key]


PIECE_RATIO = 0.6   # every sub-block is this fraction of the one before it


def piece_edges(lo: int, hi: int, pieces: int):
    """Row edges of the `pieces` contiguous sub-blocks of [lo, hi).  The sub-blocks shrink geometrically (PIECE_RATIO): only
    the transfer of the last one is exposed, so the last one is small, while the first ones stay large enough for the
    scan to work at its full rate.  Every sub-block of a non-empty range with at least `pieces` rows is non-empty."""
    n = hi - lo
    if pieces <= 1 or n <= pieces:
        return [lo + min(k, n) for k in range(pieces)] + [hi]
    w = np.array([PIECE_RATIO ** k for k in range(pieces)])
    cuts = np.floor(np.cumsum(w) / w.sum() * n + 0.5).astype(int)
    cuts = np.maximum(cuts, np.arange(1, pieces + 1))          # at least one row each
    cuts = np.minimum(cuts, n - (pieces - 1 - np.arange(pieces)))
    return [lo] + [lo + int(c) for c in cuts[:-1]] + [hi]


def n_pieces(n_lines: int, width: int, world: int, target_px: int = 16 << 20) -> int:
    """Sub-blocks per rank: as many as keep ~16 Mpx on average (enough pixels per incidence bin for the scan's fast mode
    and its pruning), at most 4; the same number on every rank."""
    per_rank = (n_lines // max(world, 1)) * max(width, 1)
    return int(max(1, min(4, per_rank // target_px)))


def invert_rows_resident(plan, blk, n_lines: int, lo: int, hi: int, *, dst=0, dsig_cr=0.1, merge_dual=False, cr_abs=False,
                         group=None, events=None, pieces=None, _invert=None):
    """This rank's rows [lo, hi) of one scene, device-resident: blk = (inc, sigma0_co, sigma0_cr, ancillary) tensors of
    those rows (None for an absent raster).  Inverts them with `plan` and gathers both results on the device into rank
    `dst`: returns (wind_co, wind_cr) of the whole scene there and (None, None) elsewhere; dst=None keeps every rank's
    own block.

    The block is inverted in `pieces` row sub-blocks (default `n_pieces`): the results of sub-block j travel on a side
    stream while sub-block j + 1 is inverted, so only the last sub-block's transfer is exposed.
    `events` = two CUDA events recorded on the current stream after the last inversion and after the gather (bench.py)."""
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
    if pieces is None:
        pieces = n_pieces(n_lines, int(np.prod(width)) if width else 1, world) if gather else 1
    if gather and rank == dst:
        full_co = torch.empty((n_lines,) + width, dtype=cdt, device=dev)
        full_cr = torch.empty((n_lines,) + width, dtype=cr_dt, device=dev)
        o_co, o_cr = full_co[lo:hi], full_cr[lo:hi]   # own rows are written in place
    else:
        full_co = full_cr = None
        o_co = torch.empty((hi - lo,) + width, dtype=cdt, device=dev)
        o_cr = torch.empty((hi - lo,) + width, dtype=cr_dt, device=dev)
    on_gpu = inc.is_cuda
    cur = torch.cuda.current_stream(dev) if on_gpu else None
    comm = _comm_stream(dev) if (on_gpu and gather and pieces > 1) else None
    if comm is not None:
        comm.wait_stream(cur)   # the output buffers were allocated on `cur`
    cont = lambda t, a, b: None if t is None else t[a - lo:b - lo].contiguous()
    mine = piece_edges(lo, hi, pieces)
    for k in range(pieces):
        a, b = mine[k], mine[k + 1]
        if b > a:
            run(cont(inc, a, b), cont(s_co, a, b), cont(s_cr, a, b), dsig_cr if np.isscalar(dsig_cr) else cont(dsig_cr, a, b),
                cont(anc, a, b), o_co[a - lo:b - lo], o_cr[a - lo:b - lo])
        if events is not None and k == pieces - 1:
            events[0].record()
        if not gather:
            continue
        # sub-block k of every rank -> rank dst (every rank computes the same edges)
        ops = []
        for r in range(world):
            if r == dst:
                continue
            rlo, rhi = row_shard(n_lines, world, r)
            e = piece_edges(rlo, rhi, pieces)
            if e[k + 1] <= e[k]:
                continue
            if rank == dst:
                ops.append(dist.P2POp(dist.irecv, _as_real(full_co[e[k]:e[k + 1]]), r, group))
                ops.append(dist.P2POp(dist.irecv, _as_real(full_cr[e[k]:e[k + 1]]), r, group))
            elif rank == r:
                ops.append(dist.P2POp(dist.isend, _as_real(o_co[a - lo:b - lo]), dst, group))
                ops.append(dist.P2POp(dist.isend, _as_real(o_cr[a - lo:b - lo]), dst, group))
        if not ops:
            continue
        if comm is None:
            for w in dist.batch_isend_irecv(ops):
                w.wait()
        else:
            ev = torch.cuda.Event()
            ev.record(cur)
            with torch.cuda.stream(comm):
                comm.wait_event(ev)   # the transfer starts when this sub-block's results exist; `cur` goes on inverting
                for w in dist.batch_isend_irecv(ops):
                    w.wait()
    if comm is not None:
        cur.wait_stream(comm)
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
