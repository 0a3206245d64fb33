"""Case sharding across the GPUs of one box (one process per GPU, torch.distributed).

The path has no data-path collective: cases are independent
(data_extract_vitaldb.py:194-199 maps them over a process pool), so every rank runs
the whole pipeline on its own shard.  Two collectives exist and both are tiny:
the variable-length all-gather of the final ``np_waves`` / ``np_info`` rows, and the
all-reduce(sum) of the per-candidate error sums of the sigma search (beats sharded).
Works with ``nccl`` (GPU tensors) and ``gloo`` (CPU tensors, used by the CPU tests).
"""
from __future__ import annotations

import numpy as np


def shard_cases(lengths, world_size: int, rank: int | None = None):
    """Greedy longest-processing-time assignment of cases to ranks by sample count.
    Returns the list of case indices per rank (or the one for ``rank``)."""
    lengths = np.asarray(lengths, dtype=np.int64)
    order = np.argsort(-lengths, kind="stable")
    load = np.zeros(world_size, dtype=np.int64)
    shards = [[] for _ in range(world_size)]
    for i in order:
        r = int(np.argmin(load))
        shards[r].append(int(i))
        load[r] += lengths[i]
    shards = [sorted(s) for s in shards]
    return shards if rank is None else shards[rank]


def allgather_rows(t, group=None):
    """Variable-length all-gather along dim 0: gathers the row counts first, pads to the
    maximum, gathers, and strips the padding.  Returns the list of per-rank tensors."""
    import torch
    import torch.distributed as dist
    ws = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    m = max(counts) if counts else 0
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad, group=group)
    return [b[:c] for b, c in zip(bufs, counts)]


def gather_rows_to(arrays, dst: int = 0, group=None, async_op: bool = False):
    """Variable-length gather of several row arrays to ONE rank: the row counts are all-gathered once, then every
    other rank sends its rows as they are (no padding) and ``dst`` receives them into exact-size buffers -- a
    quarter of the all-gather's bytes per link at eight ranks, and nothing lands where it is not needed.
    Returns, on ``dst``, one tensor per array with the ranks' rows in rank order; None elsewhere.  With ``async_op``
    the transfers are only started: returns ``(parts, requests)``; the caller calls ``wait()`` on the requests before
    it touches ``parts`` (or frees ``arrays``) -- the rows of one step then cross NVLink under the next step's kernels."""
    import torch
    import torch.distributed as dist
    ws, rank = dist.get_world_size(group), dist.get_rank(group)
    glob = (lambda r: dist.get_global_rank(group, r)) if group is not None else (lambda r: r)
    n = torch.tensor([arrays[0].shape[0]], dtype=torch.int64, device=arrays[0].device)
    counts = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c) for c in torch.cat(counts).cpu()]
    ops, parts = [], None
    if rank == dst:
        # every array is received straight into its slice of ONE buffer (no concatenation pass afterwards)
        first = [sum(counts[:r]) for r in range(ws)]
        parts = []
        for a in arrays:
            full = torch.empty((sum(counts),) + tuple(a.shape[1:]), dtype=a.dtype, device=a.device)
            full[first[rank]:first[rank] + counts[rank]].copy_(a)
            ops += [dist.P2POp(dist.irecv, full[first[r]:first[r] + counts[r]], glob(r), group)
                    for r in range(ws) if r != rank and counts[r]]
            parts.append(full)
    elif counts[rank]:
        ops = [dist.P2POp(dist.isend, a.contiguous(), glob(dst), group) for a in arrays]
    reqs = dist.batch_isend_irecv(ops) if ops else []
    if async_op:
        return parts, reqs
    for req in reqs:
        req.wait()
    return parts


def sort_case_major(waves, info, feats=None):
    """Rows of several ranks -> the case-major order of data_create_arrays.py:27-48 (sort on (case_id, idx_end))."""
    import torch
    if info.shape[0]:
        key = info[:, 0].to(torch.int64) * (1 << 40) + info[:, 1].to(torch.int64)
        order = torch.argsort(key, stable=True)
        waves, info = waves.index_select(0, order), info.index_select(0, order)
        if feats is not None:
            feats = feats.index_select(0, order)
    return (waves, info) if feats is None else (waves, info, feats)


def gather_beat_arrays(np_waves, np_info, group=None, np_features=None, dst: int | None = None):
    """All-gather the per-rank ``np_waves`` (f32) and ``np_info`` (f64) -- and ``np_features`` (f64, n x 35)
    when given -- and restore the case-major order of data_create_arrays.py:27-48 by sorting on
    (case_id, idx_end).  Returns (waves, info) or (waves, info, features).  With ``dst`` the rows are gathered to
    that rank only (point-to-point, unpadded) and the other ranks get None."""
    import torch
    if dst is not None:                    # the arrays are wanted on one rank only (the one that writes them out)
        arrays = [np_waves, np_info] + ([] if np_features is None else [np_features])
        parts = gather_rows_to(arrays, dst, group)
        if parts is None:
            return (None, None) if np_features is None else (None, None, None)
        waves, info = parts[0], parts[1]
        feats = None if np_features is None else parts[2]
    else:
        waves = torch.cat(allgather_rows(np_waves, group), dim=0)
        info = torch.cat(allgather_rows(np_info, group), dim=0)
        feats = None if np_features is None else torch.cat(allgather_rows(np_features, group), dim=0)
    return sort_case_major(waves, info, feats)


def rows_to_host(*arrays, dst_rank: int = 0, group=None):
    """The final arrays of data_create_arrays.py:56-58 live on ONE host: after the all-gather only ``dst_rank``
    moves them off its device, into reusable page-locked buffers (a pageable ``.cpu()`` of N x 190 MB on every rank
    was the slowest part of the multi-GPU end-to-end step).  Returns host tensors on ``dst_rank`` (views of the
    pool, valid until the next call), None elsewhere."""
    import torch
    import torch.distributed as dist
    from . import batched
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    if rank != dst_rank:
        return tuple(None for _ in arrays)
    out = []
    for k, a in enumerate(arrays):
        shape = tuple(a.shape) if a.dim() > 1 else (a.shape[0], 1)
        h = batched._pinned(f"gathered{k}", (max(shape[0], 1),) + shape[1:], a.dtype, torch)[:shape[0]]
        h.copy_(a.reshape(shape), non_blocking=True)
        out.append(h if a.dim() > 1 else h[:, 0])
    torch.cuda.current_stream(arrays[0].device).synchronize() if arrays else None
    return tuple(out)


def allreduce_sse(sse, group=None):
    """Sum the per-candidate error sums of all ranks (beats sharded by rank)."""
    import torch.distributed as dist
    dist.all_reduce(sse, op=dist.ReduceOp.SUM, group=group)
    return sse
