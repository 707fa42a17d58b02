"""Several GPUs, one process each: the exchange steps around the sharded C ABI.

Reads are sharded by read id (each rank is pushed its reads).  The histogram is sharded by slice
(2^22 consecutive padded bins): rank r of n owns slices [S*r/n, S*(r+1)/n) - ``owned_slices``.  After
``slimm_gpu_coverage`` a rank holds its items grouped by slice; ``exchange_items`` routes every item to
the rank that owns its slice with ONE all-to-all (4 bytes per record); the histogram itself never
crosses NVLink.  ``run_sharded`` strings the stages together (include/slimm_gpu.h, "several GPUs").

torch.distributed is plumbing here (NCCL over NVLink on the GPUs; gloo on CPU tensors in the tests).
The reference has no multi-process mode: this layer has no counterpart there.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def owned_slices(n_slices: int, rank: int, n_ranks: int) -> Tuple[int, int]:
    """Slices [lo, hi) of rank ``rank`` - the same arithmetic as owned_slices() in csrc/slimm_gpu.cu."""
    return n_slices * rank // n_ranks, n_slices * (rank + 1) // n_ranks


def send_splits(slice_counts: Sequence[int], n_ranks: int) -> List[int]:
    """Items this rank sends to every rank: the grouped buffer is ordered by slice, so the share of rank q is
    the contiguous block of the slices q owns."""
    counts = np.asarray(slice_counts, dtype=np.int64)
    out = []
    for q in range(n_ranks):
        lo, hi = owned_slices(counts.size, q, n_ranks)
        out.append(int(counts[lo:hi].sum()))
    return out


def exchange_items(items, slice_counts: Sequence[int], group=None):
    """All-to-all of the slice-grouped items (1-D integer tensor, ``sum(slice_counts)`` long).  Returns the
    tensor of items this rank owns (in source-rank order) and the per-source counts."""
    import torch
    import torch.distributed as dist
    n_ranks = dist.get_world_size(group)
    splits_out = send_splits(slice_counts, n_ranks)
    if int(items.numel()) != sum(splits_out):
        raise ValueError("items length does not match the slice counts")
    # gloo has no all_to_all for CPU tensors in every build: exchange the counts with all_gather
    mine = torch.tensor(splits_out, dtype=torch.int64, device=items.device)
    table = [torch.empty_like(mine) for _ in range(n_ranks)]
    dist.all_gather(table, mine, group=group)
    rank = dist.get_rank(group)
    splits_in = [int(v) for v in torch.stack(table)[:, rank].tolist()]   # one device->host copy
    recv = torch.empty(sum(splits_in), dtype=items.dtype, device=items.device)
    if items.is_cuda:
        dist.all_to_all_single(recv, items, output_split_sizes=splits_in, input_split_sizes=splits_out, group=group)
    else:   # CPU tests (gloo): point-to-point, same result
        offs_out = np.concatenate([[0], np.cumsum(splits_out)])
        offs_in = np.concatenate([[0], np.cumsum(splits_in)])
        reqs = []
        for q in range(n_ranks):
            if q == rank:
                recv[offs_in[q]:offs_in[q + 1]] = items[offs_out[q]:offs_out[q + 1]]
                continue
            reqs.append(dist.isend(items[offs_out[q]:offs_out[q + 1]].contiguous(), q, group=group))
            reqs.append(dist.irecv(recv[offs_in[q]:offs_in[q + 1]], q, group=group))
        for r in reqs:
            r.wait()
    return recv, splits_in


_tokens = {}
_tables = {}


def _barrier_token(device):
    import torch
    if device not in _tokens:
        _tokens[device] = torch.zeros(1, dtype=torch.int32, device=device)
    return _tokens[device]


def connect_peers(gpu, device, cap_items: int, group=None) -> bool:
    """Once per context, after ``set_shard``: every rank reserves a receive buffer of ``cap_items`` items and maps all the
    others' (CUDA IPC; one NVLink/NVSwitch box, one process per GPU).  Returns False - and leaves the NCCL all-to-all path
    in place - when peer mapping is not available."""
    import torch
    import torch.distributed as dist
    n_ranks = dist.get_world_size(group)
    ok, handle = 1, b"\0" * 64
    try:
        handle = gpu.p2p_reserve(cap_items)
    except Exception:
        ok = 0
    mine = torch.tensor(list(handle) + [ok], dtype=torch.uint8, device=device)
    table = torch.empty((n_ranks, 65), dtype=torch.uint8, device=device)
    dist.all_gather_into_tensor(table, mine, group=group)
    table = table.cpu().numpy()
    if not table[:, 64].all():
        return False
    try:
        gpu.p2p_connect(table[:, :64].tobytes(), n_ranks)
        ok = 1
    except Exception:
        ok = 0
    flag = torch.tensor([ok], dtype=torch.int32, device=device)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    if not int(flag.item()):
        if getattr(gpu, "p2p", False):
            gpu.p2p_disable()       # this rank mapped its peers but another one could not: everybody takes the all-to-all
        return False
    return True


def run_sharded(gpu, device, cov_cut_off: float, min_reads: int, global_hits: int, group=None, phase_ms=None):
    """coverage -> items all-to-all -> accumulate owned bins -> sum statistics -> filter -> assign -> sum
    assign block.  ``gpu`` is a :class:`slimm_b200.api.SlimmGpu` with ``set_shard`` done and this rank's
    records pushed; afterwards ``gpu.summary()`` / ``gpu.profile()`` give the global results on every rank.
    ``phase_ms``: optional dict that receives CUDA-event durations of the exchange steps (bench.py)."""
    import torch
    import torch.distributed as dist
    from . import api
    marks = []

    def mark(name):
        if phase_ms is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record(torch.cuda.current_stream(device))
            marks.append((name, e))

    mark("start")
    gpu.coverage()
    if getattr(gpu, "p2p", False):
        # the split writes every item straight into its owner's receive buffer (peer memory over NVLink).  Nothing here waits
        # for the GPU: the slice counts are all-gathered on the device, a small kernel turns the table into destinations, and
        # coverage, collectives and split queue up on the stream
        n_ranks = dist.get_world_size(group)
        p, ns = gpu.slice_counts_device()
        mine = api.device_tensor(p, ns, torch.int32, device)
        key = (id(gpu), n_ranks, ns)
        table = _tables.get(key)
        if table is None:
            table = _tables[key] = torch.empty((n_ranks, ns), dtype=torch.int32, device=device)
        dist.all_gather_into_tensor(table, mine, group=group)
        mark("coverage")
        gpu.split_to_peers_device(table.data_ptr())
        dist.all_reduce(_barrier_token(device), group=group)       # every rank's stores have landed before anyone accumulates
        mark("split into peers")
        gpu.accumulate_received()
    else:
        counts = gpu.slice_counts()
        n_items = int(counts.sum(dtype=np.int64))
        items = api.device_tensor(gpu.items_device(), max(n_items, 1), torch.int32, device)[:n_items]
        mark("coverage+split")
        recv, _ = exchange_items(items, counts, group)
        mark("items all-to-all")
        gpu._keep.append(recv)                       # the library reads it asynchronously on its stream
        gpu.accumulate_items(recv.data_ptr() if recv.numel() else 0, int(recv.numel()))
    mark("accumulate+stats")
    p, n = gpu.stats_device()
    dist.all_reduce(api.device_tensor(p, n, torch.int32, device), group=group)
    p, n = gpu.counters_device()
    dist.all_reduce(api.device_tensor(p, n, torch.int64, device), group=group)
    mark("stats all-reduce")
    gpu.set_global_hits(global_hits)
    gpu.filter(cov_cut_off, min_reads)
    gpu.assign()
    mark("filter+assign")
    p, n = gpu.assign_device()
    dist.all_reduce(api.device_tensor(p, n, torch.int32, device), group=group)
    mark("assign all-reduce")
    if phase_ms is not None:
        torch.cuda.synchronize(device)
        for (_, a), (name, b) in zip(marks, marks[1:]):
            phase_ms[name] = phase_ms.get(name, 0.0) + a.elapsed_time(b)
