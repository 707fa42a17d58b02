"""Device-side twin of ``synth.make_records`` for workloads too large to generate with numpy
(BASELINE.json configs 3-5: 1e8-1e9 records).  Same distributions, torch RNG; used by bench.py
and the full-size GPU tests.  torch is plumbing here (device memory + RNG), not the product."""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class DeviceRecords:
    read_id: torch.Tensor      # int32 [N] (bit pattern of the u32 id), non-decreasing
    ref_id: torch.Tensor       # int32 [N]
    begin_pos: torch.Tensor    # int32 [N]
    n_reads: int

    @property
    def n(self) -> int:
        return int(self.read_id.numel())


def make_records_device(lengths: np.ndarray, weights: np.ndarray, n_records: int, device, seed: int = 12345,
                        multi_frac: float = 0.2, k_lo: int = 2, k_hi: int = 8, neigh: int = 8,
                        repeat_frac: float = 0.002, read_len: int = 100, shuffle: bool = False,
                        neigh_mode: str = "index", read_id_base: int = 0) -> DeviceRecords:
    """``neigh_mode``: "index" - extra targets of a multi-mapped read come from the +-``neigh`` index neighbourhood;
    "taxonomy" - every multi-mapped read draws a level (species 40 %, genus 30 %, family 15 %, order 8 %, class 4 %,
    phylum 3 %) and its extra targets uniformly among the genomes that share that taxon with the primary one
    (fan-out 4/4/4/2/2/2 of ``synth.make_taxonomy``), so LCAs land on every rank.
    ``read_id_base`` is added to every read id (block-wise generation: a block's ids stay disjoint from the others')."""
    G = int(lengths.size)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    mean_k = (1.0 - multi_frac) + multi_frac * 0.5 * (k_lo + k_hi)
    R = int(math.ceil(n_records / (mean_k * (1.0 + repeat_frac)) * 1.01)) + 1024
    cdf = torch.from_numpy(np.cumsum(weights.astype(np.float64))).to(device)
    cdf[-1] = 1.0
    primary = torch.empty(R, dtype=torch.int32, device=device)
    chunk = 1 << 26
    for a in range(0, R, chunk):          # chunked: the f64 uniforms are the largest temporary
        b = min(R, a + chunk)
        u = torch.rand(b - a, generator=gen, device=device, dtype=torch.float64)
        primary[a:b] = torch.searchsorted(cdf, u, right=True).clamp_(max=G - 1).to(torch.int32)
        del u
    is_multi = torch.rand(R, generator=gen, device=device) < multi_frac
    k = torch.randint(k_lo, k_hi + 1, (R,), generator=gen, device=device, dtype=torch.int32)
    k = torch.where(is_multi, k, torch.ones_like(k))
    del is_multi
    k64 = k.to(torch.int64)
    start = torch.cumsum(k64, 0) - k64
    total = int(start[-1].item() + k64[-1].item())
    read_of = torch.repeat_interleave(torch.arange(R, dtype=torch.int32, device=device), k64, output_size=total)
    j0 = torch.arange(total, dtype=torch.int64, device=device) == start[read_of.long()]
    del start, k64, k
    p = primary[read_of.long()]
    if neigh_mode == "taxonomy":
        sizes = torch.tensor([4, 16, 64, 128, 256, 512], dtype=torch.int32, device=device)      # genomes under a species .. phylum
        cum = torch.tensor([0.40, 0.70, 0.85, 0.93, 0.97, 1.0], dtype=torch.float32, device=device)
        lvl = torch.searchsorted(cum, torch.rand(R, generator=gen, device=device)).clamp_(max=5)
        size = sizes[lvl][read_of.long()]
        pick = (torch.rand(total, generator=gen, device=device) * size.to(torch.float32)).to(torch.int32).clamp_(max=511)
        pick = torch.minimum(pick, size - 1)
        ref = torch.where(j0, p, ((p // size) * size + pick).clamp_(0, G - 1))
        del sizes, cum, lvl, size, pick
    else:
        off = torch.randint(1, neigh + 1, (total,), generator=gen, device=device, dtype=torch.int32)
        sign = torch.randint(0, 2, (total,), generator=gen, device=device, dtype=torch.int32) * 2 - 1
        ref = torch.where(j0, p, (p + off * sign).clamp_(0, G - 1))
        del off, sign
    del p, j0, primary
    # planted repeat hits directly after the original record
    times = (torch.rand(total, generator=gen, device=device) < repeat_frac).to(torch.int64) + 1
    total2 = int(times.sum().item())
    read_of = torch.repeat_interleave(read_of, times, output_size=total2)
    ref = torch.repeat_interleave(ref, times, output_size=total2)
    del times
    if total2 < n_records:
        raise RuntimeError("generator undershoot; raise the read count factor")
    read_of = read_of[:n_records].contiguous()
    ref = ref[:n_records].contiguous()
    span = (torch.from_numpy(lengths.astype(np.int64)).to(device) - read_len).clamp_(min=1).to(torch.float64)
    pos = torch.empty(n_records, dtype=torch.int32, device=device)
    for a in range(0, n_records, chunk):
        b = min(n_records, a + chunk)
        u = torch.rand(b - a, generator=gen, device=device, dtype=torch.float64)
        pos[a:b] = (u * span[ref[a:b].long()]).to(torch.int32)
        del u
    n_reads = int(read_of[-1].item()) + 1
    if read_id_base:
        read_of += read_id_base
    if shuffle:
        perm = torch.randperm(n_records, generator=gen, device=device)
        read_of, ref, pos = read_of[perm].contiguous(), ref[perm].contiguous(), pos[perm].contiguous()
    return DeviceRecords(read_of, ref, pos, n_reads)
