"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: slice ownership, the items all-to-all and the
decomposition it relies on (reads sharded by read id, bins sharded by slice, per-reference statistics summed)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from slimm_b200 import dist as sdist
from slimm_b200 import synth

SHIFT = 12   # small slices so that a small test has many of them


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _case():
    rng = np.random.default_rng(5)
    tax, accs = synth.make_taxonomy(64)
    contigs = synth.make_contigs(64, rng, accs, 20_000, 90_000)
    rec = synth.make_records(contigs, 60_000, rng, multi_frac=0.3)
    lineage = synth.database_for(tax).lineage_table(contigs.accessions)
    return contigs, rec, lineage




def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        contigs, rec, lineage = _case()
        w = 100
        full = oracle.run(contigs.lengths, lineage, w, 100, 0.95, rec.read_id, rec.ref_id, rec.begin_pos)
        mine = (rec.read_id % world) == rank                        # shard by read: all records of a read on one rank
        part = oracle.run(contigs.lengths, lineage, w, 100, 0.95, rec.read_id[mine], rec.ref_id[mine], rec.begin_pos[mine])
        # items of this rank = its partial histogram expanded; grouped by slice like k_split leaves them
        n_bins = int(full.bin_off[-1])
        n_slices = (n_bins + (1 << SHIFT) - 1) >> SHIFT
        bins = np.repeat(np.arange(n_bins, dtype=np.int64), part.cov.astype(np.int64))
        uniq_left = part.uniq_cov.astype(np.int64).copy()
        flag = np.zeros(bins.size, dtype=np.int64)
        for k, b in enumerate(bins):                                # mark as many items unique as uniq_cov says
            if uniq_left[b] > 0:
                flag[k] = 1; uniq_left[b] -= 1
        items = torch.from_numpy((bins | (flag << 31)).astype(np.int64))
        counts = np.bincount(bins >> SHIFT, minlength=n_slices)
        recv, splits_in = sdist.exchange_items(items, counts)
        lo, hi = sdist.owned_slices(n_slices, rank, world)
        got = recv.numpy()
        gb = got & 0x7FFFFFFF
        assert ((gb >> SHIFT) >= lo).all() and ((gb >> SHIFT) < hi).all(), "an item arrived at a rank that does not own its slice"
        # accumulate the owned bins, then compare with the global oracle histogram on that range
        cov = np.bincount(gb, minlength=n_bins)
        ucov = np.bincount(gb[(got >> 31) & 1 == 1], minlength=n_bins)
        a, b = lo << SHIFT, min(n_bins, hi << SHIFT)
        np.testing.assert_array_equal(cov[a:b], full.cov[a:b])
        np.testing.assert_array_equal(ucov[a:b], full.uniq_cov[a:b])
        # partial per-reference statistics over the owned bins, summed over ranks == global statistics
        G = contigs.lengths.size
        stats = np.zeros((G, 4), dtype=np.int64)
        for g in range(G):
            s, e = max(int(full.bin_off[g]), a), min(int(full.bin_off[g + 1]), b)
            if e > s:
                stats[g] = [(cov[s:e] != 0).sum(), cov[s:e].sum(), (ucov[s:e] != 0).sum(), ucov[s:e].sum()]
        t = torch.from_numpy(stats)
        dist.all_reduce(t)
        np.testing.assert_array_equal(t.numpy()[:, 0], full.nz)
        np.testing.assert_array_equal(t.numpy()[:, 1], full.reads_count)
        np.testing.assert_array_equal(t.numpy()[:, 2], full.unz)
        np.testing.assert_array_equal(t.numpy()[:, 3], full.uniq_reads_count)
        # read-level counters are additive over read shards
        c = torch.tensor([part.n_reads, part.n_uniq, part.hits], dtype=torch.int64)
        dist.all_reduce(c)
        assert c.tolist() == [full.n_reads, full.n_uniq, full.hits]
        out.put((rank, "ok"))
    except Exception as e:   # surface the failure in the parent
        out.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_owned_slices_cover_everything():
    for n_slices in (1, 7, 416, 1000):
        for n in (1, 2, 3, 8):
            spans = [sdist.owned_slices(n_slices, r, n) for r in range(n)]
            assert spans[0][0] == 0 and spans[-1][1] == n_slices
            assert all(spans[i][1] == spans[i + 1][0] for i in range(n - 1))
    assert sdist.send_splits([5, 0, 2, 9], 2) == [5, 11]
    assert sdist.send_splits([5, 0, 2, 9], 4) == [5, 0, 2, 9]


def test_items_exchange_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
