"""Two GPUs, one process each (NCCL): the sharded run - reads by read id, bins by histogram slice, items routed with
one all-to-all - gives exactly the single-GPU / oracle results on every rank.  Skipped on a one-GPU box."""
import os
import socket

import numpy as np
import pytest

import oracle
from slimm_b200 import api, synth

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _case(w):
    rng = np.random.default_rng(21)
    tax, accs = synth.make_taxonomy(1500)
    contigs = synth.make_contigs(1500, rng, accs, 200_000, 900_000)
    rec = synth.make_records(contigs, 1_500_000, rng, multi_frac=0.35)
    db = synth.database_for(tax)
    lineage = db.lineage_table(contigs.accessions)
    return contigs, rec, lineage, {t: v for t, v in db.taxid__name.items()}


_PEER_MODES = {"p2p_split": "0", "p2p_route": "1", "p2p_blocks": "2"}   # SLIMM_PEER_ROUTE (csrc/slimm_gpu.cu)


def _worker(rank, world, port, w, exchange, out):
    small_buffer = exchange.endswith("+small_buffer")
    exchange = exchange.split("+")[0]
    p2p = exchange in _PEER_MODES
    if p2p:
        os.environ["SLIMM_PEER_ROUTE"] = _PEER_MODES[exchange]
    import torch
    import torch.distributed as dist
    from slimm_b200 import dist as sdist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        contigs, rec, lineage, taxa = _case(w)
        res = oracle.run(contigs.lengths, lineage, w, 100, 0.9, rec.read_id, rec.ref_id, rec.begin_pos)
        # contiguous read ranges keep each rank's ids non-decreasing (no device sort); every read on one rank
        cut = np.searchsorted(rec.read_id, np.linspace(0, rec.n_reads, world + 1).astype(np.int64))
        a, b = int(cut[rank]), int(cut[rank + 1])
        with api.SlimmGpu(contigs.lengths, lineage, w, 100, device=rank) as gpu:
            gpu.set_stream(torch.cuda.current_stream().cuda_stream)
            gpu.set_taxa(taxa)
            for it in range(2):                      # twice: a context is reused between samples
                gpu.reset()
                gpu.set_shard(rank, world)
                if p2p and it == 0:                  # items travel as peer-to-peer stores inside the split instead of NCCL
                    if not sdist.connect_peers(gpu, dev, 1000 if small_buffer else int(rec.read_id.size)):
                        out.put((rank, "skip: CUDA IPC peer mapping is not available on this box"))
                        return
                gpu.push(rec.read_id[a:b], rec.ref_id[a:b], rec.begin_pos[a:b])
                sdist.run_sharded(gpu, dev, 0.9, 0, int(rec.read_id.size))
                if small_buffer:                     # every rank sees from the all-gathered table that a receive buffer is too small:
                    with pytest.raises(api.SlimmGpuError):   # nothing is copied, the run completes, the result is refused
                        gpu.summary()
                    continue
                s = gpu.summary()
                assert (s.hits_count, s.matches_count, s.uniq_matches_count, s.uniq_matches_count2) == \
                       (res.hits, res.n_reads, res.n_uniq, res.n_uniq2)
                assert np.float32(s.coverage_cut_off).tobytes() == np.float32(res.cut).tobytes()
                assert np.float32(s.uniq_coverage_cut_off).tobytes() == np.float32(res.ucut).tobytes()
                assert (s.n_valid, s.n_pairs) == (res.n_valid, res.n_pairs)
                st = gpu.ref_stats()
                for x, y in ((st.reads_count, res.reads_count), (st.uniq_reads_count, res.uniq_reads_count),
                             (st.uniq_reads_count2, res.uniq_reads_count2), (st.nz_bins, res.nz), (st.uniq_nz_bins, res.unz),
                             (st.valid, res.valid)):
                    np.testing.assert_array_equal(x, y)
                assert gpu.lca_counts() == res.direct
                np.testing.assert_array_equal(gpu.lca_children(), res.child_pairs)
                rows = gpu.profile(1, 0.001)
                count, children = oracle.propagate(res.direct, res.child_pairs, res.uniq_reads_count2, lineage,
                                                   {t: v[0] for t, v in taxa.items()})
                exp = oracle.profile_rows(count, children, lineage, contigs.lengths, {t: v[0] for t, v in taxa.items()},
                                          {t: v[1] for t, v in taxa.items()}, res.n_reads, 100, res.cut, 1, 0.001)
                assert [(r.taxon, r.read_count) for r in rows if r.kind == 0] == \
                       [(int(r.taxa_id), r.read_count) for r in exp if not r.taxa_id.endswith("*")]
                # the bins of a reference live on the rank that owns its slices: compare those this rank owns entirely
                nb = contigs.lengths.astype(np.int64) // w + 1
                padded = np.concatenate([[0], np.cumsum((nb + 63) & ~63)])
                n_slices = (int(padded[-1]) + (1 << 22) - 1) >> 22
                lo = (n_slices * rank // world) << 22
                hi = (n_slices * (rank + 1) // world) << 22
                checked = 0
                for g in range(0, contigs.lengths.size, 37):
                    ga, gb = int(padded[g]), int(padded[g + 1])
                    if ga >= lo and gb <= hi:
                        x, y = int(res.bin_off[g]), int(res.bin_off[g + 1])
                        np.testing.assert_array_equal(gpu.fetch_bins(0, g), res.cov[x:y])
                        np.testing.assert_array_equal(gpu.fetch_bins(1, g), res.uniq_cov[x:y])
                        checked += 1
                assert checked > 0 or n_slices < world
        out.put((rank, "ok"))
    except Exception as e:
        import traceback
        out.put((rank, traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["nccl_all_to_all", "p2p_split", "p2p_route", "p2p_blocks", "p2p_blocks+small_buffer"])
@pytest.mark.parametrize("w", [10, 1000])
def test_two_gpus_match_oracle(w, exchange):
    """The four item exchanges: one NCCL all-to-all of the slice-grouped items; peer stores inside the split (runs per slice);
    routed (ranked by owner, grouped by slice on arrival); blocks (grouped by slice locally, one contiguous copy per owner - the
    default)."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, w, exchange, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = [out.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    if all(isinstance(m, str) and m.startswith("skip:") for _, m in results):
        pytest.skip(results[0][1])
    assert sorted(results) == [(0, "ok"), (1, "ok")], results
