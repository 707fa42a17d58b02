"""GPU parity at the shapes BASELINE.json names, against the CPU oracle on the same records:

* cfg2 exactly: 1 000 genomes, 10 M records, 20 % multi-mapped, bin width 1 000 - both scatter modes;
* cfg5-shaped: fine bins (w = 100) over enough genomes that the padded histogram spans more than 64 coarse slices
  and thousands of fine slices, hot / packed / empty fine slices mixed, 40 M records from the device generator;
* cfg4-shaped: every read on 2..64 references drawn inside the primary's species .. phylum (taxonomy-aware
  neighbourhoods, so LCAs land on every rank), -cc 1.0, profiles at species .. phylum.

The oracle (oracle/liboracle.so) needs seconds to a minute at these sizes; everything is compared bit for bit."""
import numpy as np
import pytest

import oracle
from slimm_b200 import api, synth

pytestmark = pytest.mark.gpu


def _community(G, seed, len_lo=1_000_000, len_hi=6_000_000, sigma=2.0):
    rng = np.random.default_rng(seed)
    tax, accs = synth.make_taxonomy(G)
    contigs = synth.make_contigs(G, rng, accs, len_lo, len_hi, sigma=sigma)
    db = synth.database_for(tax)
    lineage = db.lineage_table(contigs.accessions)
    return rng, contigs, db, lineage


def _compare(gpu, res, profile_ranks=(), taxa=None, lineage=None, contigs=None, bins_of=()):
    s = gpu.summary()
    assert (s.hits_count, s.matches_count, s.uniq_matches_count, s.uniq_matches_count2, s.n_pairs) == \
           (res.hits, res.n_reads, res.n_uniq, res.n_uniq2, res.n_pairs)
    assert np.float32(s.coverage_cut_off).tobytes() == np.float32(res.cut).tobytes()
    assert np.float32(s.uniq_coverage_cut_off).tobytes() == np.float32(res.ucut).tobytes()
    assert (s.n_valid, s.failed_by_cov, s.failed_by_uniq_cov, s.failed_by_min_read) == \
           (res.n_valid, res.failed_by_cov, res.failed_by_uniq_cov, res.failed_by_min_read)
    st = gpu.ref_stats()
    for x, y in ((st.reads_count, res.reads_count), (st.uniq_reads_count, res.uniq_reads_count), (st.nz_bins, res.nz),
                 (st.uniq_nz_bins, res.unz), (st.uniq_reads_count2, res.uniq_reads_count2), (st.valid, res.valid)):
        np.testing.assert_array_equal(x, y)
    assert st.cov_percent.tobytes() == res.cp.tobytes() and st.uniq_cov_percent.tobytes() == res.ucp.tobytes()
    assert gpu.lca_counts() == res.direct
    np.testing.assert_array_equal(gpu.lca_children(), res.child_pairs)
    for g in bins_of:
        a, b = int(res.bin_off[g]), int(res.bin_off[g + 1])
        np.testing.assert_array_equal(gpu.fetch_bins(0, int(g)), res.cov[a:b])
        np.testing.assert_array_equal(gpu.fetch_bins(1, int(g)), res.uniq_cov[a:b])
    if profile_ranks:
        rank_of = {t: v[0] for t, v in taxa.items()}
        name_of = {t: v[1] for t, v in taxa.items()}
        count, children = oracle.propagate(res.direct, res.child_pairs, res.uniq_reads_count2, lineage, rank_of)
        for rk in profile_ranks:
            rows = gpu.profile(rk, 0.001)
            exp = oracle.profile_rows(count, children, lineage, contigs.lengths, rank_of, name_of, res.n_reads, 100, res.cut, rk, 0.001)
            assert sorted((r.taxon, r.read_count) for r in rows if r.kind == 0) == \
                   sorted((int(r.taxa_id), r.read_count) for r in exp if not r.taxa_id.endswith("*")), f"rank {rk}"
            assert len([r for r in rows if r.kind == 0]) > 0 or rk > 1


@pytest.mark.parametrize("mode", [0, 1])
def test_cfg2_exact(mode):
    """BASELINE config 2 at its full size."""
    rng, contigs, db, lineage = _community(1000, 12345)
    rec = synth.make_records(contigs, 10_000_000, rng, multi_frac=0.2, k_lo=2, k_hi=8, neigh=8)
    res = oracle.run(contigs.lengths, lineage, 1000, 100, 0.95, rec.read_id, rec.ref_id, rec.begin_pos)
    taxa = {t: v for t, v in db.taxid__name.items()}
    with api.SlimmGpu(contigs.lengths, lineage, 1000, 100) as gpu:
        gpu.set_taxa(taxa)
        gpu.set_scatter_mode(mode)
        gpu.push(rec.read_id, rec.ref_id, rec.begin_pos)
        gpu.run(0.95)
        _compare(gpu, res, profile_ranks=(1, 2), taxa=taxa, lineage=lineage, contigs=contigs, bins_of=(0, 17, 500, 999))


def test_cfg5_shaped_fine_bins():
    """Fine bins at scale: > 64 coarse slices, > 17 000 fine slices, hot and packed and empty ones, 40 M records generated on
    the device (the bench's generator), copied to the host for the oracle."""
    import torch
    from slimm_b200 import synth_torch
    G, N, w = 8000, 40_000_000, 100
    rng, contigs, db, lineage = _community(G, 777, sigma=3.5)     # a long tail of references nobody maps to: empty fine slices
    dev = torch.device("cuda", 0)
    recs = synth_torch.make_records_device(contigs.lengths, contigs.weights, N, dev, seed=4711, multi_frac=0.2, k_lo=2, k_hi=8, neigh=8)
    rid = recs.read_id.cpu().numpy().view(np.uint32)
    ref = recs.ref_id.cpu().numpy().view(np.uint32)
    pos = recs.begin_pos.cpu().numpy()
    res = oracle.run(contigs.lengths, lineage, w, 100, 0.95, rid, ref, pos)
    n_bins = int(res.bin_off[-1])
    assert n_bins > 64 * (1 << 22)
    # items per fine slice (2^14 padded bins): hot slices (wide counters), ordinary ones (packed 16+16 bit) and empty ones all occur
    nb = contigs.lengths.astype(np.int64) // w + 1
    padded = np.concatenate([[0], np.cumsum((nb + 63) & ~63)])
    r64 = ref.astype(np.int64)
    bins = padded[r64] + np.minimum(pos.astype(np.int64) + 50, contigs.lengths.astype(np.int64)[r64]) // w
    per_fine = np.bincount(bins >> 14, minlength=(int(padded[-1]) >> 14) + 1)
    assert (per_fine >= 65536).any() and (per_fine == 0).any() and ((per_fine > 0) & (per_fine < 65536)).any()
    assert (per_fine >= (1 << 19)).any() and ((per_fine >= 65536) & (per_fine < (1 << 19))).any()     # cluster-shared and single-CTA hot slices
    del bins, r64
    taxa = {t: v for t, v in db.taxid__name.items()}
    with api.SlimmGpu(contigs.lengths, lineage, w, 100) as gpu:
        gpu.set_taxa(taxa)
        gpu.push_device(recs.read_id.data_ptr(), recs.ref_id.data_ptr(), recs.begin_pos.data_ptr(), recs.n)
        gpu.run(0.95)
        hot = int(np.argmax(res.reads_count))
        _compare(gpu, res, profile_ranks=(1,), taxa=taxa, lineage=lineage, contigs=contigs, bins_of=(0, hot, G // 2, G - 1))


def test_cfg3_exact():
    """BASELINE config 3 at its full size (10 000 genomes, 100 M records, 40 % multi-mapped, bin width 1 000) on one GPU; the
    sharded run of the same shape is covered by test_gpu_multi.py and by bench.py's sharded_equals_single."""
    import torch
    from slimm_b200 import synth_torch
    G, N, w = 10_000, 100_000_000, 1000
    rng, contigs, db, lineage = _community(G, 31)
    dev = torch.device("cuda", 0)
    recs = synth_torch.make_records_device(contigs.lengths, contigs.weights, N, dev, seed=3, multi_frac=0.4, k_lo=2, k_hi=8, neigh=8)
    rid = recs.read_id.cpu().numpy().view(np.uint32)
    ref = recs.ref_id.cpu().numpy().view(np.uint32)
    pos = recs.begin_pos.cpu().numpy()
    res = oracle.run(contigs.lengths, lineage, w, 100, 0.95, rid, ref, pos)
    taxa = {t: v for t, v in db.taxid__name.items()}
    with api.SlimmGpu(contigs.lengths, lineage, w, 100) as gpu:
        gpu.set_taxa(taxa)
        gpu.push_device(recs.read_id.data_ptr(), recs.ref_id.data_ptr(), recs.begin_pos.data_ptr(), recs.n)
        gpu.run(0.95)
        _compare(gpu, res, profile_ranks=(1, 2), taxa=taxa, lineage=lineage, contigs=contigs, bins_of=(0, int(np.argmax(res.reads_count)), G - 1))


def test_cfg4_shaped_lca_stress():
    """Reads on 2..64 references inside the primary's species .. phylum, -cc 1.0: the LCAs spread over all ranks; long reads
    (more than 32 records) take the whole-warp paths of both kernels."""
    G, N = 8192, 6_000_000
    rng, contigs, db, lineage = _community(G, 99, 200_000, 900_000)
    rec = synth.make_records(contigs, N, rng, multi_frac=1.0, k_lo=2, k_hi=64, neigh_mode="taxonomy")
    res = oracle.run(contigs.lengths, lineage, 1000, 100, 1.0, rec.read_id, rec.ref_id, rec.begin_pos)
    taxa = {t: v for t, v in db.taxid__name.items()}
    rank_of = {t: v[0] for t, v in taxa.items()}
    lca_ranks = {rank_of.get(t, 0) for t in res.direct}
    assert {1, 2, 3, 4, 5, 6} <= lca_ranks, lca_ranks                     # species .. phylum all occur as LCAs
    with api.SlimmGpu(contigs.lengths, lineage, 1000, 100, flags=api.READ_RESULTS) as gpu:
        gpu.set_taxa(taxa)
        gpu.push(rec.read_id, rec.ref_id, rec.begin_pos)
        gpu.run(1.0)
        _compare(gpu, res, profile_ranks=(1, 2, 3, 4, 5, 6), taxa=taxa, lineage=lineage, contigs=contigs)
        rid, kind, val = gpu.read_results()
        o = np.argsort(rid, kind="stable")
        rid, kind, val = rid[o], kind[o], val[o]
        exp_lca = np.nonzero(res.read_n_valid >= 2)[0]
        np.testing.assert_array_equal(rid[kind == 2], exp_lca)
        np.testing.assert_array_equal(val[kind == 2], res.read_lca[exp_lca])
