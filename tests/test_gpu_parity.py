"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle and the reference
binaries' golden outputs.  Integers, index sets and the two f32 cut-offs must be bit-exact."""
import os

import numpy as np
import pytest

import oracle
from golden_util import all_runs, assert_profiles_match, load_case, runs_of
from slimm_b200 import api, report, synth

pytestmark = pytest.mark.gpu


def compare_with_oracle(gpu: api.SlimmGpu, res: oracle.OracleResult, lineage, check_bins=True, cov2=True):
    s = gpu.summary()
    assert s.hits_count == res.hits
    assert s.matches_count == res.n_reads
    assert s.uniq_matches_count == res.n_uniq
    assert s.uniq_matches_count2 == res.n_uniq2
    assert s.n_pairs == res.n_pairs
    assert s.n_bins == int(res.bin_off[-1])
    assert s.reference_count == int((res.reads_count > 0).sum())
    assert np.float32(s.coverage_cut_off).tobytes() == np.float32(res.cut).tobytes(), (s.coverage_cut_off, res.cut)
    assert np.float32(s.uniq_coverage_cut_off).tobytes() == np.float32(res.ucut).tobytes()
    assert (s.n_valid, s.failed_by_cov, s.failed_by_uniq_cov, s.failed_by_min_read, s.min_reads) == \
           (res.n_valid, res.failed_by_cov, res.failed_by_uniq_cov, res.failed_by_min_read, res.min_reads)
    st = gpu.ref_stats()
    np.testing.assert_array_equal(st.reads_count, res.reads_count)
    np.testing.assert_array_equal(st.uniq_reads_count, res.uniq_reads_count)
    np.testing.assert_array_equal(st.uniq_reads_count2, res.uniq_reads_count2)
    np.testing.assert_array_equal(st.nz_bins, res.nz)
    np.testing.assert_array_equal(st.uniq_nz_bins, res.unz)
    assert st.cov_percent.tobytes() == res.cp.tobytes()
    assert st.uniq_cov_percent.tobytes() == res.ucp.tobytes()
    np.testing.assert_array_equal(st.valid, res.valid)
    assert gpu.lca_counts() == res.direct
    np.testing.assert_array_equal(gpu.lca_children(), res.child_pairs)
    if check_bins:
        hists = [(0, res.cov), (1, res.uniq_cov)] + ([(2, res.uniq_cov2)] if cov2 else [])
        for g in range(gpu.n_refs):
            a, b = int(res.bin_off[g]), int(res.bin_off[g + 1])
            for which, h in hists:
                np.testing.assert_array_equal(gpu.fetch_bins(which, g), h[a:b], err_msg=f"bins {which} of ref {g}")
        if cov2:
            np.testing.assert_array_equal(gpu.uniq2_nz(), res.unz2)
    # per-read reassignments and LCAs
    rid, kind, val = gpu.read_results()
    exp_assigned = np.nonzero(res.read_n_valid == 1)[0]
    exp_lca = np.nonzero(res.read_n_valid >= 2)[0]
    o = np.argsort(rid, kind="stable")
    rid, kind, val = rid[o], kind[o], val[o]
    np.testing.assert_array_equal(rid[kind == 1], exp_assigned)
    np.testing.assert_array_equal(val[kind == 1], res.read_assigned[exp_assigned])
    np.testing.assert_array_equal(rid[kind == 2], exp_lca)
    np.testing.assert_array_equal(val[kind == 2], res.read_lca[exp_lca])


@pytest.mark.parametrize("case_name,run_name", all_runs())
def test_golden_case(case_name, run_name):
    case = load_case(case_name)
    run = [r for r in runs_of(case) if r.name == run_name][0]
    w = run.bin_width or case.avg_read_length
    res = oracle.run(case.ref_len, case.lineage, w, case.avg_read_length, run.cov_cut_off, case.read_id, case.ref_id,
                     case.begin_pos, run.min_reads)
    taxa = {t: (case.rank_of[t], case.name_of[t]) for t in case.rank_of}
    with api.SlimmGpu(case.ref_len, case.lineage, w, case.avg_read_length,
                      flags=api.KEEP_UNIQ_COV2 | api.READ_RESULTS) as gpu:
        gpu.set_taxa(taxa)
        for mode in (0, 1):   # direct REDs, then the bucketed multisplit scatter
            gpu.reset()
            gpu.set_scatter_mode(mode)
            # two batches, to cover a read that straddles a push boundary
            h = case.read_id.size // 2
            gpu.push(case.read_id[:h], case.ref_id[:h], case.begin_pos[:h])
            gpu.push(case.read_id[h:], case.ref_id[h:], case.begin_pos[h:])
            gpu.run(run.cov_cut_off, run.min_reads)
            compare_with_oracle(gpu, res, case.lineage, check_bins=case.ref_len.size <= 64)
            s = gpu.summary()
            # the host tail fed from the taxon lists ...
            rows = api.profile_rows(case.ref_len, case.lineage, taxa, gpu.lca_counts(), gpu.lca_children(),
                                    gpu.ref_stats().uniq_reads_count2, s.matches_count, case.avg_read_length,
                                    s.coverage_cut_off, run.abundance_cut_off, run.rank)
            lines = report.profile_lines(rows, case.lineage, case.name_of, run.rank)
            assert_profiles_match(os.path.join(run.path, "profile.tsv"), lines)
            # ... and straight from the context
            rows2 = gpu.profile(run.rank, run.abundance_cut_off)
            assert report.profile_lines(rows2, case.lineage, case.name_of, run.rank) == lines


def _synthetic(G, N, seed, **kw):
    rng = np.random.default_rng(seed)
    tax, accs = synth.make_taxonomy(G)
    contigs = synth.make_contigs(G, rng, accs, kw.pop("len_lo", 100_000), kw.pop("len_hi", 600_000))
    rec = synth.make_records(contigs, N, rng, **kw)
    lineage = synth.database_for(tax).lineage_table(contigs.accessions)
    return contigs, rec, lineage


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("shuffle", [False, True])
@pytest.mark.parametrize("G,N,w,cc,kw", [
    (1000, 1_000_000, 1000, 0.95, dict(multi_frac=0.2)),
    (1000, 300_000, 100, 0.5, dict(multi_frac=0.4)),
    (512, 400_000, 1000, 1.0, dict(multi_frac=0.6, k_lo=2, k_hi=64, neigh=64)),
    (3, 50_000, 7, 0.95, dict(multi_frac=0.5, k_lo=2, k_hi=3, neigh=2, len_lo=500, len_hi=3000)),
    (4097, 200_000, 250, 0.9, dict(multi_frac=0.3, len_lo=20_000, len_hi=50_000)),
])
def test_synthetic_vs_oracle(G, N, w, cc, kw, shuffle, mode):
    contigs, rec, lineage = _synthetic(G, N, 1234 + G, shuffle=shuffle, **kw)
    res = oracle.run(contigs.lengths, lineage, w, 100, cc, rec.read_id, rec.ref_id, rec.begin_pos)
    with api.SlimmGpu(contigs.lengths, lineage, w, 100, flags=api.KEEP_UNIQ_COV2 | api.READ_RESULTS) as gpu:
        gpu.set_scatter_mode(mode)
        gpu.push(rec.read_id, rec.ref_id, rec.begin_pos)
        gpu.run(cc)
        assert gpu.summary().input_was_sorted == (0 if shuffle else 1)
        compare_with_oracle(gpu, res, lineage, check_bins=G <= 16)
        # bins of a sample of references
        for g in np.random.default_rng(0).choice(G, size=min(G, 24), replace=False):
            a, b = int(res.bin_off[g]), int(res.bin_off[g + 1])
            np.testing.assert_array_equal(gpu.fetch_bins(0, int(g)), res.cov[a:b])
            np.testing.assert_array_equal(gpu.fetch_bins(1, int(g)), res.uniq_cov[a:b])
            np.testing.assert_array_equal(gpu.fetch_bins(2, int(g)), res.uniq_cov2[a:b])


def test_edge_cases():
    ref_len = np.array([1000, 50, 99], dtype=np.uint32)
    lineage = np.array([[11, 1, 2, 3, 4, 5, 6, 7], [12, 1, 2, 3, 4, 5, 6, 7], [0] * 8], dtype=np.uint32)
    cases = {
        "empty": (np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.int32)),
        "single": (np.array([0], np.uint32), np.array([1], np.uint32), np.array([-1], np.int32)),
        "one_read_many_repeats": (np.zeros(500, np.uint32), np.zeros(500, np.uint32), np.arange(500, dtype=np.int32)),
        "all_levels_zero_lca": (np.array([0, 0, 1, 1], np.uint32), np.array([0, 2, 1, 2], np.uint32),
                                np.array([5, 6, 7, 2 ** 31 - 1], np.int32)),
    }
    for name, (rid, ref, pos) in cases.items():
        for cc in (0.95, 1.0, 0.0):
            res = oracle.run(ref_len, lineage, 10, 100, cc, rid, ref, pos)
            for mode in (0, 1):
                with api.SlimmGpu(ref_len, lineage, 10, 100, flags=api.KEEP_UNIQ_COV2 | api.READ_RESULTS) as gpu:
                    gpu.set_scatter_mode(mode)
                    gpu.push(rid, ref, pos)
                    gpu.run(cc)
                    compare_with_oracle(gpu, res, lineage)


def test_bad_reference_id_is_rejected():
    ref_len = np.array([1000], dtype=np.uint32)
    with api.SlimmGpu(ref_len, np.zeros((1, 8), np.uint32), 10, 100) as gpu:
        gpu.push(np.array([0], np.uint32), np.array([5], np.uint32), np.array([1], np.int32))
        with pytest.raises(api.SlimmGpuError):
            gpu.run(0.95)


def test_reset_reuses_context():
    contigs, rec, lineage = _synthetic(64, 50_000, 5)
    with api.SlimmGpu(contigs.lengths, lineage, 1000, 100, flags=api.KEEP_UNIQ_COV2 | api.READ_RESULTS) as gpu:
        for w in (1000, 333):
            gpu.reset(w, 100)
            gpu.bin_width = w
            res = oracle.run(contigs.lengths, lineage, w, 100, 0.95, rec.read_id, rec.ref_id, rec.begin_pos)
            gpu.push(rec.read_id, rec.ref_id, rec.begin_pos)
            gpu.run(0.95)
            compare_with_oracle(gpu, res, lineage)


def test_bucketed_scatter_many_buckets():
    """Fine bins over long contigs: the padded histogram spans > 16 buckets of 2^22 bins."""
    contigs, rec, lineage = _synthetic(40, 2_000_000, 77, len_lo=2_000_000, len_hi=6_000_000, multi_frac=0.3)
    w = 2
    res = oracle.run(contigs.lengths, lineage, w, 100, 0.95, rec.read_id, rec.ref_id, rec.begin_pos)
    assert int(res.bin_off[-1]) > 16 * (1 << 22)
    for mode, flags in ((1, api.KEEP_UNIQ_COV2 | api.READ_RESULTS), (-1, 0)):
        with api.SlimmGpu(contigs.lengths, lineage, w, 100, flags=flags) as gpu:
            gpu.set_scatter_mode(mode)
            gpu.push(rec.read_id, rec.ref_id, rec.begin_pos)
            gpu.run(0.95)
            if flags:
                compare_with_oracle(gpu, res, lineage, check_bins=False)
                for g in (0, 7, 39):
                    a, b = int(res.bin_off[g]), int(res.bin_off[g + 1])
                    np.testing.assert_array_equal(gpu.fetch_bins(0, g), res.cov[a:b])
                    np.testing.assert_array_equal(gpu.fetch_bins(1, g), res.uniq_cov[a:b])
                    np.testing.assert_array_equal(gpu.fetch_bins(2, g), res.uniq_cov2[a:b])
            else:
                st = gpu.ref_stats()
                np.testing.assert_array_equal(st.reads_count, res.reads_count)
                np.testing.assert_array_equal(st.uniq_reads_count2, res.uniq_reads_count2)
                np.testing.assert_array_equal(st.nz_bins, res.nz)
                assert gpu.lca_counts() == res.direct


@pytest.mark.parametrize("env,value", [("SLIMM_GPU_TAIL", "host"), ("SLIMM_GPU_CUTOFF", "global")])
def test_alternative_paths_agree(env, value, monkeypatch):
    """The general host tail / the global-memory cut-off sort give the same rows and statistics as the default
    device rank reduction / cluster sort."""
    contigs, rec, lineage = _synthetic(2000, 600_000, 99, multi_frac=0.35)
    tax, accs = synth.make_taxonomy(2000)
    db = synth.database_for(tax)
    taxa = {t: v for t, v in db.taxid__name.items()}
    assert api.db_is_tree_consistent(lineage, taxa)
    res = oracle.run(contigs.lengths, lineage, 1000, 100, 0.9, rec.read_id, rec.ref_id, rec.begin_pos)
    out = []
    for use_alt in (False, True):
        if use_alt:
            monkeypatch.setenv(env, value)
        with api.SlimmGpu(contigs.lengths, lineage, 1000, 100, flags=api.KEEP_UNIQ_COV2 | api.READ_RESULTS) as gpu:
            gpu.set_taxa(taxa)
            gpu.push(rec.read_id, rec.ref_id, rec.begin_pos)
            gpu.run(0.9)
            compare_with_oracle(gpu, res, lineage, check_bins=False)
            for rank in (1, 2, 6):
                out.append([(r.taxon, r.kind, r.read_count, r.first_child, np.float64(r.abundance).tobytes())
                            for r in gpu.profile(rank, 0.001)])
    half = len(out) // 2
    assert out[:half] == out[half:]
    assert len(out[0]) > 10


@pytest.mark.parametrize("env,value", [("SLIMM_GPU_ACC", "l2"), ("SLIMM_GPU_COV", "window"), ("SLIMM_GPU_FINE", "wide"),
                                       ("SLIMM_GPU_COMPACT_BINS", "0"), ("SLIMM_GPU_FINE_CLUSTER", "0"), ("SLIMM_SPLIT_BULK", "1"),
                                       ("SLIMM_FINE_SPLIT_NT", "512"), ("SLIMM_FORCE_EXACT", "1")])
def test_alternative_kernels_agree(env, value, monkeypatch):
    """The other kernels stay selectable for A/B runs (64-bit REDs into L2-resident slices instead of the shared-memory
    fine slices; sliding-window coverage instead of warp-private tiles; wide counters only; interleaved instead of compact
    bins; hot slices without clusters; bulk-copy tile staging in the coarse split; 512-thread fine split; the EXACT instantiations of
    the split / count kernels that histograms within 2^23 bins of 2^31 take): same results, bins included."""
    contigs, rec, lineage = _synthetic(300, 1_500_000, 5, len_lo=300_000, len_hi=900_000, multi_frac=0.4)
    w = 10
    res = oracle.run(contigs.lengths, lineage, w, 100, 0.9, rec.read_id, rec.ref_id, rec.begin_pos)
    monkeypatch.setenv(env, value)
    with api.SlimmGpu(contigs.lengths, lineage, w, 100, flags=api.KEEP_UNIQ_COV2 | api.READ_RESULTS) as gpu:
        gpu.set_scatter_mode(1)
        gpu.push(rec.read_id, rec.ref_id, rec.begin_pos)
        gpu.run(0.9)
        compare_with_oracle(gpu, res, lineage, check_bins=False)
        for g in (0, 150, 299):
            a, b = int(res.bin_off[g]), int(res.bin_off[g + 1])
            np.testing.assert_array_equal(gpu.fetch_bins(0, g), res.cov[a:b])
            np.testing.assert_array_equal(gpu.fetch_bins(1, g), res.uniq_cov[a:b])


def test_skip_bins_profile_only_run():
    """SLIMM_GPU_SKIP_BINS: the bins live in shared memory only - every statistic and the profile are unchanged, the
    bins themselves cannot be fetched."""
    contigs, rec, lineage = _synthetic(300, 1_500_000, 6, len_lo=300_000, len_hi=900_000, multi_frac=0.4)
    w = 10
    res = oracle.run(contigs.lengths, lineage, w, 100, 0.9, rec.read_id, rec.ref_id, rec.begin_pos)
    with api.SlimmGpu(contigs.lengths, lineage, w, 100, flags=api.SKIP_BINS) as gpu:
        gpu.set_scatter_mode(1)
        for _ in range(2):
            gpu.reset()
            gpu.push(rec.read_id, rec.ref_id, rec.begin_pos)
            gpu.run(0.9)
            s = gpu.summary()
            assert (s.hits_count, s.matches_count, s.uniq_matches_count, s.uniq_matches_count2, s.n_valid) == \
                   (res.hits, res.n_reads, res.n_uniq, res.n_uniq2, res.n_valid)
            assert np.float32(s.coverage_cut_off).tobytes() == np.float32(res.cut).tobytes()
            st = gpu.ref_stats()
            for x, y in ((st.reads_count, res.reads_count), (st.uniq_reads_count, res.uniq_reads_count), (st.nz_bins, res.nz),
                         (st.uniq_nz_bins, res.unz), (st.uniq_reads_count2, res.uniq_reads_count2), (st.valid, res.valid)):
                np.testing.assert_array_equal(x, y)
            assert gpu.lca_counts() == res.direct
            with pytest.raises(api.SlimmGpuError):
                gpu.fetch_bins(0, 0)
    with pytest.raises(api.SlimmGpuError):
        api.SlimmGpu(contigs.lengths, lineage, w, 100, flags=api.SKIP_BINS | api.KEEP_UNIQ_COV2)


@pytest.mark.parametrize("n_records", [400_000, 1_500_000])
def test_hot_fine_slice(n_records):
    """All records in one fine slice (a few short references): more than 65535 items per slice, so the packed 16+16-bit
    counters do not apply and the wide shared-memory variant runs; with 2^19 items or more a cluster of CTAs shares the
    slice (private copies of the bins summed through distributed shared memory); bins included."""
    contigs, rec, lineage = _synthetic(4, n_records, 8, len_lo=2000, len_hi=5000, multi_frac=0.5, k_lo=2, k_hi=3, neigh=2)
    w = 5
    res = oracle.run(contigs.lengths, lineage, w, 100, 0.95, rec.read_id, rec.ref_id, rec.begin_pos)
    assert res.n_pairs > 65536 and int(res.bin_off[-1]) < 16384
    assert (res.n_pairs >= (1 << 19)) == (n_records > 1_000_000)
    with api.SlimmGpu(contigs.lengths, lineage, w, 100) as gpu:
        gpu.set_scatter_mode(1)
        gpu.push(rec.read_id, rec.ref_id, rec.begin_pos)
        gpu.run(0.95)
        st = gpu.ref_stats()
        for x, y in ((st.reads_count, res.reads_count), (st.uniq_reads_count, res.uniq_reads_count), (st.nz_bins, res.nz),
                     (st.uniq_nz_bins, res.unz), (st.uniq_reads_count2, res.uniq_reads_count2), (st.valid, res.valid)):
            np.testing.assert_array_equal(x, y)
        for g in range(4):
            a, b = int(res.bin_off[g]), int(res.bin_off[g + 1])
            np.testing.assert_array_equal(gpu.fetch_bins(0, g), res.cov[a:b])
            np.testing.assert_array_equal(gpu.fetch_bins(1, g), res.uniq_cov[a:b])
    # one bin hit by 100 000 unique reads: a counter beyond 16 bits
    n = 100_000
    rid = np.arange(n, dtype=np.uint32); ref = np.zeros(n, dtype=np.uint32); pos = np.full(n, 100, dtype=np.int32)
    res = oracle.run(contigs.lengths, lineage, w, 100, 0.95, rid, ref, pos)
    assert int(res.cov.max()) == n
    with api.SlimmGpu(contigs.lengths, lineage, w, 100) as gpu:
        gpu.set_scatter_mode(1)
        gpu.push(rid, ref, pos)
        gpu.run(0.95)
        a, b = int(res.bin_off[0]), int(res.bin_off[1])
        np.testing.assert_array_equal(gpu.fetch_bins(0, 0), res.cov[a:b])
        np.testing.assert_array_equal(gpu.fetch_bins(1, 0), res.uniq_cov[a:b])
        np.testing.assert_array_equal(gpu.ref_stats().reads_count, res.reads_count)


def _pack(read_id, ref_id):
    """The wire format of grouped input: bit i set when record i starts a read, 16-bit reference ids."""
    n = read_id.size
    new = np.ones(n, dtype=bool)
    new[1:] = read_id[1:] != read_id[:-1]
    bits = np.packbits(np.concatenate([new, np.zeros((-n) % 32, dtype=bool)]), bitorder="little").view(np.uint32)
    return bits, ref_id.astype(np.uint16)


@pytest.mark.parametrize("mode", [0, 1])
def test_push_packed_wire_format(mode):
    """slimm_gpu_push_packed (one new-read bit + u16 reference id + position per record) in uneven batches gives what
    slimm_gpu_push gives: the ids rebuilt on the device are the dense ids, continued across batches."""
    contigs, rec, lineage = _synthetic(1000, 700_000, 321, multi_frac=0.3)
    res = oracle.run(contigs.lengths, lineage, 500, 100, 0.9, rec.read_id, rec.ref_id, rec.begin_pos)
    cuts = [0, 1, 33, 100_001, 100_001, 433_333, rec.read_id.size]          # one-record, empty and odd-sized batches
    with api.SlimmGpu(contigs.lengths, lineage, 500, 100, flags=api.KEEP_UNIQ_COV2 | api.READ_RESULTS) as gpu:
        for _ in range(2):                                                   # the id counter restarts with every sample
            gpu.reset()
            gpu.set_scatter_mode(mode)
            for a, b in zip(cuts, cuts[1:]):
                # a batch's own bit array starts at bit 0; a read may straddle two batches (its later records carry bit 0)
                new = np.ones(b - a, dtype=bool)
                new[1:] = rec.read_id[a + 1:b] != rec.read_id[a:b - 1]
                if a and b > a:
                    new[0] = rec.read_id[a] != rec.read_id[a - 1]
                bits = np.packbits(np.concatenate([new, np.zeros((-(b - a)) % 32, dtype=bool)]), bitorder="little")
                bits = np.concatenate([bits, np.zeros((-bits.size) % 4, dtype=np.uint8)]).view(np.uint32)
                gpu.push_packed(bits, rec.ref_id[a:b].astype(np.uint16), rec.begin_pos[a:b])
            gpu.run(0.9)
            compare_with_oracle(gpu, res, lineage, check_bins=False)
        with pytest.raises(api.SlimmGpuError):                               # the two ingest formats do not mix inside a sample
            gpu.reset()
            gpu.push(rec.read_id[:10], rec.ref_id[:10], rec.begin_pos[:10])
            gpu.push_packed(*_pack(rec.read_id[10:20], rec.ref_id[10:20]), rec.begin_pos[10:20])
