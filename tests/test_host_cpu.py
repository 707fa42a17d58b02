"""CPU-side checks of the product library: it loads without a GPU, exports every symbol the header
declares, refuses to run without a device (no fallback), and its host tail (slimm_profile_rows)
reproduces the reference's _profile.tsv when fed the oracle's stage outputs."""
import os
import re

import numpy as np
import pytest

import oracle
from golden_util import RANKS, all_runs, assert_profiles_match, load_case, runs_of
from slimm_b200 import api, report

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols():
    lib = api.load_library()
    hdr = open(os.path.join(ROOT, "include", "slimm_gpu.h")).read()
    declared = set(re.findall(r"\b(slimm_(?:gpu|profile)_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in slimm_gpu.h but not exported"
    assert declared == set(api.EXPORTED_SYMBOLS)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(api.SlimmGpuError):
        api.SlimmGpu(np.array([1000], dtype=np.uint32), np.zeros((1, 8), dtype=np.uint32), 100, 100)


@pytest.mark.parametrize("case_name,run_name", all_runs())
def test_profile_rows_host_tail(case_name, run_name):
    case = load_case(case_name)
    run = [r for r in runs_of(case) if r.name == run_name][0]
    w = run.bin_width or case.avg_read_length
    res = oracle.run(case.ref_len, case.lineage, w, case.avg_read_length, run.cov_cut_off, case.read_id, case.ref_id,
                     case.begin_pos, run.min_reads)
    taxa = {t: (case.rank_of[t], case.name_of[t]) for t in case.rank_of}
    rows = api.profile_rows(case.ref_len, case.lineage, taxa, res.direct, res.child_pairs, res.uniq_reads_count2,
                            res.n_reads, case.avg_read_length, float(res.cut), run.abundance_cut_off, run.rank)
    lines = report.profile_lines(rows, case.lineage, case.name_of, run.rank)
    assert_profiles_match(os.path.join(run.path, "profile.tsv"), lines)


def test_tree_consistency_classification():
    """The device rank reduction is only taken for tree-consistent databases; fixtures with contigs missing from
    the database (all-zero lineages) must fall back to the general host path."""
    expected = {"synth1k": True, "lca64": True, "missing": False, "quirk": False}
    for name, want in expected.items():
        case = load_case(name)
        taxa = {t: (case.rank_of[t], case.name_of[t]) for t in case.rank_of}
        assert api.db_is_tree_consistent(case.lineage, taxa) == want, name
    # a taxon that appears on two levels, or under two parents, is not consistent
    lin = np.array([[11, 1, 2, 3, 4, 5, 6, 7], [12, 1, 2, 3, 4, 5, 6, 7]], dtype=np.uint32)
    taxa = {t: (r, "n") for r, t in enumerate([0, 1, 2, 3, 4, 5, 6, 7])}
    taxa.update({11: (0, "a"), 12: (0, "b")})
    del taxa[0]
    assert api.db_is_tree_consistent(lin, taxa)
    bad = lin.copy(); bad[1, 2] = 9
    taxa2 = dict(taxa); taxa2[9] = (2, "x")
    assert not api.db_is_tree_consistent(bad, taxa2)          # species 1 under two genera
    bad = lin.copy(); bad[1, 3] = 2
    assert not api.db_is_tree_consistent(bad, taxa)           # taxon 2 on two levels
    taxa3 = dict(taxa); taxa3[5] = (6, "wrong rank")
    assert not api.db_is_tree_consistent(lin, taxa3)


def test_traffic_json_matches_the_committed_capture(tmp_path):
    """profiles/traffic.json (what bench.py reports as roofline.traffic) is what scripts/make_traffic_json.py makes of the committed
    ncu CSV: every launch of the last complete step, per kernel group; the coverage kernel moves ~18.8 bytes per record at cfg5."""
    import json
    import shutil
    import subprocess
    import sys
    prof = os.path.join(ROOT, "profiles")
    committed = json.load(open(os.path.join(prof, "traffic.json")))
    tag = committed["capture"]
    csv5 = os.path.join(prof, f"{tag}_traffic_cfg5.csv")
    assert os.path.exists(csv5), "the capture traffic.json names is not committed"
    # run the script on a copy of the tree's profiles/ so that the tracked file stays untouched
    work = tmp_path / "repo"
    (work / "scripts").mkdir(parents=True)
    (work / "profiles").mkdir()
    shutil.copy(os.path.join(ROOT, "scripts", "make_traffic_json.py"), work / "scripts" / "make_traffic_json.py")
    r = subprocess.run([sys.executable, str(work / "scripts" / "make_traffic_json.py"), tag, f"cfg5={csv5}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    made = json.load(open(work / "profiles" / "traffic.json"))
    for group in ("coverage", "split", "accumulate", "assign"):
        assert made["cfg5"][group]["dram_bytes_per_launch"] == committed["cfg5"][group]["dram_bytes_per_launch"]
    assert 16.0 < made["cfg5"]["coverage"]["dram_bytes_per_record"] < 21.0
    assert "k_fine_accumulate_cluster" in made["cfg5"]["accumulate"]["kernel"]
