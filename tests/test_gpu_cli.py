"""The drop-in command line end to end on the GPU: slimm_b200/bin/slimm (threaded decode -> C ABI -> TSV writers)
against the outputs the UNMODIFIED reference binary wrote for the same inputs (tests/golden/*/runs/*), and - when
oracle/_ref/slimm travelled to the box - against the reference binary run side by side on a fresh synthetic sample.

Comparison rules: _raw.tsv and the three coverage files are text-equal; _profile.tsv rows
are matched by taxa_id (the reference prints them in hash-iteration order), plain rows text-equal, starred rows
with exact counts and abundances within 1e-4; the -v counters on stderr are equal."""
from __future__ import annotations

import gzip
import json
import os
import subprocess

import numpy as np
import pytest

import golden_util as gu
import sam_fixtures as sf
from slimm_b200 import build as native
from slimm_b200 import sldb, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "slimm")


@pytest.fixture(scope="module", autouse=True)
def _built():
    assert os.path.exists(native.CLI), "slimm_b200/bin/slimm is missing: run `python -m slimm_b200.build`"


def sam_for_case(c: gu.Case, tmp_path, fmt="sam"):
    src = os.path.join(c.path, "in.sam.gz")
    p = tmp_path / "in.sam"
    if os.path.exists(src):
        p.write_bytes(gzip.open(src).read())
    else:
        contigs = synth.Contigs(c.contig_names, c.accessions, c.ref_len, np.ones(len(c.contig_names)))
        synth.write_sam_for_records(str(p), contigs, synth.Records(c.read_id, c.ref_id, c.begin_pos, 0))
    return str(p)


def run_cli(binary, args, db, inp, out_dir, check=True):
    os.makedirs(out_dir, exist_ok=True)
    r = subprocess.run([binary, "-v", "-ro", "-co"] + args + ["-o", out_dir + "/", db, inp], capture_output=True, text=True)
    if check:
        assert r.returncode == 0, r.stderr[-2000:]
    return r


def compare_outputs(exp_dir_or_map, got_dir, base, with_cov):
    def exp(name):
        return exp_dir_or_map[name] if isinstance(exp_dir_or_map, dict) else os.path.join(exp_dir_or_map, name + ".tsv")
    gu.assert_profiles_match(exp("profile"), open(os.path.join(got_dir, base + "_profile.tsv")).read().splitlines())
    assert open(os.path.join(got_dir, base + "_raw.tsv")).read() == open(exp("raw")).read(), "_raw.tsv differs"
    if with_cov:
        for suf in ("coverage", "uniq_coverage", "uniq_coverage2"):
            assert open(os.path.join(got_dir, f"{base}_{suf}.tsv")).read() == open(exp(suf)).read(), f"_{suf}.tsv differs"


def stderr_stats(text, tmp_path):
    p = tmp_path / "stderr.txt"
    p.write_text(text)
    return gu.parse_stderr_stats(str(p))


@pytest.mark.parametrize("case_name,run_name", [(c, r) for c, r in gu.all_runs() if not c.startswith("adeno")])
def test_cli_reproduces_reference_outputs(case_name, run_name, tmp_path):
    c = gu.load_case(case_name)
    run = [r for r in gu.runs_of(c) if r.name == run_name][0]
    args = json.load(open(os.path.join(run.path, "args.json")))["args"]
    inp = sam_for_case(c, tmp_path)
    out = str(tmp_path / "out")
    r = run_cli(native.CLI, args, os.path.join(c.path, "db.sldb"), inp, out)
    compare_outputs(run.path, out, "in", os.path.exists(os.path.join(run.path, "coverage.tsv")))
    exp = gu.parse_stderr_stats(os.path.join(run.path, "stderr.txt"))
    got = stderr_stats(r.stderr, tmp_path)
    assert {k: got.get(k) for k in exp} == exp
    # the row count line of write_abundance (-v): "<rows> <rank> (<failed> bellow cutoff ..."
    want = [l for l in open(os.path.join(run.path, "stderr.txt")).read().splitlines() if "bellow cutoff" in l]
    assert want and want[0].strip() in [l.strip() for l in r.stderr.splitlines()]


def test_cli_bam_input_and_default_output_names(tmp_path):
    """BAM through BGZF inflate threads; without -o the outputs sit next to the input as <input>_profile.tsv."""
    c = gu.load_case("dup")
    lines = gzip.open(os.path.join(c.path, "in.sam.gz"), "rt").read().splitlines()
    body = [l.split("\t") for l in lines if not l.startswith("@")]
    qn = [f[0] for f in body]
    flag = np.asarray([int(f[1]) for f in body])
    ref = np.asarray([c.contig_names.index(f[2]) if f[2] != "*" else -1 for f in body])
    pos1 = np.asarray([int(f[3]) for f in body])
    bam = tmp_path / "sample.bam"
    bam.write_bytes(sf.bgzf_compress(sf.bam_bytes(c.contig_names, c.ref_len, qn, flag, ref, pos1, seq_every=1)))
    r = subprocess.run([native.CLI, "-v", "-ro", os.path.join(c.path, "db.sldb"), str(bam)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    run = [x for x in gu.runs_of(c) if x.name == "default"][0]
    gu.assert_profiles_match(os.path.join(run.path, "profile.tsv"), open(str(bam) + "_profile.tsv").read().splitlines())
    assert open(str(bam) + "_raw.tsv").read() == open(os.path.join(run.path, "raw.tsv")).read()


def test_cli_directory_mode(tmp_path):
    """-d: every .sam / .bam of the directory gets its own profile under the output directory."""
    d = tmp_path / "samples"
    d.mkdir()
    c = gu.load_case("dup")
    for name in ("a.sam", "b.sam"):
        (d / name).write_bytes(gzip.open(os.path.join(c.path, "in.sam.gz")).read())
    (d / "notes.txt").write_text("not an alignment file")
    out = tmp_path / "out"
    out.mkdir()
    # a third sample against the same database with OTHER @SQ lines (fewer contigs): the context cannot be reused for it
    q = gu.load_case("quirk")
    (d / "c.sam").write_bytes(gzip.open(os.path.join(q.path, "in.sam.gz")).read())
    r = subprocess.run([native.CLI, "-v", "-d", "-o", str(out) + "/", os.path.join(c.path, "db.sldb"), str(d)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    run = [x for x in gu.runs_of(c) if x.name == "default"][0]
    for base in ("a", "b"):
        gu.assert_profiles_match(os.path.join(run.path, "profile.tsv"), open(out / f"{base}_profile.tsv").read().splitlines())
    assert os.path.exists(out / "c_profile.tsv")
    # one GPU context per set of @SQ lines: a.sam creates it, b.sam reuses it (slimm_gpu_reset), c.sam needs a new one
    assert r.stderr.count("(context reused)") == 1, r.stderr[-3000:]


@pytest.mark.skipif(not os.path.exists(REF_BIN), reason="oracle/_ref/slimm was not built (needs /root/reference at build time)")
@pytest.mark.parametrize("args", [[], ["-w", "500", "-r", "genus"], ["-cc", "1.0", "-w", "2000", "-r", "family", "-ac", "0.5"]])
def test_cli_against_reference_binary_side_by_side(args, tmp_path):
    """A fresh synthetic sample (not among the committed goldens), the reference binary and this command line on the
    same SAM and database, every output file compared."""
    rng = np.random.default_rng(len(args) + 2026)
    G = 600
    missing = rng.random(G) < 0.01
    tax, accs = synth.make_taxonomy(G, missing=missing)
    contigs = synth.make_contigs(G, rng, accs, 40_000, 200_000)
    rec = synth.make_records(contigs, 150_000, rng, multi_frac=0.35, k_lo=2, k_hi=12, neigh=12)
    sam = str(tmp_path / "fresh.sam")
    synth.write_sam_for_records(sam, contigs, rec)
    db = str(tmp_path / "db.sldb")
    sldb.write_sldb(synth.database_for(tax), db)
    ref_out, got_out = str(tmp_path / "ref"), str(tmp_path / "got")
    rr = run_cli(REF_BIN, args, db, sam, ref_out)
    rg = run_cli(native.CLI, args, db, sam, got_out)
    exp = {n: os.path.join(ref_out, f"fresh_{n}.tsv") for n in ("profile", "raw", "coverage", "uniq_coverage", "uniq_coverage2")}
    compare_outputs(exp, got_out, "fresh", True)
    a, b = stderr_stats(rr.stderr, tmp_path), stderr_stats(rg.stderr, tmp_path)
    assert a == b and a["hits"] == "150000"


def test_cli_two_gpus_give_the_single_gpu_profile(tmp_path):
    """`slimm --gpus 2` (one host process, reads sharded over two devices, slimm_gpu_run_sharded_local): same _profile.tsv and
    the same -v counters as one GPU."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(4)
    G = 1500
    tax, accs = synth.make_taxonomy(G)
    contigs = synth.make_contigs(G, rng, accs, 200_000, 900_000)
    rec = synth.make_records(contigs, 1_200_000, rng, multi_frac=0.35, k_lo=2, k_hi=40, neigh=12)
    sam = str(tmp_path / "s.sam")
    synth.write_sam_for_records(sam, contigs, rec)
    db = str(tmp_path / "db.sldb")
    sldb.write_sldb(synth.database_for(tax), db)
    outs = {}
    for n in (1, 2):
        out = str(tmp_path / f"out{n}")
        os.makedirs(out)
        r = subprocess.run([native.CLI, "-v", "-w", "10", "--gpus", str(n), "-o", out + "/", db, sam], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[n] = (open(os.path.join(out, "s_profile.tsv")).read(), stderr_stats(r.stderr, tmp_path))
    assert outs[1][0] == outs[2][0]
    assert outs[1][1] == outs[2][1] and outs[1][1]["hits"] == "1200000"
