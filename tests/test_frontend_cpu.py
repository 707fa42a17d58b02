"""Host side of the drop-in command line (slimm_b200/bin/slimm), no GPU needed: the threaded SAM / BAM decoder
against the golden record arrays and against an independent Python statement of the reference's record loop
(reference src/slimm.hpp:194-213, src/misc.hpp:509-522), and the argument handling of src/slimm.cpp:60-180."""
from __future__ import annotations

import gzip
import os
import subprocess

import numpy as np
import pytest

import golden_util as gu
import sam_fixtures as sf
from slimm_b200 import build as native
from slimm_b200 import synth

CLI = native.CLI


@pytest.fixture(scope="module", autouse=True)
def _built():
    native.build()
    assert os.path.exists(CLI)


def decode(path, tmp_path, threads=4, extra=()):
    out = str(tmp_path / "dump.bin")
    r = subprocess.run([CLI, "--threads", str(threads), "--dump-records", out, *extra, "none.sldb", str(path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return sf.load_dump(out)


def assert_records(d, rec, names, lengths, avg):
    assert d["names"] == list(names)
    assert (d["ref_len"] == np.asarray(lengths, dtype=np.uint32)).all()
    assert d["avg"] == avg
    assert d["N"] == rec.read_id.size
    assert (d["read_id"] == rec.read_id).all() and (d["ref_id"] == rec.ref_id).all() and (d["begin_pos"] == rec.begin_pos).all()


@pytest.mark.parametrize("case", [c for c in gu.case_names() if os.path.exists(os.path.join(gu.GOLD, c, "in.sam.gz"))])
@pytest.mark.parametrize("packed", ["gz", "plain"])
def test_decoder_matches_golden_records(case, packed, tmp_path):
    c = gu.load_case(case)
    src = os.path.join(c.path, "in.sam.gz")
    if packed == "plain":
        p = tmp_path / "in.sam"
        p.write_bytes(gzip.open(src).read())
        src = str(p)
    d = decode(src, tmp_path)
    assert d["names"] == c.contig_names and d["avg"] == c.avg_read_length
    assert (d["read_id"] == c.read_id).all() and (d["ref_id"] == c.ref_id).all() and (d["begin_pos"] == c.begin_pos).all()


def _fixture(n_frag, seed, shuffle=False):
    """Paired and single reads, secondary hits, unmapped mates, POS 0, reads that reappear later in the file."""
    rng = np.random.default_rng(seed)
    G = 37
    names = [f"NC_{i:06d}.1 contig {i}" if i % 3 else f"gi|{i}|ref|NC_{i:06d}.1|" for i in range(G)]
    lengths = rng.integers(5_000, 90_000, G)
    qn, fl, rf, ps = [], [], [], []
    for r in range(n_frag):
        kind = r % 7
        g = int(rng.integers(0, G))
        q = f"frag{r}"
        if kind < 3:                         # single-end read with 1..3 hits
            for j in range(1 + int(rng.integers(0, 3))):
                qn.append(q); fl.append(0 if j == 0 else 256); rf.append((g + j) % G); ps.append(int(rng.integers(0, 4000)))
        elif kind < 5:                       # proper pair, both mapped: keys q.1 / q.2
            qn += [q, q]; fl += [0x1 | 0x2 | 0x40, 0x1 | 0x2 | 0x80]; rf += [g, g]; ps += [int(rng.integers(1, 4000)), int(rng.integers(1, 4000))]
        elif kind == 5:                      # pair with an unmapped mate (flag 4, RNAME set as mappers do)
            qn += [q, q]; fl += [0x1 | 0x40 | 0x8, 0x1 | 0x80 | 0x4]; rf += [g, g]; ps += [int(rng.integers(1, 4000))] * 2
        else:                                # unmapped, no reference
            qn.append(q); fl.append(4); rf.append(-1); ps.append(0)
    if n_frag > 50:                          # a name seen long before: "frag3" again at the end
        qn.append("frag3"); fl.append(256); rf.append(5); ps.append(77)
    fx = synth.SamFixture(qn, np.asarray(fl), np.asarray(rf), np.asarray(ps))
    if shuffle:
        perm = rng.permutation(len(qn))
        fx = synth.SamFixture([qn[i] for i in perm], fx.flag[perm], fx.ref_id[perm], fx.pos1[perm])
    return names, lengths, fx


@pytest.mark.parametrize("fmt", ["sam", "sam_crlf_noeol", "sam.gz", "sam.bgzf", "bam"])
@pytest.mark.parametrize("shuffle", [False, True])
def test_decoder_formats_and_chunk_boundaries(fmt, shuffle, tmp_path):
    # ~26 MB of SAM text / ~17 MB of BAM: several 4 MB pipeline chunks and hundreds of BGZF blocks
    names, lengths, fx = _fixture(70_000, 11, shuffle)
    rec = synth.records_from_sam_fixture(fx)
    if fmt == "bam":
        p = tmp_path / "in.bam"
        p.write_bytes(sf.bgzf_compress(sf.bam_bytes(names, lengths, fx.qname, fx.flag, fx.ref_id, fx.pos1), level=1))
    else:
        txt = sf.sam_text(names, lengths, fx.qname, fx.flag, fx.ref_id, fx.pos1, crlf="crlf" in fmt,
                          trailing_newline="noeol" not in fmt).encode()
        p = tmp_path / ("in.sam" if fmt.startswith("sam") and "." not in fmt else "in." + fmt)
        p.write_bytes(gzip.compress(txt, 1) if fmt == "sam.gz" else sf.bgzf_compress(txt, level=1) if fmt == "sam.bgzf" else txt)
    for threads in (1, 6):
        d = decode(p, tmp_path, threads)
        assert_records(d, rec, names, lengths, 100)
        assert d["n_records"] == len(fx.qname) and d["n_reads"] == rec.n_reads


@pytest.mark.parametrize("fmt", ["sam", "bam"])
def test_grouped_input_fast_path_gives_the_table_ids(fmt, tmp_path):
    """Input grouped by read (what mappers write): ids by counting runs, verified in parallel by the parse workers, are the
    ids of the exact read-name table; a name that comes back makes the decoder fall back to that table (same ids again)."""
    names, lengths, fx = _fixture(70_000, 13)
    assert fx.qname[-1] == "frag3"
    for grouped in (True, False):
        q, fl, rf, ps = (fx.qname[:-1], fx.flag[:-1], fx.ref_id[:-1], fx.pos1[:-1]) if grouped else (fx.qname, fx.flag, fx.ref_id, fx.pos1)
        rec = synth.records_from_sam_fixture(synth.SamFixture(list(q), np.asarray(fl), np.asarray(rf), np.asarray(ps)))
        if fmt == "bam":
            p = tmp_path / "in.bam"
            p.write_bytes(sf.bgzf_compress(sf.bam_bytes(names, lengths, q, fl, rf, ps), level=1))
        else:
            p = tmp_path / "in.sam"
            p.write_bytes(sf.sam_text(names, lengths, q, fl, rf, ps).encode())
        fast = decode(p, tmp_path, 6)
        exact = decode(p, tmp_path, 6, extra=("--exact-ids",))
        for d in (fast, exact):
            assert_records(d, rec, names, lengths, 100)
            assert d["n_records"] == len(q) and d["n_reads"] == rec.n_reads


def test_avg_read_length_uses_first_records_with_seq(tmp_path):
    # records without SEQ are skipped, unmapped ones count (src/misc.hpp:509-522); integer division
    names, lengths = ["c1"], [10_000]
    qn = [f"r{i}" for i in range(6)]
    lines = ["@SQ\tSN:c1\tLN:10000"]
    seqs = ["A" * 10, "*", "A" * 11, "A" * 12, "*", "A" * 20]
    flags = [0, 0, 4, 0, 0, 0]
    for q, s, f in zip(qn, seqs, flags):
        lines.append(f"{q}\t{f}\tc1\t5\t60\t*\t*\t0\t0\t{s}\t*")
    p = tmp_path / "in.sam"
    p.write_text("\n".join(lines) + "\n")
    d = decode(p, tmp_path)
    assert d["avg"] == (10 + 11 + 12 + 20) // 4 and d["N"] == 5


def test_unknown_reference_name_is_an_error(tmp_path):
    p = tmp_path / "in.sam"
    p.write_text("@SQ\tSN:c1\tLN:100\nr1\t0\tc2\t5\t60\t*\t*\t0\t0\tACGT\t*\n")
    r = subprocess.run([CLI, "--dump-records", str(tmp_path / "d.bin"), "x.sldb", str(p)], capture_output=True, text=True)
    assert r.returncode == 1 and "not in the header" in r.stderr


@pytest.mark.parametrize("args, code, needle", [
    ([], 1, "Not enough arguments"),
    (["db.sldb"], 1, "Not enough arguments"),
    (["db.txt", "in.sam"], 1, "valid file extensions"),
    (["-cc", "1.5", "db.sldb", "in.sam"], 1, "not in the interval"),
    (["-ac", "11", "db.sldb", "in.sam"], 1, "not in the interval"),
    (["-r", "kingdom", "db.sldb", "in.sam"], 1, "allowed values"),
    (["-w", "abc", "db.sldb", "in.sam"], 1, "cannot be casted"),
    (["--nope", "db.sldb", "in.sam"], 1, "illegal option"),
    (["db.sldb", "/nonexistent/in.sam"], 1, "is not a file use -d option"),
    (["--help"], 0, "SYNOPSIS"),
    (["--version"], 0, "0.3.4"),
])
def test_command_line_errors(args, code, needle):
    r = subprocess.run([CLI] + args, capture_output=True, text=True)
    assert r.returncode == code
    assert needle in r.stderr + r.stdout


def test_fallback_to_the_name_table_stops_the_first_pass_at_once(tmp_path):
    """A read name that comes back early in a long gzipped SAM (the usual non-grouped input is a coordinate-sorted file): the
    grouped-input attempt must stop reading and inflating there, not run to the end of the file, so the default costs about
    one pass - not the two it cost while the framer kept going after the consumer had given up."""
    import time
    n = 700_000
    lines = ["@SQ\tSN:c1\tLN:1000000"]
    lines += [f"r{i}\t0\tc1\t{1 + i % 900000}\t60\t100M\t*\t0\t0\t{'ACGT' * 25}\t{'I' * 100}" for i in range(n)]
    lines.insert(3, "r0\t256\tc1\t77\t60\t100M\t*\t0\t0\t*\t*")      # the third record repeats the first one's name
    p = tmp_path / "in.sam.gz"
    p.write_bytes(gzip.compress(("\n".join(lines) + "\n").encode(), 1))

    def best(extra):
        t = []
        for _ in range(3):
            t0 = time.perf_counter()
            d = decode(p, tmp_path, 4, extra=extra)
            t.append(time.perf_counter() - t0)
            assert d["N"] == n + 1 and d["n_reads"] == n
        return min(t)

    exact, default = best(("--exact-ids",)), best(())
    assert default < 1.45 * exact + 0.25, (default, exact)


def test_batched_parser_equals_line_parser(tmp_path):
    """The pipeline's chunk parser (eight lines at a time, prefetched contig look-ups, SWAR number parsing) gives what the line
    parser gives - records, counters and error text - on random pieces of a SAM file, most of them damaged at random
    (tests/cpp/parser_equiv.cpp)."""
    rng = np.random.default_rng(11)
    tax, accs = synth.make_taxonomy(3000)
    contigs = synth.make_contigs(3000, rng, accs, 200_000, 900_000)
    rec = synth.make_records(contigs, 300_000, rng, multi_frac=0.3)
    sam = tmp_path / "in.sam"
    synth.write_sam_for_records(str(sam), contigs, rec)
    exe = tmp_path / "parser_equiv"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = os.path.join(root, "tests", "cpp", "parser_equiv.cpp")
    inc = os.path.join(root, "slimm_b200", "csrc", "frontend")
    r = subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", inc, src, "-lz", "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), str(sam)], capture_output=True, text=True)
    assert r.returncode == 0 and "bad 0 bad_numbers 0" in r.stdout, r.stdout + r.stderr
