#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference binaries.

Run in the build container only (needs oracle/_ref/slimm and slimm_build, i.e. `make -C oracle ref`,
and /root/reference for the two SeqAn test files used as known-answer inputs):

    python tests/golden/make_golden.py [case ...]

For every case this writes

    <case>/records.npz      kept records as SoA (read_id by first appearance of qname+mate,
                            ref_id, begin_pos), ref_len, lineage[G,8], avg_read_length
    <case>/meta.json        contig names/accessions, taxid -> [rank, name]
    <case>/in.sam.gz        the SAM the reference consumed (small cases only)
    <case>/db.sldb          database written by the reference's slimm_build
    <case>/runs/<run>/      args.json + the reference's _profile.tsv, _raw.tsv, -v stderr
                            (+ the three coverage files for small cases)

The expected files are outputs of the reference itself; nothing in them is hand-edited.
"""
from __future__ import annotations

import gzip
import json
import os
import shutil
import struct
import subprocess
import sys
import tempfile
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from slimm_b200 import sldb, synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF = os.path.join(ROOT, "oracle", "_ref")
REFERENCE = "/root/reference"


def run_ref_build(tax_paths, out_db):
    subprocess.run([os.path.join(REF, "slimm_build"), "-nm", tax_paths["names"], "-nd", tax_paths["nodes"],
                    "-o", out_db, tax_paths["fasta"], tax_paths["acc2taxid"]], check=True,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def run_ref(case_dir, run_name, db_path, in_path, args, keep_cov):
    rd = os.path.join(case_dir, "runs", run_name)
    os.makedirs(rd, exist_ok=True)
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "out") + "/"
        os.makedirs(out)
        cmd = [os.path.join(REF, "slimm"), "-v", "-ro"] + (["-co"] if keep_cov else []) + args + ["-o", out, db_path, in_path]
        r = subprocess.run(cmd, capture_output=True, text=True, check=True)
        base = os.path.basename(in_path)
        for ext in (".sam", ".bam"):
            if base.endswith(ext):
                base = base[: -len(ext)]
        for suf in ["_profile", "_raw"] + (["_coverage", "_uniq_coverage", "_uniq_coverage2"] if keep_cov else []):
            src = os.path.join(out, base + suf + ".tsv")
            if os.path.exists(src):
                shutil.copy(src, os.path.join(rd, suf[1:] + ".tsv"))
        # drop the timing lines (whole seconds, machine dependent)
        err = "\n".join(l for l in r.stderr.splitlines() if " secs" not in l and "Total time" not in l)
        open(os.path.join(rd, "stderr.txt"), "w").write(err + "\n")
    json.dump({"args": args}, open(os.path.join(rd, "args.json"), "w"))


def save_case(case_dir, rec, contigs, db, avg_read_length, sam_path=None, keep_sam=True):
    os.makedirs(case_dir, exist_ok=True)
    lineage = db.lineage_table(contigs.accessions)
    np.savez_compressed(os.path.join(case_dir, "records.npz"), read_id=rec.read_id, ref_id=rec.ref_id,
                        begin_pos=rec.begin_pos, ref_len=contigs.lengths.astype(np.uint32), lineage=lineage,
                        avg_read_length=np.uint32(avg_read_length))
    # taxid -> (rank, name) restricted to taxa the lineage table can reach
    used = set(int(x) for x in np.unique(lineage))
    taxa = {str(t): [r, n] for t, (r, n) in db.taxid__name.items() if t in used}
    json.dump({"contig_names": contigs.names, "accessions": contigs.accessions, "taxa": taxa},
              open(os.path.join(case_dir, "meta.json"), "w"))
    if sam_path and keep_sam:
        with open(sam_path, "rb") as f, gzip.GzipFile(os.path.join(case_dir, "in.sam.gz"), "wb", mtime=0) as g:
            shutil.copyfileobj(f, g)


def db_via_reference(tax, contigs, case_dir):
    with tempfile.TemporaryDirectory() as td:
        p = synth.write_taxonomy_files(tax, contigs, td)
        os.makedirs(case_dir, exist_ok=True)
        out_db = os.path.join(case_dir, "db.sldb")
        run_ref_build(p, out_db)
    db = sldb.read_sldb(out_db)
    # the python DB builder must agree with the reference's slimm_build
    mine = synth.database_for(tax)
    assert set(mine.ac__taxid) == set(db.ac__taxid)
    assert all((mine.ac__taxid[k] == db.ac__taxid[k]).all() for k in db.ac__taxid)
    assert mine.taxid__name == db.taxid__name
    return db, out_db


# --------------------------------------------------------------------------------------------
def case_quirk():
    cd = os.path.join(GOLD, "quirk")
    rng = np.random.default_rng(20261017)
    G = 13
    missing = np.zeros(G, dtype=bool)
    missing[12] = True
    tax, accs = synth.make_taxonomy(G, missing=missing, fanout=(2, 2, 3, 1, 1, 1))
    contigs = synth.make_contigs(G, rng, accs, 20_000, 60_000, sigma=1.0)
    fx = synth.make_quirk_fixture(contigs, rng, 6000, unknown_ref=12)
    db, db_path = db_via_reference(tax, contigs, cd)
    with tempfile.TemporaryDirectory() as td:
        sam = os.path.join(td, "in.sam")
        synth.write_sam(sam, contigs, fx.qname, fx.flag, fx.ref_id, fx.pos1, seq_records=40)
        rec = synth.records_from_sam_fixture(fx)
        save_case(cd, rec, contigs, db, 100, sam)
        run_ref(cd, "default", db_path, sam, [], True)
        run_ref(cd, "cc050", db_path, sam, ["-cc", "0.5"], True)
        run_ref(cd, "cc100_w250", db_path, sam, ["-cc", "1.0", "-w", "250"], True)
        run_ref(cd, "genus_w1000", db_path, sam, ["-r", "genus", "-w", "1000"], True)
        run_ref(cd, "family_ac5", db_path, sam, ["-r", "family", "-ac", "5"], False)


def case_dup():
    cd = os.path.join(GOLD, "dup")
    rng = np.random.default_rng(7)
    G = 6
    tax, accs = synth.make_taxonomy(G, fanout=(2, 3, 1, 1, 1, 1))
    contigs = synth.make_contigs(G, rng, accs, 5_000, 9_000, sigma=0.5)
    qn, fl, rf, ps = [], [], [], []
    for r in range(640):
        g = int(rng.integers(0, G))
        n = 1 + (r % 80 == 0) * int(rng.integers(1, 4))          # planted repeats of (read, ref)
        for j in range(n):
            qn.append(f"read{r}"); fl.append(0 if j == 0 else 256); rf.append(g)
            ps.append(int(rng.integers(1, int(contigs.lengths[g]) - 100)))
        if r % 5 == 0:                                           # a second reference, then back to the first
            g2 = (g + 1) % G
            qn.append(f"read{r}"); fl.append(256); rf.append(g2); ps.append(int(rng.integers(1, 4000)))
            qn.append(f"read{r}"); fl.append(256); rf.append(g); ps.append(int(rng.integers(1, 4000)))
    fx = synth.SamFixture(qn, np.asarray(fl), np.asarray(rf), np.asarray(ps))
    db, db_path = db_via_reference(tax, contigs, cd)
    with tempfile.TemporaryDirectory() as td:
        sam = os.path.join(td, "in.sam")
        synth.write_sam(sam, contigs, fx.qname, fx.flag, fx.ref_id, fx.pos1, seq_records=10 ** 9)
        rec = synth.records_from_sam_fixture(fx)
        save_case(cd, rec, contigs, db, 100, sam)
        run_ref(cd, "default", db_path, sam, [], True)
        run_ref(cd, "w37", db_path, sam, ["-w", "37"], True)


def _synth_case(name, G, N, seed, runs, shuffle=False, **kw):
    cd = os.path.join(GOLD, name)
    rng = np.random.default_rng(seed)
    missing = rng.random(G) < kw.pop("missing_frac", 0.0)
    tax, accs = synth.make_taxonomy(G, missing=missing)
    contigs = synth.make_contigs(G, rng, accs, kw.pop("len_lo", 1_000_000), kw.pop("len_hi", 6_000_000))
    rec = synth.make_records(contigs, N, rng, shuffle=shuffle, **kw)
    db, db_path = db_via_reference(tax, contigs, cd)
    with tempfile.TemporaryDirectory() as td:
        sam = os.path.join(td, "in.sam")
        synth.write_sam_for_records(sam, contigs, rec)
        save_case(cd, rec, contigs, db, 100, sam, keep_sam=False)
        for rn, args in runs.items():
            run_ref(cd, rn, db_path, sam, args, False)


def case_synth1k():
    _synth_case("synth1k", 1000, 60_000, 12345, {"w1000": ["-w", "1000"], "w1000_genus": ["-w", "1000", "-r", "genus"]},
                len_lo=100_000, len_hi=600_000)


def case_synth1k_shuffled():
    _synth_case("synth1k_shuffled", 1000, 60_000, 12345, {"w1000": ["-w", "1000"]}, shuffle=True,
                len_lo=100_000, len_hi=600_000)


def case_lca64():
    runs = {f"cc1_{r}": ["-cc", "1.0", "-w", "1000", "-r", r] for r in ("species", "genus", "family", "order", "class", "phylum")}
    runs["cc095_species"] = ["-w", "1000"]
    _synth_case("lca64", 512, 40_000, 4242, runs, multi_frac=0.6, k_lo=2, k_hi=64, neigh=64,
                len_lo=50_000, len_hi=200_000)


def case_missing():
    # 2 % of the contigs absent from the database (all-zero lineages), unique-heavy
    _synth_case("missing", 200, 30_000, 99, {"w500": ["-w", "500"], "w500_cc08": ["-w", "500", "-cc", "0.8"]},
                missing_frac=0.02, len_lo=30_000, len_hi=90_000, multi_frac=0.3)


# --------------------------------------------------------------------------------------------
# known-answer inputs that live in the reference tree: SeqAn's yara gold SAM
# and rabema gold BAM.  Only derived arrays + the reference's outputs are stored, not the files.
def _adeno_db(case_dir):
    class _T:  # minimal taxonomy: one accession "gi" (SN gi|9632547|... cuts at '|')
        pass
    nodes = {1: (1, "no rank"), 2: (1, "superkingdom"), 10: (2, "phylum"), 20: (10, "class"), 30: (20, "order"),
             40: (30, "family"), 50: (40, "genus"), 100: (50, "species"), 1000: (100, "no rank")}
    names = {1: "root", 2: "Bacteria", 10: "Phy", 20: "Cls", 30: "Ord", 40: "Fam", 50: "GenusOne", 100: "Species A",
             1000: "A str1"}
    tax = synth.Taxonomy(nodes, names, {"gi": 1000})
    contigs = synth.Contigs(["gi|9632547|ref|NC_002077.1|"], ["gi"], np.asarray([4718], dtype=np.uint32), np.ones(1))
    with tempfile.TemporaryDirectory() as td:
        p = synth.write_taxonomy_files(tax, contigs, td)
        open(p["fasta"], "w").write(">gi|9632547|ref|NC_002077.1| adeno\nACGT\n")
        open(p["acc2taxid"], "w").write("gi\tgi.1\t1000\t0\n")
        os.makedirs(case_dir, exist_ok=True)
        run_ref_build(p, os.path.join(case_dir, "db.sldb"))
    return sldb.read_sldb(os.path.join(case_dir, "db.sldb")), contigs


def _parse_sam(path):
    qn, fl, rf, ps, seqlen = [], [], [], [], []
    names = []
    for line in open(path):
        if line.startswith("@"):
            if line.startswith("@SQ"):
                names.append([f[3:] for f in line.rstrip("\n").split("\t") if f.startswith("SN:")][0])
            continue
        f = line.rstrip("\n").split("\t")
        qn.append(f[0]); fl.append(int(f[1])); rf.append(names.index(f[2]) if f[2] != "*" else -1); ps.append(int(f[3]))
        seqlen.append(0 if f[9] == "*" else len(f[9]))
    return synth.SamFixture(qn, np.asarray(fl), np.asarray(rf), np.asarray(ps)), seqlen


def _parse_bam(path):
    raw = open(path, "rb").read()
    data = b""
    d = zlib.decompressobj(31)
    while raw:                                    # BGZF = concatenated gzip members
        data += d.decompress(raw)
        raw = d.unused_data
        d = zlib.decompressobj(31)
    assert data[:4] == b"BAM\1"
    (l_text,) = struct.unpack_from("<i", data, 4)
    off = 8 + l_text
    (n_ref,) = struct.unpack_from("<i", data, off); off += 4
    for _ in range(n_ref):
        (l_name,) = struct.unpack_from("<i", data, off); off += 4 + l_name + 4
    qn, fl, rf, ps, seqlen = [], [], [], [], []
    while off < len(data):
        (bs,) = struct.unpack_from("<i", data, off)
        refID, pos, l_read_name, mapq, bin_, n_cigar, flag, l_seq = struct.unpack_from("<iiBBHHHi", data, off + 4)
        name = data[off + 36: off + 36 + l_read_name - 1].decode()
        qn.append(name); fl.append(flag); rf.append(refID); ps.append(pos + 1); seqlen.append(l_seq)
        off += 4 + bs
    return synth.SamFixture(qn, np.asarray(fl), np.asarray(rf), np.asarray(ps)), seqlen


def _avg_len(seqlen):
    v = [s for s in seqlen if s > 0][:100000]          # reference src/misc.hpp:509-522
    return sum(v) // len(v)


def case_adeno():
    for name, src, parse in (("adeno_sam", f"{REFERENCE}/include/seqan/apps/yara/tests/gold/adeno-reads_1.t1.sam", _parse_sam),
                             ("adeno_bam", f"{REFERENCE}/include/seqan/apps/rabema/tests/gold-adeno-hamming-08.by_qname.bam", _parse_bam),
                             ("adeno_bam_coord", f"{REFERENCE}/include/seqan/apps/rabema/tests/gold-adeno-hamming-08.by_coordinate.bam", _parse_bam)):
        if not os.path.exists(src):
            print("skip", name, "(input not in the reference tree)")
            continue
        cd = os.path.join(GOLD, name)
        db, contigs = _adeno_db(cd)
        fx, seqlen = parse(src)
        rec = synth.records_from_sam_fixture(fx)
        save_case(cd, rec, contigs, db, _avg_len(seqlen), None)
        json.dump({"source": src.replace(REFERENCE + "/", ""), "n_records_in_file": len(fx.qname)},
                  open(os.path.join(cd, "source.json"), "w"))
        run_ref(cd, "default", os.path.join(cd, "db.sldb"), src, [], True)



# --------------------------------------------------------------------------------------------
# Config 1 of BASELINE.json: a REAL mapper's output (record order, flags, secondary-hit layout, strain-level
# multi-mapping) - reads sampled from the reference's own tests/example/toy-references.fa, mapped with yara (vendored with
# the reference under include/seqan/apps/yara; built per SURVEY.md Appendix B into $YARA_DIR, default /tmp/yara).
TOY_TAXA = [  # accession -> (strain name, species, genus, family, order, class, phylum)
    ("CP009656", "Borrelia burgdorferi B31", "Borrelia burgdorferi", "Borrelia", "Spirochaetaceae", "Spirochaetales", "Spirochaetia", "Spirochaetes"),
    ("CP007793", "Azospirillum brasilense Az39", "Azospirillum brasilense", "Azospirillum", "Rhodospirillaceae", "Rhodospirillales", "Alphaproteobacteria", "Proteobacteria"),
    ("CP007604", "Helicobacter pylori BM013A", "Helicobacter pylori", "Helicobacter", "Helicobacteraceae", "Campylobacterales", "Epsilonproteobacteria", "Proteobacteria"),
    ("CP007605", "Helicobacter pylori BM012B", "Helicobacter pylori", "Helicobacter", "Helicobacteraceae", "Campylobacterales", "Epsilonproteobacteria", "Proteobacteria"),
    ("CP007606", "Helicobacter pylori BM013B", "Helicobacter pylori", "Helicobacter", "Helicobacteraceae", "Campylobacterales", "Epsilonproteobacteria", "Proteobacteria"),
    ("AP014622", "Pseudomonas aeruginosa NCGM 1900", "Pseudomonas aeruginosa", "Pseudomonas", "Pseudomonadaceae", "Pseudomonadales", "Gammaproteobacteria", "Proteobacteria"),
    ("AP014646", "Pseudomonas aeruginosa NCGM 1984", "Pseudomonas aeruginosa", "Pseudomonas", "Pseudomonadaceae", "Pseudomonadales", "Gammaproteobacteria", "Proteobacteria"),
    ("CP008896", "Pseudomonas fluorescens UK4", "Pseudomonas fluorescens", "Pseudomonas", "Pseudomonadaceae", "Pseudomonadales", "Gammaproteobacteria", "Proteobacteria"),
    ("AJ344068", "Pseudomonas putida pWW0", "Pseudomonas putida", "Pseudomonas", "Pseudomonadaceae", "Pseudomonadales", "Gammaproteobacteria", "Proteobacteria"),
    ("CP007620", "Pseudomonas putida DLL-E4", "Pseudomonas putida", "Pseudomonas", "Pseudomonadaceae", "Pseudomonadales", "Gammaproteobacteria", "Proteobacteria"),
    ("CP007441", "Pseudomonas stutzeri 28a24", "Pseudomonas stutzeri", "Pseudomonas", "Pseudomonadaceae", "Pseudomonadales", "Gammaproteobacteria", "Proteobacteria"),
    ("AM920689", "Xanthomonas campestris B100", "Xanthomonas campestris", "Xanthomonas", "Xanthomonadaceae", "Xanthomonadales", "Gammaproteobacteria", "Proteobacteria"),
    ("EU186381", "Agrobacterium rhizogenes pRi2659", "Agrobacterium rhizogenes", "Agrobacterium", "Rhizobiaceae", "Rhizobiales", "Alphaproteobacteria", "Proteobacteria"),
    ("CP009144", "Sinorhizobium meliloti RMO17", "Sinorhizobium meliloti", "Sinorhizobium", "Rhizobiaceae", "Rhizobiales", "Alphaproteobacteria", "Proteobacteria"),
]


def case_toy_yara():
    import re
    yara = os.environ.get("YARA_DIR", "/tmp/yara")
    src = f"{REFERENCE}/tests/example/toy-references.fa"
    if not (os.path.exists(os.path.join(yara, "yara_mapper")) and os.path.exists(os.path.join(yara, "yara_indexer")) and os.path.exists(src)):
        print("skip toy_yara (needs yara_indexer / yara_mapper in $YARA_DIR, see SURVEY.md Appendix B)")
        return
    cd = os.path.join(GOLD, "toy_yara")
    # references: header -> "ACC.V text" (the accession must lead the name, src/misc.hpp:415-422)
    names, seqs, cur = [], [], []
    for line in open(src):
        if line.startswith(">"):
            if cur:
                seqs.append("".join(cur)); cur = []
            names.append(re.sub(r"^>gi\|[0-9]+\|[a-z]+\|([A-Z0-9_]+\.[0-9]+)\| ", r"\1 ", line.strip()))
        else:
            cur.append(line.strip().upper())
    seqs.append("".join(cur))
    assert len(names) == len(TOY_TAXA) == 14
    # taxonomy with all seven ranks for every genome (LCA = 0 reads cannot be pinned otherwise)
    nodes = {1: (1, "no rank"), 2: (1, "superkingdom")}
    tnames = {1: "root", 2: "Bacteria"}
    ids = {}

    def node(name, rank, parent):
        if (name, rank) not in ids:
            ids[(name, rank)] = 100 + len(ids)
            nodes[ids[(name, rank)]] = (parent, rank)
            tnames[ids[(name, rank)]] = name
        return ids[(name, rank)]

    acc_taxid = {}
    for acc, strain, species, genus, family, order, cls, phylum in TOY_TAXA:
        p = node(phylum, "phylum", 2); c = node(cls, "class", p); o = node(order, "order", c); f = node(family, "family", o)
        g = node(genus, "genus", f); sp = node(species, "species", g); st = node(strain, "no rank", sp)
        acc_taxid[acc] = st
    tax = synth.Taxonomy(nodes, tnames, acc_taxid)
    sn = [n.split(" ")[0] for n in names]
    contigs = synth.Contigs(sn, [x.split(".")[0] for x in sn], np.asarray([len(x) for x in seqs], dtype=np.uint32),
                            np.ones(14) / 14)
    rng = np.random.default_rng(20261018)
    comp = str.maketrans("ACGTN", "TGCAN")
    with tempfile.TemporaryDirectory() as td:
        ref_fa = os.path.join(td, "refs.fa")
        with open(ref_fa, "w") as f:
            for n, sq in zip(names, seqs):
                f.write(">" + n + "\n")
                for i in range(0, len(sq), 70):
                    f.write(sq[i:i + 70] + "\n")
        # 20 000 reads of 100 bp, genome ~ length, 1 % substitutions, half of them reverse-complemented, non-ACGT -> N
        total = sum(len(x) for x in seqs)
        with open(os.path.join(td, "reads.fa"), "w") as f:
            for r in range(20000):
                g = int(rng.choice(14, p=[len(x) / total for x in seqs]))
                s0 = int(rng.integers(0, len(seqs[g]) - 100))
                read = list(re.sub("[^ACGT]", "N", seqs[g][s0:s0 + 100]))
                for k in np.nonzero(rng.random(100) < 0.01)[0]:
                    read[k] = "ACGT"[int(rng.integers(0, 4))]
                read = "".join(read)
                if rng.random() < 0.5:
                    read = read.translate(comp)[::-1]
                f.write(f">read{r}\n{read}\n")
        subprocess.run([os.path.join(yara, "yara_indexer"), ref_fa, "-o", os.path.join(td, "idx")], check=True, stdout=subprocess.DEVNULL)
        subprocess.run([os.path.join(yara, "yara_mapper"), os.path.join(td, "idx"), os.path.join(td, "reads.fa"), "-sa", "record",
                        "-s", "2", "-t", "4", "-o", os.path.join(td, "mapped.sam")], check=True, stdout=subprocess.DEVNULL)
        # keep the mapper's records, order and flags; SEQ / QUAL only on the first 200 records (the reference samples the average
        # read length from records that carry a SEQ, src/misc.hpp:509-522) so that the fixture stays small
        sam = os.path.join(td, "in.sam")
        n = 0
        with open(os.path.join(td, "mapped.sam")) as f, open(sam, "w") as out:
            for line in f:
                if line.startswith("@"):
                    out.write(line); continue
                fl = line.rstrip("\n").split("\t")
                n += 1
                if n > 200:
                    fl[9] = "*"; fl[10] = "*"
                out.write("\t".join(fl[:11]) + "\n")      # optional tags dropped
        fx, seqlen = _parse_sam(sam)
        rec = synth.records_from_sam_fixture(fx)
        db, db_path = db_via_reference(tax, contigs, cd)
        save_case(cd, rec, contigs, db, _avg_len(seqlen), sam)
        json.dump({"source": "reads sampled from tests/example/toy-references.fa, mapped with the vendored yara (SURVEY.md Appendix B)",
                   "n_records_in_file": len(fx.qname), "kept_records": int(rec.read_id.size), "reads": int(rec.n_reads)},
                  open(os.path.join(cd, "source.json"), "w"))
        run_ref(cd, "default", db_path, sam, [], True)
        run_ref(cd, "cc050_genus", db_path, sam, ["-cc", "0.5", "-r", "genus"], False)
        run_ref(cd, "w1000", db_path, sam, ["-w", "1000"], False)

CASES = {"toy_yara": case_toy_yara, "quirk": case_quirk, "dup": case_dup, "synth1k": case_synth1k, "synth1k_shuffled": case_synth1k_shuffled,
         "lca64": case_lca64, "missing": case_missing, "adeno": case_adeno}

if __name__ == "__main__":
    for c in (sys.argv[1:] or list(CASES)):
        print("==", c)
        CASES[c]()
