"""Seeded synthetic SLIMM inputs: taxonomy + database, contigs, alignment records, SAM text.

One generator feeds the reference binary (text SAM + ``.sldb``), the CPU oracle and the GPU
path (struct-of-arrays ``read_id | ref_id | begin_pos``) so all three see the same records
(SURVEY.md section 8(d)).  Shapes follow BASELINE.json's configs:

* 8-level taxonomy, fan-out 4/4/4/2/2/2 from strain to phylum, 2 superkingdoms, every genome its
  own "no rank" strain taxid;
* contig lengths U[1 Mbp, 6 Mbp], community weights lognormal(0, 2), read length 100;
* a read's primary reference ~ weights, the extra references of a multi-mapped read come from
  the +-``neigh`` reference-index neighbourhood (same genus/family mostly);
* records of a read are contiguous (mapper order); ``shuffle=True`` gives a shuffled variant;
* ``repeat_frac`` of records are followed by a planted repeat hit of the same (read, ref).
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from .sldb import SlimmDatabase, build_db_from_taxonomy

# taxid bases per level (disjoint ranges, so a taxid never appears at two ranks)
_SK_IDS = (2, 2157)
_BASE = {6: 1_000, 5: 3_000, 4: 6_000, 3: 10_000, 2: 20_000, 1: 100_000, 0: 1_000_000}
_FANOUT = (4, 4, 4, 2, 2, 2)            # strain->species, species->genus, ... class->phylum
_RANK_STR = {1: "species", 2: "genus", 3: "family", 4: "order", 5: "class", 6: "phylum",
             7: "superkingdom"}


@dataclass
class Taxonomy:
    nodes: Dict[int, Tuple[int, str]]      # taxid -> (parent, rank string)  (nodes.dmp)
    names: Dict[int, str]                  # taxid -> scientific name        (names.dmp)
    acc_taxid: Dict[str, int]              # accession -> taxid              (accession2taxid)


@dataclass
class Contigs:
    names: List[str]                       # @SQ SN (accession.version)
    accessions: List[str]                  # SN cut at first whitespace / '.' / '|'
    lengths: np.ndarray                    # u32 [G]
    weights: np.ndarray                    # f64 [G], sums to 1


@dataclass
class Records:
    """Kept (mapped) alignment records in file order, as the GPU path consumes them."""
    read_id: np.ndarray                    # u32 [N]  dense id of qname(+mate suffix)
    ref_id: np.ndarray                     # u32 [N]
    begin_pos: np.ndarray                  # i32 [N]  0-based (SAM POS-1)
    n_reads: int


def make_taxonomy(n_genomes: int, missing: Optional[np.ndarray] = None,
                  extra_nodes: int = 0, seed: int = 0,
                  fanout: Tuple[int, ...] = _FANOUT) -> Tuple[Taxonomy, List[str]]:
    """Balanced taxonomy over ``n_genomes`` genomes.  ``missing`` marks genomes whose accession
    is absent from accession2taxid (their contigs get an all-zero lineage at run time).
    ``extra_nodes`` interleaves that many unranked "clade" nodes between species and genus; they
    vanish when the database is built (config 4's 2 M-node taxonomy)."""
    nodes: Dict[int, Tuple[int, str]] = {1: (1, "no rank")}
    names: Dict[int, str] = {1: "root"}
    idx = np.arange(n_genomes)
    level_idx = [idx]
    for f in fanout:
        level_idx.append(level_idx[-1] // f)
    n_phyla = int(level_idx[6].max()) + 1 if n_genomes else 0
    for i, t in enumerate(_SK_IDS):
        nodes[t] = (1, "superkingdom")
        names[t] = ("Bacteria", "Archaea")[i]

    def sk_of_phylum(p):
        return _SK_IDS[0] if 2 * p < max(n_phyla, 1) else _SK_IDS[1]

    label = {6: "Phylum", 5: "Class", 4: "Order", 3: "Family", 2: "Genus", 1: "Species"}
    clade_next = 5_000_000
    rng = np.random.default_rng(seed)
    clade_species = set()
    if extra_nodes:
        n_species = int(level_idx[1].max()) + 1
        # spread the extra nodes as chains above randomly chosen species
        chain_len = np.bincount(rng.integers(0, n_species, size=extra_nodes), minlength=n_species)
    for lvl in range(6, 0, -1):
        ids = np.unique(level_idx[lvl])
        for i in ids:
            t = _BASE[lvl] + int(i)
            if lvl == 6:
                parent = sk_of_phylum(int(i))
            else:
                parent = _BASE[lvl + 1] + int(i) // fanout[lvl]
            if lvl == 1 and extra_nodes and chain_len[int(i)]:
                for _ in range(int(chain_len[int(i)])):
                    nodes[clade_next] = (parent, "no rank")
                    names[clade_next] = f"clade_{clade_next}"
                    parent = clade_next
                    clade_next += 1
            nodes[t] = (parent, _RANK_STR[lvl])
            names[t] = f"{label[lvl]}_{int(i)}"
    acc_taxid: Dict[str, int] = {}
    accessions: List[str] = []
    for g in range(n_genomes):
        t = _BASE[0] + g
        nodes[t] = (_BASE[1] + g // fanout[0], "no rank")
        names[t] = f"Species_{g // fanout[0]} str{g}"
        acc = f"ACC{g:07d}"
        accessions.append(acc)
        if missing is None or not missing[g]:
            acc_taxid[acc] = t
    return Taxonomy(nodes, names, acc_taxid), accessions


def make_contigs(n_genomes: int, rng: np.random.Generator, accessions: List[str],
                 len_lo: int = 1_000_000, len_hi: int = 6_000_000, sigma: float = 2.0) -> Contigs:
    lengths = rng.integers(len_lo, len_hi + 1, size=n_genomes).astype(np.uint32)
    w = rng.lognormal(0.0, sigma, size=n_genomes)
    w /= w.sum()
    return Contigs([a + ".1" for a in accessions], list(accessions), lengths, w)


def make_records(contigs: Contigs, n_records: int, rng: np.random.Generator,
                 multi_frac: float = 0.2, k_lo: int = 2, k_hi: int = 8, neigh: int = 8,
                 repeat_frac: float = 0.002, read_len: int = 100, shuffle: bool = False,
                 neigh_mode: str = "index") -> Records:
    """Records of a read are contiguous and read ids ascend in file order unless ``shuffle``.
    ``neigh_mode`` "taxonomy": a multi-mapped read draws a level (species 40 %, genus 30 %, family 15 %, order 8 %,
    class 4 %, phylum 3 %) and its extra targets uniformly among the genomes sharing that taxon with the primary one
    (fan-out of ``make_taxonomy``), so that LCAs land on every rank."""
    G = len(contigs.lengths)
    mean_k = (1.0 - multi_frac) + multi_frac * 0.5 * (k_lo + k_hi)
    n_reads = max(1, int(np.ceil(n_records / (mean_k * (1.0 + repeat_frac)))) + 16)
    cdf = np.cumsum(contigs.weights)
    cdf[-1] = 1.0
    while True:
        primary = np.searchsorted(cdf, rng.random(n_reads), side="right").astype(np.int64)
        primary = np.minimum(primary, G - 1)
        k = np.where(rng.random(n_reads) < multi_frac,
                     rng.integers(k_lo, k_hi + 1, size=n_reads), 1).astype(np.int64)
        rep = rng.random(int(k.sum())) < repeat_frac
        if int(k.sum()) + int(rep.sum()) >= n_records:
            break
        n_reads = int(n_reads * 1.05) + 16
    read_of = np.repeat(np.arange(n_reads, dtype=np.int64), k)
    start = np.cumsum(k) - k
    j = np.arange(read_of.size, dtype=np.int64) - start[read_of]
    if neigh_mode == "taxonomy":
        sizes = np.array([4, 16, 64, 128, 256, 512], dtype=np.int64)
        lvl = np.minimum(np.searchsorted(np.array([0.40, 0.70, 0.85, 0.93, 0.97, 1.0]), rng.random(n_reads)), 5)
        size = sizes[lvl][read_of]
        pick = np.minimum((rng.random(read_of.size) * size).astype(np.int64), size - 1)
        ref = np.where(j == 0, primary[read_of], np.clip((primary[read_of] // size) * size + pick, 0, G - 1))
    else:
        off = rng.integers(1, neigh + 1, size=read_of.size) * rng.choice((-1, 1), size=read_of.size)
        ref = np.where(j == 0, primary[read_of], np.clip(primary[read_of] + off, 0, G - 1))
    # planted repeat hits: the repeated record directly follows the original
    times = 1 + rep.astype(np.int64)
    read_of = np.repeat(read_of, times)
    ref = np.repeat(ref, times)
    read_of, ref = read_of[:n_records], ref[:n_records]
    span = np.maximum(contigs.lengths[ref].astype(np.int64) - read_len, 1)
    pos1 = 1 + (rng.random(ref.size) * span).astype(np.int64)          # SAM POS, 1-based
    # re-densify read ids (the tail cut may have dropped whole reads)
    _, read_id = np.unique(read_of, return_inverse=True)
    n_reads = int(read_id.max()) + 1 if read_id.size else 0
    rec = Records(read_id.astype(np.uint32), ref.astype(np.uint32),
                  (pos1 - 1).astype(np.int32), n_reads)
    if shuffle:
        p = rng.permutation(rec.read_id.size)
        rec = Records(rec.read_id[p], rec.ref_id[p], rec.begin_pos[p], n_reads)
    return rec


# ------------------------------------------------------------------------------------------
# text emitters for the reference binaries
# ------------------------------------------------------------------------------------------

def write_taxonomy_files(tax: Taxonomy, contigs: Contigs, out_dir: str) -> Dict[str, str]:
    """nodes.dmp / names.dmp / accession2taxid / FASTA stub in the layouts slimm_build parses
    (reference src/slimm_build.cpp:151-190,295-322)."""
    os.makedirs(out_dir, exist_ok=True)
    p = {k: os.path.join(out_dir, v) for k, v in
         dict(nodes="nodes.dmp", names="names.dmp", acc2taxid="acc2taxid.tsv", fasta="refs.fa").items()}
    with open(p["nodes"], "w") as f:
        for t, (parent, rank) in tax.nodes.items():
            f.write(f"{t}\t|\t{parent}\t|\t{rank}\t|\t\t|\n")
    with open(p["names"], "w") as f:
        for t, name in tax.names.items():
            f.write(f"{t}\t|\t{name}\t|\t\t|\tscientific name\t|\n")
    with open(p["acc2taxid"], "w") as f:
        f.write("accession\taccession.version\ttaxid\tgi\n")
        for acc, t in tax.acc_taxid.items():
            f.write(f"{acc}\t{acc}.1\t{t}\t0\n")
    with open(p["fasta"], "w") as f:
        for name in contigs.names:
            f.write(f">{name} synthetic\nACGT\n")
    return p


def database_for(tax: Taxonomy) -> SlimmDatabase:
    return build_db_from_taxonomy(tax.acc_taxid, tax.nodes, tax.names)


def write_sam(path: str, contigs: Contigs, qname: List[str], flag: np.ndarray, ref_id: np.ndarray,
              pos1: np.ndarray, read_len: int = 100, seq_records: int = 64) -> None:
    """Plain SAM.  Only the first ``seq_records`` records carry SEQ (the reference samples the
    average read length from records that have one, src/misc.hpp:509-522); the rest use ``*`` to
    keep fixtures small.  ``ref_id`` -1 writes an unmapped record (RNAME ``*``)."""
    seq = "ACGT" * (read_len // 4) + "ACGT"[: read_len % 4]
    with open(path, "w") as f:
        f.write("@HD\tVN:1.4\tSO:unsorted\n")
        for n, ln in zip(contigs.names, contigs.lengths):
            f.write(f"@SQ\tSN:{n}\tLN:{int(ln)}\n")
        lines = []
        for i in range(len(qname)):
            s = seq if i < seq_records else "*"
            g = int(ref_id[i])
            if g < 0:
                lines.append(f"{qname[i]}\t{int(flag[i])}\t*\t0\t0\t*\t*\t0\t0\t{s}\t*\n")
            else:
                lines.append(f"{qname[i]}\t{int(flag[i])}\t{contigs.names[g]}\t{int(pos1[i])}\t60\t"
                             f"{read_len}M\t*\t0\t0\t{s}\t*\n")
            if len(lines) >= 65536:
                f.write("".join(lines))
                lines = []
        f.write("".join(lines))


def write_sam_for_records(path: str, contigs: Contigs, rec: Records, read_len: int = 100,
                          seq_records: int = 64) -> None:
    """SAM twin of a ``Records`` SoA: qname ``r<read_id>``, flag 0 (256 for a read's later hits)."""
    rid = rec.read_id
    first = np.ones(rid.size, dtype=bool)
    if rid.size:
        first[1:] = rid[1:] != rid[:-1]
    flag = np.where(first, 0, 256)
    qname = [f"r{int(r)}" for r in rid]
    write_sam(path, contigs, qname, flag, rec.ref_id.astype(np.int64),
              rec.begin_pos.astype(np.int64) + 1, read_len, seq_records)


# ------------------------------------------------------------------------------------------
# SAM-level fixtures (qnames, flags, mates, unmapped records) for decoder / quirk coverage
# ------------------------------------------------------------------------------------------

@dataclass
class SamFixture:
    qname: List[str]
    flag: np.ndarray        # SAM FLAG
    ref_id: np.ndarray      # i64, -1 = unmapped (RNAME '*')
    pos1: np.ndarray        # SAM POS (1-based; 0 allowed)


def records_from_sam_fixture(fx: SamFixture) -> Records:
    """The SoA the hot path sees for a SAM: drop records with flag&4 or no reference
    (reference src/slimm.hpp:197), key reads by qname + ".1"/".2" for first/last mates
    (:204-208), ids in order of first appearance."""
    ids: Dict[str, int] = {}
    rid, ref, pos = [], [], []
    for i, q in enumerate(fx.qname):
        fl = int(fx.flag[i])
        if (fl & 4) or fx.ref_id[i] < 0:
            continue
        key = q + (".1" if fl & 0x40 else ".2" if fl & 0x80 else "")
        rid.append(ids.setdefault(key, len(ids)))
        ref.append(int(fx.ref_id[i]))
        pos.append(int(fx.pos1[i]) - 1)
    return Records(np.asarray(rid, dtype=np.uint32), np.asarray(ref, dtype=np.uint32),
                   np.asarray(pos, dtype=np.int32), len(ids))


def make_quirk_fixture(contigs: Contigs, rng: np.random.Generator, n_fragments: int = 6000,
                       strains_per_species: int = 2, unknown_ref: Optional[int] = None) -> SamFixture:
    """Fragments over a small community that exercise the reference's load-bearing quirks
    (SURVEY.md section 7): 1/3 paired (mates are separate reads), 30 % sibling-strain multi-hits,
    10 % a random third hit, 3 % spurious unique hits anywhere, planted repeat (read, ref) hits,
    unmapped records, one POS=0 record (beginPos -1 wraps in u32) and hits on a contig that is
    missing from the database."""
    G = len(contigs.lengths)
    qn: List[str] = []
    fl: List[int] = []
    rf: List[int] = []
    ps: List[int] = []

    def emit(q, f, g, p):
        qn.append(q); fl.append(f); rf.append(g); ps.append(p)

    cdf = np.cumsum(contigs.weights)
    for f in range(n_fragments):
        q = f"frag{f}"
        paired = rng.random() < 1 / 3
        if rng.random() < 0.03:
            emit(q, 4, -1, 0)                                   # unmapped record
            continue
        for mate in ((1, 2) if paired else (0,)):
            base = {0: 0, 1: 0x41, 2: 0x81}[mate]
            if rng.random() < 0.03:
                g = int(rng.integers(0, G))                     # spurious unique
            else:
                g = min(int(np.searchsorted(cdf, rng.random(), side="right")), G - 1)
            hits = [g]
            if rng.random() < 0.30:                             # sibling strain
                sib = g ^ 1 if strains_per_species == 2 else (g // strains_per_species) * strains_per_species
                if sib < G and sib != g:
                    hits.append(sib)
            if rng.random() < 0.10:
                hits.append(int(rng.integers(0, G)))            # random third hit
            if unknown_ref is not None and rng.random() < 0.02:
                hits.append(unknown_ref)
            for j, h in enumerate(hits):
                L = int(contigs.lengths[h])
                p = int(rng.integers(1, max(L - 100, 2)))
                emit(q, base | (0x100 if j else 0), h, p)
                if rng.random() < 0.01:                         # repeat hit, other position
                    emit(q, base | 0x100, h, int(rng.integers(1, max(L - 100, 2))))
    emit("fragPOS0", 0, 0, 0)                                   # beginPos = -1
    emit("fragEND", 0, G - 1, int(contigs.lengths[G - 1]))      # centre clamps to the contig length
    return SamFixture(qn, np.asarray(fl, dtype=np.int64), np.asarray(rf, dtype=np.int64),
                      np.asarray(ps, dtype=np.int64))
