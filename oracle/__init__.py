"""CPU oracle for SLIMM's profiling hot path - TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package; nothing under ``slimm_b200/`` does.

* ``run()``            - ctypes front of ``slimm_oracle.c`` (stages A1-A7 of SURVEY.md appendix A)
* ``propagate()``      - A8, reference src/slimm.hpp:560-610
* ``profile_rows()``   - A9, reference src/slimm.hpp:733-843
* ``raw_table()``      - A10 / ``_raw.tsv`` columns, reference src/slimm.hpp:259-302,883-943

Parity pin: see the header of ``slimm_oracle.c`` (golden outputs of the reference binaries under
``tests/golden/``).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
REF_SLIMM = os.path.join(_HERE, "_ref", "slimm")
REF_SLIMM_BUILD = os.path.join(_HERE, "_ref", "slimm_build")

RANK_NAMES = ["strain", "species", "genus", "family", "order", "class", "phylum", "superkingdom"]
RANK_SHORT = ["r", "s", "g", "f", "o", "c", "p", "k"]


def build(force: bool = False) -> str:
    """Compile liboracle.so (gcc) if missing or stale."""
    src = os.path.join(_HERE, "slimm_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", _HERE, "all"], check=True)
    return _LIB_PATH


class _In(C.Structure):
    _fields_ = [("n_refs", C.c_uint32), ("ref_len", C.c_void_p), ("lineage", C.c_void_p),
                ("bin_width", C.c_uint32), ("avg_read_length", C.c_uint32), ("cov_cut_off", C.c_float),
                ("min_reads", C.c_uint32), ("n_records", C.c_uint64), ("read_id", C.c_void_p),
                ("ref_id", C.c_void_p), ("begin_pos", C.c_void_p)]


_U32P = C.POINTER(C.c_uint32)


class _Out(C.Structure):
    _fields_ = [("hits", C.c_uint32), ("n_reads", C.c_uint32), ("n_uniq", C.c_uint32), ("n_uniq2", C.c_uint32),
                ("failed_by_cov", C.c_uint32), ("failed_by_uniq_cov", C.c_uint32),
                ("failed_by_min_read", C.c_uint32), ("n_valid", C.c_uint32), ("min_reads", C.c_uint32),
                ("cut", C.c_float), ("ucut", C.c_float), ("n_bins", C.c_uint64), ("n_pairs", C.c_uint64),
                ("nb", _U32P), ("reads_count", _U32P), ("uniq_reads_count", _U32P), ("uniq_reads_count2", _U32P),
                ("nz", _U32P), ("unz", _U32P), ("unz2", _U32P),
                ("cp", C.POINTER(C.c_float)), ("ucp", C.POINTER(C.c_float)), ("valid", C.POINTER(C.c_uint8)),
                ("bin_off", C.POINTER(C.c_uint64)),
                ("cov", _U32P), ("uniq_cov", _U32P), ("uniq_cov2", _U32P),
                ("n_read_slots", C.c_uint32),
                ("read_n_targets", _U32P), ("read_n_valid", _U32P), ("read_assigned", _U32P), ("read_lca", _U32P),
                ("n_direct", C.c_uint64), ("direct_taxid", _U32P), ("direct_count", _U32P),
                ("n_child_pairs", C.c_uint64), ("child_taxid", _U32P), ("child_ref", _U32P)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_run.argtypes = [C.POINTER(_In), C.POINTER(_Out)]
        _lib.oracle_run.restype = C.c_int
        _lib.oracle_free.argtypes = [C.POINTER(_Out)]
        _lib.oracle_cov_depth.argtypes = [C.c_void_p, C.c_uint32]
        _lib.oracle_cov_depth.restype = C.c_float
        _lib.oracle_raw_abundance.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    return _lib


@dataclass
class OracleResult:
    hits: int
    n_reads: int
    n_uniq: int
    n_uniq2: int
    failed_by_cov: int
    failed_by_uniq_cov: int
    failed_by_min_read: int
    n_valid: int
    min_reads: int
    cut: np.float32
    ucut: np.float32
    n_pairs: int
    nb: np.ndarray
    reads_count: np.ndarray
    uniq_reads_count: np.ndarray
    uniq_reads_count2: np.ndarray
    nz: np.ndarray
    unz: np.ndarray
    unz2: np.ndarray
    cp: np.ndarray
    ucp: np.ndarray
    valid: np.ndarray
    bin_off: np.ndarray
    cov: np.ndarray
    uniq_cov: np.ndarray
    uniq_cov2: np.ndarray
    read_n_targets: np.ndarray
    read_n_valid: np.ndarray
    read_assigned: np.ndarray
    read_lca: np.ndarray
    direct: Dict[int, int]                 # taxon -> reads whose LCA it is
    child_pairs: np.ndarray                # [n,2] (taxon, ref), sorted, distinct


def _arr(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)


def run(ref_len, lineage, bin_width: int, avg_read_length: int, cov_cut_off: float,
        read_id, ref_id, begin_pos, min_reads: int = 0) -> OracleResult:
    ref_len = np.ascontiguousarray(ref_len, dtype=np.uint32)
    lineage = np.ascontiguousarray(lineage, dtype=np.uint32).reshape(-1, 8)
    read_id = np.ascontiguousarray(read_id, dtype=np.uint32)
    ref_id = np.ascontiguousarray(ref_id, dtype=np.uint32)
    begin_pos = np.ascontiguousarray(begin_pos, dtype=np.int32)
    G = ref_len.size
    assert lineage.shape[0] == G and read_id.size == ref_id.size == begin_pos.size
    i = _In(G, ref_len.ctypes.data, lineage.ctypes.data, bin_width, avg_read_length,
            float(cov_cut_off), min_reads, read_id.size, read_id.ctypes.data, ref_id.ctypes.data,
            begin_pos.ctypes.data)
    o = _Out()
    rc = lib().oracle_run(C.byref(i), C.byref(o))
    if rc != 0:
        lib().oracle_free(C.byref(o))
        raise ValueError(f"oracle_run failed with code {rc}")
    try:
        nbins, slots = o.n_bins, o.n_read_slots
        res = OracleResult(
            o.hits, o.n_reads, o.n_uniq, o.n_uniq2, o.failed_by_cov, o.failed_by_uniq_cov,
            o.failed_by_min_read, o.n_valid, o.min_reads, np.float32(o.cut), np.float32(o.ucut), o.n_pairs,
            _arr(o.nb, G, np.uint32), _arr(o.reads_count, G, np.uint32), _arr(o.uniq_reads_count, G, np.uint32),
            _arr(o.uniq_reads_count2, G, np.uint32), _arr(o.nz, G, np.uint32), _arr(o.unz, G, np.uint32),
            _arr(o.unz2, G, np.uint32), _arr(o.cp, G, np.float32), _arr(o.ucp, G, np.float32),
            _arr(o.valid, G, np.uint8), _arr(o.bin_off, G + 1, np.uint64),
            _arr(o.cov, nbins, np.uint32), _arr(o.uniq_cov, nbins, np.uint32), _arr(o.uniq_cov2, nbins, np.uint32),
            _arr(o.read_n_targets, slots, np.uint32), _arr(o.read_n_valid, slots, np.uint32),
            _arr(o.read_assigned, slots, np.uint32), _arr(o.read_lca, slots, np.uint32),
            dict(zip(_arr(o.direct_taxid, o.n_direct, np.uint32).tolist(),
                     _arr(o.direct_count, o.n_direct, np.uint32).tolist())),
            np.stack([_arr(o.child_taxid, o.n_child_pairs, np.uint32),
                      _arr(o.child_ref, o.n_child_pairs, np.uint32)], axis=1))
    finally:
        lib().oracle_free(C.byref(o))
    return res


# ------------------------------------------------------------------------------------------------
# A8: counts and children up the lineages (reference src/slimm.hpp:560-610)
# ------------------------------------------------------------------------------------------------
_M32 = 0xFFFFFFFF


def propagate(direct: Dict[int, int], child_pairs: np.ndarray, uniq_reads_count2: np.ndarray,
              lineage: np.ndarray, rank_of: Dict[int, int]) -> Tuple[Dict[int, int], Dict[int, set]]:
    """Returns (taxon_id__read_count, taxon_id__children).

    The reference walks a snapshot of the direct counts in libstdc++ hash order and reads
    ``children[t]`` live.  For lineage tables that are tree-consistent the result does not depend
    on that order; this restatement (and the product) visit the snapshot by
    ascending (rank, taxon)."""
    count: Dict[int, int] = dict(direct)
    children: Dict[int, set] = {}
    for t, g in np.asarray(child_pairs).reshape(-1, 2).tolist():
        children.setdefault(t, set()).add(g)
    for t, c in sorted(direct.items(), key=lambda kv: (rank_of.get(kv[0], 0), kv[0])):
        r = rank_of.get(t, 0)
        f = min(children[t])                       # first child of a std::set, :569-573
        lin = lineage[f]
        kids = set(children[t])                    # copied before the loop, :575
        for j in range(r + 1, 8):
            rec = int(lin[j])
            count[rec] = (count.get(rec, 0) + c) & _M32
            children.setdefault(rec, set()).update(kids)
    for g in range(len(uniq_reads_count2)):        # :589-610
        u2 = int(uniq_reads_count2[g])
        if u2 == 0:
            continue
        lin = lineage[g]
        kids = set(children.setdefault(int(lin[0]), set()))
        for j in range(1, 8):
            rec = int(lin[j])
            count[rec] = (count.get(rec, 0) + u2) & _M32
            s = children.setdefault(rec, set())
            s.add(g)
            s.update(kids)
    return count, children


# ------------------------------------------------------------------------------------------------
# A9: profile rows (reference src/slimm.hpp:690-843)
# ------------------------------------------------------------------------------------------------
def _lineage_string(rank: int, lin: Sequence[int], name_of: Dict[int, str]) -> str:
    parts = []
    for i in range(7, rank - 1, -1):
        nm = name_of.get(int(lin[i]), "")
        if nm == "":
            nm = "unknown_" + RANK_NAMES[i]
        parts.append(RANK_SHORT[i] + "__" + nm)
    return "|".join(parts)


def fmt_g(x) -> str:
    """ostream << float/double at default precision 6 (== printf %g)."""
    return "%g" % float(x)


@dataclass
class ProfileRow:
    taxa_id: str          # "123", "123*" or "0*"
    lineage: str
    abundance: float      # as f32 (f64 for the 0* row), before printing
    read_count: int

    def text(self, rank_name: str) -> str:
        return f"{rank_name}\t{self.taxa_id}\t{self.lineage}\t{fmt_g(self.abundance)}\t{self.read_count}"


def profile_rows(count: Dict[int, int], children: Dict[int, set], lineage: np.ndarray, ref_len: np.ndarray,
                 rank_of: Dict[int, int], name_of: Dict[int, str], n_reads: int, avg_read_length: int,
                 cut: np.float32, rank: int = 1, abundance_cut_off: float = 0.01) -> List[ProfileRow]:
    f32 = np.float32
    ac = f32(abundance_cut_off)
    pr = rank + 1
    pab: Dict[int, np.float32] = {}
    pcnt: Dict[int, int] = {}
    R = f32(np.uint32(n_reads))
    for t in sorted(count):
        if rank_of.get(t, 0) == pr:
            pab[t] = f32(f32(f32(np.uint32(count[t])) / R) * f32(100))
            pcnt[t] = count[t]
    rows: List[ProfileRow] = []
    sab: Dict[int, np.float32] = {}
    scnt: Dict[int, int] = {}
    sum_ab = f32(0)
    sum_cnt = 0
    for t in sorted(count):
        if rank_of.get(t, 0) != rank:
            continue
        c = count[t]
        kids = sorted(children[t])
        gl = (int(sum(int(ref_len[k]) for k in kids)) & _M32) // len(kids)
        with np.errstate(divide="ignore", invalid="ignore"):
            cv = f32(f32(np.uint32((c * avg_read_length) & _M32)) / f32(np.uint32(gl)))
        ab = f32(f32(f32(np.uint32(c)) / R) * f32(100))
        p = int(lineage[kids[-1]][pr])            # lineage of the LAST child iterated, :783-797
        sab[p] = f32(sab[p] + ab) if p in sab else ab
        scnt[p] = (scnt.get(p, 0) + c) & _M32
        name = name_of.get(t, "")
        if ab < ac or cv < cut or name == "":
            continue
        rows.append(ProfileRow(str(t), _lineage_string(rank, lineage[kids[0]], name_of), float(ab), c))
        sum_ab = f32(sum_ab + ab)
        sum_cnt = (sum_cnt + c) & _M32
    for p in sorted(sab):
        uab = f32(pab.get(p, f32(0)) - sab[p])
        ucnt = (pcnt.get(p, 0) - scnt[p]) & _M32
        pname = name_of.get(p, "")
        if uab > ac and pname != "":
            lin = lineage[min(children[p])] if p != 0 else np.zeros(8, dtype=np.uint32)
            ls = _lineage_string(pr, lin, name_of) + "|" + RANK_SHORT[rank] + "__" + pname + "_unclassified"
            rows.append(ProfileRow(f"{p}*", ls, float(uab), ucnt))
            sum_cnt = (sum_cnt + ucnt) & _M32
            sum_ab = f32(sum_ab + uab)
    rows.append(ProfileRow("0*", _lineage_string(rank, np.zeros(8, dtype=np.uint32), name_of),
                           100.0 - float(sum_ab), (n_reads - sum_cnt) & _M32))
    return rows


# ------------------------------------------------------------------------------------------------
# A10: _raw.tsv (reference src/slimm.hpp:259-302,883-943)
# ------------------------------------------------------------------------------------------------
RAW_HEADER = ("accesion\ttaxaid\tname\treads_count\tabundance\tuniq1_abundance\tuniq2_abundance\tgenome_length\t"
              "uniq1_reads_count\tuniq2_reads_count\tbins_count\tbins_count(>0)\tuniq1_bins_count(>0)\t"
              "uniq2_bins_count(>0)\tcoverage_depth\tuniq1_coverage_depth\tuniq2_coverage_depth\tcoverage(%)\t"
              "uniq1_coverage(%)\tuniq2_coverage(%)")


def raw_table(res: OracleResult, accessions: List[str], lineage: np.ndarray, ref_len: np.ndarray,
              name_of: Dict[int, str]) -> List[str]:
    G = len(accessions)
    ref_len = np.ascontiguousarray(ref_len, dtype=np.uint32)
    ab = np.zeros(G, dtype=np.float32)
    uab = np.zeros(G, dtype=np.float32)
    rc = np.ascontiguousarray(res.reads_count)
    urc = np.ascontiguousarray(res.uniq_reads_count)
    lib().oracle_raw_abundance(rc.ctypes.data, ref_len.ctypes.data, G, res.hits, ab.ctypes.data)
    lib().oracle_raw_abundance(urc.ctypes.data, ref_len.ctypes.data, G, res.n_uniq, uab.ctypes.data)
    lines = [RAW_HEADER]
    for g in range(G):
        a, b = int(res.bin_off[g]), int(res.bin_off[g + 1])
        depth = []
        for h in (res.cov, res.uniq_cov, res.uniq_cov2):
            seg = np.ascontiguousarray(h[a:b])
            depth.append(lib().oracle_cov_depth(seg.ctypes.data, b - a))
        nbf = np.float32(res.nb[g])
        cp2 = np.float32(res.unz2[g]) / nbf
        taxid = int(lineage[g][0])
        nm = name_of.get(taxid, "") or "no_name_found"
        cols = [accessions[g], str(taxid), nm, str(int(rc[g])), fmt_g(ab[g]), fmt_g(uab[g]), "0",
                str(int(ref_len[g])), str(int(urc[g])), str(int(res.uniq_reads_count2[g])), str(int(res.nb[g])),
                str(int(res.nz[g])), str(int(res.unz[g])), str(int(res.unz2[g])), fmt_g(depth[0]), fmt_g(depth[1]),
                fmt_g(depth[2]), fmt_g(res.cp[g]), fmt_g(res.ucp[g]), fmt_g(cp2)]
        lines.append("\t".join(cols))
    return lines


def coverage_lines(res: OracleResult, which: str, accessions: List[str], lineage: np.ndarray,
                   name_of: Dict[int, str]) -> List[str]:
    """``-co`` files (reference src/slimm.hpp:846-881): valid references ascending, comma-separated
    accession, the 8 lineage names, then every bin."""
    h = {"cov": res.cov, "uniq_cov": res.uniq_cov, "uniq_cov2": res.uniq_cov2}[which]
    out = []
    for g in np.nonzero(res.valid)[0].tolist():
        a, b = int(res.bin_off[g]), int(res.bin_off[g + 1])
        cols = [accessions[g]] + [name_of.get(int(t), "") for t in lineage[g]] + [str(int(x)) for x in h[a:b]]
        out.append(",".join(cols))
    return out
