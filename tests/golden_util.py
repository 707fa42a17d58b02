"""Shared helpers for the golden-fixture tests: load a case, parse the reference's outputs,
compare profiles the way SURVEY.md section 8(c) prescribes."""
from __future__ import annotations

import glob
import json
import os
import re
from dataclasses import dataclass
from typing import Dict, List

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RANKS = ["strain", "species", "genus", "family", "order", "class", "phylum", "superkingdom"]


@dataclass
class Case:
    name: str
    path: str
    read_id: np.ndarray
    ref_id: np.ndarray
    begin_pos: np.ndarray
    ref_len: np.ndarray
    lineage: np.ndarray
    avg_read_length: int
    contig_names: List[str]
    accessions: List[str]
    rank_of: Dict[int, int]
    name_of: Dict[int, str]


@dataclass
class Run:
    name: str
    path: str
    bin_width: int        # 0 => avg read length
    cov_cut_off: float
    abundance_cut_off: float
    rank: int
    min_reads: int


def case_names() -> List[str]:
    return sorted(os.path.basename(os.path.dirname(p)) for p in glob.glob(os.path.join(GOLD, "*", "records.npz")))


def load_case(name: str) -> Case:
    p = os.path.join(GOLD, name)
    z = np.load(os.path.join(p, "records.npz"))
    meta = json.load(open(os.path.join(p, "meta.json")))
    taxa = {int(k): v for k, v in meta["taxa"].items()}
    return Case(name, p, z["read_id"], z["ref_id"], z["begin_pos"], z["ref_len"], z["lineage"],
                int(z["avg_read_length"]), meta["contig_names"], meta["accessions"],
                {t: v[0] for t, v in taxa.items()}, {t: v[1] for t, v in taxa.items()})


def runs_of(case: Case) -> List[Run]:
    out = []
    for rd in sorted(glob.glob(os.path.join(case.path, "runs", "*"))):
        args = json.load(open(os.path.join(rd, "args.json")))["args"]
        opt = {"-w": "0", "-cc": "0.95", "-ac": "0.01", "-r": "species", "-mr": "0"}
        for k, v in zip(args[::2], args[1::2]):
            opt[k] = v
        out.append(Run(os.path.basename(rd), rd, int(opt["-w"]), float(opt["-cc"]), float(opt["-ac"]),
                       RANKS.index(opt["-r"]), int(opt["-mr"])))
    return out


def all_runs():
    return [(c, r.name) for c in case_names() for r in runs_of(load_case(c))]


def parse_profile(path_or_lines):
    lines = open(path_or_lines).read().splitlines() if isinstance(path_or_lines, str) else list(path_or_lines)
    assert lines[0] == "taxa_level\ttaxa_id\tlinage\tabundance\tread_count"
    rows = {}
    for ln in lines[1:]:
        lvl, tid, lin, ab, cnt = ln.split("\t")
        assert tid not in rows, f"duplicate row {tid}"
        rows[tid] = (lvl, lin, ab, cnt)
    return rows, [ln.split("\t")[1] for ln in lines[1:]]


def assert_profiles_match(expected_path: str, got_lines: List[str]):
    """Rows keyed by taxa_id, order ignored except: header first, starred rows after the plain rows,
    0* last.  Plain rows: everything text-equal.  Starred rows: counts exact, abundance within 1e-4
    absolute (the reference sums them in f32 in hash-iteration order, SURVEY.md hard-part 7)."""
    exp, exp_order = parse_profile(expected_path)
    got, got_order = parse_profile(got_lines)
    assert set(exp) == set(got), f"row sets differ: only ref {set(exp) - set(got)}, only ours {set(got) - set(exp)}"
    assert got_order[-1] == "0*"
    first_star = min(i for i, t in enumerate(got_order) if t.endswith("*"))
    assert all(t.endswith("*") for t in got_order[first_star:])
    for tid, (lvl, lin, ab, cnt) in exp.items():
        g = got[tid]
        assert g[0] == lvl and g[1] == lin, (tid, g, exp[tid])
        assert g[3] == cnt, f"read_count of {tid}: ref {cnt} ours {g[3]}"
        if tid.endswith("*"):
            assert abs(float(g[2]) - float(ab)) <= 1e-4, f"abundance of {tid}: ref {ab} ours {g[2]}"
        else:
            assert g[2] == ab, f"abundance of {tid}: ref {ab} ours {g[2]}"


def parse_stderr_stats(path: str) -> Dict[str, str]:
    txt = open(path).read()
    pats = {"hits": r"(\d+) records processed", "n_reads": r"(\d+) matching reads\n", "n_uniq": r"(\d+) uniquily matching reads\n",
            "cut": r"  bins coverage cut-off = (\S+)", "ucut": r"uniq bins coverage cut-off = (\S+)",
            "n_valid": r"(\d+) passed the threshould", "failed_by_cov": r"(\d+) ref's couldn't pass the coverage",
            "failed_by_uniq_cov": r"(\d+) ref's couldn't pass the uniq coverage",
            "n_uniq2": r"increased from \d+ to (\d+)", "refs_with_reads": r"references with reads = (\d+)"}
    out = {}
    for k, p in pats.items():
        m = re.search(p, txt)
        if m:
            out[k] = m.group(1)
    return out
