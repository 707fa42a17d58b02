"""Merge the ``_profile.tsv`` files of several samples into one table - the role of the reference's
``collect_profiles.py`` (reference collect_profiles.py:17-61), used after a directory run.

    python -m slimm_b200.collect_profiles [--clean] [-o merged_profile.tsv] a_profile.tsv b_profile.tsv ...

Default output: byte for byte what the reference script writes for the same files (``merged_profile.tsv`` in the working
directory), including what follows from its being written for an older column layout: rows are keyed by the lineage string,
the column called ``name`` carries the abundance of the last sample that has the row, a sample's column carries its
read_count field with the line's newline still attached (hence the quoted fields), absent rows read ``0.0``, and the rows are
sorted as STRINGS, descending, by level and lineage.

``--clean`` writes what one wants to read instead: one row per taxon (taxa_level, taxa_id, linage), and per sample its
read_count and abundance as numbers, sorted by level and total read count.
"""
from __future__ import annotations

import csv
import sys
from typing import Dict, List


def sample_name(path: str) -> str:
    """File name between the last '/' and the last '.' of the PATH (reference collect_profiles.py:17-22)."""
    return path[path.rfind("/") + 1:path.rfind(".")]


def read_rows(path: str) -> List[List[str]]:
    with open(path, "r") as f:
        lines = f.readlines()
    return [ln.split("\t") for ln in lines[1:]]               # the header line is skipped; fields keep the line's newline


def merge_reference_style(paths: List[str]) -> List[List[str]]:
    names = [sample_name(p) for p in paths]
    table: Dict[str, List[str]] = {}
    for p in paths:                                           # a row's first four columns: the last file that has it wins,
        for v in read_rows(p):                                # its place in the table: where it was first seen
            table[v[2]] = [v[0], v[1], v[3], v[2]]
    for row in table.values():
        row.extend(["0.0"] * len(paths))
    for k, p in enumerate(paths):
        for v in read_rows(p):
            table[v[2]][4 + k] = v[4]
    rows = list(table.values())
    # descending, as strings, by level, lineage and then every sample column (the reference's sort columns are its column names
    # from the fourth on, the lineage among them - which already decides the order); equal keys keep their order
    rows.sort(key=lambda r: (r[0], r[3], *r[4:]), reverse=True)
    return [["level", "taxid", "name", "linage"] + names] + rows


def merge_clean(paths: List[str]) -> List[List[str]]:
    names = [sample_name(p) for p in paths]
    table: Dict[str, List] = {}
    for k, p in enumerate(paths):
        for v in read_rows(p):
            key = v[1]
            if key not in table:
                table[key] = [v[0], v[1], v[2]] + [0, 0.0] * len(paths)
            table[key][3 + 2 * k] = int(v[4])
            table[key][4 + 2 * k] = float(v[3])
    rows = sorted(table.values(), key=lambda r: (r[0], -sum(r[3::2]), r[1]))
    header = ["taxa_level", "taxa_id", "linage"]
    for n in names:
        header += [n + ".read_count", n + ".abundance"]
    return [header] + [[str(x) for x in r] for r in rows]


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    clean, out = False, "merged_profile.tsv"
    files: List[str] = []
    while argv:
        a = argv.pop(0)
        if a == "--clean":
            clean = True
        elif a == "-o":
            out = argv.pop(0)
        else:
            files.append(a)
    if not files:
        sys.stderr.write(__doc__)
        return 1
    rows = merge_clean(files) if clean else merge_reference_style(files)
    with open(out, "w", newline="") as f:
        csv.writer(f, delimiter="\t", quoting=csv.QUOTE_MINIMAL, lineterminator="\n").writerows(rows)
    return 0


if __name__ == "__main__":
    sys.exit(main())
