"""Text side of the reference's outputs: ``*_profile.tsv`` rows and lineage strings.

Formatting follows the reference exactly: ``operator<<`` on float/double at the default precision
of 6 significant digits (== printf ``%g``), tab separated (reference src/slimm.hpp:733-843),
lineage strings ``k__..|p__..|...`` down to the requested rank with ``unknown_<rank>`` for missing
names (reference src/slimm.hpp:690-730).
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

RANK_NAMES = ["strain", "species", "genus", "family", "order", "class", "phylum", "superkingdom"]
RANK_SHORT = ["r", "s", "g", "f", "o", "c", "p", "k"]
PROFILE_HEADER = "taxa_level\ttaxa_id\tlinage\tabundance\tread_count"


def fmt_g(x) -> str:
    return "%g" % float(x)


def lineage_string(rank: int, lin: Sequence[int], name_of: Dict[int, str]) -> str:
    parts = []
    for i in range(7, rank - 1, -1):
        nm = name_of.get(int(lin[i]), "")
        parts.append(RANK_SHORT[i] + "__" + (nm if nm != "" else "unknown_" + RANK_NAMES[i]))
    return "|".join(parts)


def profile_lines(rows, lineage: np.ndarray, name_of: Dict[int, str], rank: int) -> List[str]:
    zeros = np.zeros(8, dtype=np.uint32)
    out = [PROFILE_HEADER]
    for r in rows:
        lin = lineage[r.first_child] if r.first_child != 0xFFFFFFFF else zeros
        if r.kind == 0:
            tid, ls = str(r.taxon), lineage_string(rank, lin, name_of)
        elif r.kind == 1:
            tid = f"{r.taxon}*"
            ls = (lineage_string(rank + 1, lin, name_of) + "|" + RANK_SHORT[rank] + "__" +
                  name_of.get(r.taxon, "") + "_unclassified")
        else:
            tid, ls = "0*", lineage_string(rank, zeros, name_of)
        out.append(f"{RANK_NAMES[rank]}\t{tid}\t{ls}\t{fmt_g(r.abundance)}\t{r.read_count}")
    return out
