"""Reader/writer for SLIMM's ``.sldb`` database files (host side, no cereal needed).

The reference serialises ``slimm_database`` with a cereal ``BinaryOutputArchive``
(reference: src/misc.hpp:77-100 for the struct, :178-195 for save/load).  The archive is
little-endian, has no header or version tag and lays the two maps out as

    u64 n ; n x { u64 len ; bytes accession ; u64 8 ; 8 x u32 lineage }
    u64 m ; m x { u32 taxid ; u32 rank(0..8) ; u64 len ; bytes name }

(cereal: types/concepts/pair_associative_container.hpp, types/vector.hpp for arithmetic
vectors, types/string.hpp, types/tuple.hpp - the vendored cereal is not shipped here, the
layout was pinned by round-tripping files through the reference binaries, see
tests/test_sldb.py and tests/golden/make_golden.py).

Map iteration order in the file is libstdc++ hash order; nothing on the profiling path depends
on it, so the writer emits insertion order.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Tuple

import numpy as np

LINEAGE_LENGTH = 8  # reference: src/slimm_build.cpp LINAGE_LENGTH / src/misc.hpp taxa_ranks 0..7

RANK_NAMES = ["strain", "species", "genus", "family", "order", "class", "phylum", "superkingdom",
              "intermidiate"]          # spelling as in reference src/misc.hpp:52-63
RANK_SHORT = ["r", "s", "g", "f", "o", "c", "p", "k", "i"]   # reference src/misc.hpp:65-75


def rank_from_string(s: str) -> int:
    """reference: src/misc.hpp:38-49 (to_taxa_ranks); anything unknown is 8."""
    try:
        i = RANK_NAMES.index(s)
    except ValueError:
        return 8
    return i if i < 8 else 8


@dataclass
class SlimmDatabase:
    """In-memory form of the two maps of ``slimm_database`` (reference src/misc.hpp:77-84)."""
    ac__taxid: Dict[str, np.ndarray] = field(default_factory=dict)      # accession -> u32[8]
    taxid__name: Dict[int, Tuple[int, str]] = field(default_factory=dict)  # taxid -> (rank, name)

    def lineage_table(self, accessions: List[str]) -> np.ndarray:
        """[G,8] u32 lineage table for contigs in @SQ order; unknown accessions get all zeros
        (reference src/slimm.hpp:434-442)."""
        out = np.zeros((len(accessions), LINEAGE_LENGTH), dtype=np.uint32)
        for g, acc in enumerate(accessions):
            lin = self.ac__taxid.get(acc)
            if lin is not None:
                out[g] = lin
        return out


def write_sldb(db: SlimmDatabase, path: str) -> None:
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(db.ac__taxid)))
        for acc, lin in db.ac__taxid.items():
            b = acc.encode()
            f.write(struct.pack("<Q", len(b)))
            f.write(b)
            lin = np.asarray(lin, dtype="<u4")
            f.write(struct.pack("<Q", lin.size))
            f.write(lin.tobytes())
        f.write(struct.pack("<Q", len(db.taxid__name)))
        for taxid, (rank, name) in db.taxid__name.items():
            b = name.encode()
            f.write(struct.pack("<IIQ", taxid, rank, len(b)))
            f.write(b)


def read_sldb(path: str) -> SlimmDatabase:
    with open(path, "rb") as f:
        buf = f.read()
    off = 0

    def take(fmt):
        nonlocal off
        v = struct.unpack_from(fmt, buf, off)
        off += struct.calcsize(fmt)
        return v

    db = SlimmDatabase()
    (n,) = take("<Q")
    for _ in range(n):
        (ln,) = take("<Q")
        acc = buf[off:off + ln].decode()
        off += ln
        (cnt,) = take("<Q")
        lin = np.frombuffer(buf, dtype="<u4", count=cnt, offset=off).astype(np.uint32)
        off += 4 * cnt
        db.ac__taxid[acc] = lin
    (m,) = take("<Q")
    for _ in range(m):
        taxid, rank, ln = take("<IIQ")
        name = buf[off:off + ln].decode()
        off += ln
        db.taxid__name[taxid] = (rank, name)
    if off != len(buf):
        raise ValueError(f"{path}: {len(buf) - off} trailing bytes after the two maps")
    return db


def build_db_from_taxonomy(acc_taxid: Dict[str, int], nodes: Dict[int, Tuple[int, str]],
                           names: Dict[int, str]) -> SlimmDatabase:
    """What ``slimm_build`` derives from nodes.dmp/names.dmp/accession2taxid
    (reference src/slimm_build.cpp:283-346): slot 0 is the accession's own taxid (labelled
    strain), slots 1..7 are filled while walking parent pointers to the root for ranks
    species..superkingdom; ranks never met stay 0.  ``nodes`` maps taxid -> (parent, rank string).
    """
    db = SlimmDatabase()
    for acc, tid0 in acc_taxid.items():
        lin = np.zeros(LINEAGE_LENGTH, dtype=np.uint32)
        lin[0] = tid0
        db.ac__taxid[acc] = lin
        tid = tid0
        db.taxid__name[tid] = (0, names.get(tid, ""))
        while tid != 1:
            node = nodes.get(tid)
            if node is None:
                break
            parent, rank_s = node
            r = rank_from_string(rank_s)
            if 1 <= r <= 7:
                lin[r] = tid
                db.taxid__name[tid] = (r, names.get(tid, ""))
            tid = parent
    return db
