"""Test-side writers for alignment files (independent of the C++ decoder they are used to check):
plain SAM with SEQ on every record, BGZF compression, and a minimal BAM encoder."""
from __future__ import annotations

import struct
import zlib
from typing import List

import numpy as np

BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf_compress(data: bytes, block: int = 0xFF00, level: int = 6) -> bytes:
    """Concatenated BGZF blocks (gzip members with the BC extra field) + the EOF marker block."""
    out = []
    for off in range(0, len(data), block):
        raw = data[off:off + block]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        cdata = c.compress(raw) + c.flush()
        bsize = len(cdata) + 25
        out.append(struct.pack("<BBBBIBBHBBHH", 0x1F, 0x8B, 8, 4, 0, 0, 0xFF, 6, ord("B"), ord("C"), 2, bsize))
        out.append(cdata)
        out.append(struct.pack("<II", zlib.crc32(raw) & 0xFFFFFFFF, len(raw)))
    out.append(BGZF_EOF)
    return b"".join(out)


def sam_text(contig_names: List[str], contig_lengths, qname: List[str], flag, ref_id, pos1, read_len: int = 100,
             seq_every: int = 1, crlf: bool = False, trailing_newline: bool = True) -> str:
    """SAM text; ref_id -1 -> RNAME '*'.  Every `seq_every`-th record carries SEQ, the others '*'."""
    nl = "\r\n" if crlf else "\n"
    seq = ("ACGT" * (read_len // 4 + 1))[:read_len]
    lines = ["@HD\tVN:1.4\tSO:unsorted"]
    lines += [f"@SQ\tSN:{n}\tLN:{int(l)}" for n, l in zip(contig_names, contig_lengths)]
    lines.append("@PG\tID:synth\tPN:synth")
    for i in range(len(qname)):
        s = seq if i % seq_every == 0 else "*"
        g = int(ref_id[i])
        if g < 0:
            lines.append(f"{qname[i]}\t{int(flag[i])}\t*\t0\t0\t*\t*\t0\t0\t{s}\t*")
        else:
            lines.append(f"{qname[i]}\t{int(flag[i])}\t{contig_names[g]}\t{int(pos1[i])}\t60\t{read_len}M\t*\t0\t0\t{s}\t*\tNM:i:0")
    txt = nl.join(lines)
    return txt + nl if trailing_newline else txt


def bam_bytes(contig_names: List[str], contig_lengths, qname: List[str], flag, ref_id, pos1, read_len: int = 100,
              seq_every: int = 1) -> bytes:
    """Uncompressed BAM stream (wrap with bgzf_compress)."""
    text = "@HD\tVN:1.4\tSO:unsorted\n" + "".join(f"@SQ\tSN:{n}\tLN:{int(l)}\n" for n, l in zip(contig_names, contig_lengths))
    out = [b"BAM\1", struct.pack("<i", len(text)), text.encode(), struct.pack("<i", len(contig_names))]
    for n, l in zip(contig_names, contig_lengths):
        nb = n.encode() + b"\0"
        out.append(struct.pack("<i", len(nb)) + nb + struct.pack("<i", int(l)))
    seq_packed = bytes([0x12] * ((read_len + 1) // 2))
    qual = bytes([0xFF] * read_len)
    for i in range(len(qname)):
        nm = qname[i].encode() + b"\0"
        has_seq = i % seq_every == 0
        l_seq = read_len if has_seq else 0
        g = int(ref_id[i])
        mapped = g >= 0
        cigar = struct.pack("<I", (read_len << 4) | 0) if mapped else b""
        body = struct.pack("<iiBBHHHiiii", g, int(pos1[i]) - 1, len(nm), 60 if mapped else 0, 4680, 1 if mapped else 0,
                           int(flag[i]), l_seq, -1, -1, 0)
        body += nm + cigar + (seq_packed + qual if has_seq else b"")
        out.append(struct.pack("<i", len(body)) + body)
    return b"".join(out)


def load_dump(path: str):
    """Reads the --dump-records file of slimm_b200/bin/slimm."""
    b = open(path, "rb").read()
    G, N, avg, n_records, n_reads = struct.unpack_from("<5Q", b, 0)
    off = 40
    lens = np.frombuffer(b, "<u4", G, off)
    off += 4 * G
    names = []
    for _ in range(G):
        (l,) = struct.unpack_from("<I", b, off)
        off += 4
        names.append(b[off:off + l].decode())
        off += l
    rid = np.frombuffer(b, "<u4", N, off); off += 4 * N
    ref = np.frombuffer(b, "<u4", N, off); off += 4 * N
    pos = np.frombuffer(b, "<i4", N, off); off += 4 * N
    assert off == len(b)
    return dict(G=G, N=N, avg=avg, n_records=n_records, n_reads=n_reads, ref_len=lens, names=names, read_id=rid, ref_id=ref,
                begin_pos=pos)
