"""Shared test helpers (FASTX reading with needletail's seq() semantics, flat buffers)."""
import gzip
import os

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DATA = os.path.join(REPO, "tests", "golden", "data")


def read_fastx(path):
    """-> list of (header_line_without_marker, sequence_bytes). Line endings stripped, multi-line
    FASTA joined, no case/alphabet normalisation (needletail 0.5.1 seq()/id() semantics)."""
    op = gzip.open if open(path, "rb").read(2) == b"\x1f\x8b" else open
    with op(path, "rb") as f:
        data = f.read()
    recs = []
    if not data:
        return recs
    lines = data.split(b"\n")
    if data[:1] == b">":
        name, chunks = None, []
        for ln in lines:
            ln = ln.rstrip(b"\r")
            if ln[:1] == b">":
                if name is not None:
                    recs.append((name, b"".join(chunks)))
                name, chunks = ln[1:], []
            else:
                chunks.append(ln)
        if name is not None:
            recs.append((name, b"".join(chunks)))
    elif data[:1] == b"@":
        i = 0
        while i + 3 < len(lines) + 1 and i < len(lines):
            if not lines[i]:
                i += 1
                continue
            name = lines[i].rstrip(b"\r")[1:]
            seq = lines[i + 1].rstrip(b"\r")
            recs.append((name, seq))
            i += 4
    else:
        raise ValueError("not fasta/fastq: " + path)
    return recs


def flatten(seqs):
    """list of bytes -> (uint8 buffer, uint64 offsets[n+1])"""
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if seqs:
        off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    buf = np.frombuffer(b"".join(seqs), dtype=np.uint8).copy() if seqs else np.zeros(0, dtype=np.uint8)
    return buf, off
