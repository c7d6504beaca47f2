"""sylph's on-disk sketches (.syldb / .sylsp) in Python: bincode 1.3 default configuration of
src/types.rs:145-173 (see host/sketch_io.hpp for the C++ twin).  Used by tests and by Python
callers that want to feed real sylph files to the library."""
import struct

import numpy as np


def _rd_u64(b, o):
    return struct.unpack_from("<Q", b, o)[0], o + 8


def _rd_str(b, o):
    n, o = _rd_u64(b, o)
    return b[o:o + n].decode("utf-8", "replace"), o + n


def _rd_vec64(b, o):
    n, o = _rd_u64(b, o)
    return np.frombuffer(b, dtype="<u8", count=n, offset=o).copy(), o + 8 * n


def read_syldb(path):
    """-> list of dicts (genome_kmers, tracked or None, file_name, first_contig_name, c, k, gn_size, min_spacing)"""
    b = open(path, "rb").read()
    n, o = _rd_u64(b, 0)
    out = []
    for _ in range(n):
        g = {}
        g["genome_kmers"], o = _rd_vec64(b, o)
        tag = b[o]
        o += 1
        g["tracked"] = None
        if tag == 1:
            g["tracked"], o = _rd_vec64(b, o)
        g["file_name"], o = _rd_str(b, o)
        g["first_contig_name"], o = _rd_str(b, o)
        for f in ("c", "k", "gn_size", "min_spacing"):
            g[f], o = _rd_u64(b, o)
        out.append(g)
    assert o == len(b), "trailing bytes in " + path
    return out


def write_syldb(path, genomes):
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(genomes)))
        for g in genomes:
            km = np.ascontiguousarray(g["genome_kmers"], dtype="<u8")
            f.write(struct.pack("<Q", len(km)) + km.tobytes())
            if g.get("tracked") is None:
                f.write(b"\x00")
            else:
                tr = np.ascontiguousarray(g["tracked"], dtype="<u8")
                f.write(b"\x01" + struct.pack("<Q", len(tr)) + tr.tobytes())
            for s in (g["file_name"], g["first_contig_name"]):
                e = s.encode()
                f.write(struct.pack("<Q", len(e)) + e)
            f.write(struct.pack("<QQQQ", g["c"], g["k"], g["gn_size"], g["min_spacing"]))


_PAIR = np.dtype([("hash", "<u8"), ("count", "<u4")])


def read_sylsp(path):
    """-> dict (hashes, counts, c, k, file_name, sample_name or None, paired, mean_read_length)"""
    b = open(path, "rb").read()
    n, o = _rd_u64(b, 0)
    pairs = np.frombuffer(b, dtype=_PAIR, count=n, offset=o)
    o += 12 * n
    s = {"hashes": pairs["hash"].copy(), "counts": pairs["count"].copy()}
    s["c"], o = _rd_u64(b, o)
    s["k"], o = _rd_u64(b, o)
    s["file_name"], o = _rd_str(b, o)
    tag = b[o]
    o += 1
    s["sample_name"] = None
    if tag == 1:
        s["sample_name"], o = _rd_str(b, o)
    s["paired"] = bool(b[o])
    o += 1
    s["mean_read_length"] = struct.unpack_from("<d", b, o)[0]
    o += 8
    assert o == len(b), "trailing bytes in " + path
    return s


def write_sylsp(path, s):
    pairs = np.empty(len(s["hashes"]), dtype=_PAIR)
    pairs["hash"] = s["hashes"]
    pairs["count"] = s["counts"]
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", len(pairs)) + pairs.tobytes())
        f.write(struct.pack("<QQ", s["c"], s["k"]))
        e = s["file_name"].encode()
        f.write(struct.pack("<Q", len(e)) + e)
        if s.get("sample_name") is None:
            f.write(b"\x00")
        else:
            e = s["sample_name"].encode()
            f.write(b"\x01" + struct.pack("<Q", len(e)) + e)
        f.write(b"\x01" if s.get("paired") else b"\x00")
        f.write(struct.pack("<d", s["mean_read_length"]))
