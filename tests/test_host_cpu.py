"""CPU suite: host layer — the C++ driver builds and refuses to run without a GPU; the Python
.syldb/.sylsp codec round-trips and matches the documented bincode byte layout."""
import os
import struct
import subprocess

import numpy as np

from tests.util import REPO


def test_host_driver_builds_and_fails_loudly_without_gpu(tmp_path):
    from sylph_b200 import build
    build.build()
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(["make", "-C", os.path.join(REPO, "host"), "-s"], env=env)
    exe = os.path.join(REPO, "host", "sylph-b200")
    assert os.path.exists(exe)
    import torch
    if not torch.cuda.is_available():
        r = subprocess.run([exe, "query", "a.fa", "b.fq"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stderr


def test_formats_roundtrip_and_layout(tmp_path):
    from sylph_b200 import formats as F
    g = [dict(genome_kmers=np.array([5, 7, 9], np.uint64), tracked=np.array([11], np.uint64), file_name="a.fa",
              first_contig_name="chr1 desc", c=200, k=31, gn_size=1234, min_spacing=30),
         dict(genome_kmers=np.array([], np.uint64), tracked=None, file_name="b.fa", first_contig_name="", c=100, k=21,
              gn_size=0, min_spacing=5)]
    p = str(tmp_path / "x.syldb")
    F.write_syldb(p, g)
    raw = open(p, "rb").read()
    # u64 n_genomes | u64 3 | 3 x u64 | tag 1 | u64 1 | u64 11 | str | str | 4 x u64 | ...
    assert raw[:8] == struct.pack("<Q", 2) and raw[8:16] == struct.pack("<Q", 3)
    assert raw[16:40] == struct.pack("<QQQ", 5, 7, 9) and raw[40] == 1
    back = F.read_syldb(p)
    assert back[0]["genome_kmers"].tolist() == [5, 7, 9] and back[0]["tracked"].tolist() == [11]
    assert back[1]["tracked"] is None and back[1]["k"] == 21 and back[0]["first_contig_name"] == "chr1 desc"
    s = dict(hashes=np.array([3, 1], np.uint64), counts=np.array([2, 9], np.uint32), c=200, k=31, file_name="r.fq",
             sample_name=None, paired=False, mean_read_length=150.5)
    p2 = str(tmp_path / "r.sylsp")
    F.write_sylsp(p2, s)
    raw = open(p2, "rb").read()
    assert raw[:8] == struct.pack("<Q", 2) and raw[8:20] == struct.pack("<QI", 3, 2)  # (u64, u32) pairs, no padding
    b = F.read_sylsp(p2)
    assert b["hashes"].tolist() == [3, 1] and b["counts"].tolist() == [2, 9] and b["mean_read_length"] == 150.5
    s["sample_name"] = "S1"
    F.write_sylsp(p2, s)
    assert F.read_sylsp(p2)["sample_name"] == "S1"


def _fnv(recs):
    h = 1469598103934665603
    M = (1 << 64) - 1
    for _, s in recs:
        for b in s:
            h = ((h ^ b) * 1099511628211) & M
    for _, s in recs:
        h = ((h ^ len(s)) * 1099511628211) & M
    for name, _ in recs:
        for b in name:
            h = ((h ^ b) * 1099511628211) & M
    return h


def test_fastx_parser_matches_needletail_semantics(tmp_path):
    """C++ ingest (host/fastx.hpp) vs the Python reader: multi-line FASTA, CRLF, lower case / N kept as is,
    empty records, FASTQ, gzip, and the real config-1 files."""
    import gzip
    from tests.util import DATA, read_fastx
    exe = os.path.join(REPO, "host", "sylph-b200")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(["make", "-C", os.path.join(REPO, "host"), "-s"], env=env)
    fa = tmp_path / "m.fa"
    fa.write_bytes(b">c1 first contig\r\nACGTNNacgt\r\nTTTT\r\n>c2\n\n>c3 x\nAC\nGT\nA")
    fq = tmp_path / "r.fq"
    fq.write_bytes(b"@r1 d\nACGTN\n+\nIIIII\n@r2\nacg\n+r2\nIII\n")
    gz = tmp_path / "m.fa.gz"
    with gzip.open(gz, "wb") as f:
        f.write(fa.read_bytes())
    files = [str(fa), str(fq), str(gz), os.path.join(DATA, "e.coli-o157.fasta.gz"), os.path.join(DATA, "t1.fq"),
             os.path.join(DATA, "k12_R1.fq")]
    out = subprocess.run([exe, "fastx-stats"] + files, stdout=subprocess.PIPE, text=True, check=True).stdout.strip().split("\n")
    assert len(out) == len(files)
    for line, f in zip(out, files):
        name, n, nb, h, first = line.split("\t", 4)
        recs = read_fastx(f)
        assert (int(n), int(nb)) == (len(recs), sum(len(s) for _, s in recs)), f
        if sum(len(s) for _, s in recs) < 600000:  # the pure-Python FNV is slow on Mbp inputs
            assert int(h, 16) == _fnv(recs), f
        assert first == recs[0][0].decode(), f
    assert read_fastx(str(fa))[0] == (b"c1 first contig", b"ACGTNNacgtTTTT") and read_fastx(str(fa))[1] == (b"c2", b"")


def _bgzf(data, block=40000):
    """bgzip's container: gzip members of <= 64 KiB, each with its compressed size in a 'BC' extra field, + the EOF block."""
    import zlib
    out = bytearray()
    chunks = [data[i:i + block] for i in range(0, len(data), block)] + [b""]
    for ch in chunks:
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = co.compress(ch) + co.flush()
        bsize = len(body) + 25   # 12 header + 6 extra + body + 8 trailer - 1
        out += struct.pack("<BBBBIBBH", 0x1f, 0x8b, 8, 4, 0, 0, 0xff, 6) + b"BC" + struct.pack("<HH", 2, bsize)
        out += body + struct.pack("<II", zlib.crc32(ch) & 0xFFFFFFFF, len(ch))
    return bytes(out)


def test_bgzf_members_inflated_in_parallel_give_the_same_records(tmp_path):
    """A bgzip-compressed FASTQ goes through the block-parallel inflater (any -t > 1); plain gzip, the uncompressed
    file and -t 1 (zlib's gzread on the same BGZF bytes) must all give the same records.  A damaged block makes the
    file invalid instead of silently shortening it."""
    import gzip
    exe = os.path.join(REPO, "host", "sylph-b200")
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(["make", "-C", os.path.join(REPO, "host"), "-s"], env=env)
    rng = np.random.default_rng(11)
    recs = []
    for i in range(6000):
        L = int(rng.integers(0, 300))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTN", np.uint8), size=L))
        recs.append(b"@r%d some description\n" % i + seq + b"\n+\n" + b"I" * L + b"\n")
    data = b"".join(recs)
    assert len(data) > 1_000_000            # dozens of BGZF blocks, records straddle block borders
    plain, gz, bg = tmp_path / "r.fq", tmp_path / "r.fq.gz", tmp_path / "r.bgzf.fq.gz"
    plain.write_bytes(data)
    with gzip.open(gz, "wb") as f:
        f.write(data)
    bg.write_bytes(_bgzf(data))
    assert gzip.decompress(bg.read_bytes()) == data   # the container is valid multi-member gzip

    def stats(path, t):
        out = subprocess.run([exe, "fastx-stats", "-t", str(t), str(path)], stdout=subprocess.PIPE, text=True, check=True).stdout
        return out.strip().split("\t")[1:]

    want = stats(plain, 1)
    assert int(want[0]) == 6000
    for t in (1, 2, 8):
        assert stats(bg, t) == want and stats(gz, t) == want, t
    bad = bytearray(bg.read_bytes())
    bad[len(bad) // 2] ^= 0x55
    (tmp_path / "bad.fq.gz").write_bytes(bytes(bad))
    r = subprocess.run([exe, "fastx-stats", "-t", "4", str(tmp_path / "bad.fq.gz")], stdout=subprocess.PIPE, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("INVALID")


def test_pack_pool_scheduling_never_deadlocks(tmp_path):
    """The worker pool behind the host-memory read path (gate on the pinned staging ring, chunks taken over by the
    caller from the back, forced alternation): a scheduler stress test without a GPU — a host-side deadlock would
    otherwise only show up as a hung GPU test."""
    import subprocess
    src = os.path.join(REPO, "tests", "cpp", "pack_pool_stress.cpp")
    exe = str(tmp_path / "pack_pool_stress")
    csrc = os.path.join(REPO, "sylph_b200", "csrc")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", "-I", csrc, "-o", exe, src, os.path.join(csrc, "host_pack.cpp")])
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0 and "ok" in r.stdout
