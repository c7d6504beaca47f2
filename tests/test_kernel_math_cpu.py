"""CPU suite: the arithmetic identities the CUDA kernels rely on, checked in numpy / pure Python.

These do not run device code.  Each test restates one reformulation used by a kernel (cited) and
compares it with the direct definition, so that a future change of the kernel's reasoning has a
CPU-side reference to be checked against."""
import numpy as np
import pytest

from oracle import pyref as R

MASK64 = (1 << 64) - 1


def test_fp64_compare_orders_kmers_like_integers():
    """seed_kernel.cuh, canonical k-mer: values < 2^62 reinterpreted as IEEE doubles are finite,
    non-negative and ordered like the integers, so one DSETP replaces the 64-bit integer compare."""
    rng = np.random.default_rng(1)
    edge = np.array([0, 1, 2, (1 << 52) - 1, 1 << 52, (1 << 52) + 1, (1 << 61), (1 << 62) - 1, (1 << 62) - 2,
                     0x3FF0000000000000, 0x3FEFFFFFFFFFFFFF, 0x3FFFFFFFFFFFFFFF], dtype=np.uint64)
    rnd = rng.integers(0, 1 << 62, size=200000, dtype=np.uint64)
    near = rnd[:50000] ^ rng.integers(0, 4, size=50000, dtype=np.uint64)  # pairs that differ in the low bits only
    a = np.concatenate([edge, rnd, rnd[:50000], np.repeat(edge, len(edge))])
    b = np.concatenate([edge[::-1], rnd[::-1], near, np.tile(edge, len(edge))])
    fa, fb = a.view(np.float64), b.view(np.float64)
    assert np.all(np.isfinite(fa)) and np.all(np.isfinite(fb)) and np.all(fa >= 0) and np.all(fb >= 0)
    assert np.array_equal(fa < fb, a < b)
    # k = 21 k-mers (42 bits) are denormals or tiny normals: same property
    a21 = rng.integers(0, 1 << 42, size=100000, dtype=np.uint64)
    b21 = rng.integers(0, 1 << 42, size=100000, dtype=np.uint64)
    assert np.array_equal(a21.view(np.float64) < b21.view(np.float64), a21 < b21)
    # and the claim fails exactly where the kernel does not use it: bit 62 set
    assert not np.isfinite(np.array([0x7FF0000000000001], dtype=np.uint64).view(np.float64))[0]


def _even_fields(x):
    """seed_kernel.cuh even_fields(): the 16 even 2-bit fields (MSB-first) of a 64-bit word -> 32 bits."""
    x &= 0xCCCCCCCCCCCCCCCC
    x = (x | (x << 2)) & 0xF0F0F0F0F0F0F0F0 & MASK64
    x = (x | (x << 4)) & 0xFF00FF00FF00FF00 & MASK64
    x = (x | (x << 8)) & 0xFFFF0000FFFF0000 & MASK64
    x = (x | (x << 16)) & 0xFFFFFFFF00000000 & MASK64
    return x >> 32


def test_pair_keys_from_the_packed_stream():
    """seed_resolve<EMIT=1>: pair_kmer_single's four keys (src/sketch.rs:624-656) taken from two
    64-bit extracts of the MSB-first 2-bit stream (read start, read middle) by even/odd compression."""
    rng = np.random.default_rng(2)
    for L in (66, 67, 100, 150, 151, 400):
        for _ in range(20):
            s = bytes(rng.choice(list(b"ACGTNacgtn"), size=L).astype(np.uint8))
            codes = [R.BYTE_TO_SEQ[c] for c in s]

            def extract(q):  # 32 bases starting at q, first base in the top bits
                v = 0
                for c in codes[q:q + 32]:
                    v = (v << 2) | c
                return v

            a, b = extract(0), extract(L // 2)
            kf, kg = _even_fields(a), _even_fields((a << 2) & MASK64)
            kr, kt = _even_fields(b), _even_fields((b << 2) & MASK64)
            (f, r), (g, t) = R.pair_kmer_single(s)
            assert (kf, kr, kg, kt) == (f, r, g, t)


def _replay_direct(events, no_dedup=False):
    """dup_removal_lsh_full_exact for ONE k-mer (src/sketch.rs:690-731): events = [(read, pair or None)]
    in read order -> (count, dups)."""
    c, dups, seen = 0, 0, set()
    for _, pair in events:
        if not no_dedup and c < 4 and pair is not None:
            ret = False
            for pk in pair:
                if pk in seen:
                    if c > 0:
                        ret = True
                else:
                    seen.add(pk)
            if ret:
                dups += 1
                continue
        c += 1
    return c, dups


def _replay_kernel(events):
    """k_group_dedup's per-thread replay: keep the four smallest (read, event) keys of the unordered
    event list, replay them with 'set = every key of the earlier paired events'; a k-mer with more
    than four events and fewer than four counted goes to the full replay."""
    keyed = sorted(range(len(events)), key=lambda i: (events[i][0], i))[:4]
    c, dups, earlier = 0, 0, []
    for i in keyed:
        pair = events[i][1]
        if pair is None:
            c += 1
            continue
        ka, kb = pair
        found = (kb == ka) or any(ka in p or kb in p for p in earlier)
        earlier.append(pair)
        if found and c > 0:
            dups += 1
        else:
            c += 1
    if len(events) > 4:
        if c >= 4:
            return c + len(events) - 4, dups
        return _replay_direct(sorted(events, key=lambda e: e[0]))  # warp-cooperative path: whole k-mer in read order
    return c, dups


def test_dedup_replay_reformulation():
    rng = np.random.default_rng(3)
    for trial in range(3000):
        n = int(rng.integers(1, 12))
        reads = sorted(int(x) for x in rng.integers(0, 40, size=n))
        ev = []
        for r in reads:
            if rng.random() < 0.15:
                pair = None                      # read without pair keys (len < 66 or > 400)
            else:
                ka = int(rng.integers(0, 5))     # few distinct keys => many duplicates
                kb = ka if rng.random() < 0.1 else int(rng.integers(0, 5))
                pair = (ka, kb)
            ev.append((r, pair))
        # equal read index = the same read contributing the k-mer twice: identical events
        for i in range(1, n):
            if ev[i][0] == ev[i - 1][0]:
                ev[i] = ev[i - 1]
        direct = _replay_direct(ev)
        shuffled = [ev[i] for i in rng.permutation(n)]
        assert _replay_kernel(shuffled) == direct, (ev,)


def test_bucket_function_is_monotone():
    """sample.cu / seed_kernel.cuh BucketHist: bucket = min(mulhi(h, Mb), nbk-1) is monotone in the hash,
    so concatenating the buckets in order yields a sorted sketch."""
    rng = np.random.default_rng(4)
    for c in (1, 3, 200, 1000):
        thr = MASK64 // c
        for nbk in (4096, 131072):
            Mb = min((nbk << 64) // (thr + 1), MASK64)
            h = sorted(int(x) for x in rng.integers(0, thr, size=2000, dtype=np.uint64)) + [thr - 1]
            b = [min((x * Mb) >> 64, nbk - 1) for x in h]
            assert all(b[i] <= b[i + 1] for i in range(len(b) - 1))
            assert 0 <= b[0] and b[-1] <= nbk - 1 and b[-1] >= nbk - 2


def _stats_from_hist(values, glen):
    """k_stats_hist: everything get_stats needs from the 256-bin histogram of the hit counts."""
    H = [0] * 256
    for v in values:
        assert 0 < v < 256
        H[v] += 1
    n = sum(H)
    kth, acc, median = n // 2, 0, None
    for v in range(256):
        if acc <= kth < acc + H[v]:
            median = v
        acc += H[v]
    cut = [11, 15, 18, 21, 24, 26, 28, 31, 33, 35, 37, 39, 41, 43, 45, 46, 48, 50, 52, 53, 55, 57, 58, 60, 62, 63, 65, 67, 68]
    max_cov = cut[median - 1] if median < 30 else 1 << 40
    nz = sum(H[v] for v in range(256) if v <= max_cov)
    total = sum(v * H[v] for v in range(256) if v <= max_cov) & 0xFFFFFFFF
    return n, median, nz, total, [H[v] if v <= max_cov else 0 for v in range(17)]


@pytest.mark.parametrize("lam", [0.3, 1.5, 8.0, 40.0, 120.0])
def test_histogram_formulation_of_get_stats(lam):
    """contain.cu: median / Poisson cut / sums from a histogram equal the sorted-list definition
    (src/contain.rs:657-690)."""
    rng = np.random.default_rng(int(lam * 10))
    for _ in range(20):
        vals = [int(v) for v in rng.poisson(lam, size=int(rng.integers(1, 400))) if 0 < v < 256]
        if not vals:
            continue
        if rng.random() < 0.5:
            vals += [int(v) for v in rng.integers(1, 255, size=5)]  # outliers above the cut-off
        n, median, nz, total, h16 = _stats_from_hist(vals, 1000)
        covs = sorted(vals)
        assert n == len(covs) and median == covs[len(covs) // 2]
        max_cov = float("inf")
        if median < 30:
            for cv in covs[len(covs) // 2:]:
                if R.poisson_cdf(float(median), cv) < 0.9999999999:
                    max_cov = cv
                else:
                    break
        kept = [c for c in covs if c <= max_cov]
        assert nz == len(kept) and total == sum(kept) & 0xFFFFFFFF
        assert h16[1:] == [kept.count(v) for v in range(1, 17)]


@pytest.mark.parametrize("isa", ["best", "avx2", "scalar"])
def test_host_packer_every_instruction_set(isa):
    """the three code paths of the host packer (AVX-512 VBMI table look-up, AVX2 nibble tables, scalar table) are
    selected once per process: run the exactness test below in a child process for each"""
    import os
    import subprocess
    import sys
    env = dict(os.environ)
    env.pop("SYL_PACK_AVX2", None)
    env.pop("SYL_PACK_SCALAR", None)
    if isa == "avx2":
        env["SYL_PACK_AVX2"] = "1"
    if isa == "scalar":
        env["SYL_PACK_SCALAR"] = "1"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", "import sys; sys.path.insert(0, %r); from tests.test_kernel_math_cpu import "
                        "test_host_packer_is_exact_byte_to_seq as t; t(); print('ok')" % root],
                       env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300, cwd=root)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout[-2000:]


def test_host_packer_is_exact_byte_to_seq():
    """syl_pack2 (SIMD classification + scalar tail) against BYTE_TO_SEQ (src/types.rs:50-59) written
    out directly: every byte value, every length mod 64, multi-threaded == single-threaded.  Runs without a GPU."""
    import numpy as np
    from sylph_b200.api import pack2
    lut = np.zeros(256, np.uint64)
    for ch, v in ((b"C", 1), (b"c", 1), (b"G", 2), (b"g", 2), (b"T", 3), (b"t", 3), (b"U", 3), (b"u", 3)):
        lut[ch[0]] = v
    lut[1], lut[2], lut[3] = 1, 2, 3

    def ref(b):
        nw = (len(b) + 15) // 16
        c = np.zeros(nw * 16, np.uint64)
        c[:len(b)] = lut[b]
        return (c.reshape(nw, 16) << (30 - 2 * np.arange(16, dtype=np.uint64))).sum(axis=1).astype(np.uint32)

    rng = np.random.default_rng(1)
    every = np.arange(256, dtype=np.uint8).repeat(3)
    assert np.array_equal(pack2(every, 1), ref(every))
    for n in list(range(0, 140)) + [1000, 4097, (1 << 21) + 5]:
        b = rng.integers(0, 256, n, dtype=np.uint8)
        assert np.array_equal(pack2(b, 1), ref(b)), n
        assert np.array_equal(pack2(b, 3), ref(b)), n


def _lemire_threshold(c, n):
    """k_boot_iter_p's lemire_threshold: ceil(c * 2^64 / n) by two 64/32 division steps."""
    if c >= n:
        return (1 << 64) - 1
    d1 = c << 32
    q1, r1 = divmod(d1, n)
    q0, r0 = divmod(r1 << 32, n)
    assert q1 < (1 << 32) and q0 < (1 << 32) and d1 < (1 << 64) and (r1 << 32) < (1 << 64)
    return (q1 << 32) + q0 + (1 if r0 else 0)


def test_bootstrap_class_boundaries_on_the_raw_draw():
    """The bootstrap kernel never forms Lemire's index for the common classes: floor(x*n / 2^64) >= c
    is decided as x >= ceil(c * 2^64 / n), and that compare is taken from the carry of x + (2^64 - T)."""
    rng = np.random.default_rng(5)
    M = 1 << 64
    for _ in range(400):
        n = int(rng.integers(1, 1 << 32)) if rng.random() < 0.5 else int(rng.integers(1, 40000))
        for c in {0, 1, n - 1, n, int(rng.integers(0, n + 1)), int(rng.integers(0, n + 1))}:
            if c < 0:
                continue
            T = _lemire_threshold(c, n)
            if c < n:
                assert T == -((-c * M) // n)
            xs = [0, M - 1, T, max(T - 1, 0), min(T + 1, M - 1)] + [int(v) for v in rng.integers(0, M, 8, dtype=np.uint64)]
            for x in xs:
                hi = (x * n) >> 64
                if c < n:
                    assert (hi >= c) == (x >= T), (n, c, x)
                    if T >= 1:
                        assert ((x + (M - T)) >> 64) == (1 if x >= T else 0)
                else:
                    assert hi < c                    # nothing is drawn at or above n: the kernel zeroes this count
                # fastrand's redraw needs lo < n < 2^32, which needs the low word of x_lo * n to be < n
                lo = (x * n) % M
                if lo < n:
                    assert ((x & 0xFFFFFFFF) * n) & 0xFFFFFFFF < n


def test_tiled_join_ranges_partition_every_sample():
    """for_each_key's tiled mapping (contain.cu): rb[s][r] = first key of sample s whose db bucket is >= r * bpr, found by
    binary search on the sample's sorted hashes.  The R cells of a sample must be disjoint, cover all its keys, and put a
    key into the cell of its own bucket — whatever the sample size, including empty samples and R > number of buckets."""
    rng = np.random.default_rng(17)
    M64 = (1 << 64) - 1
    for trial in range(60):
        NB = int(rng.integers(1, 5000))
        M = int(rng.integers(1, 1 << 40))                       # bucket = min(mulhi(key, M), NB - 1), monotone in the key
        max_n = int(rng.integers(1, 4000))
        R = min(max(max_n // 100, 1), max(NB, 1)) if trial % 3 else int(rng.integers(1, 2 * NB + 2))
        bpr = (NB + R - 1) // R

        def bucket(k):
            return min((k * M) >> 64, NB - 1)

        for n in (0, 1, max_n, int(rng.integers(0, max_n + 1))):
            keys = np.sort(rng.integers(0, M64, size=n, dtype=np.uint64)).tolist()
            rb = []
            for r in range(R + 1):
                want = r * bpr
                lo, hi = (n, n) if r == R else (0, n)
                while lo < hi:
                    mid = (lo + hi) >> 1
                    if bucket(keys[mid]) < want:
                        lo = mid + 1
                    else:
                        hi = mid
                rb.append(lo)
            assert rb[0] == 0 and rb[R] == n and all(a <= b for a, b in zip(rb, rb[1:]))
            for r in range(R):
                for i in range(rb[r], rb[r + 1]):
                    assert r * bpr <= bucket(keys[i]) < (r + 1) * bpr
