"""GPU parity of the 2-bit packed ingest path: packed == ASCII == oracle.
syl_pack2 (host packer, exact BYTE_TO_SEQ) -> syl_seed_batch_packed2 / syl_sketch_reads_packed2, host and
device memory, every byte value, ragged records, tile edges of the warp-tile (4096) and the chunk ring."""
import numpy as np
import pytest

from tests.test_seed_gpu import oracle_survivors, random_records
from tests.util import flatten

pytestmark = pytest.mark.gpu

ALL_BYTES = bytes(range(256))


def survivors_packed(ctx, buf, off, k, c, sem, with_pos, device):
    from sylph_b200.api import pack2
    words = pack2(buf)
    if device:
        import torch
        w = torch.from_numpy(words.view(np.int32)).cuda()
        o = torch.from_numpy(off.astype(np.int64)).cuda()
        return ctx.extract_markers_batch(w, o, k=k, c=c, sem=sem, with_pos=with_pos, packed_bases=len(buf))
    return ctx.extract_markers_batch(words, off, k=k, c=c, sem=sem, with_pos=with_pos, packed_bases=len(buf))


@pytest.mark.parametrize("k", [31, 21])
@pytest.mark.parametrize("sem", [1, 0])
@pytest.mark.parametrize("device", [False, True])
def test_seeding_packed_equals_oracle_all_byte_values(ctx, k, sem, device):
    rng = np.random.default_rng(3 + k + sem)
    lengths = [0, 1, k - 1, k, k + 1, k + 3, 2 * k - 1, 2 * k, 66, 150, 151, 400, 401, 4095, 4096, 4097, 4096 + 47, 4096 + 48,
               3 * 4096 + 5, 0, 17, 40000] + list(rng.integers(0, 500, size=600))
    buf, off = random_records(rng, lengths, alphabet=ALL_BYTES)
    sv = survivors_packed(ctx, buf, off, k, 5, sem, True, device)
    exp = oracle_survivors(buf, off, k, 5, sem, True)
    got = sorted((int(a), int(b), int(h)) for h, a, b in zip(sv["hash"], sv["rec"], sv["pos"]))
    assert got == sorted(exp) and len(exp) > 1000
    sv2 = ctx.extract_markers_batch(buf, off, k=k, c=5, sem=sem, with_pos=True)      # ASCII path, same survivors
    assert got == sorted((int(a), int(b), int(h)) for h, a, b in zip(sv2["hash"], sv2["rec"], sv2["pos"]))


@pytest.mark.parametrize("device", [False, True])
@pytest.mark.parametrize("no_dedup", [False, True])
def test_read_sketch_packed_equals_ascii_equals_oracle(ctx, device, no_dedup):
    """Pair keys come from the packed stream (in-tile) or from packed global memory (reads cut by a warp-tile
    edge: k_events_fix<packed>); duplicates make the dedup state machine depend on them."""
    from oracle import oracle as O
    from sylph_b200.api import pack2
    rng = np.random.default_rng(77)
    genome = bytes(rng.choice(list(b"ACGT"), size=120000).astype(np.uint8))
    seqs = []
    for _ in range(9000):
        st = int(rng.integers(0, 120000 - 450))
        ln = int(rng.choice([60, 66, 70, 100, 149, 150, 150, 151, 250, 400, 401]))
        s = genome[st:st + ln]
        if rng.random() < 0.05:
            s = bytes(rng.choice(list(ALL_BYTES), size=ln).astype(np.uint8))   # arbitrary bytes
        seqs.append(s)
        if rng.random() < 0.3:
            seqs.append(s)
    seqs += [b"", b"A", b"N" * 150, b"acgtn" * 30]
    buf, off = flatten([seqs[i] for i in rng.permutation(len(seqs))])
    eh, ec, _, nd = O.sketch_reads(buf, off, c=5, no_dedup=no_dedup)
    words = pack2(buf)
    if device:
        import torch
        sp = ctx.sketch_sequences(torch.from_numpy(words.view(np.int32)).cuda(), torch.from_numpy(off.astype(np.int64)).cuda(),
                                  c=5, no_dedup=no_dedup, packed_bases=len(buf))
        sa = ctx.sketch_sequences(torch.from_numpy(buf).cuda(), torch.from_numpy(off.astype(np.int64)).cuda(), c=5, no_dedup=no_dedup)
    else:
        sp = ctx.sketch_sequences(words, off, c=5, no_dedup=no_dedup, packed_bases=len(buf))
        sa = ctx.sketch_sequences(buf, off, c=5, no_dedup=no_dedup)
    for s in (sp, sa):
        h, c = s.download()
        assert np.array_equal(h, eh) and np.array_equal(c, ec) and s.num_dup_removed == nd
    if not no_dedup:
        assert nd > 1000


def test_host_ingest_chunk_ring(ctx, monkeypatch):
    """Host ASCII -> worker-pool packer -> pinned ring -> device: chunks far smaller than the input so that every
    staging slot is recycled many times, records larger than a chunk, and 1-thread / many-thread pools agree."""
    from oracle import oracle as O
    rng = np.random.default_rng(5)
    lengths = list(rng.integers(0, 400, size=4000)) + [70000, 150, 150, 33, 0, 0, 9000]
    buf, off = random_records(rng, [lengths[i] for i in rng.permutation(len(lengths))], alphabet=b"ACGTNacgt")
    eh, ec, _, nd = O.sketch_reads(buf, off, c=11)
    for chunk in ("4096", "65536", None):
        if chunk:
            monkeypatch.setenv("SYL_INGEST_CHUNK", chunk)
        else:
            monkeypatch.delenv("SYL_INGEST_CHUNK", raising=False)
        s = ctx.sketch_sequences(buf, off, c=11)
        h, c = s.download()
        assert np.array_equal(h, eh) and np.array_equal(c, ec) and s.num_dup_removed == nd, chunk


@pytest.mark.parametrize("tw", [64, 128, 1024, 3072])
def test_warp_tile_lengths(ctx, monkeypatch, tw):
    """The launcher sizes the warp-tile from the mean record length; force odd lengths so that tile edges fall
    everywhere inside records, halos and pair-key windows (ASCII and packed input, survivors and read sketches)."""
    from oracle import oracle as O
    from sylph_b200.api import pack2
    monkeypatch.setenv("SYL_SEED_TW", str(tw))
    monkeypatch.setenv("SYL_SEED_IMPL", "warp")
    rng = np.random.default_rng(100 + tw)
    lengths = list(rng.integers(0, 420, size=1500)) + [5000, 31, 32, 61, 62, 66]
    buf, off = random_records(rng, [lengths[i] for i in rng.permutation(len(lengths))], alphabet=b"ACGTNacgt\x00\x03")
    exp = sorted(oracle_survivors(buf, off, 31, 7, 1, True))
    for packed in (False, True):
        sv = survivors_packed(ctx, buf, off, 31, 7, 1, True, False) if packed else ctx.extract_markers_batch(buf, off, k=31, c=7, with_pos=True)
        assert sorted((int(a), int(b), int(h)) for h, a, b in zip(sv["hash"], sv["rec"], sv["pos"])) == exp
    eh, ec, _, nd = O.sketch_reads(buf, off, c=7)
    for s in (ctx.sketch_sequences(buf, off, c=7), ctx.sketch_sequences(pack2(buf), off, c=7, packed_bases=len(buf))):
        h, c = s.download()
        assert np.array_equal(h, eh) and np.array_equal(c, ec) and s.num_dup_removed == nd


def test_host_ingest_mixed_packed_and_ascii_chunks(ctx, monkeypatch):
    """Pinned caller memory: chunks the packers have not started may cross the link as ASCII (taken from the back
    of the sample, out of order — the read indices travel with the chunk).  Forced here to alternate."""
    import torch
    from oracle import oracle as O
    rng = np.random.default_rng(6)
    lengths = list(rng.integers(0, 400, size=6000)) + [150] * 3000 + [70000, 33, 0]
    buf, off = random_records(rng, [lengths[i] for i in rng.permutation(len(lengths))], alphabet=b"ACGTNacgt")
    seqs_dup = np.concatenate([buf, buf[: len(buf) // 3]])   # duplicated reads across distant chunks: order-dependent dedup
    off_dup = np.concatenate([off, off[-1] + off[1:np.searchsorted(off, len(buf) // 3)]]).astype(np.uint64)
    seqs_dup = seqs_dup[: int(off_dup[-1])]
    eh, ec, _, nd = O.sketch_reads(seqs_dup, off_dup, c=11)
    hb = torch.empty(len(seqs_dup), dtype=torch.uint8, pin_memory=True)
    hb.numpy()[:] = seqs_dup
    ho = torch.empty(len(off_dup), dtype=torch.int64, pin_memory=True)
    ho.numpy()[:] = off_dup.astype(np.int64)
    monkeypatch.setenv("SYL_INGEST_CHUNK", "16384")
    for force in (True, False):
        if force:
            monkeypatch.setenv("SYL_INGEST_FORCE_STEAL", "1")
        else:
            monkeypatch.delenv("SYL_INGEST_FORCE_STEAL", raising=False)
        s = ctx.sketch_sequences(hb.numpy(), ho.numpy().view(np.uint64), c=11)
        h, c = s.download()
        assert np.array_equal(h, eh) and np.array_equal(c, ec) and s.num_dup_removed == nd, force
        h2d, n_packed, n_ascii = ctx.ingest_stats()   # what the call moved (bench.py's e2e.h2d_bytes_per_step)
        assert n_packed + n_ascii >= len(seqs_dup) // 16384 // 2 and n_packed > 0   # chunks end on record boundaries; one record is 70 kb
        if force:
            assert n_ascii > 0
        assert len(seqs_dup) // 4 <= h2d <= len(seqs_dup) + 8 * (len(off_dup) + n_packed + n_ascii)
    assert nd > 100
    monkeypatch.setenv("SYL_HOST_INGEST", "ascii")
    s = ctx.sketch_sequences(hb.numpy(), ho.numpy().view(np.uint64), c=11)
    h, c = s.download()
    assert np.array_equal(h, eh) and np.array_equal(c, ec)
    h2d, n_packed, n_ascii = ctx.ingest_stats()
    assert n_packed == 0 and n_ascii >= 1 and h2d >= len(seqs_dup) + 8 * len(off_dup)
