"""GPU: BASELINE.json full-size workload (1 Gbp of 150 bp reads) through size-independent properties
(the oracle cannot finish 1 Gbp in seconds, so this complements the bit-exact small-size tests):
  * the sketch is strictly ascending in hash and every count >= 1
  * conservation: sum(count) + num_dup_removed == number of seeding survivors
  * determinism: two runs are identical; the host-chunked (H2D pipelined) path == the device path
  * a 1/300 sub-sample of the reads sketched by the ORACLE without dedup is dominated by the full
    no-dedup GPU sketch (same hashes present, counts <=), and equals the GPU sketch of that sub-sample
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_one_gbp_properties(ctx):
    import torch
    from oracle import oracle as O
    from sylph_b200 import synth
    n_reads = 6_666_667
    b, o = synth.reads(n_reads, device="cuda")
    torch.cuda.synchronize()
    s1 = ctx.sketch_sequences(b, o)
    h1, c1 = s1.download()
    assert len(h1) > 1_000_000
    assert np.all(h1[1:] > h1[:-1]) and c1.min() >= 1
    surv = torch.empty(int(b.numel() / 200 * 1.3 + 65536) * 2, dtype=torch.int64, device="cuda")
    n_surv = ctx.extract_markers_batch(b, o, out=surv)
    del surv
    assert int(c1.astype(np.int64).sum()) + s1.num_dup_removed == n_surv
    assert abs(s1.mean_read_length - 150.0) < 1e-9
    s2 = ctx.sketch_sequences(b, o)
    h2, c2 = s2.download()
    assert np.array_equal(h1, h2) and np.array_equal(c1, c2) and s1.num_dup_removed == s2.num_dup_removed
    hb, ho = b.cpu().numpy(), o.cpu().numpy().astype(np.uint64)
    s3 = ctx.sketch_sequences(hb, ho)   # host buffers: 128 MB chunks, copies overlapped with seeding
    h3, c3 = s3.download()
    assert np.array_equal(h1, h3) and np.array_equal(c1, c3) and s3.num_dup_removed == s1.num_dup_removed
    # sub-sample vs oracle (no dedup => counts are plain multiplicities, so sub-sample <= full)
    sub = 22_222
    r0 = 1_234_567
    sb, so = hb[r0 * 150:(r0 + sub) * 150], (ho[r0:r0 + sub + 1] - ho[r0]).astype(np.uint64)
    eh, ec, _, _ = O.sketch_reads(sb, so, no_dedup=True, nthreads=4)
    gs = ctx.sketch_sequences(sb, so, no_dedup=True)
    gh, gc = gs.download()
    assert np.array_equal(gh, eh) and np.array_equal(gc, ec)
    full_nd = ctx.sketch_sequences(b, o, no_dedup=True)
    fh, fc = full_nd.download()
    pos = np.searchsorted(fh, eh)
    assert np.all(pos < len(fh)) and np.array_equal(fh[pos], eh) and np.all(fc[pos] >= ec)
    for s in (s1, s2, s3, gs, full_nd):
        s.free()
