"""Scratch: time the seeding kernel on synthetic 150 bp reads (device-resident)."""
import sys, time
import torch
sys.path.insert(0, ".")
import sylph_b200
from sylph_b200 import synth

n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
ctx = sylph_b200.Context(0, stream=torch.cuda.current_stream().cuda_stream)
t = time.time()
buf, off = synth.reads(n_reads, device="cuda")
torch.cuda.synchronize()
print("gen %.2fs, %d bases" % (time.time() - t, buf.numel()))
out = torch.empty(int(buf.numel() / 200 * 1.3 + 4096) * 2, dtype=torch.int64, device="cuda")
for it in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n = ctx.extract_markers_batch(buf, off, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print("iter %d: %d survivors  %.3f ms  %.1f Gbase/s" % (it, n, ms, buf.numel() / ms / 1e6))
