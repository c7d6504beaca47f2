"""Scratch: N sketch steps on device-resident synthetic reads (for ncu launch lists)."""
import sys, time
import torch
sys.path.insert(0, ".")
import sylph_b200
from sylph_b200 import synth
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 6_666_667
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
ctx = sylph_b200.Context(0, stream=torch.cuda.current_stream().cuda_stream)
buf, off = synth.reads(n_reads, device="cuda")
torch.cuda.synchronize()
for it in range(steps):
    t = time.perf_counter()
    s = ctx.sketch_sequences(buf, off)
    torch.cuda.synchronize()
    print("step %d: %d entries %.3f ms" % (it, len(s), (time.perf_counter() - t) * 1e3))
    s.free()
