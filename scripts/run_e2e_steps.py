"""Scratch: e2e sketch steps from pinned host memory, with library debug timing."""
import sys, time, os
import numpy as np
import torch
sys.path.insert(0, ".")
import sylph_b200
from sylph_b200 import synth
n_reads = int(sys.argv[1]) if len(sys.argv) > 1 else 6_666_667
ctx = sylph_b200.Context(0, stream=torch.cuda.current_stream().cuda_stream)
buf, off = synth.reads(n_reads, device="cuda")
hb = torch.empty(buf.numel(), dtype=torch.uint8, pin_memory=True); hb.copy_(buf)
ho = torch.empty(off.numel(), dtype=torch.int64, pin_memory=True); ho.copy_(off)
torch.cuda.synchronize()
# raw H2D bandwidth reference
d = torch.empty_like(buf)
for _ in range(3):
    t = time.perf_counter(); d.copy_(hb, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t
print("raw pinned H2D: %.1f GB/s" % (buf.numel() / dt / 1e9))
hb_np, ho_np = hb.numpy(), ho.numpy().view(np.uint64)
oh = torch.empty(buf.numel() // 100, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
oc = torch.empty(buf.numel() // 100, dtype=torch.int32, pin_memory=True).numpy().view(np.uint32)
for it in range(4):
    t = time.perf_counter()
    s = ctx.sketch_sequences(hb_np, ho_np)
    h, c = s.download(oh, oc)
    dt = time.perf_counter() - t
    print("e2e step %d: %d entries %.2f ms  %.1f Gbase/s" % (it, len(h), dt * 1e3, buf.numel() / dt / 1e9))
    s.free()
