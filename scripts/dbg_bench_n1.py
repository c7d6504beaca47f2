import sys, time, os
import torch
sys.path.insert(0, ".")
import sylph_b200, bench
from sylph_b200 import synth
use_sampler = int(sys.argv[1]); timing = int(sys.argv[2])
torch.cuda.set_device(0)
ctx = sylph_b200.Context(0, stream=torch.cuda.current_stream().cuda_stream)
bases, off = synth.reads(6666667, device="cuda"); torch.cuda.synchronize()
def step():
    s = ctx.sketch_sequences(bases, off); n = len(s); s.free()
cs = bench.ClockSampler(0)
if use_sampler: cs.start()
t = time.perf_counter(); n = 0
while n < 3 or time.perf_counter() - t < 0.5:
    step(); n += 1
print("warmup steps", n)
if timing: ctx.enable_timing(True)
for rep in range(3):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    ts = []
    for i in range(10):
        a = time.perf_counter(); step(); ts.append((time.perf_counter() - a) * 1e3)
    e1.record(); torch.cuda.synchronize()
    print("sampler %d timing %d: events %.2f ms/step wall %.2f; per-step %s" % (use_sampler, timing, e0.elapsed_time(e1) / 10, (time.perf_counter() - t0) * 100, " ".join("%.1f" % x for x in ts)))
if use_sampler: print(cs.stop())
