"""Scratch: query steps (1 sample vs G synthetic genomes) for ncu launch lists / timing."""
import sys, time
import torch
sys.path.insert(0, ".")
import sylph_b200
from sylph_b200 import synth
from sylph_b200.api import contain_params
sys.path.insert(0, ".")
import bench
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
ctx = sylph_b200.Context(0, stream=torch.cuda.current_stream().cuda_stream)
genomes = bench.build_db(ctx, 0, G)
db = ctx.build_db(genomes)
b, o = synth.reads(n_reads, device="cuda")
smp = ctx.sketch_sequences(b, o)
torch.cuda.synchronize()
for kw in ({}, {"no_ci": 1}):
    P = contain_params(pseudotax=False, **kw)
    for it in range(4):
        t = time.perf_counter()
        rows = ctx.query(db, [smp], P)
        dt = time.perf_counter() - t
    print(kw, "rows", len(rows), "%.3f ms" % (dt * 1e3))
P = contain_params(pseudotax=True)
for it in range(3):
    t = time.perf_counter(); rows = ctx.profile(db, [smp], P); dt = time.perf_counter() - t
print("profile rows", len(rows), "%.3f ms" % (dt * 1e3))
