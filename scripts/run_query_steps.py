"""Scratch: query / profile steps (1 sample vs G synthetic genomes) for ncu launch lists / timing."""
import sys, time
import torch
sys.path.insert(0, ".")
import sylph_b200
from sylph_b200 import synth
from sylph_b200.api import contain_params
import bench
G = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
n_reads = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ctx = sylph_b200.Context(0, stream=torch.cuda.current_stream().cuda_stream)
genomes = synth.sketch_db_range(ctx, 0, G)
db = ctx.build_db(genomes)
b, o = synth.reads(n_reads, device="cuda")
smp = ctx.sketch_sequences(b, o)
torch.cuda.synchronize()


def best_of(fn, n):
    ts = []
    for _ in range(n):
        t = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t)
    return r, min(ts) * 1e3, sorted(ts)[len(ts) // 2] * 1e3


for kw in ({}, {"no_ci": 1}):
    P = contain_params(pseudotax=False, **kw)
    rows, lo, med = best_of(lambda: ctx.query(db, [smp], P), reps)
    print(kw, "rows", len(rows), "min %.3f ms median %.3f ms" % (lo, med))
P = contain_params(pseudotax=True)
rows, lo, med = best_of(lambda: ctx.profile(db, [smp], P), reps)
print("profile rows", len(rows), "min %.3f ms median %.3f ms" % (lo, med))
