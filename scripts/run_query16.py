"""Scratch: 16-sample query vs G genomes: timing + launch list."""
import sys, time
import torch
sys.path.insert(0, ".")
import sylph_b200, bench
from sylph_b200 import synth
from sylph_b200.api import contain_params
G = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
ctx = sylph_b200.Context(0, stream=torch.cuda.current_stream().cuda_stream)
genomes = bench.build_db(ctx, 0, G)
db = ctx.build_db(genomes)
samples = []
for si in range(16):
    b, o = synth.reads(800000, seed=synth.SEED_READS + 0x10 + si, device="cuda")
    samples.append(ctx.sketch_sequences(b, o))
torch.cuda.synchronize()
for kw in ({}, {"no_ci": 1}):
    P = contain_params(pseudotax=False, **kw)
    for it in range(4):
        t = time.perf_counter(); rows = ctx.query(db, samples, P); dt = time.perf_counter() - t
    print(kw, "rows", len(rows), "%.3f ms" % (dt * 1e3))
