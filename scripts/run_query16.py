"""Scratch: 16 full-depth samples vs G genomes in ONE profile call (config 4 shape on one GPU): timing; used under
ncu for the join kernels.  SYL_JOIN_PLAIN=1 selects the sample-major mapping for comparison."""
import sys, time
import torch
sys.path.insert(0, ".")
import sylph_b200
from sylph_b200 import synth
from sylph_b200.api import contain_params
G = int(sys.argv[1]) if len(sys.argv) > 1 else 5000
reads = int(sys.argv[2]) if len(sys.argv) > 2 else 6_666_667
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
ctx = sylph_b200.Context(0, stream=torch.cuda.current_stream().cuda_stream)
genomes = synth.sketch_db_range(ctx, 0, G)
db = ctx.build_db(genomes)
samples = []
for si in range(16):
    seed = synth.SEED_READS + 0x10 + si
    b, o = synth.reads(reads, seed=seed, device="cuda", comm=synth.community_ids(64, G, seed=seed))
    samples.append(ctx.sketch_sequences(b, o))
    del b, o
torch.cuda.synchronize()
P = contain_params(pseudotax=True)
for it in range(reps):
    t = time.perf_counter(); rows = ctx.profile(db, samples, P); dt = time.perf_counter() - t
    print("profile rows", len(rows), "%.3f ms" % (dt * 1e3), {k: round(ctx.kernel_time(k, reset=True)[0], 4) for k in ("join", "join2", "stats", "boot")})
