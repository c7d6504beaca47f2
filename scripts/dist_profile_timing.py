"""Scratch: time the phases of dist.profile_sharded under torchrun."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sylph_b200, bench
from sylph_b200 import dist as D, synth
from sylph_b200.api import contain_params
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = sylph_b200.Context(local, stream=torch.cuda.current_stream().cuda_stream)
G = 10000
genomes = bench.build_db(ctx, rank * G, (rank + 1) * G)
db = ctx.build_db(genomes, genome_base=rank * G)
samples = []
for si in range(16):
    b, o = synth.reads(833333, seed=synth.SEED_READS + 0x1000 + si, device="cuda")
    samples.append(ctx.sketch_sequences(b, o))
P = contain_params(pseudotax=True)
def T():
    torch.cuda.synchronize(); return time.perf_counter()
for it in range(3):
    t0 = T()
    p1 = contain_params(pseudotax=True); p1.no_ci = 1
    rows1 = ctx.query(db, samples, p1); t1 = T()
    local_ids = np.unique(rows1["genome"].astype(np.int64) - rank * G).astype(np.uint32)
    sub = ctx.select_genomes(genomes, local_ids); t2 = T()
    merged, gids = D.gather_survivor_genomes(sub.device_tensors(), local_ids.astype(np.uint64) + np.uint64(rank * G)); t3 = T()
    g = ctx.upload_genomes(merged["kmers"], merged["kmer_off"], merged["tracked"], merged["tracked_off"], merged["gn_size"]); t4 = T()
    sdb = ctx.build_db(g); t5 = T()
    mine = [i for i in range(16) if i % world == rank]
    out = ctx.profile(sdb, [samples[i] for i in mine], P); t6 = T()
    parts = D.all_gather_bytes(out); t7 = T()
    if rank == 0:
        print("rows1 %d survivors(local) %d merged genomes %d kmers %d | query %.2f select %.2f gather %.2f upload %.2f build %.2f profile %.2f rows-gather %.2f ms" % (
            len(rows1), len(local_ids), len(gids), merged["kmers"].numel(), (t1-t0)*1e3, (t2-t1)*1e3, (t3-t2)*1e3, (t4-t3)*1e3, (t5-t4)*1e3, (t6-t5)*1e3, (t7-t6)*1e3))
    sdb.free(); g.free(); sub.free()
dist.destroy_process_group()
