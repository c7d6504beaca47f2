"""Run under torchrun (N ranks, one GPU each): sharded query/profile must equal the single-GPU result.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sylph_b200  # noqa: E402
from sylph_b200 import dist as D  # noqa: E402
from sylph_b200 import synth  # noqa: E402
from sylph_b200.api import contain_params  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sylph_b200.Context(local, stream=torch.cuda.current_stream().cuda_stream)
    G, glen, c = 200, 120000, 20
    b0, b1 = D.shard_range(G, rank, world)
    bases, off = synth.db_chunk(b0, b1, glen, device="cuda")
    goff = torch.arange(b1 - b0 + 1, dtype=torch.int64, device="cuda")
    genomes = ctx.sketch_genomes(bases, off, goff, c=c)
    db = ctx.build_db(genomes, genome_base=b0)
    samples = []
    for si in range(3):
        rb, ro = synth.reads(60000, n_comm=120, genome_len=glen, seed=synth.SEED_READS + si, device="cuda")
        samples.append(ctx.sketch_sequences(rb, ro, c=c))
    q = D.query_sharded(ctx, db, samples)
    p = D.profile_sharded(ctx, genomes, db, samples, b0)
    ok = True
    if rank == 0:  # single-GPU reference with the whole db
        fb, fo = synth.db_chunk(0, G, glen, device="cuda")
        fg = ctx.sketch_genomes(fb, fo, torch.arange(G + 1, dtype=torch.int64, device="cuda"), c=c)
        fdb = ctx.build_db(fg)
        q1 = ctx.query(fdb, samples, contain_params(pseudotax=False))
        p1 = ctx.profile(fdb, samples, contain_params(pseudotax=True))
        for name, a, b in (("query", q, q1), ("profile", p, p1)):
            same = len(a) == len(b)
            if same:
                for f in ("sample", "genome", "contain", "glen", "kmers_lost", "lambda_status", "ci_valid"):
                    same &= bool(np.array_equal(a[f], b[f]))
                for f in ("final_est_ani", "final_est_cov", "naive_ani", "rel_abund", "seq_abund"):
                    same &= bool(np.allclose(a[f], b[f], rtol=1e-9, atol=0))
                same &= bool(np.allclose(a["ci"], b["ci"], rtol=1e-9, atol=0))
            print("%s: sharded(%d ranks) rows=%d single rows=%d equal=%s" % (name, world, len(a), len(b), same))
            ok &= same
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
