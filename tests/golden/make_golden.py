"""Regenerates tests/golden/config1.json: the CPU oracle's outputs on BASELINE config 1
(test_files/e.coli-*.fasta.gz vs o157_reads.fastq.gz, copied to tests/golden/data/).

The reference is a Rust crate and cannot be run in the build image (no cargo/rustc), so these
values are ORACLE-generated regression pins, not reference-generated vectors; what the
reference's own tests do pin (1 profile row, 3 query rows) is asserted separately in
tests/test_oracle_cpu.py.  Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests.util import DATA, flatten, read_fastx  # noqa: E402

GENOMES = ["e.coli-EC590.fasta.gz", "e.coli-o157.fasta.gz", "e.coli-K12.fasta.gz"]


def digest(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:32]


def config1():
    out = {"k": 31, "c": 200, "genomes": [], "reads": {}}
    db = []
    for g in GENOMES:
        recs = read_fastx(os.path.join(DATA, g))
        buf, off = flatten([s for _, s in recs])
        km, tr, gs = O.sketch_genome(buf, off)
        db.append((km, tr, gs))
        out["genomes"].append({"file": g, "contigs": len(recs), "gn_size": int(gs), "n_kmers": len(km),
                               "n_tracked": len(tr), "kmers_sha": digest(km), "tracked_sha": digest(tr),
                               "first_contig": recs[0][0].decode()})
    recs = read_fastx(os.path.join(DATA, "o157_reads.fastq.gz"))
    buf, off = flatten([s for _, s in recs])
    h, c, mean, nd = O.sketch_reads(buf, off)
    out["reads"] = {"file": "o157_reads.fastq.gz", "n_reads": len(recs), "n_keys": len(h), "sum_counts": int(c.sum()),
                    "hash_sha": digest(h), "count_sha": digest(c), "mean_read_length": mean, "num_dup_removed": int(nd)}
    recs = read_fastx(os.path.join(DATA, "k12_R1.fq"))
    buf, off = flatten([s for _, s in recs])
    h2, c2, mean2, nd2 = O.sketch_reads(buf, off, c=20)
    out["k12_R1_c20"] = {"n_reads": len(recs), "n_keys": len(h2), "sum_counts": int(c2.sum()), "hash_sha": digest(h2),
                         "count_sha": digest(c2), "num_dup_removed": int(nd2)}
    smp = O.Sample(h, c)

    def run(sel, pseudotax):
        kmers = np.concatenate([db[i][0] for i in sel])
        koff = np.cumsum([0] + [len(db[i][0]) for i in sel]).astype(np.uint64)
        tr = np.concatenate([db[i][1] for i in sel])
        toff = np.cumsum([0] + [len(db[i][1]) for i in sel]).astype(np.uint64)
        gs = np.array([db[i][2] for i in sel], dtype=np.uint64)
        res = O.contain_sample(O.default_params(pseudotax=pseudotax), kmers, koff, tr, toff, gs, smp)
        return [O.format_row(r, pseudotax, "o157_reads.fastq.gz", GENOMES[sel[r.genome]],
                             out["genomes"][sel[r.genome]]["first_contig"]) for r in res]

    out["profile_vs_EC590"] = run([0], True)
    out["query_vs_all"] = run([0, 1, 2], False)
    out["profile_vs_all"] = run([0, 1, 2], True)
    return out


if __name__ == "__main__":
    res = config1()
    with open(os.path.join(ROOT, "tests", "golden", "config1.json"), "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1)[:1500])
