"""GPU: the C++ driver end to end on BASELINE config 1 — sketch files on disk, query/profile TSV.
Checks (i) the reference's own self-consistency property: output from raw files == output from
pre-sketched files (tests/integration_test.rs:286-292), (ii) rows == the oracle's golden rows,
(iii) the .syldb/.sylsp files it writes decode to the oracle's sketches."""
import json
import os
import subprocess

import numpy as np
import pytest

from tests.util import DATA, REPO, flatten, read_fastx

pytestmark = pytest.mark.gpu
GENOMES = ["e.coli-EC590.fasta.gz", "e.coli-o157.fasta.gz", "e.coli-K12.fasta.gz"]


@pytest.fixture(scope="module")
def exe():
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    subprocess.check_call(["make", "-C", os.path.join(REPO, "host"), "-s"], env=env)
    return os.path.join(REPO, "host", "sylph-b200")


def run(exe, args, cwd):
    r = subprocess.run([exe] + args, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    return r.stdout


def test_driver_config1(exe, tmp_path):
    from oracle import oracle as O
    from sylph_b200 import formats as F
    gold = json.load(open(os.path.join(REPO, "tests", "golden", "config1.json")))
    out = str(tmp_path)
    run(exe, ["sketch"] + GENOMES + ["o157_reads.fastq.gz", "-o", out + "/db", "-d", out], DATA)
    db = F.read_syldb(out + "/db.syldb")
    assert [g["file_name"] for g in db] == GENOMES
    for g, gg in zip(db, gold["genomes"]):
        assert (len(g["genome_kmers"]), len(g["tracked"]), g["gn_size"]) == (gg["n_kmers"], gg["n_tracked"], gg["gn_size"])
        assert g["first_contig_name"] == gg["first_contig"] and (g["c"], g["k"], g["min_spacing"]) == (200, 31, 30)
        recs = read_fastx(os.path.join(DATA, g["file_name"]))
        km, tr, _ = O.sketch_genome(*flatten([s for _, s in recs]))
        assert np.array_equal(g["genome_kmers"], km) and np.array_equal(g["tracked"], tr)
    sp = F.read_sylsp(out + "/o157_reads.fastq.gz.sylsp")
    assert len(sp["hashes"]) == gold["reads"]["n_keys"] and int(sp["counts"].sum()) == gold["reads"]["sum_counts"]
    assert abs(sp["mean_read_length"] - gold["reads"]["mean_read_length"]) < 1e-6 and not sp["paired"]

    def rows(txt):
        lines = txt.strip().split("\n")
        assert lines[0].startswith("Sample_file\tGenome_file")
        return lines[1:]

    # profile vs EC590 -> 1 row; query vs the three -> 3 rows (tests/integration_test.rs:117-140)
    p1 = rows(run(exe, ["profile", "o157_reads.fastq.gz", GENOMES[0]], DATA))
    q3 = rows(run(exe, ["query", "o157_reads.fastq.gz"] + GENOMES, DATA))
    p3 = rows(run(exe, ["profile", "o157_reads.fastq.gz"] + GENOMES, DATA))
    assert p1 == gold["profile_vs_EC590"] and q3 == gold["query_vs_all"] and p3 == gold["profile_vs_all"]
    # raw inputs == pre-sketched inputs (sample name = the file_name stored in the sketch)
    q3s = rows(run(exe, ["query", out + "/o157_reads.fastq.gz.sylsp", out + "/db.syldb"], DATA))
    p3s = rows(run(exe, ["profile", out + "/o157_reads.fastq.gz.sylsp", out + "/db.syldb", "-o", out + "/p.tsv"], DATA) or
               open(out + "/p.tsv").read())
    assert q3s == q3 and p3s == p3
