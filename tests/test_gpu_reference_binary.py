"""GPU results against the reference BINARY itself, run on the GPU box's host cores
(oracle/_ref/oswald_ref: the reference's unmodified host sources behind the inert OpenCL shim,
built by oracle/Makefile in the build container and shipped with the snapshot).

TEST INFRASTRUCTURE.  Skipped when the binary is not there (it cannot be built on the GPU box:
/root/reference does not exist on it)."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

import bench
import oswald_b200 as ob

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "oswald_ref")


def run_reference(db, queries_list, run, top, dump=False):
    """Preprocess + search with the reference binary; returns (printed top lists, titles, raw score matrix or None)."""
    name, go, ge = run
    seqs = [db.sequence(i) for i in range(db.n_seqs)]
    with tempfile.TemporaryDirectory() as tmp:
        bench.write_fasta(os.path.join(tmp, "db.fasta"), seqs, "s")
        bench.write_fasta(os.path.join(tmp, "q.fasta"), queries_list, "q")
        subprocess.run([REF, "-O", "preprocess", "-i", "db.fasta", "-o", "db", "-c", "4"], cwd=tmp, check=True, capture_output=True)
        env = dict(os.environ)
        if dump:
            env["OSWALD_ORACLE_DUMP"] = os.path.join(tmp, "dump.bin")
        out = subprocess.run([REF, "-O", "search", "-q", "q.fasta", "-d", "db", "-m", "1", "-v", "32", "-c", str(os.cpu_count() or 1),
                              "-p", "0.2", "-r", str(top), "-s", name, "-g", str(go), "-e", str(ge)], cwd=tmp, check=True,
                             capture_output=True, env=env, timeout=1500).stdout.decode(errors="replace")
        titles = [l.rstrip("\n").rstrip("\x00")[1:] for l in open(os.path.join(tmp, "db.desc"), errors="replace")]
        raw = np.fromfile(os.path.join(tmp, "dump.bin"), dtype=np.int32).reshape(len(queries_list), db.n_seqs) if dump else None
    tops = []
    for b in out.split("Query no.")[1:]:
        lines = b.split("Score\tSequence description\n")[1].split("\n")[:top]
        tops.append([(int(l.split("\t")[0]), l.split("\t")[1].rstrip("\x00")) for l in lines if "\t" in l])
    return tops, titles, raw


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/oswald_ref not built")
def test_config1_top10_equals_the_reference_binarys(built):
    """BASELINE.json config 1 at full size (144-residue query vs 10 000 sequences, BLOSUM62 10/2):
    the printed top 10 of the reference's host AVX2 path."""
    wl = bench.make_workload(1)
    db, q = wl["db"], wl["queries"]
    tops, titles, _ = run_reference(db, [q.query(0)], wl["runs"][0], 10)
    with ob.Searcher(1) as s:
        s.load_db(db)
        hits, tm = s.search(q, ob.matrix("blosum62"), 10, 2, top=10)
    assert [(sc, titles[i]) for sc, i in hits[0]] == tops[0]


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/oswald_ref not built")
def test_every_raw_score_of_the_reference_binary(built):
    """20 queries x 20 000 Swiss-Prot-shaped sequences: all 400 000 scores the reference computes
    (dumped by the shim's sort_scores hook) and its ranking."""
    wl = bench.make_workload(2, n_override=20000)
    db, q = wl["db"], wl["queries"]
    _, _, want = run_reference(db, [q.query(i) for i in range(q.n)], wl["runs"][0], 10, dump=True)
    with ob.Searcher(1) as s:
        s.load_db(db)
        hits, tm, scores = s.search(q, ob.matrix("blosum62"), 10, 2, top=10, all_scores=True)
    assert int((scores != want).sum()) == 0
    assert bench.host_ranking(want, 10) == hits
