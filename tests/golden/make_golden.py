#!/usr/bin/env python3
"""Generate the golden fixtures in tests/golden/ by RUNNING THE REFERENCE ITSELF.

Run in the build container (needs /root/reference; `make -C oracle ref` builds
oracle/_ref/oswald_ref from the reference's own host/src, compiled in place).  For each case
the reference preprocesses a small synthetic FASTA database and searches it with its host
AVX2 path (-m 1 -v 32); the shim's sort_scores hook dumps every raw score row.  Stored per
case: the FASTA inputs (gzip), the parameters, the full int32 score matrix in canonical
order (.npy), the top-r lines the reference printed, and the .desc title order.

usage: python tests/golden/make_golden.py
"""
import gzip, json, os, re, shutil, subprocess, sys, tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "oswald_ref")
SYNTH = os.path.join(ROOT, "tools", "osw_synth")

KAT_DB = ["PAWHEAE", "ACDEFGHIKLMNPQRSTVWY", "W" * 10 + "AAA" + "H" * 10, "W" * 10 + "A" + "H" * 10]
KAT_Q = ["HEAGAWGHEE", "ACDEFGHIKLMNPQRSTVWY", "W" * 10 + "H" * 10]


def sh(*a, **kw):
    return subprocess.run(a, check=True, capture_output=True, text=True, **kw).stdout


def run_case(name, db_fasta, q_fasta, settings, top, pct):
    """settings: list of (matrix, go, ge).  Returns metadata dict; writes fixture files."""
    tmp = tempfile.mkdtemp()
    try:
        sh(REF, "-O", "preprocess", "-i", db_fasta, "-o", os.path.join(tmp, "db"), "-c", "2")
        n_seqs = int(open(os.path.join(tmp, "db.info")).read().split()[0])
        desc = [l.rstrip("\n").rstrip("\x00")[1:] for l in open(os.path.join(tmp, "db.desc"), errors="replace")]
        assert len(desc) == n_seqs
        meta = {"name": name, "n_seqs": n_seqs, "top": top, "runs": []}
        for k, (mat, go, ge) in enumerate(settings):
            dump = os.path.join(tmp, "dump%d.bin" % k)
            env = dict(os.environ, OSWALD_ORACLE_DUMP=dump)
            out = subprocess.run([REF, "-O", "search", "-q", q_fasta, "-d", os.path.join(tmp, "db"), "-m", "1",
                                  "-v", "32", "-c", "2", "-p", str(pct), "-r", str(top), "-s", mat,
                                  "-g", str(go), "-e", str(ge)], check=True, capture_output=True, env=env,
                                 timeout=600).stdout.decode(errors="replace")
            scores = np.fromfile(dump, dtype=np.int32).reshape(-1, n_seqs)
            np.save(os.path.join(HERE, "%s_%s_%d_%d.npy" % (name, mat, go, ge)), scores)
            blocks = out.split("Query no.")[1:]
            hits = []
            for b in blocks:
                qlen = int(re.search(r"Query length:\s+(\d+)", b).group(1))
                lines = b.split("Score\tSequence description\n")[1].split("\n")[:top]
                hits.append({"query_length": qlen,
                             "top": [[int(l.split("\t")[0]), l.split("\t")[1].rstrip("\x00")] for l in lines if "\t" in l]})
            meta["runs"].append({"matrix": mat, "gap_open": go, "gap_extend": ge,
                                 "scores": "%s_%s_%d_%d.npy" % (name, mat, go, ge), "hits": hits})
        for src, dst in ((db_fasta, name + "_db.fasta.gz"), (q_fasta, name + "_q.fasta.gz")):
            with open(src, "rb") as f, gzip.GzipFile(os.path.join(HERE, dst), "wb", mtime=0) as g:
                shutil.copyfileobj(f, g)
        meta["desc"] = desc if n_seqs <= 4000 else None
        json.dump(meta, open(os.path.join(HERE, name + ".json"), "w"), indent=1)
        print(name, "ok:", n_seqs, "sequences,", len(settings), "runs")
    finally:
        shutil.rmtree(tmp)


def append_records(path, recs):
    with open(path, "a") as f:
        for t, s in recs:
            f.write(">%s\n" % t)
            for k in range(0, len(s), 60):
                f.write(s[k:k + 60] + "\n")


def case_g4(w):
    """G4: scores beyond 16 bits - the reference's own 32-bit stage (HybridSearch.c:1046-1134) pins
    what the GPU's 32-bit re-score returns.  Tryptophan-rich queries and database sequences under
    PAM30 9/1 (W/W = 13): exact and mutated copies, all-W runs, and pairs on both sides of the packed
    16-bit kernel's flag threshold (65 504 - bias)."""
    rng = np.random.default_rng(44)
    aa = "ACDEFGHIKLMNPQRSTVWY"

    def rand(n):
        return "".join(aa[i] for i in rng.integers(0, 20, size=n))

    def w_rich(n, frac):
        return "".join("W" if x < frac else aa[i] for x, i in zip(rng.random(n), rng.integers(0, 20, size=n)))

    q_rich = w_rich(5400, 0.93)
    q_all = "W" * 5300
    queries = [("q_rand144", rand(144)), ("q_allW5300", q_all), ("q_rich5400", q_rich), ("q_rand3000", rand(3000))]
    open(w + "/q4.fasta", "w").close()
    append_records(w + "/q4.fasta", queries)
    sh(SYNTH, "db", "-n", "600", "-mu", "4.8", "-sigma", "0.6", "-seed", "23", "-o", w + "/db4.fasta")
    recs = [("copy_rich", q_rich), ("rich_indel", q_rich[:2600] + rand(7) + q_rich[2600:]), ("rich_other", w_rich(6000, 0.9)),
            ("copy_rand3000", queries[3][1])]
    recs += [("allW_%d" % n, "W" * n) for n in (5020, 5030, 5035, 5036, 5037, 5040, 5100, 5300, 7000, 20000)]
    recs += [("allW_rand_%d" % k, "W" * int(n)) for k, n in enumerate(rng.integers(4800, 9000, size=12))]
    append_records(w + "/db4.fasta", recs)
    run_case("g4_wide", w + "/db4.fasta", w + "/q4.fasta", [("pam30", 9, 1), ("blosum62", 10, 2)], 15, 0.3)


def main():
    if "--only-g4" in sys.argv:
        w = tempfile.mkdtemp()
        case_g4(w)
        shutil.rmtree(w)
        return
    if not os.path.exists(REF):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    if not os.path.exists(SYNTH):
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-o", SYNTH, os.path.join(ROOT, "tools", "osw_synth.c"), "-lm"])
    w = tempfile.mkdtemp()
    # G1: plain random database, two queries, default scoring; many tied scores.
    sh(SYNTH, "queries", "-lengths", "144,464", "-seed", "7", "-o", w + "/q1.fasta")
    sh(SYNTH, "db", "-n", "2000", "-mu", "5.0", "-sigma", "0.7", "-seed", "11", "-o", w + "/db1.fasta")
    run_case("g1_random", w + "/db1.fasta", w + "/q1.fasta", [("blosum62", 10, 2)], 25, 0.2)
    # G2: overflow cascade - planted copies of long queries, a titin-length random sequence and a
    # tandem repeat of the longest query (scores beyond 32767), three scoring systems.
    sh(SYNTH, "queries", "-lengths", "144,1000,3005,5478", "-seed", "5", "-o", w + "/q2.fasta")
    sh(SYNTH, "db", "-n", "1500", "-mu", "4.8", "-sigma", "0.6", "-seed", "13", "-plant", w + "/q2.fasta",
       "-plantmin", "1000", "-long", "1", "-longmin", "36000", "-longmax", "36000",
       "-tandem", w + "/q2.fasta", "-copies", "6", "-o", w + "/db2.fasta")
    run_case("g2_overflow", w + "/db2.fasta", w + "/q2.fasta",
             [("pam30", 9, 1), ("blosum45", 14, 2), ("blosum62", 10, 2)], 12, 0.3)
    # G3: known-answer sequences of SURVEY.md section 8(c) appended to a random database,
    # all eight matrices at their usual penalties.
    sh(SYNTH, "db", "-n", "1200", "-mu", "4.6", "-sigma", "0.5", "-seed", "17", "-o", w + "/db3.fasta")
    append_records(w + "/db3.fasta", [("kat%d" % i, s) for i, s in enumerate(KAT_DB)])
    open(w + "/q3.fasta", "w").close()
    append_records(w + "/q3.fasta", [("katq%d" % i, s) for i, s in enumerate(KAT_Q)])
    run_case("g3_kat", w + "/db3.fasta", w + "/q3.fasta",
             [("blosum62", 10, 2), ("blosum50", 10, 2), ("pam30", 9, 1), ("blosum45", 14, 2),
              ("blosum80", 10, 2), ("blosum90", 10, 2), ("pam70", 10, 1), ("pam250", 12, 2)], 8, 0.3)
    case_g4(w)
    shutil.rmtree(w)


if __name__ == "__main__":
    main()
