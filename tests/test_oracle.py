"""CPU tests: the oracle (oracle/sw_oracle.c) against the golden vectors produced by the
reference's own host AVX2 path, and against the known answers of SURVEY.md section 8(c)."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_lib as O
from golden_util import CASES, GOLDEN, load_case
import oswald_b200 as ob


def _inputs(meta):
    titles, seqs = O.read_fasta_gz(meta["db_fasta"])
    ctitles, codes, off = O.canonical(titles, seqs)
    qt, qs = O.read_fasta_gz(meta["q_fasta"])
    order = np.argsort([len(s) for s in qs], kind="stable")          # reference sorts queries too
    qcodes = [O.encode(qs[i]) for i in order]
    q_off = np.zeros(len(qs) + 1, dtype=np.uint32)
    q_off[1:] = np.cumsum([len(c) for c in qcodes])
    return ctitles, codes, off, np.concatenate(qcodes), q_off


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_scores(name):
    meta = load_case(name)
    ctitles, codes, off, q, q_off = _inputs(meta)
    assert len(ctitles) == meta["n_seqs"]
    if meta["desc"] is not None:
        assert ctitles == meta["desc"], "canonical order differs from the reference's .desc"
    for run in meta["runs"]:
        got = O.search(q, q_off, codes, off, O.matrix(run["matrix"]), run["gap_open"], run["gap_extend"])
        assert np.array_equal(got, run["score_matrix"]), (name, run["matrix"])
        for qi, h in enumerate(run["hits"]):
            assert h["query_length"] == int(q_off[qi + 1] - q_off[qi])
            idx, sc = O.top_r(run["score_matrix"][qi], meta["top"])
            assert [[int(s), ctitles[i]] for s, i in zip(sc, idx)] == h["top"]


def test_known_answers():
    """values printed by the reference binary."""
    kat = {("HEAGAWGHEE", "PAWHEAE"): (17, 24, 26, 23),
           ("HEAGAWGHEE", "ACDEFGHIKLMNPQRSTVWY"): (15, 19, 15, 18),
           ("ACDEFGHIKLMNPQRSTVWY", "ACDEFGHIKLMNPQRSTVWY"): (116, 150, 164, 141),
           ("W" * 10 + "H" * 10, "W" * 10 + "AAA" + "H" * 10): (174, 234, 208, 230),
           ("W" * 10 + "H" * 10, "W" * 10 + "A" + "H" * 10): (180, 238, 210, 238),
           ("W" * 10 + "H" * 10, "PAWHEAE"): (19, 25, 22, 25)}
    settings = [("blosum62", 10, 2), ("blosum50", 10, 2), ("pam30", 9, 1), ("blosum45", 14, 2)]
    for (a, b), want in kat.items():
        for (mat, go, ge), w in zip(settings, want):
            assert O.sw_score(O.encode(a), O.encode(b), O.matrix(mat), go, ge) == w, (a, b, mat)


def test_matrices_pinned():
    pins = json.load(open(os.path.join(GOLDEN, "submat.json")))
    assert sorted(pins) == sorted(ob.matrix_names())
    for name, pin in pins.items():
        for m in (O.matrix(name), ob.matrix(name)):       # oracle copy and product copy
            assert hashlib.sha256(m.tobytes()).hexdigest() == pin["sha256"], name
            assert int(m.min()) == pin["min"] and int(m.max()) == pin["max"]


def test_ranking_is_the_reference_merge_sort():
    rng = np.random.default_rng(5)
    for n in (1, 2, 3, 7, 64, 1000, 4097):
        row = rng.integers(0, 12, size=n).astype(np.int32)           # many ties
        order, sc = O.ref_mergesort(row)
        idx, s2 = O.top_r(row, n)
        assert np.array_equal(order, idx) and np.array_equal(sc, s2)


def test_length_sort_is_stable():
    rng = np.random.default_rng(6)
    lens = rng.integers(1, 40, size=5000).astype(np.uint32)
    assert np.array_equal(O.sort_by_length(lens), np.argsort(lens, kind="stable"))


def test_encoding():
    s = "ABCDEFGHIJKLMNOPQRSTUVWXYZ"
    want = [0, 1, 2, 3, 4, 5, 6, 7, 8, 23, 9, 10, 11, 12, 23, 13, 14, 15, 16, 17, 23, 18, 19, 20, 21, 22]
    assert list(O.encode(s)) == want
    assert list(ob.encode(s)) == want
