"""Loads the golden fixtures (tests/golden/, produced by running the reference itself)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["g1_random", "g2_overflow", "g3_kat", "g4_wide"]


def load_case(name):
    meta = json.load(open(os.path.join(GOLDEN, name + ".json")))
    meta["db_fasta"] = os.path.join(GOLDEN, name + "_db.fasta.gz")
    meta["q_fasta"] = os.path.join(GOLDEN, name + "_q.fasta.gz")
    for run in meta["runs"]:
        run["score_matrix"] = np.load(os.path.join(GOLDEN, run["scores"]))
    return meta
