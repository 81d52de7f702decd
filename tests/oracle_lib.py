"""ctypes binding to oracle/liboswald_oracle.so (the CPU parity checker).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg
may import this module.  Nothing under oswald_b200/ does.
"""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(ROOT, "oracle", "liboswald_oracle.so")
        if not os.path.exists(path):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboswald_oracle.so"])
        L = C.CDLL(path)
        L.orc_encode.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
        L.orc_matrix.argtypes = [C.c_char_p, C.c_void_p]
        L.orc_matrix.restype = C.c_int
        L.orc_sort_by_length.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.orc_sw_score.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.orc_sw_score.restype = C.c_int32
        L.orc_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t,
                                 C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_top_r.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]
        L.orc_top_r.restype = C.c_size_t
        L.orc_ref_mergesort.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def encode(letters):
    if isinstance(letters, str):
        letters = letters.encode()
    out = np.empty(len(letters), dtype=np.uint8)
    lib().orc_encode(letters, len(letters), _p(out))
    return out


def matrix(name):
    out = np.zeros(24 * 32, dtype=np.int8)
    if lib().orc_matrix(name.encode(), _p(out)) != 0:
        raise KeyError(name)
    return out


def sort_by_length(lengths):
    lengths = np.ascontiguousarray(lengths, dtype=np.uint32)
    perm = np.empty(len(lengths), dtype=np.uint32)
    lib().orc_sort_by_length(_p(lengths), len(lengths), _p(perm))
    return perm


def sw_score(a, b, mat, go, ge):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    return int(lib().orc_sw_score(_p(a), len(a), _p(b), len(b), _p(mat), go, ge))


def search(queries, q_off, db, db_off, mat, go, ge, threads=None):
    """scores[nq, n_seqs] int32 (canonical order), every pair scored."""
    queries = np.ascontiguousarray(queries, dtype=np.uint8)
    q_off = np.ascontiguousarray(q_off, dtype=np.uint32)
    db = np.ascontiguousarray(db, dtype=np.uint8)
    db_off = np.ascontiguousarray(db_off, dtype=np.uint64)
    nq, n = len(q_off) - 1, len(db_off) - 1
    out = np.zeros((nq, n), dtype=np.int32)
    lib().orc_search(_p(queries), _p(q_off), nq, _p(db), _p(db_off), n, _p(mat), go, ge, _p(out),
                     threads or os.cpu_count() or 1)
    return out


def top_r(row, r):
    row = np.ascontiguousarray(row, dtype=np.int32)
    k = min(r, len(row))
    idx = np.empty(k, dtype=np.uint32)
    sc = np.empty(k, dtype=np.int32)
    lib().orc_top_r(_p(row), len(row), r, _p(idx), _p(sc))
    return idx, sc


def ref_mergesort(row):
    s = np.array(row, dtype=np.int32)
    x = np.arange(len(s), dtype=np.uint32)
    lib().orc_ref_mergesort(_p(s), _p(x), len(s))
    return x, s


# ---- FASTA helpers (host-side semantics of the reference, used to feed the oracle) -----
def read_fasta(path):
    titles, seqs, cur = [], [], []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                if titles:
                    seqs.append("".join(cur))
                titles.append(line[1:])
                cur = []
            else:
                cur.append(line)
    if titles:
        seqs.append("".join(cur))
    return titles, seqs


def read_fasta_gz(path):
    import gzip, shutil, tempfile
    with tempfile.NamedTemporaryFile(suffix=".fasta") as tmp:
        with gzip.open(path, "rb") as g:
            shutil.copyfileobj(g, tmp)
        tmp.flush()
        return read_fasta(tmp.name)


def canonical(titles, seqs):
    """Reference preprocessing: stable ascending length sort, encode, concatenate."""
    lens = np.array([len(s) for s in seqs], dtype=np.uint32)
    perm = sort_by_length(lens)
    titles = [titles[i] for i in perm]
    seqs = [seqs[i] for i in perm]
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens[perm])
    codes = encode("".join(seqs))
    return titles, codes, off
