"""Host-side mirror of the reference's interface for the hot path (names follow the reference).

  encode                residue codes            reference sequences.c:163-175, :361-371
  preprocess_db         FASTA -> canonical DB    reference sequences.c:4-220 (stable length sort :1130-1225)
  load_query_sequences  query FASTA, sorted      reference sequences.c:223-391
  matrix                -s name -> 24x32 table   reference submat.c, arguments.c:94-111
  Searcher              init / db load / search  reference main.c:46-62 + hybrid_search_avx2()

All scoring happens in liboswald_cuda.so through the C ABI (capi.py).
"""
import ctypes as C
import numpy as np

from . import capi

ALPHABET = "ABCDEFGHIKLMNPQRSTVWXYZ"      # code -> letter; 23 = J/O/U/padding

_ENC = np.full(256, 255, dtype=np.uint8)
for _c in range(ord('A'), ord('Z') + 2):
    _x = ord('Z') + 1 if chr(_c) in "JOU" else _c
    _ENC[_c] = _x - ord('A') - (_x > ord('J')) - (_x > ord('O')) - (_x > ord('U'))


def encode(letters):
    """Residue codes of a letter string (reference sequences.c:163-175)."""
    if isinstance(letters, str):
        letters = letters.encode()
    return _ENC[np.frombuffer(letters, dtype=np.uint8)]


def read_fasta(path):
    """(titles, sequences) in file order; a title is the header line without '>'."""
    import gzip
    titles, seqs, cur = [], [], []
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rt") as f:
        for line in f:
            line = line.rstrip("\r\n")
            if line.startswith(">"):
                if titles:
                    seqs.append("".join(cur))
                titles.append(line[1:])
                cur = []
            elif titles:
                cur.append(line)
    if titles:
        seqs.append("".join(cur))
    return titles, seqs


class Database:
    """Canonical database: sequences in stable ascending length order (reference sequences.c:125)."""

    def __init__(self, residues, offsets, titles=None):
        self.residues = np.ascontiguousarray(residues, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        self.titles = titles
        self.n_seqs = len(self.offsets) - 1
        self.n_residues = int(self.offsets[-1]) if self.n_seqs >= 0 else 0

    @classmethod
    def from_lengths(cls, lengths, residues_in_order, titles=None, presorted=False):
        """lengths/residues in input (FASTA) order -> canonical order."""
        lengths = np.asarray(lengths, dtype=np.uint64)
        if presorted:
            perm = np.arange(len(lengths))
        else:
            perm = np.argsort(lengths, kind="stable")
        in_off = np.zeros(len(lengths) + 1, dtype=np.uint64)
        in_off[1:] = np.cumsum(lengths)
        out_off = np.zeros(len(lengths) + 1, dtype=np.uint64)
        out_off[1:] = np.cumsum(lengths[perm])
        if presorted:
            res = np.asarray(residues_in_order, dtype=np.uint8)
        else:
            res = np.empty(int(in_off[-1]), dtype=np.uint8)
            src = np.asarray(residues_in_order, dtype=np.uint8)
            # gather sequence by sequence through a flat index (vectorised)
            idx = np.repeat(in_off[:-1][perm].astype(np.int64) - out_off[:-1].astype(np.int64), lengths[perm].astype(np.int64))
            idx += np.arange(int(in_off[-1]), dtype=np.int64)
            res[:] = src[idx]
        t = [titles[i] for i in perm] if titles is not None else None
        return cls(res, out_off, t)

    def sequence(self, i):
        return self.residues[int(self.offsets[i]):int(self.offsets[i + 1])]


def preprocess_db(fasta_path):
    """FASTA -> canonical Database (the in-memory equivalent of `-O preprocess`)."""
    titles, seqs = read_fasta(fasta_path)
    lengths = np.array([len(s) for s in seqs], dtype=np.uint64)
    return Database.from_lengths(lengths, encode("".join(seqs)), titles)


class Queries:
    def __init__(self, residues, offsets, titles=None):
        self.residues = np.ascontiguousarray(residues, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        self.titles = titles
        self.n = len(self.offsets) - 1
        self.total_length = int(self.offsets[-1])

    def lengths(self):
        return np.diff(self.offsets.astype(np.int64))

    def query(self, i):
        return self.residues[int(self.offsets[i]):int(self.offsets[i + 1])]

    @classmethod
    def from_list(cls, seqs, titles=None, sort=True):
        """seqs: list of code arrays.  The reference sorts queries by length (sequences.c:342)."""
        order = np.argsort([len(s) for s in seqs], kind="stable") if sort else np.arange(len(seqs))
        seqs = [np.asarray(seqs[i], dtype=np.uint8) for i in order]
        off = np.zeros(len(seqs) + 1, dtype=np.uint32)
        off[1:] = np.cumsum([len(s) for s in seqs])
        res = np.concatenate(seqs) if seqs else np.zeros(0, dtype=np.uint8)
        return cls(res, off, [titles[i] for i in order] if titles is not None else None)


def load_query_sequences(fasta_path):
    titles, seqs = read_fasta(fasta_path)
    return Queries.from_list([encode(s) for s in seqs], titles)


def matrix_names():
    L = capi.lib()
    return [L.osw_matrix_name(k).decode() for k in range(L.osw_matrix_count())]


def matrix(name):
    out = np.zeros(24 * 32, dtype=np.int8)
    if capi.lib().osw_matrix_by_name(name.encode(), out.ctypes.data_as(C.c_void_p)) != 0:
        raise KeyError("unknown substitution matrix %r" % name)
    return out


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


class Searcher:
    """One context = this process's GPUs.  shard_rank/shard_count split the database between
    processes (one process per GPU under torchrun); a single process uses the defaults."""

    def __init__(self, n_gpus=1, devices=None):
        self._L = capi.lib()
        self._ctx = C.c_void_p()
        dev = None
        if devices is not None:
            n_gpus = len(devices)
            dev = (C.c_int * n_gpus)(*devices)
        capi.check(self._L.osw_init(n_gpus, dev, C.byref(self._ctx)), "osw_init")
        self.n_seqs = 0
        self.n_gpus = n_gpus

    def close(self):
        if self._ctx:
            self._L.osw_free(self._ctx)
            self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_kernels(self, mask):
        capi.check(self._L.osw_set_kernels(self._ctx, mask), "osw_set_kernels")

    def set_device_window(self, n_bytes):
        """Stream databases larger than n_bytes per GPU through two device windows (0 = resident)."""
        capi.check(self._L.osw_set_device_window(self._ctx, n_bytes), "osw_set_device_window")

    def load_db(self, db, shard_rank=0, shard_count=1, max_chunk_residues=0):
        capi.check(self._L.osw_db_load(self._ctx, _vp(db.residues), _vp(db.offsets), db.n_seqs,
                                       shard_rank, shard_count, max_chunk_residues), "osw_db_load")
        self.n_seqs = db.n_seqs

    def load_db_file(self, path, shard_rank=0, shard_count=1):
        """Load a database from its X.osw file (written by write_db_file / `-O preprocess`)."""
        capi.check(self._L.osw_db_load_file(self._ctx, str(path).encode(), shard_rank, shard_count), "osw_db_load_file")
        self.n_seqs = db_file_info(path)["n_seqs"]

    def upload_db(self):
        """Host -> device copy of the resident chunk streams again; returns bytes copied."""
        n = C.c_uint64()
        capi.check(self._L.osw_db_upload(self._ctx, C.byref(n)), "osw_db_upload")
        return n.value

    def stats(self):
        a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
        capi.check(self._L.osw_db_stats(self._ctx, C.byref(a), C.byref(b), C.byref(c)), "osw_db_stats")
        return {"n_seqs": a.value, "residues": b.value, "chunks": c.value}

    def search(self, queries, mat, gap_open=10, gap_extend=2, top=10, all_scores=False):
        """Returns (hits, timing[, scores]): hits[q] = list of (score, canonical index)."""
        nq = queries.n
        hits = (capi.OswHit * max(1, nq * top))()
        n_hits = np.zeros(nq, dtype=np.uint32)
        scores = np.zeros((nq, self.n_seqs), dtype=np.int32) if all_scores else None
        tm = capi.OswTiming()
        mat = np.ascontiguousarray(mat, dtype=np.int8)
        capi.check(self._L.osw_search(self._ctx, _vp(queries.residues), _vp(queries.offsets), nq, _vp(mat),
                                      gap_open, gap_extend, top, C.cast(hits, C.c_void_p), _vp(n_hits),
                                      _vp(scores) if all_scores else None, C.byref(tm)), "osw_search")
        out = [[(hits[q * top + k].score, hits[q * top + k].index) for k in range(int(n_hits[q]))] for q in range(nq)]
        if all_scores:
            return out, tm.as_dict(), scores
        return out, tm.as_dict()


def write_db_file(path, db, max_chunk_residues=0):
    """The canonical database in its device layout, on disk (X.osw).  Host only: needs no GPU."""
    capi.check(capi.lib().osw_db_write_file(str(path).encode(), _vp(db.residues), _vp(db.offsets), db.n_seqs, max_chunk_residues),
               "osw_db_write_file")


def db_file_info(path):
    a, b, c, m, v = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint32(), C.c_uint32()
    capi.check(capi.lib().osw_db_file_info(str(path).encode(), C.byref(a), C.byref(b), C.byref(c), C.byref(m), C.byref(v)), "osw_db_file_info")
    return {"n_seqs": a.value, "n_residues": b.value, "n_chunks": c.value, "max_len": m.value, "version": v.value}


def merge_hits(lists, top):
    """Host-side top-r merge of several shards' hit lists (reference order)."""
    L = capi.lib()
    n = len(lists)
    arrs = [(capi.OswHit * max(1, len(l)))(*[capi.OswHit(s, i) for s, i in l]) for l in lists]
    ptrs = (C.c_void_p * n)(*[C.cast(a, C.c_void_p) for a in arrs])
    counts = (C.c_uint32 * n)(*[len(l) for l in lists])
    out = (capi.OswHit * max(1, top))()
    k = L.osw_merge_hits(ptrs, counts, n, top, C.cast(out, C.c_void_p))
    return [(out[i].score, out[i].index) for i in range(k)]


def calibrate(device=0):
    out = (C.c_double * 12)()
    capi.check(capi.lib().osw_calibrate(device, out), "osw_calibrate")
    return {"viaddmnmx_u16x2_per_sm_clk": out[0], "vimnmx3_u16x2_per_sm_clk": out[1],
            "cells_per_sm_clk_step_imad": out[2], "imad_per_sm_clk": out[3], "sm_mhz": out[4],
            "cells_per_sm_clk_step_plain_sub": out[5], "cells_per_sm_clk_step_with_lds": out[6],
            "cells_per_sm_clk_step_signed_relu": out[7], "cells_per_sm_clk_step_signed_relu_nomax": out[8],
            "cells_per_sm_clk_step_nomax": out[9], "cells_per_sm_clk_step_clamped_e": out[10]}


MIX_CLASSES = ["VIADDMNMX.U16x2", "VIMNMX3.U16x2", "VIADD", "IMAD", "HMNMX2", "VIMNMX.U16x2", "VIMNMX.U32", "LOP3",
               "FMNMX", "PRMT", "SHF", "HADD2", "IMAD.HI", "LEA.HI", "IMAD(x65536)"]


def calibrate_mix(device=0):
    """Issue rates (thread instructions per SM-cycle): alone / with 8 DPX + 8 / with 8 DPX + 4."""
    out = (C.c_double * 45)()
    capi.check(capi.lib().osw_calibrate_mix(device, out, 45), "osw_calibrate_mix")
    return {name: {"alone": out[3 * k], "dpx8_plus8_total": out[3 * k + 1], "dpx8_plus4_total": out[3 * k + 2]}
            for k, name in enumerate(MIX_CLASSES)}
