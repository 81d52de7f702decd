"""world_size-2 test of the multi-rank path on CPU (gloo): every rank builds ITS shard of the
chunk streams with the product's host code (osw_shard_build), scores only its own sequences
(here with the oracle standing in for the GPU - this is a test of the sharding, gathering and
host-side top-r merge, not of the kernels), the ranks gather their r hits per query and rank 0
merges them with osw_merge_hits.  The merged lists must equal the unsharded ranking.  Every rank also
takes its shard from the one X.osw file the ranks share: it must be the shard built from the arrays."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle_lib as O
    import oswald_b200 as ob
    from oswald_b200 import capi
    from oswald_b200.host import merge_hits
    from test_host import Shard, build_shard

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(2026)                       # same database on every rank
    aa = np.array([0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 21], dtype=np.uint8)
    lens = rng.integers(5, 120, size=900).astype(np.uint64)
    db = ob.Database.from_lengths(lens, aa[rng.integers(0, 20, size=int(lens.sum()))])
    queries = ob.Queries.from_list([aa[rng.integers(0, 20, size=m)] for m in (30, 64)])
    mat = O.matrix("blosum62")
    top = 12
    L = capi.lib()
    shard = build_shard(L, db, rank, world, 512)
    local = [int(shard.canon[i]) for i in range(shard.n_seqs)]
    # the same shard from the ONE X.osw file all ranks share (rank 0 writes it)
    import ctypes as C
    from test_host import DbFile, shard_fields
    path = os.path.join(out_dir, "db.osw")
    if rank == 0:
        ob.write_db_file(path, db, max_chunk_residues=512)
    dist.barrier()
    L.osw_dbfile_open.argtypes = [C.c_char_p, C.POINTER(DbFile)]
    L.osw_shard_from_file.argtypes = [C.POINTER(DbFile), C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(Shard)]
    f, from_file = DbFile(), Shard()
    assert L.osw_dbfile_open(path.encode(), C.byref(f)) == 0
    assert L.osw_shard_from_file(C.byref(f), rank, world, None, None, C.byref(from_file)) == 0
    assert shard_fields(from_file) == shard_fields(build_shard(L, db, rank, world, f.h.chunk_cols))      # (the file's own work-unit size)
    # this rank's score rows over its own sequences only
    hits = []
    for q in range(queries.n):
        sc = np.array([O.sw_score(queries.query(q), db.sequence(i), mat, 10, 2) for i in local], dtype=np.int32)
        idx, val = O.top_r(sc, top)
        # local order -> canonical indices; ties must still be by canonical index (ascending local == ascending canonical)
        hits.append([(int(v), local[int(i)]) for v, i in zip(val, idx)])
    assert local == sorted(local)
    parts = [None] * world
    dist.all_gather_object(parts, hits)
    n_local = [None] * world
    dist.all_gather_object(n_local, (int(shard.n_seqs), int(shard.n_residues)))
    if rank == 0:
        assert sum(n for n, _ in n_local) == db.n_seqs
        res = [r for _, r in n_local]
        assert max(res) - min(res) <= 2 * 512 + int(lens.max())
        want = O.search(queries.residues, queries.offsets, db.residues, db.offsets, mat, 10, 2)
        for q in range(queries.n):
            merged = merge_hits([p[q] for p in parts], top)
            idx, val = O.top_r(want[q], top)
            assert merged == [(int(v), int(i)) for v, i in zip(val, idx)]
        open(os.path.join(out_dir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_merge(built, tmp_path):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
