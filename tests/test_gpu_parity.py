"""GPU parity tests (B200): every score and every top-r list through the C ABI must equal the
reference's own output (golden fixtures) and the oracle, bit for bit."""
import numpy as np
import pytest

import oracle_lib as O
from golden_util import CASES, load_case
import oswald_b200 as ob
from oswald_b200 import capi

pytestmark = pytest.mark.gpu
AA = np.array([0, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 21], dtype=np.uint8)


@pytest.fixture(scope="module")
def searcher(built):
    s = ob.Searcher(1)
    yield s
    s.close()


def rand_seqs(rng, n, lo, hi):
    return [AA[rng.integers(0, 20, size=int(l))] for l in rng.integers(lo, hi, size=n)]


def make_db(seqs):
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)
    return ob.Database.from_lengths(lens, np.concatenate(seqs) if len(seqs) else np.zeros(0, np.uint8))


def oracle_scores(q, db, name, go, ge):
    return O.search(q.residues, q.offsets, db.residues, db.offsets, O.matrix(name), go, ge)


def check(searcher, db, q, name, go, ge, top, mask=capi.OSW_K_DEFAULT, want=None):
    searcher.set_kernels(mask)
    hits, tm, scores = searcher.search(q, ob.matrix(name), go, ge, top=top, all_scores=True)
    if want is None:
        want = oracle_scores(q, db, name, go, ge)
    bad = np.argwhere(scores != want)
    assert bad.size == 0, "first mismatches (query, index): %s got %s want %s" % (
        bad[:5].tolist(), scores[tuple(bad[:5].T)].tolist(), want[tuple(bad[:5].T)].tolist())
    for qi in range(q.n):
        idx, sc = O.top_r(want[qi], top)
        assert hits[qi] == [(int(s), int(i)) for s, i in zip(sc, idx)], "top-r of query %d" % qi
    return tm


@pytest.mark.parametrize("mask", [capi.OSW_K_I32, capi.OSW_K_DEFAULT, capi.OSW_K_DEFAULT | capi.OSW_K_TWO_TRACK,
                                  capi.OSW_K_DEFAULT | capi.OSW_K_PAIR_DB])
@pytest.mark.parametrize("name", CASES)
def test_golden_cases(searcher, name, mask):
    """Score matrices and printed top-r lists of the reference's host AVX2 path."""
    meta = load_case(name)
    db = ob.preprocess_db(meta["db_fasta"])
    q = ob.load_query_sequences(meta["q_fasta"])
    searcher.load_db(db)
    searcher.set_kernels(mask)
    for run in meta["runs"]:
        hits, tm, scores = searcher.search(q, ob.matrix(run["matrix"]), run["gap_open"], run["gap_extend"],
                                           top=meta["top"], all_scores=True)
        assert np.array_equal(scores, run["score_matrix"]), (name, run["matrix"])
        for qi, h in enumerate(run["hits"]):
            got = [[s, db.titles[i]] for s, i in hits[qi]]
            assert got == h["top"], (name, run["matrix"], qi)
        if mask != capi.OSW_K_I32:
            # g2: the reference needed its 16- and 32-bit stages (scores up to 44 783 > 32 767); the biased
            # unsigned 16-bit kernel holds them, so nothing is re-scored.  g4 (up to 68 900): exactly the
            # pairs at or above the flag threshold go to the GPU's 32-bit stage, and what comes back is
            # what the reference's own 32-bit stage (HybridSearch.c:1046-1134) computed.
            bias = run["gap_open"] + 2 * run["gap_extend"] + 32
            assert tm["rescored_pairs"] == int((run["score_matrix"] + bias >= 65504).sum())
            assert name != "g4_wide" or run["matrix"] != "pam30" or tm["rescored_pairs"] >= 20


MODES = {"auto": capi.OSW_K_DEFAULT, "two_track": capi.OSW_K_DEFAULT | capi.OSW_K_TWO_TRACK,
         "pair_db": capi.OSW_K_DEFAULT | capi.OSW_K_PAIR_DB,
         # database residues as rows, the query as the column stream (sw_t16.cu); falls back to the passes
         # above for queries that do not fit it, and is what "auto" picks for short queries on small databases
         "transposed": capi.OSW_K_DEFAULT | capi.OSW_K_TRANSPOSED}


@pytest.mark.parametrize("mode", ["auto", "two_track", "pair_db", "transposed"])
@pytest.mark.parametrize("lengths", [[144], [5, 37, 144, 189], [1, 1, 2], [144, 189, 222, 375, 464, 567, 657],
                                      [1000, 1500], [2005], [1537, 3005]])
def test_random_db_all_query_geometries(searcher, lengths, mode):
    """Every (G, R, passes) plan the query lengths select, odd and even query counts, with the two
    halves of the packed words used as two query tracks or as two database sequences."""
    rng = np.random.default_rng(sum(lengths))
    seqs = rand_seqs(rng, 700, 1, 500)
    q = ob.Queries.from_list([AA[rng.integers(0, 20, size=m)] for m in lengths])
    seqs[3] = q.query(q.n - 1).copy()                                  # an exact copy of the longest query
    seqs[5] = np.concatenate([q.query(0), rng.permutation(q.query(q.n - 1))])
    db = make_db(seqs)
    searcher.load_db(db)
    check(searcher, db, q, "blosum62", 10, 2, 10, mask=MODES[mode])


@pytest.mark.parametrize("name,go,ge", [("blosum45", 14, 2), ("blosum50", 10, 2), ("blosum80", 10, 2), ("blosum90", 10, 2),
                                        ("pam30", 9, 1), ("pam70", 10, 1), ("pam250", 12, 2), ("blosum62", 0, 0),
                                        ("blosum62", 255, 127), ("blosum62", 1, 0)])
def test_matrices_and_gap_penalties(searcher, name, go, ge):
    rng = np.random.default_rng(go * 131 + ge)
    seqs = rand_seqs(rng, 400, 1, 300)
    q = ob.Queries.from_list([AA[rng.integers(0, 20, size=m)] for m in (60, 144, 333)])
    for k in range(6):                                                  # homologs with indels
        src = q.query(k % 3)
        cut = rng.integers(1, len(src))
        seqs[k] = np.concatenate([src[:cut], AA[rng.integers(0, 20, size=rng.integers(0, 6))], src[cut:]])
    db = make_db(seqs)
    searcher.load_db(db)
    check(searcher, db, q, name, go, ge, 15)


def test_edge_shapes(searcher):
    rng = np.random.default_rng(9)
    # single sequence, single residue; database smaller than top; sequences of length 1
    q = ob.Queries.from_list([AA[rng.integers(0, 20, size=50)]])
    for seqs in ([AA[:1]], [AA[:1], AA[3:4], AA[5:9]], rand_seqs(rng, 33, 1, 3)):
        db = make_db(seqs)
        searcher.load_db(db)
        check(searcher, db, q, "blosum62", 10, 2, 10)
    # residues outside the 20 standard ones (B, X, Z and the J/O/U code 23) in the database
    seqs = [rng.integers(0, 24, size=int(l)).astype(np.uint8) for l in rng.integers(1, 200, size=300)]
    db = make_db(seqs)
    searcher.load_db(db)
    check(searcher, db, ob.Queries.from_list([AA[rng.integers(0, 20, size=m)] for m in (77, 200)]), "blosum62", 10, 2, 5)


def test_all_ties_rank_by_higher_index(searcher):
    seq = AA[[0, 1, 2, 3, 4, 5, 6, 7]]
    db = make_db([seq.copy() for _ in range(500)])
    searcher.load_db(db)
    q = ob.Queries.from_list([seq])
    hits, tm = searcher.search(q, ob.matrix("blosum62"), 10, 2, top=7)
    assert [i for _, i in hits[0]] == [499, 498, 497, 496, 495, 494, 493]
    assert len({s for s, _ in hits[0]}) == 1


def test_long_sequences_and_chunk_boundaries(searcher):
    """Sequences longer than a chunk (incl. the u16 maximum 65535) and tiny chunks."""
    rng = np.random.default_rng(21)
    q = ob.Queries.from_list([AA[rng.integers(0, 20, size=m)] for m in (144, 464)])
    seqs = rand_seqs(rng, 200, 1, 200) + [AA[rng.integers(0, 20, size=n)] for n in (4095, 4096, 4097, 9000, 36000, 65535)]
    seqs[-2][1000:1464] = q.query(1)
    db = make_db(seqs)
    want = oracle_scores(q, db, "blosum62", 10, 2)
    for k in (0, 64, 300):
        searcher.load_db(db, max_chunk_residues=k)
        for mode in MODES.values():
            check(searcher, db, q, "blosum62", 10, 2, 10, mask=mode, want=want)


def test_overflow_rescore_exact(searcher):
    """Scores beyond 16 bits (tryptophan-rich query against copies of itself, PAM30: 13 per
    W/W pair) are flagged by the 16-bit kernel and must come back exact from the 32-bit one,
    including scores just below / above the flag threshold and a wrapped-around maximum."""
    rng = np.random.default_rng(4)
    W = np.uint8(19)
    rich = np.where(rng.random(5478) < 0.9, W, AA[rng.integers(0, 20, size=5478)]).astype(np.uint8)
    mid = np.full(5000, W, dtype=np.uint8)             # self score 65 000: just below the threshold
    big = np.full(12000, W, dtype=np.uint8)            # 156 000 against itself: wraps twice
    q = ob.Queries.from_list([AA[rng.integers(0, 20, size=144)], mid, rich, big])
    seqs = rand_seqs(rng, 100, 50, 400) + [rich.copy(), np.concatenate([rich[:5200], AA[:9], rich[5200:]]),
                                           np.full(5036, W, dtype=np.uint8), np.full(5040, W, dtype=np.uint8),
                                           np.full(5050, W, dtype=np.uint8), np.full(20000, W, dtype=np.uint8), big.copy()]
    db = make_db(seqs)
    searcher.load_db(db)
    want = oracle_scores(q, db, "pam30", 9, 1)
    assert want.max() > 2 * 65535 and ((want > 64000) & (want < 65400)).any()
    for mode in MODES.values():
        tm = check(searcher, db, q, "pam30", 9, 1, 10, mask=mode, want=want)
        assert tm["rescored_pairs"] == int((want + 9 + 1 + 1 + 32 >= 65504).sum())      # bias = go + 2*ge + 32


def overflow_case(rng):
    """Tryptophan-rich queries and sequences under PAM30 9/1: dozens of pairs beyond 16 bits."""
    W = np.uint8(19)
    qs = [AA[rng.integers(0, 20, size=120)], np.full(5300, W, dtype=np.uint8),
          np.where(rng.random(5400) < 0.93, W, AA[rng.integers(0, 20, size=5400)]).astype(np.uint8)]
    seqs = rand_seqs(rng, 300, 20, 300) + [np.full(int(n), W, dtype=np.uint8) for n in rng.integers(5100, 9000, size=24)]
    return make_db(seqs), ob.Queries.from_list(qs)


def test_rescore_in_rounds_when_the_flag_list_is_small(built, monkeypatch):
    """More flagged pairs than the list holds: the 32-bit stage runs in rounds (re-score what is
    listed, scan again) instead of failing - any input the reference accepts is accepted.  Also
    with the database streamed through device windows (staged re-score)."""
    monkeypatch.setenv("OSW_FLAG_CAP", "5")
    rng = np.random.default_rng(66)
    db, q = overflow_case(rng)
    want = oracle_scores(q, db, "pam30", 9, 1)
    n_over = int((want + 9 + 1 + 1 + 32 >= 65504).sum())
    assert n_over > 3 * 5
    for window in (0, 1 << 20):
        with ob.Searcher(1) as s:
            s.set_device_window(window)
            s.load_db(db)
            for mode in MODES.values():
                tm = check(s, db, q, "pam30", 9, 1, 10, mask=mode, want=want)
                assert tm["rescored_pairs"] == n_over


def test_scores_only_and_streamed_file(built, tmp_path):
    """top = 0 (no hit list, every score on request): the flagged pairs are then found by a scan of their
    own and re-scored all the same; and a database loaded from X.osw into a context that streams it through
    device windows."""
    rng = np.random.default_rng(68)
    db, q = overflow_case(rng)
    want = oracle_scores(q, db, "pam30", 9, 1)
    path = tmp_path / "db.osw"
    ob.write_db_file(path, db)
    for window in (0, 1 << 20):
        with ob.Searcher(1) as s:
            s.set_device_window(window)
            s.load_db_file(path)
            hits, tm, scores = s.search(q, ob.matrix("pam30"), 9, 1, top=0, all_scores=True)
            assert np.array_equal(scores, want) and all(h == [] for h in hits)
            assert tm["rescored_pairs"] == int((want + 9 + 1 + 1 + 32 >= 65504).sum()) > 0
            check(s, db, q, "pam30", 9, 1, 10, want=want)


def test_empty_shard_after_an_overflow_search(built):
    """A context whose shard is empty (fewer chunks than shards) must not re-score from the flagged
    count a previous search left behind."""
    rng = np.random.default_rng(67)
    db, q = overflow_case(rng)
    tiny = make_db([AA[rng.integers(0, 20, size=50)]])
    with ob.Searcher(1) as s:
        s.load_db(db)
        tm = check(s, db, q, "pam30", 9, 1, 10)
        assert tm["rescored_pairs"] > 0
        s.load_db(tiny, shard_rank=1, shard_count=4)            # one chunk, dealt to shard 0: this shard holds nothing
        assert s.stats()["n_seqs"] == 0
        hits, tm = s.search(q, ob.matrix("pam30"), 9, 1, top=10)
        assert tm["rescored_pairs"] == 0 and all(h == [] for h in hits)
        s.load_db(tiny)
        check(s, tiny, q, "pam30", 9, 1, 10)


def test_sharded_contexts_merge_to_the_same_hits(built):
    """Two shards (as two ranks would hold them) merged on the host == one unsharded search."""
    from oswald_b200.host import merge_hits
    rng = np.random.default_rng(77)
    db = make_db(rand_seqs(rng, 3000, 20, 300))
    q = ob.Queries.from_list([AA[rng.integers(0, 20, size=m)] for m in (100, 200, 300)])
    want = oracle_scores(q, db, "blosum62", 10, 2)
    parts, all_scores = [], np.zeros_like(want)
    for rank in range(2):
        with ob.Searcher(1) as s:
            s.load_db(db, shard_rank=rank, shard_count=2, max_chunk_residues=512)
            st = s.stats()
            assert abs(st["residues"] - db.n_residues / 2) < 0.02 * db.n_residues
            hits, tm, sc = s.search(q, ob.matrix("blosum62"), 10, 2, top=10, all_scores=True)
            parts.append(hits)
            all_scores += sc            # untouched entries stay 0
    assert np.array_equal(all_scores, want)
    for qi in range(q.n):
        idx, sc = O.top_r(want[qi], 10)
        assert merge_hits([parts[0][qi], parts[1][qi]], 10) == [(int(s), int(i)) for s, i in zip(sc, idx)]


def test_swissprot_shape_properties(searcher):
    """A slice of the Swiss-Prot-shaped workload too big for the scalar oracle: the packed
    16-bit kernel and the 32-bit kernel (different code, different geometry) must agree on every
    score, and planted exact copies must score their self-score."""
    rng = np.random.default_rng(123)
    lens = np.clip(np.round(np.exp(rng.normal(5.6, 0.6, size=20000))), 10, 5000).astype(np.uint64)
    seqs = [AA[rng.integers(0, 20, size=int(l))] for l in lens]
    qs = [AA[rng.integers(0, 20, size=m)] for m in (144, 375, 1000, 2005)]
    for k, qq in enumerate(qs):
        seqs[100 + k] = qq.copy()
    db = make_db(seqs)
    q = ob.Queries.from_list(qs)
    searcher.load_db(db)
    searcher.set_kernels(capi.OSW_K_DEFAULT)
    h1, tm1, s1 = searcher.search(q, ob.matrix("blosum62"), 10, 2, top=10, all_scores=True)
    searcher.set_kernels(capi.OSW_K_I32)
    h2, tm2, s2 = searcher.search(q, ob.matrix("blosum62"), 10, 2, top=10, all_scores=True)
    assert np.array_equal(s1, s2) and h1 == h2
    m = ob.matrix("blosum62").reshape(24, 32)
    for k in range(q.n):
        self_score = int(sum(int(m[c, c]) for c in q.query(k)))
        assert h1[k][0][0] == self_score


def test_cli_report_matches_the_reference(built, tmp_path):
    """The command-line tool end to end: preprocess + search, `Score<TAB>title` blocks and the
    raw score dump must equal what the reference binary printed / computed (golden g1, g3)."""
    import gzip, os, re, shutil, subprocess
    cli = os.path.join(os.path.dirname(capi.LIB_PATH), "oswald")
    for name in ("g1_random", "g3_kat"):
        meta = load_case(name)
        for src, dst in ((meta["db_fasta"], "db.fasta"), (meta["q_fasta"], "q.fasta")):
            with gzip.open(src, "rb") as g, open(tmp_path / dst, "wb") as f:
                shutil.copyfileobj(g, f)
        subprocess.run([cli, "-O", "preprocess", "-i", "db.fasta", "-o", "db"], cwd=tmp_path, check=True)
        for run in meta["runs"][:3]:
            out = subprocess.run([cli, "-O", "search", "-q", "q.fasta", "-d", "db", "-s", run["matrix"], "-g", str(run["gap_open"]),
                                  "-e", str(run["gap_extend"]), "-r", str(meta["top"]), "--dump-scores", "dump.bin"]
                                 + (["-k", "1048576"] if run is meta["runs"][0] else []),
                                 cwd=tmp_path, check=True, capture_output=True, text=True).stdout
            blocks = out.split("Query no.")[1:]
            assert len(blocks) == len(run["hits"])
            for b, h in zip(blocks, run["hits"]):
                assert int(re.search(r"Query length:\s+(\d+)", b).group(1)) == h["query_length"]
                lines = b.split("Score\tSequence description\n")[1].split("\n")[:meta["top"]]
                assert [[int(l.split("\t")[0]), l.split("\t")[1]] for l in lines] == h["top"]
            dump = np.fromfile(tmp_path / "dump.bin", dtype=np.int32).reshape(-1, meta["n_seqs"])
            assert np.array_equal(dump, run["score_matrix"])
            assert "X.osw" in out                      # the search used the device layout -O preprocess wrote
        # without X.osw (a database preprocessed by the reference itself): laid out from X.seq, same report
        run = meta["runs"][0]
        legacy = subprocess.run([cli, "-O", "search", "-q", "q.fasta", "-d", "db", "-s", run["matrix"], "-g", str(run["gap_open"]),
                                 "-e", str(run["gap_extend"]), "-r", str(meta["top"])], cwd=tmp_path, check=True, capture_output=True,
                                text=True, env=dict(os.environ, OSW_NO_DBFILE="1")).stdout
        first = subprocess.run([cli, "-O", "search", "-q", "q.fasta", "-d", "db", "-s", run["matrix"], "-g", str(run["gap_open"]),
                                "-e", str(run["gap_extend"]), "-r", str(meta["top"])], cwd=tmp_path, check=True, capture_output=True, text=True).stdout
        assert "X.seq" in legacy and legacy.split("\nSearch date:")[0] == first.split("\nSearch date:")[0]


def test_database_from_its_file(built, tmp_path):
    """osw_db_load_file (X.osw: the chunk streams from disk) gives the same scores and hits as
    osw_db_load from the canonical arrays, unsharded and as two ranks' shards."""
    from oswald_b200.host import merge_hits
    rng = np.random.default_rng(515)
    seqs = rand_seqs(rng, 2500, 0, 400) + [AA[rng.integers(0, 20, size=n)] for n in (3000, 20000)]
    q = ob.Queries.from_list([AA[rng.integers(0, 20, size=m)] for m in (64, 300, 1400)])
    seqs[11] = q.query(1).copy()
    db = make_db(seqs)
    want = oracle_scores(q, db, "blosum62", 10, 2)
    path = tmp_path / "db.osw"
    ob.write_db_file(path, db, max_chunk_residues=2048)
    with ob.Searcher(1) as s:
        s.load_db_file(path)
        assert s.stats()["n_seqs"] == db.n_seqs and s.stats()["residues"] == db.n_residues
        for mode in MODES.values():
            check(s, db, q, "blosum62", 10, 2, 10, mask=mode, want=want)
        s.set_kernels(capi.OSW_K_DEFAULT)
        s.upload_db()
        check(s, db, q, "blosum62", 10, 2, 10, want=want)
    parts, total = [], np.zeros_like(want)
    for rank in range(2):
        with ob.Searcher(1) as s:
            s.load_db_file(path, shard_rank=rank, shard_count=2)
            hits, tm, sc = s.search(q, ob.matrix("blosum62"), 10, 2, top=10, all_scores=True)
            parts.append(hits)
            total += sc
    assert np.array_equal(total, want)
    for qi in range(q.n):
        idx, sc = O.top_r(want[qi], 10)
        assert merge_hits([parts[0][qi], parts[1][qi]], 10) == [(int(a), int(b)) for a, b in zip(sc, idx)]
    with ob.Searcher(1) as s:
        (tmp_path / "bad.osw").write_bytes(path.read_bytes()[:5000])
        with pytest.raises(capi.OswError):
            s.load_db_file(tmp_path / "bad.osw")
        s.load_db(db)                              # the context stays usable
        check(s, db, q, "blosum62", 10, 2, 10, want=want)


def test_two_gpus_in_one_context(built):
    """osw_init(2): the library deals chunks to both GPUs and merges their hit lists itself."""
    import ctypes as C
    n = C.c_int(0)
    built.osw_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs two GPUs")
    rng = np.random.default_rng(31)
    db = make_db(rand_seqs(rng, 4000, 10, 400))
    q = ob.Queries.from_list([AA[rng.integers(0, 20, size=m)] for m in (90, 250, 700)])
    with ob.Searcher(2) as s:
        s.load_db(db, max_chunk_residues=1024)
        assert s.stats()["n_seqs"] == db.n_seqs
        check(s, db, q, "blosum62", 10, 2, 10)


def test_two_gpus_pipelined_and_long_chunk_launches(built, monkeypatch):
    """The concurrent launch forms (pipelined passes over the longest chunks, the long-chunk launch of
    small databases) with two GPUs driven by one host thread: every GPU has its own side streams."""
    import ctypes as C
    n = C.c_int(0)
    built.osw_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs two GPUs")
    monkeypatch.setenv("OSW_PIPE_CHUNKS", "3")
    monkeypatch.setenv("OSW_LONG_CHUNKS", "5")
    monkeypatch.setenv("OSW_EXPRESS_RATIO", "0")
    rng = np.random.default_rng(32)
    seqs = rand_seqs(rng, 3000, 0, 400) + [AA[rng.integers(0, 20, size=k)] for k in (9000, 30000, 65535, 50000)]
    db = make_db(seqs)
    with ob.Searcher(2) as s:
        s.load_db(db, max_chunk_residues=1024)
        for lens in ([1500, 2900, 1400], [144], [90, 100], [5478]):
            q = ob.Queries.from_list([AA[rng.integers(0, 20, size=m)] for m in lens])
            want = oracle_scores(q, db, "blosum62", 10, 2)
            for mode in MODES.values():
                check(s, db, q, "blosum62", 10, 2, 10, mask=mode, want=want)


def test_invalid_inputs_are_rejected(searcher):
    rng = np.random.default_rng(2)
    db = make_db(rand_seqs(rng, 50, 5, 60))
    searcher.load_db(db)
    searcher.set_kernels(capi.OSW_K_DEFAULT)
    q = ob.Queries.from_list([AA[rng.integers(0, 20, size=30)]])
    bad_q = ob.Queries(np.full(30, 24, dtype=np.uint8), q.offsets)
    with pytest.raises(capi.OswError):
        searcher.search(bad_q, ob.matrix("blosum62"), 10, 2, top=5)
    big = ob.matrix("blosum62").copy()
    big[0] = 100
    with pytest.raises(capi.OswError):
        searcher.search(q, big, 10, 2, top=5)
    with pytest.raises(capi.OswError):
        searcher.search(q, ob.matrix("blosum62"), 256, 2, top=5)
    bad_db = ob.Database(np.full(40, 30, dtype=np.uint8), np.array([0, 40], dtype=np.uint64))
    with pytest.raises(capi.OswError):
        searcher.load_db(bad_db)
    # not in canonical (ascending length) order / offsets going backwards: rejected, not mis-scored
    unsorted = ob.Database(AA[rng.integers(0, 20, size=60)], np.array([0, 40, 60], dtype=np.uint64))
    with pytest.raises(capi.OswError):
        searcher.load_db(unsorted)
    backwards = ob.Database(AA[rng.integers(0, 20, size=60)], np.array([0, 40, 30, 60], dtype=np.uint64))
    with pytest.raises(capi.OswError):
        searcher.load_db(backwards)
    searcher.load_db(db)                       # the context stays usable
    check(searcher, db, q, "blosum62", 10, 2, 5)


def test_fuzz_small_cases(searcher):
    """Randomised sweep: query counts and lengths (including empty queries and empty database
    sequences), matrices, gap penalties, chunk sizes, top-r and all three first-stage modes."""
    rng = np.random.default_rng(20261018)
    names = ob.matrix_names()
    for case in range(40):
        n = int(rng.integers(1, 400))
        lo, hi = (0, 8) if case % 7 == 0 else (1, int(rng.integers(5, 700)))
        seqs = [rng.integers(0, 24, size=int(l)).astype(np.uint8) if case % 5 == 0 else AA[rng.integers(0, 20, size=int(l))]
                for l in rng.integers(lo, hi + 1, size=n)]
        nq = int(rng.integers(1, 7))
        qlens = [int(rng.choice([0, 1, 7, 33, 150, 400, 1300, 2900])) if rng.random() < 0.5 else int(rng.integers(1, 600)) for _ in range(nq)]
        qs = [AA[rng.integers(0, 20, size=m)] for m in qlens]
        if len(seqs) > 3 and qlens[0] > 3:
            seqs[2] = np.concatenate([qs[0][: qlens[0] // 2], seqs[2]])[:65535]
        db = make_db(seqs)
        q = ob.Queries.from_list(qs)
        name = names[int(rng.integers(0, len(names)))]
        go, ge = int(rng.integers(0, 30)), int(rng.integers(0, 6))
        top = int(rng.choice([1, 3, 10, 50, n + 5]))
        searcher.load_db(db, max_chunk_residues=int(rng.choice([0, 0, 32, 200, 1000])))
        mode = list(MODES.values())[case % len(MODES)]
        check(searcher, db, q, name, go, ge, top, mask=mode)


def test_cli_two_gpus_same_report(built, tmp_path):
    """`-f 2` (two GPUs in one process) prints the same hit blocks as `-f 1`."""
    import ctypes as C, gzip, os, shutil, subprocess
    n = C.c_int(0)
    built.osw_device_count(C.byref(n))
    if n.value < 2:
        pytest.skip("needs two GPUs")
    cli = os.path.join(os.path.dirname(capi.LIB_PATH), "oswald")
    meta = load_case("g2_overflow")
    for src, dst in ((meta["db_fasta"], "db.fasta"), (meta["q_fasta"], "q.fasta")):
        with gzip.open(src, "rb") as g, open(tmp_path / dst, "wb") as f:
            shutil.copyfileobj(g, f)
    subprocess.run([cli, "-O", "preprocess", "-i", "db.fasta", "-o", "db"], cwd=tmp_path, check=True)
    outs = []
    for f in ("1", "2"):
        out = subprocess.run([cli, "-O", "search", "-q", "q.fasta", "-d", "db", "-s", "pam30", "-g", "9", "-e", "1", "-r", "12", "-f", f],
                             cwd=tmp_path, check=True, capture_output=True, text=True).stdout
        outs.append(out.split("\nSearch date:")[0].split("Query filename:")[1])
    assert outs[0] == outs[1]


def test_segmented_bottom_rows(built, monkeypatch):
    """Several passes with a bottom-row buffer that holds only a few chunks at a time."""
    monkeypatch.setenv("OSW_PIPE_CHUNKS", "0")                  # (launch counts below: without the pipelined launches)
    rng = np.random.default_rng(88)
    seqs = rand_seqs(rng, 1500, 1, 400) + [AA[rng.integers(0, 20, size=n)] for n in (3000, 70000)]
    q = ob.Queries.from_list([AA[rng.integers(0, 20, size=m)] for m in (1500, 2900, 1400)])
    seqs[7] = q.query(2).copy()
    db = make_db(seqs)
    want = oracle_scores(q, db, "blosum62", 10, 2)
    unsegmented = {}
    with ob.Searcher(1) as s:                                   # (the switch is read when a context is made)
        s.load_db(db, max_chunk_residues=512)
        for name, mode in MODES.items():
            unsegmented[name] = check(s, db, q, "blosum62", 10, 2, 10, mask=mode, want=want)["score_launches"]
    monkeypatch.setenv("OSW_BOUND_BUDGET_COLS", "2048")
    with ob.Searcher(1) as s:
        s.load_db(db, max_chunk_residues=512)
        for name, mode in MODES.items():
            tm = check(s, db, q, "blosum62", 10, 2, 10, mask=mode, want=want)
            assert tm["score_launches"] >= 2 * unsegmented[name]          # at least two segments x passes


def test_query_batches(built, monkeypatch):
    """Many queries are searched in batches that bound the score matrix; results are unchanged."""
    monkeypatch.setenv("OSW_SCORE_BUDGET_KB", "8")
    rng = np.random.default_rng(5)
    db = make_db(rand_seqs(rng, 700, 1, 300))
    q = ob.Queries.from_list([AA[rng.integers(0, 20, size=m)] for m in (20, 60, 61, 150, 400, 900, 1700)])
    with ob.Searcher(1) as s:
        s.load_db(db)
        tm = check(s, db, q, "blosum62", 10, 2, 10)
        assert tm["launches"] >= 4 * 3            # four batches, each with its own profile, scoring and top-r launches


def test_streamed_database_windows(built):
    """osw_set_device_window: the column stream is not resident but copied segment by segment
    through two device windows (the reference's -k); results, overflow re-score included, are
    the same."""
    rng = np.random.default_rng(404)
    W = np.uint8(19)
    seqs = rand_seqs(rng, 9000, 1, 600) + [AA[rng.integers(0, 20, size=n)] for n in (30000, 65535)] + [np.full(6000, W, dtype=np.uint8)]
    q_sets = [[AA[rng.integers(0, 20, size=m)] for m in (144, 189, 222, 1500, 2900)],
              [AA[rng.integers(0, 20, size=700)]],
              [np.full(5400, W, dtype=np.uint8), AA[rng.integers(0, 20, size=100)]]]         # 70 200 > 16 bits: staged re-score
    db = make_db(seqs)
    assert db.n_residues > 2.5 * (1 << 20)
    with ob.Searcher(1) as s:
        s.set_device_window(1 << 20)
        s.load_db(db)
        for qs in q_sets:
            q = ob.Queries.from_list(qs)
            for mode in MODES.values():
                tm = check(s, db, q, "pam30", 9, 1, 10, mask=mode)
        assert tm["rescored_pairs"] >= 1
        s.set_kernels(capi.OSW_K_I32)
        with pytest.raises(capi.OswError):
            s.search(q, ob.matrix("pam30"), 9, 1, top=5)


@pytest.mark.parametrize("min_g", ["4", "8", "16", "32"])
def test_forced_group_widths(built, monkeypatch, min_g):
    """Single-pass plans: every group width the small-database estimate can pick (G = 4 ... 32, down
    to 8 rows per lane) gives the same scores; a tiny database deals its first chunks statically."""
    monkeypatch.setenv("OSW_MIN_G", min_g)
    rng = np.random.default_rng(31 + int(min_g))
    seqs = rand_seqs(rng, 1200, 0, 500) + [AA[rng.integers(0, 20, size=n)] for n in (2500, 9000)]
    db = make_db(seqs)
    with ob.Searcher(1) as s:
        s.load_db(db)
        for lens in ([144], [30], [90, 100], [33, 150, 7, 61]):
            q = ob.Queries.from_list([AA[rng.integers(0, 20, size=m)] for m in lens])
            for mode in MODES.values():
                check(s, db, q, "blosum62", 10, 2, 10, mask=mode)


@pytest.mark.parametrize("transpose", ["0", "2"])
@pytest.mark.parametrize("n_long", ["3", "100000"])
def test_long_chunk_launch(built, monkeypatch, n_long, transpose):
    """Small databases: the longest chunks of a chain-bound single-pass launch go to a second, concurrent
    launch - of the same kernel with the widest array (32 lanes x 8..16 rows, every CTA an express CTA), or
    (transpose = 2, mode "auto", queries that fit it) in the transposed form, which scores the pairs of
    sequences those chunks hold - forced here to take a few chunks / every chunk: one or several queries, both
    uses of the packed halves, empty and long sequences, overflow, tiny chunks."""
    monkeypatch.setenv("OSW_LONG_CHUNKS", n_long)
    monkeypatch.setenv("OSW_TRANSPOSE", transpose)
    monkeypatch.setenv("OSW_EXPRESS_RATIO", "0")            # every launch counts as chain-bound
    rng = np.random.default_rng(77 + int(n_long))
    W = np.uint8(19)
    seqs = rand_seqs(rng, 900, 0, 500) + [AA[rng.integers(0, 20, size=n)] for n in (2500, 9000, 65535)] + [np.full(7000, W, dtype=np.uint8)]
    db = make_db(seqs)
    with ob.Searcher(1) as s:
        s.load_db(db)
        for lens in ([144], [30], [90, 100], [33, 150, 7, 61], [400], [500, 480], [256, 255, 1], [97] * 9):
            q = ob.Queries.from_list([AA[rng.integers(0, 20, size=m)] for m in lens])
            for mode in MODES.values():
                check(s, db, q, "blosum62", 10, 2, 10, mask=mode)
        q = ob.Queries.from_list([np.full(400, W, dtype=np.uint8), AA[rng.integers(0, 20, size=300)]])
        for mode in MODES.values():
            check(s, db, q, "pam30", 9, 1, 10, mask=mode)
        s.load_db(db, max_chunk_residues=64)
        check(s, db, ob.Queries.from_list([AA[rng.integers(0, 20, size=m)] for m in (144, 222)]), "blosum62", 10, 2, 10)


@pytest.mark.parametrize("n_pipe", ["1", "5", "16"])
def test_pipelined_passes_over_the_longest_chunks(built, monkeypatch, n_pipe):
    """Several passes, the longest chunks walked by all passes at once (each pass on SMs of its own, a few
    dozen columns apart, bottom rows handed over in place under progress counters): two-track and
    pair-database plans, passes of different heights, passes that do and do not continue a query, titin-length
    and ordinary chunks, overflow."""
    monkeypatch.setenv("OSW_PIPE_CHUNKS", n_pipe)
    rng = np.random.default_rng(300 + int(n_pipe))
    W = np.uint8(19)
    seqs = rand_seqs(rng, 1200, 0, 400) + [AA[rng.integers(0, 20, size=n)] for n in (3000, 20000, 40000, 65535)] + [np.full(8000, W, dtype=np.uint8)]
    q_sets = [[1500, 2900, 1400], [3005], [5478, 144], [1280, 1280, 1280, 1280], [2560, 100, 100, 2000], [1281, 1279, 35, 2600, 7]]
    db = make_db(seqs)
    with ob.Searcher(1) as s:
        s.load_db(db, max_chunk_residues=2048)
        for lens in q_sets:
            q = ob.Queries.from_list([AA[rng.integers(0, 20, size=m)] for m in lens])
            seqs_q = q.query(q.n - 1)
            want = oracle_scores(q, db, "blosum62", 10, 2)
            for mode in MODES.values():
                tm = check(s, db, q, "blosum62", 10, 2, 10, mask=mode, want=want)
        q = ob.Queries.from_list([np.full(5300, W, dtype=np.uint8), AA[rng.integers(0, 20, size=1400)]])        # 68 900 > 16 bits
        tm = check(s, db, q, "pam30", 9, 1, 10)
        assert tm["rescored_pairs"] >= 1


def test_many_tiny_queries(built):
    """20 000 queries of 1..6 residues: more than one batch (16 384 queries per plan), hundreds of
    passes of one-query lanes; every score and hit list still equals the oracle's."""
    rng = np.random.default_rng(2024)
    db = make_db(rand_seqs(rng, 60, 0, 40))
    q = ob.Queries.from_list([AA[rng.integers(0, 20, size=int(m))] for m in rng.integers(1, 7, size=20000)])
    with ob.Searcher(1) as s:
        s.load_db(db)
        tm = check(s, db, q, "blosum62", 10, 2, 3)
        assert tm["score_launches"] >= 2


@pytest.mark.parametrize("transpose", ["1", "-1"])
def test_transposed_form(built, monkeypatch, transpose):
    """The transposed first stage (sw_t16.cu): one warp per pair of sequences, gangs of 2, 4 and 8 warps for
    the long ones (bottom rows handed from block to block through tagged ring entries), 8- and 4-warp CTAs
    (queries up to 248 / 1024 residues), several queries, empty sequences and queries, an odd number of
    sequences, W-rich sequences - forced (OSW_TRANSPOSE=1) and as the model picks it (-1)."""
    monkeypatch.setenv("OSW_TRANSPOSE", transpose)
    rng = np.random.default_rng(404)
    W = np.uint8(19)
    seqs = ([np.zeros(0, np.uint8)] * 3 + rand_seqs(rng, 1200, 1, 700) +
            [AA[rng.integers(0, 20, size=n)] for n in (255, 256, 257, 511, 513, 1023, 1025, 2047, 2049, 2500, 4097, 9000, 16385, 36000, 65535)] +
            [np.full(7000, W, dtype=np.uint8)])
    db = make_db(seqs)
    with ob.Searcher(1) as s:
        s.load_db(db)
        for lens in ([144], [1], [31], [32], [33], [248], [249], [90, 100], [33, 150, 7, 61, 0, 200], [400], [1000], [1024, 3], [97] * 9, [500] * 8):
            q = ob.Queries.from_list([AA[rng.integers(0, 20, size=m)] for m in lens])
            tm = check(s, db, q, "blosum62", 10, 2, 10)
        # (with queries of at most 1024 residues and substitution scores of at most 31 no score of this form reaches
        # the 16-bit flag threshold; the largest it sees here is 1000 x 13)
        q = ob.Queries.from_list([np.full(1000, W, dtype=np.uint8), AA[rng.integers(0, 20, size=77)]])
        check(s, db, q, "pam30", 9, 1, 5)
    # tiny databases: fewer pairs than warps, a single sequence
    with ob.Searcher(1) as s:
        for n in (1, 2, 3, 40):
            db = make_db(rand_seqs(rng, n, 1, 3000))
            s.load_db(db)
            check(s, db, ob.Queries.from_list([AA[rng.integers(0, 20, size=m)] for m in (144, 20)]), "blosum50", 10, 2, 4)
