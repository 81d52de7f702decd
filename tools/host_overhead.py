import sys, time, numpy as np
sys.path.insert(0, '.')
import bench, oswald_b200 as ob
for cfg, kw in ((1, {}), (2, {"n_override": 50000})):
    wl = bench.make_workload(cfg, **kw)
    db, q = wl["db"], wl["queries"]
    mat = ob.matrix("blosum62")
    with ob.Searcher(1) as s:
        s.load_db(db)
        for _ in range(5): s.search(q, mat, 10, 2, top=10)
        t_up = t_s = dev = wall = 0.0
        n = 200 if cfg == 1 else 20
        for _ in range(n):
            t0 = time.perf_counter(); s.upload_db(); t1 = time.perf_counter()
            hits, tm = s.search(q, mat, 10, 2, top=10); t2 = time.perf_counter()
            t_up += t1 - t0; t_s += t2 - t1; dev += tm["device_ms"]; wall += tm["wall_ms"]
        print("config %d (%d seqs, %d queries): upload %.3f ms, search call %.3f ms (library wall %.3f, device %.3f, h2d phase %.3f) launches %d" % (
            cfg, db.n_seqs, q.n, 1e3 * t_up / n, 1e3 * t_s / n, wall / n, dev / n, tm["h2d_ms"], tm["launches"]))
