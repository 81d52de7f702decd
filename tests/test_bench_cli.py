"""bench.py's reference arm (CPU only) and the profile tooling: the JSON contract the driver parses."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line(built):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--ref-sample", "3000"], cwd=ROOT, check=True, capture_output=True, text=True, timeout=600).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "GCUPS" and d["unit"] == "GCUPS" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_other_ranks_stay_silent(built):
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       cwd=ROOT, env=env, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_ncu_traffic_summary(tmp_path):
    hdr = '"ID","Process ID","Process Name","Host Name","Kernel Name","Context","Stream","Block Size","Grid Size","Device","CC","Section Name","Metric Name","Metric Unit","Metric Value"'
    rows = [hdr]
    for i, (rd, wr) in enumerate([(200.0, 1600.0), (1900.0, 1600.0)]):
        base = '"%d","1","python","h","void osw_u16::sw_u16_kernel<32, 40, 384, 0>(osw_u16::KArgs)","1","7","(384, 1, 1)","(148, 1, 1)","0","10.0","Command line profiler metrics",' % i
        rows.append(base + '"dram__bytes_read.sum","Mbyte","%s"' % rd)
        rows.append(base + '"dram__bytes_write.sum","Mbyte","%s"' % wr)
        rows.append(base + '"gpu__time_duration.sum","ms","75.0"')
    src = tmp_path / "t.csv"
    src.write_text("==PROF== noise line\n" + "\n".join(rows) + "\n")
    dst = tmp_path / "t.json"
    subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_traffic.py"), str(src), str(dst), "cmd"], check=True, capture_output=True)
    d = json.loads(dst.read_text())
    assert d["launches_captured"] == 2
    assert abs(d["dram_bytes_per_launch"] - (200e6 + 1600e6 + 1900e6 + 1600e6) / 2) < 1
    assert abs(d["ncu_ms_per_launch"] - 75.0) < 1e-9
