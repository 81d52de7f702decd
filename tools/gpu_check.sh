#!/bin/bash
# Quick GPU check (one B200): smoke, GPU tests, the default bench.   usage: bash tools/gpu_check.sh <tag> [bench args]
TAG=${1:-chk}; shift
mkdir -p gpurun_out
nvidia-smi -L | head -2; lscpu | grep -E "^CPU\(s\)|Model name" 
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -q -x -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/${TAG}_pytest_gpu.log
( time timeout 1500 python bench.py "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err ) 2> gpurun_out/${TAG}_bench.time; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.time
tail -3 gpurun_out/${TAG}_bench.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json"))
print("value %.1f GCUPS  e2e %.1f  frac %.3f  launches %d  ms/step %.1f  clocks %s" % (d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"], d["ms_per_step"], d["clocks"]))
print("verified", d["verified"], "sp_multi", d["single_process_multi_gpu_ok"], "setup", d["setup_seconds"])
print("cpu", d.get("cpu_baseline"))
for c in d["extra"]["configs"]:
    for r in c["runs"]:
        print("config", c["config"], r["matrix"], "%.1f GCUPS dev, %.1f e2e, %.2f ms/step, rescored %d (%.2f ms), launches %d, verified %s, %ss" % (
            r["gcups_device"], r["gcups_e2e"], r["ms_per_step"], r["rescored_pairs_max_rank"], r["rescore_ms"], r["launches"], r["verified"], c["seconds_total"]))
PY
