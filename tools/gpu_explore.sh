#!/bin/bash
# One-off measurements on one B200 (under gpurun): usage  bash tools/gpu_explore.sh <tag>
TAG=${1:-explore}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -x -m gpu -p no:cacheprovider 2>&1 | tail -3
run() {   # label, bench args...
  local label=$1; shift
  OSW_TRACE=1 timeout 300 python bench.py --no-cpu-baseline "$@" > gpurun_out/${TAG}_$label.json 2> gpurun_out/${TAG}_$label.err
  python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_$label.json"))
print("$label: %.1f GCUPS  %.3f ms/step  e2e %.1f  launches %d  padded/useful %.2f breakdown %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["padded_over_useful_cells"], d["breakdown_ms"]))
PY
  grep "osw trace" gpurun_out/${TAG}_$label.err | tail -1
}
run c1 --steps 5 --warmup 3 --seqs-per-gpu 10000 --query-lengths 144
run c1_2q --steps 5 --warmup 3 --seqs-per-gpu 10000 --query-lengths 144,189
run s50k --steps 5 --warmup 3 --seqs-per-gpu 50000 --query-lengths 144
run s50k_1000 --steps 5 --warmup 3 --seqs-per-gpu 50000 --query-lengths 1000
run q144 --steps 2 --warmup 1 --query-lengths 144
run q5478 --steps 2 --warmup 1 --query-lengths 5478
run c2 --steps 3 --warmup 2
