#!/bin/bash
# A/B on one box of an environment switch (default: OSW_NO_EXPRESS=1 as the baseline arm).
# usage: bash tools/gpu_ab.sh <tag> [VAR=value of the baseline arm]
TAG=${1:-ab}; BASE=${2:-OSW_NO_EXPRESS=1}; NEW=${3:-OSW_DUMMY=1}
mkdir -p gpurun_out
run() {
  local label=$1; shift
  for V in base new; do
    if [ $V = base ]; then PRE="env $BASE"; else PRE="env $NEW"; fi
    OSW_TRACE=1 timeout 300 $PRE python bench.py --no-cpu-baseline "$@" 2> gpurun_out/${TAG}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$label $V: %.1f GCUPS  score %.3f ms  busy-clk %.2f' % (d['value'], d['breakdown_ms']['score'], d['roofline']['achieved_cells_per_busy_sm_clk']))"
    grep "osw trace" gpurun_out/${TAG}.err | tail -1 | cut -c1-110
  done
}
run c1 --steps 5 --warmup 3 --seqs-per-gpu 10000 --query-lengths 144
run s50k_4q --steps 5 --warmup 3 --seqs-per-gpu 50000 --query-lengths 144,189,222,375
run s100k --steps 3 --warmup 2 --seqs-per-gpu 100000 --query-lengths 144
run s100k_375 --steps 3 --warmup 2 --seqs-per-gpu 100000 --query-lengths 375
run s200k --steps 3 --warmup 2 --seqs-per-gpu 200000 --query-lengths 144
run q144 --steps 2 --warmup 1 --query-lengths 144
