#!/bin/bash
# A/B of two builds (or of an environment switch) on ONE box: box-to-box spread is about 3 %, the
# steps worth finding are below 1 %.   usage (under gpurun):
#   bash tools/gpu_ab.sh <tag> <baseline.so | VAR=value> [bench arguments ...]
# baseline.so: a library from tools/build_at.sh, loaded through OSWALD_CUDA_LIB; VAR=value: an
# environment switch of the library set for the baseline arm only.
TAG=${1:-ab}; BASE=${2:?baseline}; shift 2
case "$BASE" in *.so) BASE="OSWALD_CUDA_LIB=$(readlink -f "$BASE")";; esac
mkdir -p gpurun_out
run() {
  local label=$1; shift
  for rep in 1 2; do
    for V in new base; do
      if [ $V = base ]; then PRE="env $BASE"; else PRE="env"; fi
      OSW_TRACE=1 timeout 300 $PRE python bench.py --no-cpu-baseline --no-extra --no-verify "$@" 2> gpurun_out/${TAG}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$label $V: %.1f GCUPS  e2e %.1f  score %.3f ms  busy-clk %.2f' % (d['value'], d['e2e']['value'], d['breakdown_ms']['score'], d['roofline']['achieved_cells_per_busy_sm_clk']))"
    done
  done
}
if [ $# -gt 0 ]; then run custom "$@"; exit 0; fi
run c2 --config 2 --steps 3 --warmup 2
run q5478 --config 2 --steps 2 --warmup 1 --query-lengths 5478
run q144 --config 2 --steps 2 --warmup 1 --query-lengths 144
run q1000 --config 2 --steps 2 --warmup 1 --query-lengths 1000
run c1 --config 1 --steps 10 --warmup 3
run s100k --config 2 --steps 3 --warmup 2 --seqs 100000 --query-lengths 144
