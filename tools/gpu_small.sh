#!/bin/bash
# Small / mid-size database timings on one B200 (device-timed GCUPS, traces of the launches).
# usage (under gpurun): bash tools/gpu_small.sh <tag> [VAR=value ...]
TAG=${1:-small}; shift
mkdir -p gpurun_out
run() {
  local label=$1; shift
  OSW_TRACE=1 timeout 300 env "${EXTRA[@]}" python bench.py --no-cpu-baseline --no-extra --no-verify "$@" 2> gpurun_out/${TAG}_${label}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$label: %.1f GCUPS  e2e %.1f  %.3f ms/step (score %.3f topr %.3f) launches/step %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['breakdown_ms']['score'], d['breakdown_ms']['topr'], d['gpu_launches']/d['steps']))"
  grep "osw trace" gpurun_out/${TAG}_${label}.err | tail -1
}
EXTRA=("$@"); [ ${#EXTRA[@]} -eq 0 ] && EXTRA=(OSW_DUMMY=1)
# (CASES = a subset of the labels below, default all)
want() { [ -z "$CASES" ] || [[ " $CASES " == *" $1 "* ]]; }
want c1 && run c1 --config 1 --steps 20 --warmup 5
want c1_two && run c1_two --config 1 --steps 20 --warmup 5 --query-lengths 144,189
want s50k && run s50k --config 2 --seqs 50000 --steps 5 --warmup 2 --query-lengths 144
want s100k && run s100k --config 2 --seqs 100000 --steps 5 --warmup 2 --query-lengths 144
want s200k && run s200k --config 2 --seqs 200000 --steps 3 --warmup 2 --query-lengths 144
want q144 && run q144 --config 2 --steps 2 --warmup 1 --query-lengths 144
want q400 && run q400 --config 2 --steps 2 --warmup 1 --query-lengths 400
want q5478 && run q5478 --config 2 --steps 2 --warmup 1 --query-lengths 5478
want q1000_10k && run q1000_10k --config 1 --steps 10 --warmup 3 --query-lengths 1000
true
