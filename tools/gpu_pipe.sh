#!/bin/bash
# Pipelined passes over the longest chunks: config 5 (titin-length sequences) at 1/8 and at full size, on / off.
TAG=${1:-pipe}
mkdir -p gpurun_out
run() {
  local label=$1; shift
  OSW_TRACE=1 timeout 600 env "${EXTRA[@]}" python bench.py --no-cpu-baseline --no-extra "$@" 2> gpurun_out/${TAG}_${label}.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$label: %.1f GCUPS  e2e %.1f  %.3f ms/step (score %.3f) launches/step %d verified %s' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['breakdown_ms']['score'], d['gpu_launches']/d['steps'], d['verified'] and (d['verified']['mismatches'], d['verified']['topr_ok'], d['verified'].get('planted_ok'))))"
  grep -c "pipelined=[1-9]" gpurun_out/${TAG}_${label}.err
}
for mode in on off; do
  if [ $mode = off ]; then EXTRA=(OSW_PIPE_CHUNKS=0); else EXTRA=(OSW_DUMMY=1); fi
  echo "== pipelining $mode"
  run c5_eighth_$mode --config 5 --seqs 71250 --steps 3 --warmup 1
  run c5_full_$mode --config 5 --steps 2 --warmup 1
done
