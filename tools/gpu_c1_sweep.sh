#!/bin/bash
# Config 1: what the existing mechanisms can reach - bulk geometry (OSW_MIN_G) x long-chunk launch (chunks, CTAs).
run() { env "$@" OSW_EXPRESS_RATIO=0 python bench.py --no-cpu-baseline --no-extra --no-verify --config 1 --steps 20 --warmup 5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$*: %.1f GCUPS  %.3f ms' % (d['value'], d['ms_per_step']))"; }
run OSW_LONG_CHUNKS=0
for g in ${GS:-8 16 32}; do
  for n in "8 8" "24 24" "48 48" "72 72" "96 48"; do set -- $n; run OSW_WAVE=${WAVE:-1} OSW_MIN_G=$g OSW_LONG_CHUNKS=$1 OSW_LONG_CTAS=$2; done
done
