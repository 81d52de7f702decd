#!/bin/bash
# Round-trip check on a B200 box: smoke, GPU parity tests, calibration, a short bench.
# usage (under gpurun): bash tools/gpu_check.sh [bench args...]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" > gpurun_out/cpu.txt 2>&1
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "
from oswald_b200.host import calibrate
import json; print(json.dumps(calibrate(0)))" > gpurun_out/calib.json 2> gpurun_out/calib.err; echo "calib rc=$?"
cat gpurun_out/calib.json
timeout 1200 python bench.py "$@" > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
