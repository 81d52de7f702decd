#!/bin/bash
# Multi-GPU check on one box with N GPUs: the one-process tests (osw_init(N), CLI -f N) and the
# strong-scaling bench under torchrun.   usage (gpurun --gpus N): bash tools/gpu_multi.sh <tag> N [bench args]
TAG=${1:-multi}; N=${2:-2}; shift; shift
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -k "two_gpus or sharded or empty_shard" -rs > gpurun_out/${TAG}_pytest_multigpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/${TAG}_pytest_multigpu.log
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N "$@" \
    > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err ) 2> gpurun_out/${TAG}_bench_${N}gpu.time; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench_${N}gpu.time
tail -3 gpurun_out/${TAG}_bench_${N}gpu.err
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench_${N}gpu.json"))
print("N=%d value %.1f GCUPS  e2e %.1f  frac %.3f  launches %d  ms/step %.1f  clocks %s" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"], d["ms_per_step"], d["clocks"]))
print("verified", d["verified"], "sp_multi", d["single_process_multi_gpu_ok"], "setup", d["setup_seconds"], "breakdown", d["breakdown_ms"])
for c in d["extra"]["configs"]:
    for r in c["runs"]:
        v = r["verified"]
        print("config", c["config"], r["matrix"], "%.1f GCUPS dev, %.1f e2e, %.2f ms/step, launches %d, mismatches %s topr_ok %s planted %s" % (
            r["gcups_device"], r["gcups_e2e"], r["ms_per_step"], r["launches"], v["mismatches"], v["topr_ok"], v.get("planted_ok")))
PY
