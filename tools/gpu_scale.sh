#!/bin/bash
# Strong scaling of the default bench (config 3, one fixed database split across the ranks) on ONE box:
# N = 1, 2, 4 quick (no extra configs, no verification), N = 8 complete.   usage (gpurun --gpus 8): bash tools/gpu_scale.sh <tag>
TAG=${1:-scale}
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc
for N in ${NS:-1 2 4 8}; do
  ARGS="--steps 3 --warmup 2 --no-cpu-baseline"
  [ $N -lt 8 ] && ARGS="$ARGS --no-extra --no-verify"
  if [ $N -eq 1 ]; then CMD="python bench.py --gpus 1 $ARGS"
  else CMD="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29520+N)) bench.py --gpus $N $ARGS"; fi
  ( time timeout 900 $CMD > gpurun_out/${TAG}_bench_${N}gpu.json 2> gpurun_out/${TAG}_bench_${N}gpu.err ) 2> gpurun_out/${TAG}_bench_${N}gpu.time
  echo "N=$N rc=$? $(grep real gpurun_out/${TAG}_bench_${N}gpu.time)"
done
python - <<PY
import json
v = {}
for n in [int(x) for x in "${NS:-1 2 4 8}".split()]:
    try:
        d = json.load(open("gpurun_out/${TAG}_bench_%dgpu.json" % n))
    except Exception as e:
        print(n, "no result", e); continue
    v[n] = d["value"]
    print("N=%d  %.1f GCUPS device  %.1f e2e  %.1f ms/step  eff %.4f  e2e-eff %.4f  breakdown %s" % (
        n, d["value"], d["e2e"]["value"], d["ms_per_step"], d["value"] / (n * v.get(1, d["value"] / n)), d["e2e"]["value"] / (n * v.get(1, d["value"] / n)), d["breakdown_ms"]))
    if d.get("verified"): print("   verified", d["verified"], "single-process multi-GPU", d["single_process_multi_gpu_ok"])
    for c in d["extra"]["configs"]:
        for r in c["runs"]:
            w = r["verified"]
            print("   config", c["config"], r["matrix"], "%.1f GCUPS dev, %.1f e2e, %.2f ms/step, mismatches %s topr_ok %s planted %s" % (
                r["gcups_device"], r["gcups_e2e"], r["ms_per_step"], w["mismatches"], w["topr_ok"], w.get("planted_ok")))
PY
