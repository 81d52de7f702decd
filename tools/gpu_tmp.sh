nvidia-smi -L
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 -k "two_gpus or sharded or empty_shard" -rs > gpurun_out/t10_pytest_multigpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/t10_pytest_multigpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config 1 --no-extra --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/t10_bench_c1_2gpu.json 2> gpurun_out/t10_bench_c1_2gpu.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/t10_bench_c1_2gpu.json"))
print("N=%d value %.1f GCUPS  e2e %.1f ms/step %.3f" % (d["n_gpus"], d["value"], d["e2e"]["value"], d["ms_per_step"]))
print("verified", d["verified"], "sp_multi", d["single_process_multi_gpu_ok"])
PY
