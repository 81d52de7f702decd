#!/bin/bash
# The transposed first stage (sw_t16.cu) on one B200: memcheck of a small case, its parity tests, small-database timings
# with and without it.   usage (under gpurun): bash tools/gpu_t16.sh <tag> [ncu] [tests] [bench] [bench_on]
TAG=${1:-t16}; shift
mkdir -p gpurun_out
for what in "$@"; do case $what in
ncu)
  # full captures of the kernel: BASELINE.json config 1 (latency / tail) and 50 000 sequences (steady state)
  C1="python bench.py --config 1 --steps 3 --warmup 1 --no-cpu-baseline --no-extra --no-verify"
  C2="python bench.py --config 2 --seqs 50000 --query-lengths 144 --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-verify"
  OSW_TRANSPOSE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:sw_t16 -s 2 -c 1 -f -o gpurun_out/${TAG}_c1 $C1 > gpurun_out/${TAG}_ncu_c1.log 2>&1; echo "ncu c1 rc=$?"
  OSW_TRANSPOSE=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:sw_t16 -s 1 -c 1 -f -o gpurun_out/${TAG}_s50k $C2 > gpurun_out/${TAG}_ncu_s50k.log 2>&1; echo "ncu s50k rc=$?" ;;
tests)
  timeout 900 python -m pytest tests -q -x -m gpu -p no:cacheprovider --timeout 300 -k "transposed or fuzz or edge_shapes or overflow or long_sequences or ties or golden or tiny" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/${TAG}_pytest.log ;;
bench)
  echo "== transposed forced"; bash tools/gpu_small.sh ${TAG}_on OSW_TRANSPOSE=1
  echo "== transposed never";  bash tools/gpu_small.sh ${TAG}_off OSW_TRANSPOSE=0
  echo "== by the model";      bash tools/gpu_small.sh ${TAG}_auto ;;
bench_on)
  echo "== transposed forced"; bash tools/gpu_small.sh ${TAG}_on OSW_TRANSPOSE=1 ;;
esac; done
