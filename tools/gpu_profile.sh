#!/bin/bash
# ncu evidence for the current build (one GPU): launch list + one full capture of the top kernel.
# usage (under gpurun): bash tools/gpu_profile.sh <tag>
TAG=${1:-prof}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --seqs-per-gpu 30000 --no-cpu-baseline"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sw_u16 -s 14 -c 2 -f -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/
