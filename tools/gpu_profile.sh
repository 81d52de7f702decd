#!/bin/bash
# ncu evidence for the current build (one GPU): launch list + one full capture of the top kernel
# (on a 30 000-sequence slice: ~40 replays per kernel) + DRAM bytes of the 17 first-stage launches of
# one full-size search.   usage (under gpurun): bash tools/gpu_profile.sh <tag>
TAG=${1:-prof}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 0 --config 2 --seqs 30000 --no-cpu-baseline --no-extra --no-verify"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sw_u16_kernel -s 14 -c 2 -f -o gpurun_out/${TAG} $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"
# the transposed form (sw_t16.cu) on BASELINE.json config 1: launch list of three searches, full capture of the kernel
C1="python bench.py --config 1 --steps 3 --warmup 1 --no-cpu-baseline --no-extra --no-verify"
$C1 > gpurun_out/${TAG}_plain5.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${TAG}_launches_c1.csv $C1 > gpurun_out/${TAG}_ncu5.log 2>&1
echo "config-1 launch list rc=$?"
$C1 > gpurun_out/${TAG}_plain6.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sw_t16 -s 2 -c 1 -f -o gpurun_out/${TAG}_t16 $C1 > gpurun_out/${TAG}_ncu6.log 2>&1
echo "config-1 full capture rc=$?"
FULL="python bench.py --steps 1 --warmup 0 --config ${CFG:-3} --no-cpu-baseline --no-extra --no-verify"
$FULL > gpurun_out/${TAG}_plain4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches_full.csv $FULL > gpurun_out/${TAG}_ncu4.log 2>&1
echo "full-size launch list rc=$?"
$FULL > gpurun_out/${TAG}_plain3.log 2>&1 &&
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:sw_u16_kernel -c 17 --csv \
    --log-file gpurun_out/${TAG}_traffic.csv $FULL > gpurun_out/${TAG}_ncu3.log 2>&1
echo "traffic capture rc=$?"
python tools/ncu_traffic.py gpurun_out/${TAG}_traffic.csv gpurun_out/${TAG}_traffic.json "$FULL" ${CFG:-3}
ls -la gpurun_out/
