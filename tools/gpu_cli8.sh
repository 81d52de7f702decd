#!/bin/bash
# The reference's own deployment form at full size: ONE process, `-f N` GPUs (arguments.c:108-112), config-3-sized
# database from X.osw.   usage (gpurun --gpus 8): bash tools/gpu_cli8.sh <tag>
TAG=${1:-cli8}; N=${2:-6900000}; MU=${3:-5.056}
mkdir -p gpurun_out
T=$(mktemp -d); OUT=gpurun_out/${TAG}_cli_multi_gpu.txt; : > $OUT
./tools/osw_synth db -n $N -mu $MU -sigma 0.6 -seed 5 -o $T/db.fasta
./tools/osw_synth queries -lengths 144,189,222,375,464,567,657,727,850,1000,1500,2005,2504,3005,3564,4061,4548,4743,5147,5478 -seed 9 -o $T/q.fasta
( time ./oswald_b200/oswald -O preprocess -i $T/db.fasta -o $T/db -c $(nproc) ) 2>&1 | grep real | sed "s/^/preprocess -c $(nproc): /" >> $OUT
for f in 8 1; do
  ( time OSW_TRACE=1 ./oswald_b200/oswald -O search -q $T/q.fasta -d $T/db -r 10 -f $f > $T/search_$f.txt 2> $T/search_$f.err ) 2>&1 | grep real | sed "s/^/search -f $f: /" >> $OUT
  grep "oswald trace" $T/search_$f.err | sed "s/^/   [-f $f] /" >> $OUT
  grep -E "Search time|Search speed|GPU time|GPU speed|Number of GPUs|Kernel launches|Database layout" $T/search_$f.txt | sed "s/^/   [-f $f] /" >> $OUT
done
cmp <(sed -n '/Query no/,/Search date/p' $T/search_8.txt | grep -v 'Search date') <(sed -n '/Query no/,/Search date/p' $T/search_1.txt | grep -v 'Search date') && echo "hit blocks of -f 8 and -f 1 identical" >> $OUT
rm -rf $T
cat $OUT
