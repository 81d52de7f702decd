#!/bin/bash
# N1 at config-3 scale on the GPU box: `-O preprocess` (ours, -c threads; the reference's) and a cold
# `-O search` from X.osw / from X.seq.   usage (under gpurun): bash tools/gpu_n1.sh <tag> [n_seqs] [mu]
TAG=${1:-n1}; N=${2:-6900000}; MU=${3:-5.056}
mkdir -p gpurun_out
T=$(mktemp -d); OUT=gpurun_out/${TAG}_n1.txt; : > $OUT
CORES=$(nproc)
( time ./tools/osw_synth db -n $N -mu $MU -sigma 0.6 -seed 5 -o $T/db.fasta ) 2>&1 | grep real | sed 's/^/generate FASTA: /' >> $OUT
ls -la $T/db.fasta | awk '{print "FASTA bytes:", $5}' >> $OUT
./tools/osw_synth queries -lengths 144,189,222,375,464,567,657,727,850,1000,1500,2005,2504,3005,3564,4061,4548,4743,5147,5478 -seed 9 -o $T/q.fasta
for c in 1 $CORES; do
  ( time ./oswald_b200/oswald -O preprocess -i $T/db.fasta -o $T/db -c $c ) 2>&1 | grep real | sed "s/^/ours preprocess -c $c (X.info X.seq X.desc X.osw): /" >> $OUT
done
if [ -x oracle/_ref/oswald_ref ]; then
  ( time ./oracle/_ref/oswald_ref -O preprocess -i $T/db.fasta -o $T/ref -c $CORES ) 2>&1 | grep real | sed "s/^/reference preprocess -c $CORES: /" >> $OUT
  for ext in info seq desc; do cmp $T/db.$ext $T/ref.$ext && echo "X.$ext identical to the reference's" >> $OUT; done
fi
ls -la $T | awk '{print $5, $9}' >> $OUT
for mode in osw seq; do
  if [ $mode = seq ]; then export OSW_NO_DBFILE=1; else unset OSW_NO_DBFILE; fi
  sync; echo 3 > /proc/sys/vm/drop_caches 2>/dev/null
  ( time OSW_TRACE=1 ./oswald_b200/oswald -O search -q $T/q.fasta -d $T/db -r 5 > $T/search_$mode.txt 2> $T/search_$mode.err ) 2>&1 | grep real | sed "s/^/ours cold search from X.$mode: /" >> $OUT
  grep "oswald trace" $T/search_$mode.err | sed "s/^/   [$mode] /" >> $OUT
  grep -E "Search time|GPU time|GPU speed|Database layout" $T/search_$mode.txt | sed "s/^/   [$mode] /" >> $OUT
done
cmp <(sed -n '/Query no/,/Search date/p' $T/search_osw.txt | grep -v 'Search date') <(sed -n '/Query no/,/Search date/p' $T/search_seq.txt | grep -v 'Search date') && echo "reports identical (X.osw vs X.seq)" >> $OUT
rm -rf $T
cat $OUT
