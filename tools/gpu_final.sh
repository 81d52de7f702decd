#!/bin/bash
# End-of-round validation on one B200: smoke, GPU tests, bench (both arms; the bench carries all five configs), the
# command-line tool at a moderate size, ncu evidence.   usage: bash tools/gpu_final.sh <tag>
TAG=${1:-final}
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest_gpu.log
timeout 900 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "bench reference rc=$?"
timeout 900 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.load(open("gpurun_out/${TAG}_bench.json")); r = json.load(open("gpurun_out/${TAG}_bench_reference.json"))
print("ours %.1f GCUPS  e2e %.1f  frac %.3f  launches %d  clocks %s | reference %.1f GCUPS on %d cores" % (
    d["value"], d["e2e"]["value"], d["roofline"]["frac"], d["gpu_launches"], d["clocks"], r["value"], r["cpu_baseline"]["cores"]))
PY
# command-line tool: 100k sequences, 3 queries
T=$(mktemp -d)
./tools/osw_synth db -n 100000 -mu 5.706 -sigma 0.6 -seed 5 -o $T/db.fasta
./tools/osw_synth queries -lengths 144,1000,5478 -seed 9 -o $T/q.fasta
( time ./oswald_b200/oswald -O preprocess -i $T/db.fasta -o $T/db ) 2> gpurun_out/${TAG}_cli_preprocess.time
( time ./oswald_b200/oswald -O search -q $T/q.fasta -d $T/db -r 5 ) > gpurun_out/${TAG}_cli_search.txt 2> gpurun_out/${TAG}_cli_search.time
grep -E "Search|GPU|real" gpurun_out/${TAG}_cli_search.txt gpurun_out/${TAG}_cli_search.time gpurun_out/${TAG}_cli_preprocess.time | head -12
rm -rf $T
bash tools/gpu_profile.sh ${TAG}
