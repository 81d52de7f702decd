for f in 0.7 0.5 0.35; do echo "== long frac $f"; CASES="s50k s100k s200k" bash tools/gpu_small.sh t13_$f OSW_TRANSPOSE=2 OSW_T16_LONG_FRAC=$f; done
