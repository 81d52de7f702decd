timeout 1500 python -m pytest tests -q -x -m gpu -p no:cacheprovider --timeout 600 > gpurun_out/t8_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/t8_pytest_gpu.log
CASES="c1 c1_two s50k" bash tools/gpu_small.sh t8
