#!/bin/bash
mkdir -p gpurun_out
echo "=== quick gelu kernel test"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "quick_gelu" -x --no-header -p no:cacheprovider 2>&1 | grep -E "Error|assert|passed|failed" | head -8
echo "=== clip text / vision tests"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -k "clip" -x -s --no-header -p no:cacheprovider 2>&1 | tail -12
echo "=== full gpu test-suite"; timeout 1200 python -m pytest tests/ -q -m gpu 2>&1 | tail -6 | tee gpurun_out/pytest_gpu.log
echo "=== op bench"; timeout 300 python tools/op_bench.py 2>&1 | tail -12 | tee gpurun_out/op_bench_new.txt
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_new.log | cut -c1-200
