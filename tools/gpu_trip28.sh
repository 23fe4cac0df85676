#!/bin/bash
mkdir -p gpurun_out
echo "=== new kernel tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "quick_gelu or fp32_upstream or layernorm" -x --no-header -p no:cacheprovider 2>&1 | tail -6
echo "=== clip vision tests"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu -k "clip" -x -s --no-header -p no:cacheprovider 2>&1 | tail -12
echo "=== full gpu test-suite"; timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu.log
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_new.log | cut -c1-200
