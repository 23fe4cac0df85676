#!/bin/bash
mkdir -p gpurun_out
echo "=== quick gelu kernel test"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "quick_gelu" -x --no-header -p no:cacheprovider 2>&1 | tail -4
echo "=== op bench NEW"; timeout 300 python tools/op_bench.py 2>&1 | tail -12 | tee gpurun_out/op_bench_new.txt
echo "=== op bench PREV (before the act flag)"; B200_LIB_PATH=build/libb200_prev.so timeout 300 python tools/op_bench.py 2>&1 | tail -12 | tee gpurun_out/op_bench_prev.txt
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_new.log | cut -c1-200
