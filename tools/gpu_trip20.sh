#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests"; bash tools/gpu_kernel_tests.sh 2>&1 | tail -16
echo "=== op bench NEW"; timeout 300 python tools/op_bench.py 2>&1 | tail -12 | tee gpurun_out/op_bench_new.txt
echo "=== op bench PREV"; B200_LIB_PATH=build/libb200_prev.so timeout 300 python tools/op_bench.py 2>&1 | tail -12 | tee gpurun_out/op_bench_prev.txt
echo "=== model tests"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/model_tests.log
echo "=== bench NEW"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_new.log | cut -c1-200
echo "=== bench PREV"; B200_LIB_PATH=build/libb200_prev.so timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_prev.log | cut -c1-200
echo "=== bench NEW again"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_new2.log | cut -c1-200
