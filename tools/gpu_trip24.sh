#!/bin/bash
mkdir -p gpurun_out
echo "=== gemm tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm or multicast or wgrad or epilogue" -x --no-header -p no:cacheprovider 2>&1 | tail -5
echo "=== op bench NEW (relaxed tmem_empty arrive)"; timeout 300 python tools/op_bench.py 2>&1 | tail -12 | tee gpurun_out/op_bench_new.txt
echo "=== op bench PREV (release.cluster arrive)"; B200_LIB_PATH=build/libb200_prev.so timeout 300 python tools/op_bench.py 2>&1 | tail -12 | tee gpurun_out/op_bench_prev.txt
echo "=== model tests"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -3
for r in 1 2; do
echo "=== bench NEW $r"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_new$r.log | cut -c1-200
echo "=== bench PREV $r"; B200_LIB_PATH=build/libb200_prev.so timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_prev$r.log | cut -c1-200
done
