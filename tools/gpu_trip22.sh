#!/bin/bash
mkdir -p gpurun_out
echo "=== gemm tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "gemm or multicast or wgrad or epilogue" -x --no-header -p no:cacheprovider 2>&1 | tail -5
echo "=== gemm bench (modes 1, 2)"; timeout 600 python tools/gemm_bench.py 2>&1 | tail -10 | tee gpurun_out/gemm_bench.txt
for r in 1 2; do for m in 1 2; do
echo "=== bench mode $m run $r"; B200_GEMM_MULTICAST=$m timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_mode${m}_r$r.log | cut -c1-200
done; done
echo "=== op bench mode 2"; B200_GEMM_MULTICAST=2 timeout 300 python tools/op_bench.py 2>&1 | tail -12 | tee gpurun_out/op_bench_mode2.txt
