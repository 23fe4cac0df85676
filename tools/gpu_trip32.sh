#!/bin/bash
mkdir -p gpurun_out
echo "=== attention tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention" -x --no-header -p no:cacheprovider 2>&1 | tail -3
echo "=== op bench NEW"; timeout 300 python tools/op_bench.py 2>&1 | tail -3
echo "=== op bench PREV"; B200_LIB_PATH=build/libb200_prev.so timeout 300 python tools/op_bench.py 2>&1 | tail -3
echo "=== op bench NEW"; timeout 300 python tools/op_bench.py 2>&1 | tail -3
echo "=== full gpu test-suite"; timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200
echo "=== memcheck (attention + layernorm + fcnn tests, small shapes)"
timeout 170 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention_fwd or layernorm_fwd" -x --no-header -p no:cacheprovider 2>&1 | tail -6 | tee gpurun_out/memcheck.log
