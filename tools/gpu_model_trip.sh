#!/bin/bash
# model parity tests, smoke, short bench, launch list (ncu) -> gpurun_out/
mkdir -p gpurun_out
echo "=== kernel wgrad re-test"; timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "wgrad" -x --no-header -p no:cacheprovider 2>&1 | tail -5
echo "=== model tests"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --no-header -p no:cacheprovider -s 2>&1 | tail -40 | tee gpurun_out/model_tests.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench_first.log
