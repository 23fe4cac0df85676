#!/bin/bash
mkdir -p gpurun_out
echo "=== full gpu test-suite (as the driver runs it)"; timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "=== fcnn bench"; timeout 300 python tools/fcnn_bench.py 2>&1 | tail -2 | tee gpurun_out/fcnn_bench.txt
echo "=== bench (default flags)"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_full.log
