#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/eager_gpu_bench.py 256 2>&1 | tail -3 | tee gpurun_out/eager_gpu.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_t13.log | cut -c1-200
