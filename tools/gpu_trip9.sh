#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests"; bash tools/gpu_kernel_tests.sh 2>&1 | tail -16
echo "=== model tests"; timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -4
echo "=== bench"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_t9.log | cut -c1-200
echo "=== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2 > gpurun_out/launches.log 2>&1; tail -1 gpurun_out/launches.log
