#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests"; bash tools/gpu_kernel_tests.sh 2>&1 | tail -16
echo "=== model tests"; timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -6
echo "=== bench mode1 (multicast)"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_m1.log | cut -c1-330
echo "=== bench mode2 (cta_group::2)"; B200_GEMM_MULTICAST=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_m2.log | cut -c1-330
echo "=== launch list mode1"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2 > gpurun_out/launches.log 2>&1; tail -1 gpurun_out/launches.log
echo "=== launch list mode2"
B200_GEMM_MULTICAST=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_m2.csv python tools/profile_step.py 2 > gpurun_out/launches_m2.log 2>&1; tail -1 gpurun_out/launches_m2.log
