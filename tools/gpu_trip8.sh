#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests (default mode 1)"; bash tools/gpu_kernel_tests.sh 2>&1 | tail -16
echo "=== gemm + model tests in cta_group::2 mode"; B200_GEMM_MULTICAST=2 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -m gpu -k "gemm or wgrad or epilogue or train_step or golden" --no-header -p no:cacheprovider 2>&1 | tail -6
echo "=== model tests (mode 1)"; timeout 600 python -m pytest tests/test_model_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -4
for m in 1 2 0; do
  echo "=== bench mode $m"; B200_GEMM_MULTICAST=$m timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_mode$m.log | cut -c1-200
done
echo "=== launch list mode2"
B200_GEMM_MULTICAST=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_m2.csv python tools/profile_step.py 2 > gpurun_out/launches_m2.log 2>&1; tail -1 gpurun_out/launches_m2.log
echo "=== launch list mode1"
B200_GEMM_MULTICAST=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2 > gpurun_out/launches.log 2>&1; tail -1 gpurun_out/launches.log
