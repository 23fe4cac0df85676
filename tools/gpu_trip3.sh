#!/bin/bash
mkdir -p gpurun_out
echo "=== parity report"; timeout 600 python tools/parity_report.py 2>&1 | tee gpurun_out/parity_report.txt | tail -80
echo "=== model tests"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/model_tests.log
echo "=== launch list (2 steps under ncu; per-launch times are cold-cache/serialised)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2 > gpurun_out/launches.log 2>&1; tail -2 gpurun_out/launches.log
echo "=== ncu full on the GEMM kernels (3 launches each of two epilogues)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 30 -c 6 -o gpurun_out/prof_gemm -f python tools/profile_step.py 1 > gpurun_out/prof_gemm.log 2>&1; tail -2 gpurun_out/prof_gemm.log
ls -la gpurun_out
