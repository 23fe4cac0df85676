#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests"; bash tools/gpu_kernel_tests.sh 2>&1 | tail -16
echo "=== model tests"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -30 | tee gpurun_out/model_tests.log
echo "=== parity report"; timeout 600 python tools/parity_report.py > gpurun_out/parity_report.txt 2>&1; grep -E "^##|eager-bf16\(|logits:|median" gpurun_out/parity_report.txt
echo "=== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_r01_b.log
echo "=== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2 > gpurun_out/launches.log 2>&1; tail -1 gpurun_out/launches.log
echo "=== ncu full: attention + layernorm_bwd"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"attn_|layernorm_bwd" -s 8 -c 4 -o gpurun_out/prof_attn_ln -f python tools/profile_step.py 1 > gpurun_out/prof_attn_ln.log 2>&1; tail -1 gpurun_out/prof_attn_ln.log
ls -la gpurun_out | head -30
