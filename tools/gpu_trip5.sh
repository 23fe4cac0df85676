#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests"; bash tools/gpu_kernel_tests.sh 2>&1 | tail -16
echo "=== model tests"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/model_tests.log
echo "=== bench (graph, multicast)"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_graph_mc.log | cut -c1-400
echo "=== bench (no graph, multicast)"; timeout 900 python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_nograph_mc.log | cut -c1-330
echo "=== bench (graph, unicast)"; B200_GEMM_MULTICAST=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_graph_uc.log | cut -c1-330
echo "=== launch list (multicast)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2 > gpurun_out/launches.log 2>&1; tail -1 gpurun_out/launches.log
echo "=== ncu full: gemm<0>, <1>, attn_bwd"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"gemm_bf16_kernel<0>|gemm_bf16_kernel<1>|attn_bwd" -s 6 -c 5 -o gpurun_out/prof_gemm_mc -f python tools/profile_step.py 1 > gpurun_out/prof_gemm_mc.log 2>&1; tail -1 gpurun_out/prof_gemm_mc.log
