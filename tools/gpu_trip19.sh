#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests"; bash tools/gpu_kernel_tests.sh 2>&1 | tail -16
echo "=== op bench NEW"; timeout 300 python tools/op_bench.py 2>&1 | tail -12 | tee gpurun_out/op_bench_new.txt
echo "=== op bench BASE"; B200_LIB_PATH=build/libb200_base.so timeout 300 python tools/op_bench.py 2>&1 | tail -12 | tee gpurun_out/op_bench_base.txt
echo "=== model tests"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -5 | tee gpurun_out/model_tests.log
echo "=== bench NEW"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_new.log | cut -c1-200
echo "=== bench NEW again"; timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_new2.log | cut -c1-200
echo "=== ncu full: layer-0 forward GEMMs (qkv <0>, out-proj <2>, ff1 <1>, ff2 <2>) and the first dGELU dgrad <3>"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 1 -c 4 -o gpurun_out/prof_r01_gemm_fwd3 -f python tools/profile_step.py 1 > gpurun_out/prof_r01_gemm_fwd3.log 2>&1; tail -1 gpurun_out/prof_r01_gemm_fwd3.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 52 -c 1 -o gpurun_out/prof_r01_gemm_bwd3 -f python tools/profile_step.py 1 > gpurun_out/prof_r01_gemm_bwd3.log 2>&1; tail -1 gpurun_out/prof_r01_gemm_bwd3.log
ls -la gpurun_out/*.ncu-rep
