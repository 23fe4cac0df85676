#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 1 -c 4 -o gpurun_out/prof_r01_gemm_fwd4 -f python tools/profile_step.py 1 > gpurun_out/prof_r01_gemm_fwd4.log 2>&1; tail -1 gpurun_out/prof_r01_gemm_fwd4.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 52 -c 1 -o gpurun_out/prof_r01_gemm_bwd4 -f python tools/profile_step.py 1 > gpurun_out/prof_r01_gemm_bwd4.log 2>&1; tail -1 gpurun_out/prof_r01_gemm_bwd4.log
ls -la gpurun_out/*4.ncu-rep
