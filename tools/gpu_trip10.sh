#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 1 -c 4 -o gpurun_out/prof_gemm_epi -f python tools/profile_step.py 1 > gpurun_out/prof_gemm_epi.log 2>&1; tail -2 gpurun_out/prof_gemm_epi.log
ls -la gpurun_out/prof_gemm_epi.ncu-rep
