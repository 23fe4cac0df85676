#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd|gemm_bf16" -s 2 -c 3 -o gpurun_out/prof_r01_fwd -f python tools/profile_step.py 1 > gpurun_out/prof_r01_fwd.log 2>&1; tail -1 gpurun_out/prof_r01_fwd.log
