#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/gemm_bench.py 2>&1 | tee gpurun_out/gemm_bench.txt
