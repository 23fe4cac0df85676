#!/bin/bash
# Runs the kernel parity tests group by group, each in its own process under a timeout, so a trapped kernel
# (bounded mbarrier wait -> __trap) only takes its own group down.  Logs go to gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() { # name, -k expression
  echo "=== $1" | tee -a gpurun_out/kernel_tests.log
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$2" -x --no-header -p no:cacheprovider 2>&1 | tail -25 | tee -a gpurun_out/kernel_tests.log
}
: > gpurun_out/kernel_tests.log
run rowops "cast or layernorm or colsum or patch or softmax"
run gemm_kmajor "gemm_kmajor"
run gemm_mn "gemm_mn_major"
run gemm_persist "gemm_persistent or multicast"
run gemm_epi "gelu_epilogue or residual_epilogue or dgelu_epilogue or wgrad"
run attn_fwd "attention_fwd"
run attn_bwd "attention_bwd"
grep -E "===|passed|failed|error" gpurun_out/kernel_tests.log | tail -40
