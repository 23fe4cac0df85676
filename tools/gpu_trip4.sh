#!/bin/bash
# round-2 trip 4 (1 GPU): third cut of the attention v2 kernels (loop-based two-pass softmax; lean MMA thread + ping-pong compute
# groups + packed fp32x2 arithmetic in the backward), isolated tests first, A/B timing of every version, sanitizers, then the
# whole GPU suite, the bench, and the ncu captures (attention kernels with source, the wgrad GEMM for its DRAM traffic).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "attention_fwd" > gpurun_out/pytest_attn_fwd.log 2>&1; A=$?; echo "attention fwd tests rc=$A"; tail -3 gpurun_out/pytest_attn_fwd.log
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "attention_bwd" > gpurun_out/pytest_attn_bwd.log 2>&1; A2=$?; echo "attention bwd tests rc=$A2"; tail -3 gpurun_out/pytest_attn_bwd.log
timeout 300 python tools/op_bench.py attn > gpurun_out/r02_op_bench_attn_c3.txt 2>&1; cat gpurun_out/r02_op_bench_attn_c3.txt
[ $A -ne 0 ] && export B200_ATTN_FWD=1
[ $A2 -ne 0 ] && export B200_ATTN_BWD=1
echo "running the rest with B200_ATTN_FWD=${B200_ATTN_FWD:-auto} B200_ATTN_BWD=${B200_ATTN_BWD:-auto}"
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_cases.py > gpurun_out/r02_memcheck_final.txt 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02_memcheck_final.txt
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize_cases.py attn > gpurun_out/r02_racecheck_final.txt 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r02_racecheck_final.txt
timeout 300 python tools/probe_trainer_seam.py > gpurun_out/r02_probe_trainer_seam.txt 2>&1; echo "probe seam rc=$?"
timeout 900 python -m pytest tests/test_taps_gpu.py -q -s > gpurun_out/pytest_taps.log 2>&1; echo "taps rc=$?"; tail -4 gpurun_out/pytest_taps.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_taps_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_c.log 2>&1; echo "bench rc=$?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_r02_c.log | head -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2 -s 2 -c 1 -o gpurun_out/r02c_attn_fwd2 python tools/op_bench.py attn > /dev/null 2>&1; echo "ncu fwd2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd2 -s 2 -c 1 -o gpurun_out/r02c_attn_bwd2 python tools/op_bench.py attn > /dev/null 2>&1; echo "ncu bwd2 rc=$?"
timeout 600 ncu --set full --clock-control none --kernel-name-base demangled -k 'regex:gemm_bf16_kernel<\(int\)4' -s 3 -c 1 -o gpurun_out/r02_gemm_wgrad python tools/op_bench.py > gpurun_out/r02_prof_gemm.log 2>&1; echo "ncu wgrad rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:gemm_bf16_kernel<\(int\)1' -s 3 -c 1 -o gpurun_out/r02_gemm_gelu python tools/op_bench.py > /dev/null 2>&1; echo "ncu gelu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:gemm_bf16_kernel<\(int\)3' -s 3 -c 1 -o gpurun_out/r02_gemm_dgelu python tools/op_bench.py > /dev/null 2>&1; echo "ncu dgelu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:layernorm_bwd -s 2 -c 1 -o gpurun_out/r02_ln_bwd python bench.py --steps 1 --warmup 1 --no-eager-baseline --no-cpu-baseline > /dev/null 2>&1; echo "ncu ln_bwd rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -8
