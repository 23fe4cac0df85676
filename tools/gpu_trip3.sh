#!/bin/bash
# round-2 trip 3 (1 GPU): second cut of the attention v2 kernels -- tests in isolated processes, A/B timing, sanitizers, then the
# whole GPU suite and the bench with whichever versions passed.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "attention_fwd" > gpurun_out/pytest_attn_fwd.log 2>&1; A=$?; echo "attention fwd tests rc=$A"; tail -3 gpurun_out/pytest_attn_fwd.log
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "attention_bwd" > gpurun_out/pytest_attn_bwd.log 2>&1; A2=$?; echo "attention bwd tests rc=$A2"; tail -3 gpurun_out/pytest_attn_bwd.log
FW=1; BW=1; [ $A -eq 0 ] && FW=2; [ $A2 -eq 0 ] && BW=2
B200_ATTN_FWD=$FW B200_ATTN_BWD=$BW timeout 300 python tools/op_bench.py attn > gpurun_out/r02_op_bench_attn_v2b.txt 2>&1; cat gpurun_out/r02_op_bench_attn_v2b.txt
B200_ATTN_FWD=1 B200_ATTN_BWD=1 timeout 300 python tools/op_bench.py attn > gpurun_out/r02_op_bench_attn_v1.txt 2>&1; cat gpurun_out/r02_op_bench_attn_v1.txt
export B200_ATTN_FWD=$FW B200_ATTN_BWD=$BW
echo "running the rest with B200_ATTN_FWD=$B200_ATTN_FWD B200_ATTN_BWD=$B200_ATTN_BWD"
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_cases.py > gpurun_out/r02_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize_cases.py > gpurun_out/r02_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02_racecheck.txt
timeout 900 python -m pytest tests/test_taps_gpu.py -q -s > gpurun_out/pytest_taps.log 2>&1; echo "taps rc=$?"; tail -4 gpurun_out/pytest_taps.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_taps_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -6 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_b.log 2>&1; echo "bench rc=$?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_r02_b.log | head -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2 -s 2 -c 1 -o gpurun_out/r02b_attn_fwd2 python tools/op_bench.py attn > /dev/null 2>&1; echo "ncu fwd2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd2 -s 2 -c 1 -o gpurun_out/r02b_attn_bwd2 python tools/op_bench.py attn > /dev/null 2>&1; echo "ncu bwd2 rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:gemm_bf16_kernel -s 12 -c 7 -o gpurun_out/r02_gemm python tools/op_bench.py > gpurun_out/r02_prof_gemm.log 2>&1; echo "ncu gemm rc=$?"
