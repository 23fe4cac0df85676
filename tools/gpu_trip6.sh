#!/bin/bash
# round-2 trip 6 (1 GPU, short): attention v2 fourth cut (separate masked / unmasked softmax paths; backward: early kvfree / dqfree
# arrivals, S^T|dP^T queued ahead of dQ) -- tests, A/B timing, suite, bench, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -k "attention" > gpurun_out/pytest_attn.log 2>&1; A=$?; echo "attention tests rc=$A"; tail -3 gpurun_out/pytest_attn.log
timeout 300 python tools/op_bench.py attn > gpurun_out/r02_op_bench_attn_c4.txt 2>&1; cat gpurun_out/r02_op_bench_attn_c4.txt
[ $A -ne 0 ] && export B200_ATTN_FWD=1 B200_ATTN_BWD=1
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_cases.py attn > gpurun_out/r02_memcheck_c4.txt 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r02_memcheck_c4.txt
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_d.log 2>&1; echo "bench rc=$?"; grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*' gpurun_out/bench_r02_d.log | head -3
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2 -s 2 -c 1 -o gpurun_out/r02d_attn_fwd2 python tools/op_bench.py attn > /dev/null 2>&1; echo "ncu fwd2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd2 -s 2 -c 1 -o gpurun_out/r02d_attn_bwd2 python tools/op_bench.py attn > /dev/null 2>&1; echo "ncu bwd2 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02d_launches.csv python tools/profile_step.py > gpurun_out/r02d_launches.log 2>&1; echo "launch list rc=$?"
