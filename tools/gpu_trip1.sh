#!/bin/bash
# round-2 trip 1: new parity tests, table, smoke, bench (1 GPU); the new attention forward (v2) is tested in its own process first
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "attention_fwd" > gpurun_out/pytest_attn_fwd.log 2>&1; A=$?; echo "attention fwd tests rc=$A"
tail -5 gpurun_out/pytest_attn_fwd.log
if [ $A -eq 0 ]; then
  B200_ATTN_BWD=1 timeout 300 python tools/op_bench.py attn > gpurun_out/op_bench_attn_v2.txt 2>&1; B200_ATTN_BWD=1 B200_ATTN_FWD=1 timeout 300 python tools/op_bench.py attn > gpurun_out/op_bench_attn_v1.txt 2>&1
  cat gpurun_out/op_bench_attn_v2.txt gpurun_out/op_bench_attn_v1.txt
  export B200_ATTN_FWD=2
else
  export B200_ATTN_FWD=1
fi
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "attention_bwd" > gpurun_out/pytest_attn_bwd.log 2>&1; A2=$?; echo "attention bwd tests rc=$A2"
tail -5 gpurun_out/pytest_attn_bwd.log
if [ $A2 -eq 0 ]; then
  B200_ATTN_FWD=$B200_ATTN_FWD timeout 300 python tools/op_bench.py attn > gpurun_out/op_bench_attn_bwd_v2.txt 2>&1; cat gpurun_out/op_bench_attn_bwd_v2.txt
  export B200_ATTN_BWD=2
else
  export B200_ATTN_BWD=1
fi
echo "running the rest with B200_ATTN_FWD=$B200_ATTN_FWD B200_ATTN_BWD=$B200_ATTN_BWD"
timeout 900 python tools/tap_parity.py vit_b16 256 > gpurun_out/tap_parity_b256.txt 2>&1; echo "tap_parity rc=$?"
timeout 900 python -m pytest tests/test_taps_gpu.py -q -s > gpurun_out/pytest_taps.log 2>&1; echo "taps rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_taps_gpu.py -k "not attention_fwd and not attention_bwd" > gpurun_out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_a.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/tap_parity_b256.txt; tail -5 gpurun_out/pytest_taps.log; tail -5 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; tail -c 1500 gpurun_out/bench_r02_a.log
