#!/bin/bash
# round-2 trip 2: re-run of what trip 1 left red (graph capture fix, tap tests, bench), conv-bias probe, ncu of the new attention
# kernels, launch list, sanitizers.  1 GPU.
mkdir -p gpurun_out
export B200_ATTN_FWD=2 B200_ATTN_BWD=2
timeout 120 python tools/probe_conv_bias.py > gpurun_out/r02_probe_conv_bias.txt 2>&1; cat gpurun_out/r02_probe_conv_bias.txt
timeout 300 python tools/probe_layout.py > gpurun_out/r02_probe_layout.txt 2>&1; cat gpurun_out/r02_probe_layout.txt
timeout 900 python tools/tap_parity.py vit_b16 256 > gpurun_out/tap_parity_b256.txt 2>&1; echo "tap_parity rc=$?"; grep -n "worst op\|stem" gpurun_out/tap_parity_b256.txt
timeout 900 python -m pytest tests/test_taps_gpu.py -q -s > gpurun_out/pytest_taps.log 2>&1; echo "taps rc=$?"; tail -8 gpurun_out/pytest_taps.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_taps_gpu.py > gpurun_out/pytest_gpu.log 2>&1; echo "gpu tests rc=$?"; tail -12 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02_v2.log 2>&1; echo "bench(v2 attention) rc=$?"; tail -c 2500 gpurun_out/bench_r02_v2.log
B200_ATTN_FWD=1 B200_ATTN_BWD=1 timeout 900 python bench.py --steps 10 --warmup 3 --no-eager-baseline --no-cpu-baseline > gpurun_out/bench_r02_v1.log 2>&1; echo "bench(v1 attention) rc=$?"; tail -c 600 gpurun_out/bench_r02_v1.log
timeout 600 python bench.py --config clip --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r02_clip.log 2>&1; echo "bench clip rc=$?"; tail -c 1200 gpurun_out/bench_r02_clip.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2 -s 2 -c 1 -o gpurun_out/r02_attn_fwd2 python tools/op_bench.py attn > gpurun_out/r02_prof_attn_fwd2.log 2>&1; echo "ncu fwd2 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd2 -s 2 -c 1 -o gpurun_out/r02_attn_bwd2 python tools/op_bench.py attn > gpurun_out/r02_prof_attn_bwd2.log 2>&1; echo "ncu bwd2 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python tools/profile_step.py 2 > gpurun_out/r02_launches.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/r02_launches.csv > gpurun_out/r02_step_launches.txt 2>&1; head -24 gpurun_out/r02_step_launches.txt
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_cases.py > gpurun_out/r02_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/r02_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize_cases.py > gpurun_out/r02_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r02_racecheck.txt
ls -la gpurun_out/*.ncu-rep 2>/dev/null | tail -3
