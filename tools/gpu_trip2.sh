#!/bin/bash
# round-2 trip 2: launch list of one step, full ncu captures of the top kernels, sanitizers (1 GPU)
mkdir -p gpurun_out
export B200_ATTN_FWD=${B200_ATTN_FWD:-2} B200_ATTN_BWD=${B200_ATTN_BWD:-2}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python tools/profile_step.py 2 > gpurun_out/r02_launches.log 2>&1; echo "launch list rc=$?"
python tools/launch_summary.py gpurun_out/r02_launches.csv > gpurun_out/r02_step_launches.txt 2>&1; head -30 gpurun_out/r02_step_launches.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2 -s 2 -c 1 -o gpurun_out/r02_attn_fwd2 python tools/op_bench.py attn > gpurun_out/r02_prof_attn_fwd2.log 2>&1; echo "ncu fwd2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_bwd2 -s 2 -c 1 -o gpurun_out/r02_attn_bwd2 python tools/op_bench.py attn > gpurun_out/r02_prof_attn_bwd2.log 2>&1; echo "ncu bwd2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_kernel -s 12 -c 6 -o gpurun_out/r02_gemm python tools/op_bench.py > gpurun_out/r02_prof_gemm.log 2>&1; echo "ncu gemm rc=$?"
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_cases.py > gpurun_out/r02_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r02_memcheck.txt
timeout 1500 compute-sanitizer --tool racecheck python tools/sanitize_cases.py > gpurun_out/r02_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/r02_racecheck.txt
ls -la gpurun_out/*.ncu-rep
