#!/bin/bash
# Round-end rehearsal: exactly what the driver runs (full GPU suite, smoke, bench, reference arm) + the launch list.
mkdir -p gpurun_out
echo "=== full gpu test-suite"; timeout 1200 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/pytest_gpu.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench (default flags)"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_full.log
echo "=== reference arm"; timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.log | cut -c1-400
echo "=== op bench"; timeout 300 python tools/op_bench.py 2>&1 | tail -12 | tee gpurun_out/op_bench_new.txt
echo "=== launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2 > gpurun_out/launches.log 2>&1; tail -1 gpurun_out/launches.log
