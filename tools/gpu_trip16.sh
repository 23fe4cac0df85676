#!/bin/bash
mkdir -p gpurun_out
echo "=== model tests"; timeout 900 python -m pytest tests/test_model_gpu.py -q -m gpu --no-header -p no:cacheprovider 2>&1 | tail -8 | tee gpurun_out/model_tests.log
for r in 1 2; do
for m in 0 1; do
echo "=== bench side_colsum=$m run $r"; B200_SIDE_COLSUM=$m timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_side${m}_r${r}.log | cut -c1-200
done; done
echo "=== full bench"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_full.log
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.log
echo "=== launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/profile_step.py 2 > gpurun_out/launches.log 2>&1; tail -1 gpurun_out/launches.log
