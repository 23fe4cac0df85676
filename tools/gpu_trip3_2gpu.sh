#!/bin/bash
# round-2 trip 3 (2 GPUs): data-parallel parity through the library-owned NCCL communicator (eager / graphed / flat) and the
# three gradient-exchange schedules of bench.py
mkdir -p gpurun_out
timeout 300 python tools/probe_trainer_seam.py > gpurun_out/r02_probe_trainer_seam.txt 2>&1; echo "probe seam rc=$?"; cut -c1-420 gpurun_out/r02_probe_trainer_seam.txt | tail -7
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "attention tests rc=$?"; tail -2 gpurun_out/pytest_attn.log
timeout 300 python tools/op_bench.py attn > gpurun_out/r02_op_bench_attn_auto.txt 2>&1; cat gpurun_out/r02_op_bench_attn_auto.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
timeout 600 $TR tools/dp_check.py > gpurun_out/r02_dp_check.log 2>&1; echo "dp_check rc=$?"; tail -4 gpurun_out/r02_dp_check.log
for mode in graph flat torch; do
  timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 --dp-mode $mode --no-eager-baseline > gpurun_out/bench_r02_n2_$mode.log 2>&1; echo "bench n2 $mode rc=$?"
  grep -o '"value": [0-9.]*, "unit": "samples/s", "n_gpus": 2\|"ms_per_step": [0-9.]*\|"dp_parity_rel": [0-9.e-]*' gpurun_out/bench_r02_n2_$mode.log | head -4
done
timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-eager-baseline --no-cpu-baseline > gpurun_out/bench_r02_n1_ref.log 2>&1; grep -o '"ms_per_step": [0-9.]*' gpurun_out/bench_r02_n1_ref.log | head -1
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 --dp-mode graph > gpurun_out/bench_r02_n2_graph_full.log 2>&1; echo "bench n2 graph (with DDP eager leg) rc=$?"; tail -c 1500 gpurun_out/bench_r02_n2_graph_full.log
timeout 600 $TR bench.py --gpus 2 --config clip --steps 10 --warmup 3 --dp-mode graph > gpurun_out/bench_r02_n2_clip.log 2>&1; echo "bench n2 clip rc=$?"; tail -c 600 gpurun_out/bench_r02_n2_clip.log
timeout 300 python -m pytest tests/test_model_gpu.py -q -k two_gpu > gpurun_out/pytest_2gpu.log 2>&1; echo "2-gpu test rc=$?"; tail -3 gpurun_out/pytest_2gpu.log
