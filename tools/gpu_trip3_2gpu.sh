#!/bin/bash
# round-2 trip 3 (2 GPUs): data-parallel parity through the library-owned NCCL communicator (eager / graphed / flat) and the
# three gradient-exchange schedules of bench.py
mkdir -p gpurun_out
export B200_ATTN_FWD=${B200_ATTN_FWD:-2} B200_ATTN_BWD=${B200_ATTN_BWD:-2}
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
