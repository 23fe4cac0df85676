#!/bin/bash
# round-2 trip 8 (8 GPUs, minimal): does the in-graph gradient exchange on the library-owned communicator work at N = 8
# (NVLS, 8-way unique-id broadcast), and what does one bench line look like.  The driver measures the 1 -> 8 scaling itself.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
timeout 150 $TR tools/dp_check.py > gpurun_out/r02_dp_check_n8.log 2>&1; echo "dp_check rc=$?"; tail -1 gpurun_out/r02_dp_check_n8.log
timeout 200 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-eager-baseline --no-cpu-baseline > gpurun_out/r02_bench_n8.log 2>&1; echo "bench n8 rc=$?"
grep -o '"value": [0-9.]*, "unit": "samples/s", "n_gpus": 8\|"ms_per_step": [0-9.]*\|"dp_parity_rel": [0-9.e-]*' gpurun_out/r02_bench_n8.log | head -4
timeout 120 python bench.py --gpus 1 --steps 10 --warmup 3 --no-eager-baseline --no-cpu-baseline > gpurun_out/r02_bench_n1_8box.log 2>&1; echo "n1: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_n1_8box.log | head -1)"
