#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
echo "=== attention + gemm kernel tests"; timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "attention or multicast or gemm_persistent" --no-header -p no:cacheprovider 2>&1 | tail -5
echo "=== dp check (2 GPUs)"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py 2>&1 | tail -5
echo "=== bench N=2 (eager)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_n2.log | cut -c1-400
echo "=== bench N=2 (graph incl. NCCL)"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --graph-dp 2>&1 | tail -2 | tee gpurun_out/bench_n2_graph.log | cut -c1-400
echo "=== bench N=1"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_n1.log | cut -c1-400
echo "=== reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_ref.log | cut -c1-500
