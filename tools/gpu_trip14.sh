#!/bin/bash
mkdir -p gpurun_out
echo "=== bench N=2 (graphed fwd/bwd + flat all-reduce + adam)"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 8 --warmup 3 2>&1 | tail -2 | tee gpurun_out/bench_n2_graph.log | cut -c1-300
echo "=== bench N=2 (eager, overlapped buckets)"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 8 --warmup 3 --overlap-dp 2>&1 | tail -1 | tee gpurun_out/bench_n2_overlap.log | cut -c1-300
echo "=== reference arm under torchrun N=2"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300
