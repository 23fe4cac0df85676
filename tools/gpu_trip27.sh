#!/bin/bash
mkdir -p gpurun_out
echo "=== dp_check N=2"; timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 tools/dp_check.py 2>&1 | tail -4 | tee gpurun_out/dp_check.log
echo "=== 2-GPU pytest (bucketed all-reduce test)"; timeout 300 python -m pytest tests/test_model_gpu.py -q -m gpu -k two_gpu --no-header -p no:cacheprovider 2>&1 | tail -3
echo "=== bench N=2"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 2>&1 | tail -1 | tee gpurun_out/bench_n2.log | cut -c1-600
