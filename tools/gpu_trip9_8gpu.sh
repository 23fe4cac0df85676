#!/bin/bash
# round-2 trip 9 (8 GPUs): how many persistent launches behind a bucket should leave the communicator's SMs free at N = 8
# (trip 8: 36.2 ms at N = 8 against 32.4 ms at N = 1 with only the next GEMM shrunk -- the 8-rank all-reduces outlast it)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541"
B="bench.py --gpus 8 --steps 10 --warmup 3 --no-eager-baseline --no-cpu-baseline"
for cfg in "16 next" "16 next2" "16 next4" "8 next3"; do
  set -- $cfg
  B200_COMM_CTAS=$1 B200_DP_SHRINK=$2 timeout 150 $TR $B > gpurun_out/r02_n8_c$1_$2.log 2>&1
  echo "n8 ctas=$1 shrink=$2 rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_n8_c$1_$2.log | head -1) $(grep -o '"dp_parity_rel": [0-9.e-]*' gpurun_out/r02_n8_c$1_$2.log)"
done
