#!/bin/bash
# round-2 trip 5 (2 GPUs, short): the in-graph gradient exchange with every persistent kernel shrunk while buckets are in flight;
# communicator CTA count swept; N=1 on the same box first (box-to-box clocks differ by ~8 %).
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -k "attention" > gpurun_out/pytest_attn.log 2>&1; echo "attention tests rc=$?"; tail -1 gpurun_out/pytest_attn.log
timeout 200 python tools/op_bench.py attn > gpurun_out/r02_op_bench_attn_final.txt 2>&1; head -8 gpurun_out/r02_op_bench_attn_final.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541"
B="bench.py --steps 10 --warmup 3 --no-eager-baseline --no-cpu-baseline"
timeout 300 python $B --gpus 1 > gpurun_out/r02_sweep_n1.log 2>&1; echo "n1: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_sweep_n1.log | head -1)"
timeout 240 $TR tools/dp_check.py > gpurun_out/r02_dp_check.log 2>&1; echo "dp_check rc=$?"; tail -1 gpurun_out/r02_dp_check.log
for cfg in "8 all" "4 all" "16 all" "16 next"; do
  set -- $cfg
  B200_COMM_CTAS=$1 B200_DP_SHRINK=$2 timeout 300 $TR $B --gpus 2 --dp-mode graph > gpurun_out/r02_sweep_n2_c$1_$2.log 2>&1
  echo "n2 ctas=$1 shrink=$2 rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_sweep_n2_c$1_$2.log | head -1) $(grep -o '"dp_parity_rel": [0-9.e-]*' gpurun_out/r02_sweep_n2_c$1_$2.log)"
done
timeout 300 $TR $B --gpus 2 --dp-mode flat > gpurun_out/r02_sweep_n2_flat.log 2>&1; echo "n2 flat rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_sweep_n2_flat.log | head -1)"
timeout 300 $TR bench.py --gpus 2 --config clip --steps 10 --warmup 3 > gpurun_out/r02_sweep_n2_clip.log 2>&1; echo "n2 clip rc=$?: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_sweep_n2_clip.log | head -1)"
timeout 420 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_n2_default_full.log 2>&1; echo "n2 default (with DDP eager leg) rc=$?"; grep -o '"eager_gpu": {[^}]*}' gpurun_out/r02_n2_default_full.log
timeout 200 python -m pytest tests/test_model_gpu.py -q -k two_gpu > gpurun_out/pytest_2gpu.log 2>&1; echo "2-gpu test rc=$?"
