"""BASELINE.json configs[0] (FCNN 10 -> 32 -> 32 -> 1, batch 128, multi_task[mae, mse]): steps/s of the fused B200 step
(one launch + two tiny reductions, CUDA-event timed) next to the fp32 oracle on this box's host cores."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cflearn_b200  # noqa: F401,E402
import fcnn_oracle as fo  # noqa: E402
from cflearn_b200 import registry  # noqa: E402

dev = torch.device("cuda", 0)
x_all, y_all = fo.toy_data()
x, y = x_all[:128], y_all[:128]
sd = fo.init_state_dict(10, 1, seed=0)
m = registry.build_module("fcnn", config=dict(input_dim=10, output_dim=1)).to(dev)
m.load_state_dict(sd)
xd, yd = x.to(dev), y.to(dev)
for _ in range(20):
    m.train_step(xd, yd)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 2000
e0.record()
for _ in range(n):
    m.train_step(xd, yd)
e1.record()
torch.cuda.synchronize()
gpu_us = e0.elapsed_time(e1) / n * 1e3
torch.set_num_threads(1)
for _ in range(20):
    fo.train_step(sd, x, y)
t0 = time.perf_counter()
k = 500
for _ in range(k):
    fo.train_step(sd, x, y)
cpu_us = (time.perf_counter() - t0) / k * 1e6
print(f"FCNN 10-32-32-1, batch 128, fwd + mae/mse + bwd: B200 fused step {gpu_us:.1f} us/step ({128e6 / gpu_us:.0f} samples/s, launch-bound) | "
      f"oracle (torch fp32, 1 host thread) {cpu_us:.1f} us/step ({128e6 / cpu_us:.0f} samples/s)")
