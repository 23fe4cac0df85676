"""Micro-benchmark of the tcgen05 GEMM: operand majors x shapes x cluster modes, on one box (same clocks)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cflearn_b200  # noqa: F401,E402
from cflearn_b200 import _cabi, ops  # noqa: E402

dev = "cuda"


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


shapes = [(50432, 2304, 768), (50432, 768, 768), (50432, 3072, 768), (50432, 768, 3072)]
for mode in (1, 2):
    _cabi.lib().b200_set_gemm_multicast(mode)
    for (M, N, K) in shapes:
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        b = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
        at, bt = a.t().contiguous(), b.t().contiguous()
        bias = torch.zeros(N, device=dev, dtype=torch.bfloat16)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        res = []
        for a_mn, b_mn in ((0, 0), (0, 1), (1, 0), (1, 1)):
            ms = timeit(lambda: ops.gemm(at if a_mn else a, bt if b_mn else b, a_mn_major=bool(a_mn), b_mn_major=bool(b_mn), bias=bias, out0=out))
            res.append(f"A{'mn' if a_mn else 'k'}B{'mn' if b_mn else 'k'} {ms*1e3:6.1f}us {2*M*N*K/ms/1e9:6.0f}TF")
        ref = timeit(lambda: torch.matmul(a, b.t()))
        print(f"mode{mode} {M}x{N}x{K}: " + " | ".join(res) + f" | cuBLAS {ref*1e3:6.1f}us {2*M*N*K/ref/1e9:6.0f}TF", flush=True)
