"""Small invocations of the kernels with hand-rolled cross-CTA / mbarrier protocols, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py
Covers the 2-SM (cta_group::2) GEMM with every epilogue, MN-major operands and ragged tiles, the split-K wgrad, the implicit-
GEMM convolution, and both attention kernels of version 2 (persistent; several items per CTA) and version 1."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cflearn_b200  # noqa: F401,E402
from cflearn_b200 import _cabi, ops  # noqa: E402

dev = "cuda"
torch.manual_seed(0)


def bf(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "gemm"):
    M, N, K = 304, 520, 232  # ragged in every dimension: 3 m-tiles (pairs 2 + 1), 3 n-tiles, 4 k-blocks (strides stay 16-byte multiples)
    a, b = bf(M, K), bf(N, K, scale=0.1)
    bias = bf(N, scale=0.1)
    h = ops.gemm(a, b, bias=bias)
    act = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ops.gemm(a, b, bias=bias, epilogue=ops.EPI_BIAS_GELU_BF16, out1=act)
    ops.gemm(a, b, bias=bias, epilogue=ops.EPI_BIAS_RESID_F32, aux=torch.randn(M, N, device=dev))
    ops.gemm(a, b, epilogue=ops.EPI_DGELU_BF16, aux=h)
    ops.gemm(a, b, bias=bias, epilogue=ops.EPI_BIAS_QGELU_BF16, out1=act)
    ops.gemm(a, b, epilogue=ops.EPI_DQGELU_BF16, aux=h)
    at, bt = bf(K, M), bf(K, N, scale=0.1)  # MN-major operands (dgrad / wgrad layouts)
    ops.gemm(at, bt, a_mn_major=True, b_mn_major=True)
    dy, x = bf(1000, 264), bf(1000, 136)
    ops.wgrad(dy, x, torch.empty(264, 136, device=dev))
    torch.cuda.synchronize()
    print("gemm cases done")
if which in ("all", "conv"):
    x = bf(2, 16, 16, 128)
    w = bf(72, 128, 3, 3, scale=0.05)
    ops.conv3x3(x, ops.pack_conv3x3_weight(w), bf(72, scale=0.1))
    torch.cuda.synchronize()
    dyc = bf(2, 16, 16, 72)
    ops.conv3x3_wgrad(dyc, x)
    torch.cuda.synchronize()
    print("conv case done")
if which in ("all", "attn"):
    for ver in (2, 3, 1):
        _cabi.lib().b200_set_attention_fwd_version(min(ver, 2))
        _cabi.lib().b200_set_attention_bwd_version(ver)
        for (B, T, H, causal) in ((14, 197, 12, False), (20, 77, 8, True), (3, 50, 2, False)):
            qkv = bf(B, T, 3 * H * 64)
            out, lse = ops.attention_fwd(qkv, B, T, H, causal=causal)
            dbias = torch.zeros(3 * H * 64, device=dev)
            ops.attention_bwd(qkv, out, bf(B * T, H * 64), lse, B, T, H, causal=causal, dbias=dbias)
        torch.cuda.synchronize()
        print(f"attention version {ver} cases done")
