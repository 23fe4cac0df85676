"""Per-op timings of the ViT-B/16 step's kernels at the bench shapes (B = 256), CUDA events, one box.

Run twice with B200_LIB_PATH pointing at two builds to A/B a kernel change under the same clocks:
    python tools/op_bench.py; B200_LIB_PATH=build/libb200_base.so python tools/op_bench.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cflearn_b200  # noqa: F401,E402
from cflearn_b200 import ops  # noqa: E402

dev = "cuda"
B, T, H, D, FF = 256, 197, 12, 768, 3072
M = B * T


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
    return e0.elapsed_time(e1) / reps * 1e3


def bf(*shape, scale=1.0):
    return (torch.randn(*shape, device=dev) * scale).to(torch.bfloat16)


x768, x3072 = bf(M, D), bf(M, FF)
w_qkv, w_o, w_1, w_2 = bf(3 * D, D, scale=0.02), bf(D, D, scale=0.02), bf(FF, D, scale=0.02), bf(D, FF, scale=0.02)
b_qkv, b_o, b_1 = bf(3 * D, scale=0.02), bf(D, scale=0.02), bf(FF, scale=0.02)
resid = torch.randn(M, D, device=dev)
h = bf(M, FF)
act = torch.empty(M, FF, device=dev, dtype=torch.bfloat16)
gw = torch.empty(FF, D, device=dev)
rows = []
only_attn = len(sys.argv) > 1 and sys.argv[1] == "attn"
if only_attn:
    qkv = bf(B, T, 3 * D)
    out, lse = ops.attention_fwd(qkv, B, T, H)
    dout = bf(B, T, D)
    rows.append(("attention fwd  B256 T197 H12", timeit(lambda: ops.attention_fwd(qkv, B, T, H)), 4 * B * H * T * T * 64))
    rows.append(("attention bwd  B256 T197 H12", timeit(lambda: ops.attention_bwd(qkv, out, dout, lse, B, T, H)), 10 * B * H * T * T * 64))
    from cflearn_b200 import _cabi
    for ver in (1, 2, 3):  # forced kernels (the first two rows use the per-shape default unless B200_ATTN_* is set)
        prev = _cabi.lib().b200_set_attention_bwd_version(ver)
        rows.append((f"attention bwd  B256 T197 H12, forced version {ver}", timeit(lambda: ops.attention_bwd(qkv, out, dout, lse, B, T, H)), 10 * B * H * T * T * 64))
        _cabi.lib().b200_set_attention_bwd_version(prev)
    for ver in (1, 2):
        prev = _cabi.lib().b200_set_attention_fwd_version(ver)
        rows.append((f"attention fwd  B256 T197 H12, forced version {ver}", timeit(lambda: ops.attention_fwd(qkv, B, T, H)), 4 * B * H * T * T * 64))
        _cabi.lib().b200_set_attention_fwd_version(prev)
    qk = bf(256, 77, 3 * 512)
    rows.append(("attention fwd  B256 T77 H8 causal (CLIP text)", timeit(lambda: ops.attention_fwd(qk, 256, 77, 8, causal=True)), 4 * 256 * 8 * 77 * 77 * 64))
    qv = bf(256, 50, 3 * 768)
    rows.append(("attention fwd  B256 T50 H12 (CLIP vision)", timeit(lambda: ops.attention_fwd(qv, 256, 50, 12)), 4 * 256 * 12 * 50 * 50 * 64))
    for name, us, fl in rows:
        print(f"{name:48s} {us:8.1f} us  {fl / us / 1e6:7.0f} TFLOP/s", flush=True)
    sys.exit(0)
rows.append(("gemm qkv       bias      50432x2304x768", timeit(lambda: ops.gemm(x768, w_qkv, bias=b_qkv)), 2 * M * 3 * D * D))
rows.append(("gemm out-proj  resid f32 50432x768x768", timeit(lambda: ops.gemm(x768, w_o, bias=b_o, epilogue=ops.EPI_BIAS_RESID_F32, aux=resid)), 2 * M * D * D))
rows.append(("gemm ff1       gelu      50432x3072x768", timeit(lambda: ops.gemm(x768, w_1, bias=b_1, epilogue=ops.EPI_BIAS_GELU_BF16, out1=act)), 2 * M * FF * D))
rows.append(("gemm ff2       resid f32 50432x768x3072", timeit(lambda: ops.gemm(x3072, w_2, bias=b_o, epilogue=ops.EPI_BIAS_RESID_F32, aux=resid)), 2 * M * FF * D))
rows.append(("gemm ff2-dgrad dgelu     50432x3072x768", timeit(lambda: ops.gemm(x768, w_2, b_mn_major=True, epilogue=ops.EPI_DGELU_BF16, aux=h)), 2 * M * FF * D))
rows.append(("gemm ff1-dgrad plain     50432x768x3072", timeit(lambda: ops.gemm(x3072, w_1, b_mn_major=True)), 2 * M * FF * D))
rows.append(("wgrad ff1      split-K   3072x768x50432", timeit(lambda: ops.wgrad(x3072, x768, gw)), 2 * M * FF * D))
qkv = bf(B, T, 3 * D)
out, lse = ops.attention_fwd(qkv, B, T, H)
dout = bf(B, T, D)
rows.append(("attention fwd  B256 T197 H12", timeit(lambda: ops.attention_fwd(qkv, B, T, H)), 4 * B * H * T * T * 64))
rows.append(("attention bwd  B256 T197 H12", timeit(lambda: ops.attention_bwd(qkv, out, dout, lse, B, T, H)), 10 * B * H * T * T * 64))
for name, us, fl in rows:
    print(f"{name:42s} {us:8.1f} us  {fl / us / 1e6:7.0f} TFLOP/s", flush=True)
