"""The reference path in PyTorch eager on the SAME GPU ('the real bar'): the oracle restatement
(bit-identical to the reference's modules) under torch.autocast(bf16), fwd + CE + bwd + fused torch Adam, B = 256.
Also times F.scaled_dot_product_attention fwd / fwd+bwd at the ViT-B/16 attention shape for kernel-level context."""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import vit_oracle as vo  # noqa: E402

dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cfg = vo.vit_config("vit_b16")
sd = vo.init_state_dict(cfg, seed=0, perturb=False)
params = {k: v.to(dev).requires_grad_(True) for k, v in sd.items()}
opt = torch.optim.Adam(list(params.values()), lr=1e-3, fused=True)
x = torch.randn(B, 3, 224, 224, device=dev)
y = torch.randint(0, 1000, (B, 1), device=dev)


def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = vo.cross_entropy(vo.classifier_forward(params, x, cfg), y)
    loss.backward()
    opt.step()
    return loss


def timeit(fn, warm=3, reps=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timeit(step)
print(f"eager autocast-bf16 reference path on GPU: B={B}  {ms:.2f} ms/step  {B / ms * 1e3:.0f} samples/s  (torch {torch.__version__})")
q = torch.randn(B, 12, 197, 64, device=dev, dtype=torch.bfloat16, requires_grad=True)
k = torch.randn_like(q, requires_grad=True)
v = torch.randn_like(q, requires_grad=True)
do = torch.randn(B, 12, 197, 64, device=dev, dtype=torch.bfloat16)
f_ms = timeit(lambda: F.scaled_dot_product_attention(q, k, v))


def fb():
    o = F.scaled_dot_product_attention(q, k, v)
    o.backward(do)


fb_ms = timeit(fb)
print(f"torch SDPA [B={B},12,197,64] bf16: fwd {f_ms * 1e3:.0f} us, fwd+bwd {fb_ms * 1e3:.0f} us (bwd ~{(fb_ms - f_ms) * 1e3:.0f} us)")
