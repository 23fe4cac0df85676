"""Is the attention HBM-access-pattern bound?  Same FLOPs and bytes, different DRAM locality:
  (a) qkv [B=256, T, 3*12*64]  -- the model's layout: every 128-byte head slice of a token row is fetched by a different SM
  (b) qkv [B=3072, T, 3*1*64]  -- one head per "image": the q|k|v slices of a token are adjacent (384-byte rows)
plus two torch copies that bracket what the memory system gives for 128-byte pieces at a 4608-byte stride vs contiguous."""
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
    return e0.elapsed_time(e1) / reps * 1e3


B, T, H = 256, 197, 12
qa = (torch.randn(B, T, 3 * H * 64, device=dev)).to(torch.bfloat16)
qb = (torch.randn(B * H, T, 3 * 64, device=dev)).to(torch.bfloat16)
for ver, pf in ((2, 1), (2, 0), (1, 0)):
    _cabi.lib().b200_set_attention_fwd_version(ver)
    _cabi.lib().b200_set_attention_bwd_version(ver)
    _cabi.lib().b200_set_attention_prefetch(pf)
    ta = timeit(lambda: ops.attention_fwd(qa, B, T, H))
    tb = timeit(lambda: ops.attention_fwd(qb, B * H, T, 1))
    oa, la = ops.attention_fwd(qa, B, T, H)
    ob, lb = ops.attention_fwd(qb, B * H, T, 1)
    da, db = torch.randn_like(oa), torch.randn_like(ob)
    tba = timeit(lambda: ops.attention_bwd(qa, oa, da, la, B, T, H))
    tbb = timeit(lambda: ops.attention_bwd(qb, ob, db, lb, B * H, T, 1))
    print(f"attention v{ver} prefetch={pf}: fwd model layout {ta:.1f} us | one head per image {tb:.1f} us || bwd {tba:.1f} us | {tbb:.1f} us")
x = qa.view(B, T, 36, 64)
t_strided = timeit(lambda: x.permute(0, 2, 1, 3).contiguous())
t_contig = timeit(lambda: qa.clone())
nbytes = qa.numel() * 2
print(f"copy 232 MB: 128-byte pieces at 4608-byte stride -> contiguous {t_strided:.1f} us ({2 * nbytes / t_strided / 1e6:.2f} TB/s r+w); "
      f"contiguous clone {t_contig:.1f} us ({2 * nbytes / t_contig / 1e6:.2f} TB/s)")
