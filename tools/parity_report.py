"""Prints relative-L2 parity numbers of the B200 training step against the oracle (eager bf16 autocast on the same
GPU, and fp32) for a few configurations.  Run on the GPU box; output is pasted into profiles/parity_rNN.md."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import vit_oracle as vo  # noqa: E402

import cflearn_b200  # noqa: F401,E402
from cflearn_b200 import registry, vit  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def report(name, batch):
    cfg = vo.vit_config(name)
    sd = vo.init_state_dict(cfg, seed=0)
    x, y = vo.synthetic_batch(cfg, batch, seed=1)
    x, y = x.to(DEV), y.to(DEV)
    m = registry.build_module("cv_clf", config=dict(in_channels=3, num_classes=cfg["num_classes"], img_size=cfg["img_size"],
                                                    latent_dim=cfg["latent_dim"], encoder="vit",
                                                    encoder_config=dict(patch_size=cfg["patch_size"], num_layers=cfg["num_layers"])))
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV)
    logits = m(x)[vit.PREDICTIONS_KEY]
    loss = vit.cross_entropy(logits, y)
    loss.backward()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().clone() for k, p in m.named_arena_parameters()}
    sdg = {k: v.to(DEV) for k, v in sd.items()}
    e_loss, e_grads, e_taps = vo.train_step(sdg, x, y, cfg, autocast_bf16=True, want_taps=True)
    f_loss, f_grads, f_taps = vo.train_step(sdg, x, y, cfg, autocast_bf16=False, want_taps=True)
    print(f"## {name} B={batch}: loss ours {loss.item():.6f} eager-bf16 {e_loss.item():.6f} fp32 {f_loss.item():.6f}")
    # the reference's own eager bf16 path under a second, equally legitimate SDPA backend: its self-consistency floor
    try:
        from torch.nn.attention import SDPBackend, sdpa_kernel

        with sdpa_kernel(SDPBackend.MATH):
            m_loss, m_grads, m_taps = vo.train_step(sdg, x, y, cfg, autocast_bf16=True, want_taps=True)
        import statistics as _st

        print(f"eager-bf16(default SDPA) vs eager-bf16(math SDPA): logits {rel(e_taps['logits'], m_taps['logits']):.3e}, "
              f"median grad {_st.median(rel(e_grads[k], m_grads[k]) for k in e_grads):.3e}, "
              f"worst grad {max(rel(e_grads[k], m_grads[k]) for k in e_grads):.3e}")
    except Exception as err:  # pragma: no cover
        print("eager-vs-eager skipped:", err)
    print(f"logits: ours-vs-eager {rel(logits, e_taps['logits']):.3e}  ours-vs-fp32 {rel(logits, f_taps['logits']):.3e}  eager-vs-fp32 {rel(e_taps['logits'], f_taps['logits']):.3e}")
    rows = []
    for k in grads:
        rows.append((rel(grads[k], e_grads[k]), rel(grads[k], f_grads[k]), rel(e_grads[k], f_grads[k]), k))
    rows.sort(reverse=True)
    print("grad: ours-vs-eager  ours-vs-fp32  eager-vs-fp32  key   (worst 12 by ours-vs-eager)")
    for r in rows[:12]:
        print(f"  {r[0]:.3e}  {r[1]:.3e}  {r[2]:.3e}  {r[3]}")
    import statistics

    print(f"  median ours-vs-eager {statistics.median(r[0] for r in rows):.3e}; median ours-vs-fp32 {statistics.median(r[1] for r in rows):.3e}; "
          f"median eager-vs-fp32 {statistics.median(r[2] for r in rows):.3e}; worst ratio ours/eager vs fp32 {max(r[1] / max(r[2], 1e-12) for r in rows):.2f}")


if __name__ == "__main__":
    for name, b in (("vit_tiny", 4), ("vit_small", 6), ("vit_b16", 8), ("vit_b16", 64)):
        report(name, b)
