"""Lock-step diagnosis of tests/test_model_gpu.py::test_trainer_seam_matches_reference_update_sequence.

Three optimisers walk over the same batches from the same initial weights:
  A  B200TrainStep (device-side clip coefficient + fused arena Adam)
  T  torch.nn.utils.clip_grad_norm_ + torch.optim.Adam on the gradients the B200 step exposes through ``p.grad``
  M  torch's clip on the arena's flat gradient, then the fused arena Adam with grad_scale = 1  (separates clip from Adam)
and after every step the script prints the gradient norm each one saw, its clip coefficient, and how far the parameters
have drifted apart relative to one lr-sized step.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import vit_oracle as vo  # noqa: E402

import cflearn_b200  # noqa: F401,E402
from cflearn_b200 import registry  # noqa: E402
from cflearn_b200.optim import ArenaAdam  # noqa: E402
from cflearn_b200.trainer import B200TrainStep  # noqa: E402

DEV = "cuda"


def build(cfg, sd):
    m = registry.build_module("cv_clf", config=dict(in_channels=cfg["in_channels"], num_classes=cfg["num_classes"], img_size=cfg["img_size"],
                                                    latent_dim=cfg["latent_dim"], encoder="vit",
                                                    encoder_config=dict(patch_size=cfg["patch_size"], num_layers=cfg["num_layers"])))
    m.load_state_dict(sd, strict=True)
    return m.to(DEV)


def main():
    cfg = vo.vit_config("vit_tiny")
    sd = vo.init_state_dict(cfg, seed=0)
    batches = [vo.synthetic_batch(cfg, 4, seed=20 + i) for i in range(6)]
    lam = lambda k: 1.0 + 0.5 * min(k, 3) - 0.2 * max(k - 3, 0)  # noqa: E731
    clip = 0.05
    m_a = build(cfg, sd)
    o_a = ArenaAdam(m_a, lr=1e-3, capturable=True)
    s_a = B200TrainStep(m_a, o_a, scheduler=torch.optim.lr_scheduler.LambdaLR(o_a, lam), clip_norm=clip)
    m_t = build(cfg, sd)
    p_t = list(m_t.parameters())
    o_t = torch.optim.Adam(p_t, lr=1e-3)
    sc_t = torch.optim.lr_scheduler.LambdaLR(o_t, lam)
    m_m = build(cfg, sd)
    o_m = ArenaAdam(m_m, lr=1e-3, capturable=True)
    sc_m = torch.optim.lr_scheduler.LambdaLR(o_m, lam)
    for k, (x, y) in enumerate(batches):
        x, y = x.to(DEV), y.to(DEV)
        lr = o_t.param_groups[0]["lr"]
        # A
        la = s_a.step(x, y).item()
        coef_a = o_a.hyper_dev[5].item()
        norm_a = torch.linalg.vector_norm(m_a.arena.grad).item()
        # T
        for p in p_t:
            p.grad = None
        lt = m_t.train_step(x, y).item()
        aliased = sum(int(p.grad.data_ptr() == m_t.arena.g(kk).data_ptr()) for kk, p in m_t.named_arena_parameters())
        g_flat_norm = torch.linalg.vector_norm(m_t.arena.grad).item()
        norm_t = torch.nn.utils.clip_grad_norm_(p_t, clip).item()
        o_t.step()
        sc_t.step()
        # M
        o_m.zero_grad()
        lm = m_m.train_step(x, y).item()
        norm_m = torch.linalg.vector_norm(m_m.arena.grad).item()
        m_m.arena.grad.mul_(min(1.0, clip / (norm_m + 1e-6)))
        o_m.step()
        sc_m.step()
        torch.cuda.synchronize()
        d_at = max((p - q).abs().max().item() for p, q in zip(m_a.parameters(), m_t.parameters()))
        d_am = max((p - q).abs().max().item() for p, q in zip(m_a.parameters(), m_m.parameters()))
        d_tm = max((p - q).abs().max().item() for p, q in zip(m_t.parameters(), m_m.parameters()))
        print(f"step {k}: lr {lr:.2e} loss A {la:.6f} T {lt:.6f} M {lm:.6f} | grad norm A(after step) {norm_a:.6e} T(flat) {g_flat_norm:.6e} "
              f"T(clip_grad_norm_) {norm_t:.6e} M {norm_m:.6e} | coef A {coef_a:.6e} expected {min(1.0, clip / (norm_t + 1e-6)):.6e} | "
              f"p.grad aliases arena: {aliased}/{len(p_t)} | max |dparam| A-T {d_at:.3e} A-M {d_am:.3e} T-M {d_tm:.3e} (in lr steps: {d_at / lr:.4f})")


if __name__ == "__main__":
    main()
