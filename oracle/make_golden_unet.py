"""TEST INFRASTRUCTURE ONLY -- pins ``oracle/unet_oracle.py`` against the real reference ``UNetDiffuser``
(cflearn/modules/multimodal/diffusion/unet.py, imported unmodified through oracle/load_reference.py) and writes
tests/golden/unet_tiny_reference.pt.  Run in the build container:  python oracle/make_golden_unet.py [--full]"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import unet_oracle as uo  # noqa: E402
from load_reference import load_reference_modules  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def reference_unet(cfg):
    load_reference_modules()
    from cflearn.modules.multimodal.diffusion.unet import UNetDiffuser

    return UNetDiffuser(cfg["in_channels"], cfg["out_channels"], num_heads=cfg["num_heads"], use_spatial_transformer=True,
                        context_dim=cfg["context_dim"], start_channels=cfg["start_channels"], num_res_blocks=cfg["num_res_blocks"],
                        attention_downsample_rates=cfg["attention_downsample_rates"], channel_multipliers=cfg["channel_multipliers"])


def pin(name, batch, size, ctx_len, autocast_bf16):
    cfg = uo.unet_config(name)
    m = reference_unet(cfg)
    shapes = [(k, tuple(v.shape)) for k, v in m.state_dict().items()]
    sd = uo.synthetic_state_dict(shapes, seed=0)
    m.load_state_dict(sd, strict=True)
    m.train()
    g = torch.Generator().manual_seed(11)
    x = torch.randn(batch, cfg["in_channels"], size, size, generator=g)
    ts = torch.randint(0, 1000, (batch,), generator=g)
    ctx = torch.randn(batch, ctx_len, cfg["context_dim"], generator=g)
    up = torch.randn(batch, cfg["out_channels"], size, size, generator=g)
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast_bf16):
        out = m(x, timesteps=ts, context=ctx)
    (out.float() * up).sum().backward()
    r_grads = {k: p.grad for k, p in m.named_parameters()}
    o_out, o_grads = uo.train_step(sd, x, ts, ctx, up, cfg, autocast_bf16=autocast_bf16)
    assert torch.equal(o_out, out.detach()), "output differs from the reference"
    for k, gr in r_grads.items():
        assert torch.equal(o_grads[k], gr), f"grad {k} differs from the reference"
    n = sum(p.numel() for p in m.parameters())
    print(f"pinned {name} ({n / 1e6:.1f} M parameters) B={batch} {size}x{size} {'bf16' if autocast_bf16 else 'fp32'}: "
          f"output + {len(r_grads)} grads bit-identical to the reference UNetDiffuser")
    return cfg, shapes, x, ts, ctx, up, out.detach(), r_grads


def main():
    if "--full" in sys.argv:  # the real SD-v1.5 layout (859.5 M parameters), reduced resolution: needs ~20 GB of RAM
        pin("sd_v1_5", 1, 16, 7, False)
    ref = {}
    for mode in (False, True):
        cfg, shapes, x, ts, ctx, up, out, grads = pin("unet_tiny", 2, 16, 5, mode)
        keep = [k for k in grads if k.startswith("time_embedding") or k.startswith("head.") or ".attn2.to_k" in k or k.endswith("conv2.bias")]
        ref["bf16" if mode else "fp32"] = {"out": out, "grads": {k: grads[k].clone() for k in keep}}
    os.makedirs(GOLDEN, exist_ok=True)
    torch.save({"config_name": "unet_tiny", "shapes": shapes, "weights_seed": 0, "x": x, "timesteps": ts, "context": ctx, "upstream": up,
                "reference": ref}, os.path.join(GOLDEN, "unet_tiny_reference.pt"))
    print("wrote tests/golden/unet_tiny_reference.pt")


if __name__ == "__main__":
    main()
