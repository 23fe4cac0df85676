"""TEST INFRASTRUCTURE ONLY -- pins ``oracle/clip_oracle.py`` against the real reference ``CLIP`` module
(cflearn/modules/multimodal/clip.py, imported unmodified through oracle/load_reference.py) and writes
tests/golden/clip_tiny_reference.pt.  Run in the build container:  python oracle/make_golden_clip.py"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import clip_oracle as co  # noqa: E402
from load_reference import load_reference_modules  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def reference_clip(cfg):
    load_reference_modules()
    from cflearn.modules.multimodal.clip import CLIP

    v, t = cfg["vision"], cfg["text"]
    return CLIP(img_size=v["img_size"], latent_dim=cfg["latent_dim"], in_channels=v["in_channels"], vision_latent_dim=v["latent_dim"],
                vision_patch_size=v["patch_size"], vision_num_heads=v["latent_dim"] // 64, vision_num_layers=v["num_layers"],
                vocab_size=cfg["vocab_size"], context_length=t["context_length"], text_latent_dim=t["latent_dim"],
                text_num_heads=t["latent_dim"] // 64, text_num_layers=t["num_layers"])


def pin(name, batch, autocast_bf16):
    cfg = co.clip_config(name)
    m = reference_clip(cfg)
    ref_keys = [(k, tuple(v.shape)) for k, v in m.state_dict().items() if k != "text_transformer.attention_mask"]
    assert ref_keys == co.state_dict_spec(cfg), "state_dict keys / order / shapes"
    sd = co.init_state_dict(cfg, seed=0)
    m.load_state_dict(sd, strict=False)
    m.train()
    x, ids = co.synthetic_batch(cfg, batch, seed=3)
    up = torch.randn(batch, batch, generator=torch.Generator().manual_seed(9))
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast_bf16):
        logits = m(x, ids)
    (logits.float() * up).sum().backward()
    r_grads = {k: p.grad for k, p in m.named_parameters()}
    o_logits, o_grads = co.train_step(sd, x, ids, up, cfg, autocast_bf16=autocast_bf16)
    assert torch.equal(o_logits, logits.detach()), "logits differ from the reference"
    for k, g in r_grads.items():
        assert torch.equal(o_grads[k], g), f"grad {k} differs from the reference"
    print(f"pinned {name} B={batch} {'bf16' if autocast_bf16 else 'fp32'}: logits + {len(r_grads)} grads bit-identical to the reference CLIP")
    return cfg, x, ids, up, logits.detach(), r_grads


def main():
    pin("clip", 2, True)  # the real ViT-B/32 + 512x77 text tower, once
    ref = {}
    for mode in (False, True):
        cfg, x, ids, up, logits, grads = pin("clip_tiny", 4, mode)
        keep = ("logit_scale", "token_embedding.weight", "text_projection.weight", "text_projection.bias", "vit.output_projection",
                "vit.encoder.embedding_norm.weight", "text_transformer.encoder.pos_encoding.pos_encoding")
        ref["bf16" if mode else "fp32"] = {"logits": logits, "grads": {k: grads[k].clone() for k in keep}}
    os.makedirs(GOLDEN, exist_ok=True)
    torch.save({"config_name": "clip_tiny", "weights_seed": 0, "x": x, "ids": ids, "upstream": up, "reference": ref},
               os.path.join(GOLDEN, "clip_tiny_reference.pt"))
    print("wrote tests/golden/clip_tiny_reference.pt")


if __name__ == "__main__":
    main()
