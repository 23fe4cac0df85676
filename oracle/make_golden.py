"""TEST INFRASTRUCTURE ONLY -- pins ``oracle/vit_oracle.py`` against the real reference and writes tests/golden/*.

Run in the build container (needs /root/reference):  python oracle/make_golden.py
  1. builds the reference's own ``ViTEncoder`` + ``Linear`` head + ``CrossEntropyLoss`` (imported unmodified through
     oracle/load_reference.py), loads the oracle's synthetic weights into them with ``load_state_dict`` (so the
     state_dict keys/shapes are checked too), and asserts the oracle's forward taps, loss and every parameter
     gradient are BIT-IDENTICAL to the reference's, in fp32 and under bf16 autocast;
  2. re-runs the reference's own known-answer tests for this path against the oracle's ops
     (tests/test_blocks.py:147-176 Attention == nn.MultiheadAttention, atol 1e-4);
  3. stores small golden fixtures (inputs, weights seed, reference outputs) for the CPU test-suite.
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import vit_oracle as vo  # noqa: E402
from load_reference import load_reference_modules  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def reference_step(cfg, sd, x, labels, autocast_bf16):
    load_reference_modules()
    from cflearn.modules.cv.encoder.transformer import ViTEncoder
    from cflearn.modules.core.customs import Linear
    from cflearn.losses.basic import CrossEntropyLoss

    enc = ViTEncoder(img_size=cfg["img_size"], patch_size=cfg["patch_size"], in_channels=cfg["in_channels"],
                     latent_dim=cfg["latent_dim"], num_layers=cfg["num_layers"])
    head = Linear(cfg["latent_dim"], cfg["num_classes"])
    enc_sd = {k: v for k, v in sd.items() if not k.startswith("head.linear")}
    missing = enc.load_state_dict(enc_sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    head.load_state_dict({"linear.weight": sd["head.linear.weight"], "linear.bias": sd["head.linear.bias"]}, strict=True)
    loss_fn = CrossEntropyLoss()
    enc.train()
    head.train()
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast_bf16):
        encoded = enc(x)
        logits = head(encoded)
        loss = loss_fn._reduce(loss_fn(logits, labels))
    loss.backward()
    grads = {k: p.grad for k, p in enc.named_parameters()}
    grads["head.linear.weight"] = head.linear.weight.grad
    grads["head.linear.bias"] = head.linear.bias.grad
    return encoded.detach(), logits.detach(), loss.detach(), grads


def pin(name, batch, autocast_bf16):
    cfg = vo.vit_config(name)
    sd = vo.init_state_dict(cfg, seed=0)
    x, y = vo.synthetic_batch(cfg, batch, seed=1)
    r_enc, r_logits, r_loss, r_grads = reference_step(cfg, sd, x, y, autocast_bf16)
    o_loss, o_grads, taps = vo.train_step(sd, x, y, cfg, autocast_bf16=autocast_bf16, want_taps=True)
    assert torch.equal(taps["encoded"], r_enc), "encoder output differs from the reference"
    assert torch.equal(taps["logits"], r_logits), "logits differ from the reference"
    assert torch.equal(o_loss, r_loss), "loss differs from the reference"
    assert set(o_grads) == set(r_grads), (set(o_grads) ^ set(r_grads))
    for k in r_grads:
        assert torch.equal(o_grads[k], r_grads[k]), f"grad {k} differs from the reference"
    mode = "bf16" if autocast_bf16 else "fp32"
    print(f"pinned {name} B={batch} {mode}: loss {r_loss.item():.6f}; {len(r_grads)} grads bit-identical to the reference")
    return cfg, sd, x, y, r_enc, r_logits, r_loss, r_grads


def reference_clip_vision(cfg):
    """The ViTEncoder that CLIP._init_vision builds (multimodal/clip.py:121-135), from the unmodified reference."""
    load_reference_modules()
    from cflearn.modules.cv.encoder.transformer import ViTEncoder

    d = cfg["latent_dim"]
    return ViTEncoder(img_size=cfg["img_size"], patch_size=cfg["patch_size"], in_channels=cfg["in_channels"], latent_dim=d,
                      to_patches_config={"bias": False}, num_layers=cfg["num_layers"], norm_kwargs={"eps": cfg["eps"]},
                      embedding_norm=torch.nn.LayerNorm(d, cfg["eps"]), attention_kwargs={"num_heads": d // 64},
                      feedforward_kwargs={"activation": "quick_gelu"}, norm_after_head=True, output_dim=cfg["output_dim"])


def pin_clip_vision(name, batch, autocast_bf16):
    """Encoder-only pin (the reference defines no loss for CLIP): output and every gradient for a seeded
    upstream gradient must be bit-identical between the oracle and the reference module."""
    cfg = vo.vit_config(name)
    enc = reference_clip_vision(cfg)
    assert [(k, tuple(v.shape)) for k, v in enc.state_dict().items()] == vo.state_dict_spec(cfg), "state_dict keys / order / shapes"
    sd = vo.init_state_dict(cfg, seed=0)
    enc.load_state_dict(sd, strict=True)
    enc.train()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(batch, cfg["in_channels"], cfg["img_size"], cfg["img_size"], generator=g)
    up = torch.randn(batch, cfg["output_dim"], generator=g)
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast_bf16):
        out = enc(x)
    (out.float() * up).sum().backward()
    r_grads = {k: p.grad for k, p in enc.named_parameters()}
    o_out, o_grads, _ = vo.encoder_train_step(sd, x, up, cfg, autocast_bf16=autocast_bf16)
    assert torch.equal(o_out, out.detach()), "encoder output differs from the reference"
    for k in r_grads:
        assert torch.equal(o_grads[k], r_grads[k]), f"grad {k} differs from the reference"
    print(f"pinned {name} B={batch} {'bf16' if autocast_bf16 else 'fp32'}: {len(r_grads)} grads bit-identical to the reference")
    return cfg, x, up, out.detach(), r_grads


def pin_tet(name, batch, autocast_bf16):
    """Text tower stack: the oracle vs the reference's TeTEncoder as CLIP._init_text configures it (clip.py:175-188)."""
    load_reference_modules()
    from cflearn.modules.nlp.encoder.transformer import TeTEncoder

    cfg = vo.tet_config(name)
    d, t = cfg["latent_dim"], cfg["context_length"]
    enc = TeTEncoder(d, t, use_triu_attn_mask=True, num_layers=cfg["num_layers"], norm_kwargs={"eps": cfg["eps"]},
                     attention_kwargs={"num_heads": d // 64}, feedforward_kwargs={"activation": "quick_gelu"}, head_pooler=None)
    keys = [(k, tuple(v.shape)) for k, v in enc.state_dict().items()]
    assert keys[0] == ("attention_mask", (t, t)) and keys[1:] == vo.tet_state_dict_spec(cfg), "state_dict keys / order / shapes"
    sd = vo.tet_init_state_dict(cfg, seed=0)
    enc.load_state_dict(sd, strict=False)
    enc.train()
    g = torch.Generator().manual_seed(7)
    x = torch.randn(batch, t, d, generator=g) * 0.5
    up = torch.randn(batch, t, d, generator=g)
    xin = x.clone().requires_grad_(True)
    with torch.autocast("cpu", dtype=torch.bfloat16, enabled=autocast_bf16):
        out = enc(xin)
    (out.float() * up).sum().backward()
    r_grads = {k: p.grad for k, p in enc.named_parameters()}
    o_out, o_dx, o_grads = vo.tet_train_step(sd, x, up, cfg, autocast_bf16=autocast_bf16)
    assert torch.equal(o_out, out.detach()) and torch.equal(o_dx, xin.grad), "output / input gradient differ from the reference"
    for k in r_grads:
        assert torch.equal(o_grads[k], r_grads[k]), f"grad {k} differs from the reference"
    print(f"pinned {name} B={batch} {'bf16' if autocast_bf16 else 'fp32'}: {len(r_grads)} grads + dx bit-identical to the reference")
    return cfg, x, up, out.detach(), xin.grad.detach(), r_grads


def known_answer_attention():
    """tests/test_blocks.py:147-176 logic, pointed at the oracle's attention(): == nn.MultiheadAttention, atol 1e-4."""
    torch.manual_seed(0)
    d, heads, b, t = 256, 4, 3, 11
    mha = torch.nn.MultiheadAttention(d, heads, batch_first=True)
    sd = {"p.in_w": mha.in_proj_weight.detach(), "p.qkv_bias": mha.in_proj_bias.detach(),
          "p.out_linear.linear.weight": mha.out_proj.weight.detach(), "p.out_linear.linear.bias": mha.out_proj.bias.detach()}
    x = torch.randn(b, t, d)
    ours = vo.attention(sd, "p.", x, heads)
    ref = mha(x, x, x, need_weights=False)[0]
    err = (ours - ref).abs().max().item()
    assert err < 1e-4, err
    print(f"known-answer: oracle attention == nn.MultiheadAttention, max |diff| = {err:.2e}")


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    known_answer_attention()
    for mode in (False, True):
        pin("vit_small", 3, mode)
    # ViT-B/16 itself, fp32, B=2 (the big one takes ~10 s on CPU)
    pin("vit_b16", 2, False)
    pin("vit_b16", 2, True)
    out = {}
    for mode in (False, True):
        cfg, sd, x, y, r_enc, r_logits, r_loss, r_grads = pin("vit_tiny", 4, mode)
        tag = "bf16" if mode else "fp32"
        out[tag] = {"encoded": r_enc, "logits": r_logits, "loss": r_loss,
                    "grads": {k: v.clone() for k, v in r_grads.items()}}
    cfg = vo.vit_config("vit_tiny")
    x, y = vo.synthetic_batch(cfg, 4, seed=1)
    fixture = {"config_name": "vit_tiny", "batch": 4, "weights_seed": 0, "data_seed": 1, "x": x, "labels": y,
               "reference": out,
               "torch_version": torch.__version__,
               "generator": "oracle/make_golden.py (reference @ ca5ced1 imported through oracle/load_reference.py)"}
    torch.save(fixture, os.path.join(GOLDEN, "vit_tiny_reference.pt"))
    # ViT-B/16 key/shape list of the real reference module: the checkpoint-compatibility contract
    load_reference_modules()
    from cflearn.modules.cv.encoder.transformer import ViTEncoder

    big = ViTEncoder(img_size=224, patch_size=16, in_channels=3, latent_dim=768)
    keys = {k: list(v.shape) for k, v in big.state_dict().items()}
    import json

    with open(os.path.join(GOLDEN, "vit_b16_state_dict_keys.json"), "w") as f:
        json.dump({"num_params": sum(p.numel() for p in big.parameters()), "keys": keys}, f, indent=1)
    # CLIP's vision tower: real ViT-B/32 shapes pinned once, a tiny one stored as a fixture
    pin_clip_vision("clip_vision_b32", 2, True)
    ref = {}
    for mode in (False, True):
        cfg, x, up, r_out, r_grads = pin_clip_vision("clip_vision_tiny", 3, mode)
        ref["bf16" if mode else "fp32"] = {"out": r_out, "grads": {k: v.clone() for k, v in r_grads.items()}}
    torch.save({"config_name": "clip_vision_tiny", "weights_seed": 0, "x": x, "upstream": up, "reference": ref},
               os.path.join(GOLDEN, "clip_vision_tiny_reference.pt"))
    tiny = reference_clip_vision(vo.vit_config("clip_vision_tiny"))
    with open(os.path.join(GOLDEN, "clip_vision_tiny_keys.json"), "w") as f:
        json.dump({"keys": [[k, list(v.shape)] for k, v in tiny.state_dict().items()]}, f, indent=1)
    # CLIP's text tower stack (TeTEncoder): real shape pinned once, a tiny one stored
    pin_tet("clip_text", 2, True)
    ref = {}
    for mode in (False, True):
        cfg, x, up, r_out, r_dx, r_grads = pin_tet("clip_text_tiny", 3, mode)
        ref["bf16" if mode else "fp32"] = {"out": r_out, "dx": r_dx, "grads": {k: v.clone() for k, v in r_grads.items()}}
    torch.save({"config_name": "clip_text_tiny", "weights_seed": 0, "x": x, "upstream": up, "reference": ref},
               os.path.join(GOLDEN, "clip_text_tiny_reference.pt"))
    print("wrote", os.listdir(GOLDEN))


if __name__ == "__main__":
    main()
