"""TEST INFRASTRUCTURE ONLY -- plain-PyTorch restatement of the reference's CLIP forward (BASELINE.json configs[3];
SURVEY.md 8a row a16): both towers (oracle/vit_oracle.py: ``encoder_forward`` with CLIP's vision options, ``tet_forward``)
plus the glue of ``CLIP.encode_image / encode_text`` (cflearn/modules/multimodal/clip.py:209-256) and
``IPerceptor.forward`` (multimodal/schema.py:25-30).  Never imported by the product package.

Pinned bit-for-bit (fp32 and bf16 autocast: logits, every parameter gradient for a seeded upstream gradient) against the
reference's own ``CLIP`` module by ``oracle/make_golden_clip.py``; fixture ``tests/golden/clip_tiny_reference.pt``.
The reference defines NO contrastive loss or CLIP training step: gradients are pinned through
``sum(logits * upstream)``; ``symmetric_cross_entropy`` below is this repo's definition for the future bench, loss parity
unpinned.  ``cftool.array.l2_normalize`` is not in the container; it is restated as ``t / t.norm(dim=-1, keepdim=True)``
 -- the same stub the loader gives the reference, so the pin covers the maths but not cftool itself.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

import vit_oracle as vo

StateDict = Dict[str, Tensor]


def clip_config(name: str = "clip") -> Dict[str, Dict]:
    table = {
        # CLIP() defaults (clip.py:31-62): ViT-B/32 vision tower, 512-wide 12-layer text tower, 49,408 tokens, context 77
        "clip": dict(vision=vo.vit_config("clip_vision_b32"), text=vo.tet_config("clip_text"), vocab_size=49408, latent_dim=512),
        "clip_tiny": dict(vision=vo.vit_config("clip_vision_tiny"), text=vo.tet_config("clip_text_tiny"), vocab_size=100, latent_dim=64),
    }
    return table[name]


def state_dict_spec(cfg: Dict) -> List[Tuple[str, Tuple[int, ...]]]:
    """Parameter keys in the reference's state_dict order (the bool buffer ``text_transformer.attention_mask`` sits between
    ``token_embedding.weight`` and the text tower's parameters; it is not a parameter)."""
    dt, lat = cfg["text"]["latent_dim"], cfg["latent_dim"]
    spec: List[Tuple[str, Tuple[int, ...]]] = [("logit_scale", ())]
    spec += [("vit." + k, s) for k, s in vo.state_dict_spec(cfg["vision"])]
    spec += [("token_embedding.weight", (cfg["vocab_size"], dt))]
    spec += [("text_transformer." + k, s) for k, s in vo.tet_state_dict_spec(cfg["text"])]
    spec += [("text_projection.weight", (lat, dt)), ("text_projection.bias", (lat,))]
    return spec


def init_state_dict(cfg: Dict, seed: int = 0) -> StateDict:
    g = torch.Generator().manual_seed(seed + 17)
    sd: StateDict = {"logit_scale": torch.tensor(2.6592600)}  # log(1 / 0.07), schema.py:15
    sd.update({"vit." + k: v for k, v in vo.init_state_dict(cfg["vision"], seed=seed).items()})
    sd["token_embedding.weight"] = 0.02 * torch.randn(cfg["vocab_size"], cfg["text"]["latent_dim"], generator=g)  # clip.py:192
    sd.update({"text_transformer." + k: v for k, v in vo.tet_init_state_dict(cfg["text"], seed=seed).items()})
    dt = cfg["text"]["latent_dim"]
    sd["text_projection.weight"] = torch.randn(cfg["latent_dim"], dt, generator=g) * dt ** -0.5  # clip.py:205
    sd["text_projection.bias"] = 0.02 * torch.randn(cfg["latent_dim"], generator=g)
    return sd


def synthetic_batch(cfg: Dict, batch: int, seed: int = 0) -> Tuple[Tensor, Tensor]:
    """images ~ N(0,1); ids uniform in [1, V-2] with the EOS id V-1 (the arg-max token) forced at a
    random position >= 1, so the ``indices.argmax(-1)`` pooling of clip.py:250 is well defined."""
    g = torch.Generator().manual_seed(seed)
    v = cfg["vision"]
    t = cfg["text"]["context_length"]
    x = torch.randn(batch, v["in_channels"], v["img_size"], v["img_size"], generator=g)
    ids = torch.randint(1, cfg["vocab_size"] - 1, (batch, t), generator=g)
    pos = torch.randint(1, t, (batch,), generator=g)
    ids[torch.arange(batch), pos] = cfg["vocab_size"] - 1
    return x, ids


def l2_normalize(t: Tensor) -> Tensor:
    return t / t.norm(dim=-1, keepdim=True)


def _sub(sd: StateDict, prefix: str) -> StateDict:
    return {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)}


def encode_image(sd: StateDict, image: Tensor, cfg: Dict) -> Tensor:
    """clip.py:209-216: ``l2_normalize(self.vit(image))``."""
    return l2_normalize(vo.encoder_forward(_sub(sd, "vit."), image, cfg["vision"]))


def encode_text(sd: StateDict, indices: Tensor, cfg: Dict) -> Tensor:
    """clip.py:218-256: embedding lookup (integer gather; padding_idx=0 only zeroes that row's gradient), text tower,
    pooling at the arg-max token id (:248-250; integer ops), Dropout(0), text_projection (a bf16 Linear under autocast),
    l2_normalize."""
    net = F.embedding(indices, sd["token_embedding.weight"], padding_idx=0)
    net = vo.tet_forward(_sub(sd, "text_transformer."), net, cfg["text"])
    net = net[torch.arange(net.shape[0], device=net.device), indices.argmax(dim=-1)]
    net = F.linear(net, sd["text_projection.weight"], sd["text_projection.bias"])
    return l2_normalize(net)


def forward(sd: StateDict, image: Tensor, indices: Tensor, cfg: Dict) -> Tensor:
    """IPerceptor.forward (schema.py:25-30): ``logit_scale.exp() * image_features @ text_features.t()``."""
    return sd["logit_scale"].exp() * encode_image(sd, image, cfg) @ encode_text(sd, indices, cfg).t()


def symmetric_cross_entropy(logits_per_image: Tensor) -> Tensor:
    """NOT in the reference (no contrastive loss exists there): the usual CLIP objective, defined here for a future step."""
    target = torch.arange(logits_per_image.shape[0], device=logits_per_image.device)
    lg = logits_per_image.float()
    return 0.5 * (F.cross_entropy(lg, target) + F.cross_entropy(lg.t(), target))


def train_step(sd: StateDict, image: Tensor, indices: Tensor, upstream: Tensor, cfg: Dict, *, autocast_bf16: bool):
    """Forward + backward with ``upstream`` [B, B] as the gradient of the logits; returns (logits, grads)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    with torch.autocast(image.device.type, dtype=torch.bfloat16, enabled=autocast_bf16):
        logits = forward(params, image, indices, cfg)
    (logits.float() * upstream).sum().backward()
    return logits.detach(), {k: v.grad for k, v in params.items()}
