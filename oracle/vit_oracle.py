"""TEST INFRASTRUCTURE ONLY -- a plain-PyTorch restatement of the reference's ViT classifier training step.

This is the ORACLE for the B200 kernels: it restates, op by op and in the reference's order, what
``cflearn.modules`` computes on the hot path named by BASELINE.json.  It is never imported by the product package
(``carefree-learn_b200/``); only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may use it.

Pinning: ``oracle/make_golden.py`` (run in the build container, where /root/reference exists) checks this file
bit-for-bit against the reference's own ``ViTEncoder`` / ``Linear`` / ``CrossEntropyLoss`` code imported through
``oracle/load_reference.py`` and writes the golden fixtures in ``tests/golden/`` that ``tests/test_oracle.py``
re-checks on every run.  The reference ships no golden vectors of its own for this path.

Every function cites the reference lines it follows (paths relative to /root/reference/cflearn/).
Run it under ``torch.autocast(device, dtype=torch.bfloat16)`` to get the reference's mixed-precision path
(accelerate's ``mixed_precision="bf16"`` wraps the same ops, schema.py:1260-1276); without autocast it is the fp32 path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

StateDict = Dict[str, Tensor]


def vit_config(name: str = "vit_b16") -> Dict[str, int]:
    """Shapes of the named configurations (BASELINE.json configs[1] is ``vit_b16``)."""
    table = {
        # modules/cv/encoder/transformer.py:19-43 defaults with latent_dim=768 -> 12 heads (:61), FF ratio 4 (:34)
        "vit_b16": dict(img_size=224, patch_size=16, in_channels=3, latent_dim=768, num_layers=12, num_classes=1000),
        "vit_tiny": dict(img_size=32, patch_size=16, in_channels=3, latent_dim=128, num_layers=2, num_classes=10),
        "vit_small": dict(img_size=64, patch_size=16, in_channels=3, latent_dim=256, num_layers=3, num_classes=24),
        # the ViTEncoder CLIP._init_vision builds (multimodal/clip.py:121-135; SURVEY.md 8a row a16, vision half): no conv bias,
        # embedding_norm, QuickGELU, LayerNorm eps 1e-5, head_norm after the cls pick, output_projection [D, latent_dim]
        "clip_vision_b32": dict(img_size=224, patch_size=32, in_channels=3, latent_dim=768, num_layers=12, conv_bias=False,
                                emb_norm=True, norm_after_head=True, output_dim=512, activation="quick_gelu", eps=1e-5),
        "clip_vision_tiny": dict(img_size=64, patch_size=32, in_channels=3, latent_dim=128, num_layers=2, conv_bias=False,
                                 emb_norm=True, norm_after_head=True, output_dim=64, activation="quick_gelu", eps=1e-5),
        "clip_vision_small": dict(img_size=96, patch_size=32, in_channels=3, latent_dim=256, num_layers=3, conv_bias=False,
                                  emb_norm=True, norm_after_head=True, output_dim=128, activation="quick_gelu", eps=1e-5),
    }
    return dict(table[name])


def state_dict_spec(cfg: Dict[str, int]) -> List[Tuple[str, Tuple[int, ...]]]:
    """(key, shape) of ``cv_clf``-style parameters: ViTEncoder keys + ``head.linear.*``."""
    d, p, c = cfg["latent_dim"], cfg["patch_size"], cfg["in_channels"]
    n_tok = (cfg["img_size"] // p) ** 2 + 1
    ff = 4 * d
    spec: List[Tuple[str, Tuple[int, ...]]] = []
    if cfg.get("output_dim") is not None:  # ViTEncoder's own parameter: first in state_dict order (transformer.py:78-82)
        spec.append(("output_projection", (d, cfg["output_dim"])))
    spec.append(("to_patches.projection.weight", (d, c, p, p)))
    if cfg.get("conv_bias", True):
        spec.append(("to_patches.projection.bias", (d,)))
    spec += [("encoder.head_token", (1, 1, d)), ("encoder.pos_encoding.pos_encoding", (1, n_tok, d))]
    if cfg.get("emb_norm", False):
        spec += [("encoder.embedding_norm.weight", (d,)), ("encoder.embedding_norm.bias", (d,))]
    for i in range(cfg["num_layers"]):
        b = f"encoder.mixing_blocks.{i}."
        spec += [
            (b + "token_norm.weight", (d,)), (b + "token_norm.bias", (d,)),
            (b + "token_mixing.net.in_w", (3 * d, d)), (b + "token_mixing.net.qkv_bias", (3 * d,)),
            (b + "token_mixing.net.out_linear.linear.weight", (d, d)), (b + "token_mixing.net.out_linear.linear.bias", (d,)),
            (b + "channel_norm.weight", (d,)), (b + "channel_norm.bias", (d,)),
            (b + "channel_mixing.net.0.linear.weight", (ff, d)), (b + "channel_mixing.net.0.linear.bias", (ff,)),
            (b + "channel_mixing.net.3.linear.weight", (d, ff)), (b + "channel_mixing.net.3.linear.bias", (d,)),
        ]
    hn = "encoder.head_norm." if cfg.get("norm_after_head", False) else "encoder.head.norms.0."
    spec += [(hn + "weight", (d,)), (hn + "bias", (d,))]
    if cfg.get("num_classes") is not None:
        spec += [("head.linear.weight", (cfg["num_classes"], d)), ("head.linear.bias", (cfg["num_classes"],))]
    return spec


def init_state_dict(cfg: Dict[str, int], seed: int = 0, *, perturb: bool = True) -> StateDict:
    """Synthetic weights with the reference's initialisation statistics (mixed_stacks/api.py:405-417,205;
    convs/basic.py:94-97).  ``perturb`` moves biases / LayerNorm affine parameters off their 0 / 1 initial values so
    that every term of every kernel is exercised.  Parity is always checked with IDENTICAL injected weights."""
    g = torch.Generator().manual_seed(seed)
    sd: StateDict = {}
    for key, shape in state_dict_spec(cfg):
        if key.endswith("norm.weight") or key.endswith("norms.0.weight"):
            t = torch.ones(shape) + (0.1 * torch.randn(shape, generator=g) if perturb else 0)
        elif key.endswith("bias"):
            t = 0.02 * torch.randn(shape, generator=g) if perturb else torch.zeros(shape)
        elif key == "output_projection":
            t = (shape[0] ** -0.5) * torch.randn(shape, generator=g)  # transformer.py:81
        elif key == "to_patches.projection.weight":
            fan_in = shape[1] * shape[2] * shape[3]
            fan_out = shape[0] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * (2.0 / (fan_in + fan_out)) ** 0.5  # xavier_normal
        else:
            t = torch.nn.init.trunc_normal_(torch.empty(shape), std=0.02, generator=g)
        sd[key] = t.float()
    return sd


def synthetic_batch(cfg: Dict[str, int], batch: int, seed: int = 0) -> Tuple[Tensor, Tensor]:
    """images ~ N(0,1) fp32, labels uniform int64 [B, 1]."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(batch, cfg["in_channels"], cfg["img_size"], cfg["img_size"], generator=g)
    y = torch.randint(0, cfg["num_classes"], (batch, 1), generator=g)
    return x, y


# -----------------------------------------------------------------------------------------------------------------
# forward, op by op
# -----------------------------------------------------------------------------------------------------------------
def patch_embed(sd: StateDict, x: Tensor) -> Tensor:
    """VanillaPatchEmbed.forward (modules/core/high_level.py:181-188): Conv2d k = s = patch, pad 0
    (convs/basic.py:155-174), then _flatten (high_level.py:143-149): [B,C,h,w] -> [B, h*w, C] contiguous."""
    w = sd["to_patches.projection.weight"]
    p = w.shape[-1]
    net = F.conv2d(x, w, sd.get("to_patches.projection.bias"), stride=p, padding=0)
    b, c, h, ww = net.shape
    return net.view(b, c, h * ww).transpose(1, 2).contiguous()


def pre_process(sd: StateDict, patches: Tensor, emb_eps: float = 1e-5) -> Tensor:
    """MixedStackedEncoder.pre_process (mixed_stacks/api.py:419-438): cat head token, add learned positional
    encoding (PositionalEncoding.forward early-return path, api.py:209-228,244-245).  Dropout(0) is the identity."""
    n = patches.shape[0]
    head_tokens = sd["encoder.head_token"].repeat([n, 1, 1])
    net = torch.cat([head_tokens, patches], dim=1)
    net = net + sd["encoder.pos_encoding.pos_encoding"]
    if "encoder.embedding_norm.weight" in sd:  # api.py:433-434 (CLIP: nn.LayerNorm(D, 1e-5), clip.py:131)
        net = F.layer_norm(net, (net.shape[-1],), sd["encoder.embedding_norm.weight"], sd["encoder.embedding_norm.bias"], emb_eps)
    return net


def attention(sd: StateDict, prefix: str, net: Tensor, num_heads: int, causal: bool = False) -> Tensor:
    """Attention.forward, qkv_same branch (modules/core/attentions.py:213-277): packed projection, chunk(3),
    _to_heads (:180-185), sdp_attn -> F.scaled_dot_product_attention(q,k,v,mask,0.0) (toolkit.py:953-963),
    transpose+contiguous+view (:270-275), out_linear (:277; Linear.forward customs.py:85-89)."""
    b, t, d = net.shape
    qkv = F.linear(net, sd[prefix + "in_w"], sd[prefix + "qkv_bias"])
    q, k, v = qkv.chunk(3, dim=-1)
    q, k, v = (z.view(b, t, num_heads, d // num_heads).permute(0, 2, 1, 3).contiguous() for z in (q, k, v))
    mask = None
    if causal:  # nlp/encoder/transformer.py:42-48 builds the upper-triangular "zeroed" mask; attentions.py:251-253 inverts it
        mask = ~torch.ones(t, t, dtype=torch.bool, device=net.device).triu(1)
    out = F.scaled_dot_product_attention(q, k, v, mask, 0.0)
    out = out.transpose(1, 2).contiguous().view(-1, t, d)
    return F.linear(out, sd[prefix + "out_linear.linear.weight"], sd[prefix + "out_linear.linear.bias"])


def feed_forward(sd: StateDict, prefix: str, net: Tensor, activation: str = "GELU") -> Tensor:
    """FeedForward (mixed_stacks/channel_mixers.py:29-36): Linear -> nn.GELU() (exact erf) -> Dropout(0) -> Linear;
    activation "quick_gelu" = QuickGELU, ``net * torch.sigmoid(1.702 * net)`` (core/activations.py:151-153)."""
    h = F.linear(net, sd[prefix + "0.linear.weight"], sd[prefix + "0.linear.bias"])
    a = h * torch.sigmoid(1.702 * h) if activation == "quick_gelu" else F.gelu(h)
    return F.linear(a, sd[prefix + "3.linear.weight"], sd[prefix + "3.linear.bias"])


def mixing_block(sd: StateDict, i: int, net: Tensor, num_heads: int, eps: float, activation: str = "GELU") -> Tensor:
    """MixingBlock._pre_norm_forward (mixed_stacks/api.py:130-158); DropPath / Dropout are identities at rate 0."""
    b = f"encoder.mixing_blocks.{i}."
    d = net.shape[-1]
    t = F.layer_norm(net, (d,), sd[b + "token_norm.weight"], sd[b + "token_norm.bias"], eps)
    net = net + attention(sd, b + "token_mixing.net.", t, num_heads)
    c = F.layer_norm(net, (d,), sd[b + "channel_norm.weight"], sd[b + "channel_norm.bias"], eps)
    return net + feed_forward(sd, b + "channel_mixing.net.", c, activation)


def mixing_block_ops(sd: StateDict, i: int, net: Tensor, num_heads: int, eps: float) -> Dict[str, Tensor]:
    """The SAME op sequence as ``mixing_block`` (GELU FeedForward, no mask) with every op boundary returned, for the per-op
    teacher-forced parity tests (tests/_taps.py).  ``out`` is bit-identical to ``mixing_block(...)`` (checked in
    tests/test_oracle.py).  Lines cited there: mixed_stacks/api.py:130-158, attentions.py:213-277, channel_mixers.py:29-36."""
    b = f"encoder.mixing_blocks.{i}."
    d = net.shape[-1]
    bsz, t, _ = net.shape
    o: Dict[str, Tensor] = {"x": net}
    o["ln1"] = F.layer_norm(net, (d,), sd[b + "token_norm.weight"], sd[b + "token_norm.bias"], eps)
    p = b + "token_mixing.net."
    o["qkv"] = F.linear(o["ln1"], sd[p + "in_w"], sd[p + "qkv_bias"])
    q, k, v = o["qkv"].chunk(3, dim=-1)
    q, k, v = (z.view(bsz, t, num_heads, d // num_heads).permute(0, 2, 1, 3).contiguous() for z in (q, k, v))
    att = F.scaled_dot_product_attention(q, k, v, None, 0.0)
    o["attn"] = att.transpose(1, 2).contiguous().view(-1, t, d)
    o["proj"] = F.linear(o["attn"], sd[p + "out_linear.linear.weight"], sd[p + "out_linear.linear.bias"])
    o["mid"] = net + o["proj"]
    o["ln2"] = F.layer_norm(o["mid"], (d,), sd[b + "channel_norm.weight"], sd[b + "channel_norm.bias"], eps)
    c = b + "channel_mixing.net."
    o["h"] = F.linear(o["ln2"], sd[c + "0.linear.weight"], sd[c + "0.linear.bias"])
    o["act"] = F.gelu(o["h"])
    o["ff"] = F.linear(o["act"], sd[c + "3.linear.weight"], sd[c + "3.linear.bias"])
    o["out"] = o["mid"] + o["ff"]
    return o


def encoder_forward(sd: StateDict, x: Tensor, cfg: Dict[str, int], taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """ViTEncoder.forward (modules/cv/encoder/transformer.py:88-100) -> [B, latent_dim]."""
    d = cfg["latent_dim"]
    heads = d // 64  # transformer.py:61
    eps = cfg.get("eps", 1e-6)  # norms.py:118-119; CLIP passes norm_kwargs={"eps": 1e-5} (clip.py:130)
    act = cfg.get("activation", "GELU")
    net = pre_process(sd, patch_embed(sd, x), eps)
    if taps is not None:
        taps["tokens"] = net
    for i in range(cfg["num_layers"]):
        net = mixing_block(sd, i, net, heads, eps, act)
        if taps is not None:
            taps[f"block{i}"] = net
    # head = PreNorm(LayerNorm, Lambda(x[:, 0])) (api.py:365,397-402; high_level.py:42-45): normalise ALL tokens, take token 0
    if cfg.get("norm_after_head", False):  # api.py:391-393,440-444: head (cls pick) first, then head_norm
        net = F.layer_norm(net[:, 0], (d,), sd["encoder.head_norm.weight"], sd["encoder.head_norm.bias"], eps)
    else:
        net = F.layer_norm(net, (d,), sd["encoder.head.norms.0.weight"], sd["encoder.head.norms.0.bias"], eps)[:, 0]
    if taps is not None:
        taps["encoded"] = net
    if cfg.get("output_dim") is not None:  # transformer.py:93-94
        net = net @ sd["output_projection"]
        if taps is not None:
            taps["projected"] = net
    return net


def classifier_forward(sd: StateDict, x: Tensor, cfg: Dict[str, int], taps: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """VanillaClassifier.forward (modules/cv/classifier/vanilla.py:57-61) with the ViT encoder (SURVEY finding 4:
    ``encoder.encode`` is what it would call) and head = Linear(latent_dim, num_classes) (vanilla.py:43)."""
    enc = encoder_forward(sd, x, cfg, taps)
    logits = F.linear(enc, sd["head.linear.weight"], sd["head.linear.bias"])
    if taps is not None:
        taps["logits"] = logits
    return logits


def cross_entropy(logits: Tensor, labels: Tensor) -> Tensor:
    """CrossEntropyLoss._get_stat + ILoss._reduce('mean') (losses/basic.py:137-141; schema.py:767-775).
    labels: int64 [B, 1]; the gather is an integer index op."""
    log_prob = F.log_softmax(logits, dim=1)
    return (-log_prob.gather(dim=1, index=labels)).mean()


def train_step(sd: StateDict, x: Tensor, labels: Tensor, cfg: Dict[str, int], *, autocast_bf16: bool,
               want_taps: bool = False) -> Tuple[Tensor, StateDict, Dict[str, Tensor]]:
    """One fwd + loss + bwd (IDLModel.train, schema.py:1266-1276 forward/loss under autocast; :980 backward).
    Returns (loss, grads keyed like the state_dict, taps)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    taps: Dict[str, Tensor] = {}
    dev = x.device.type
    with torch.autocast(dev, dtype=torch.bfloat16, enabled=autocast_bf16):
        logits = classifier_forward(params, x, cfg, taps if want_taps else None)
        loss = cross_entropy(logits, labels)
    loss.backward()
    grads = {k: v.grad for k, v in params.items()}
    return loss.detach(), grads, {k: v.detach() for k, v in taps.items()}


def encoder_train_step(sd: StateDict, x: Tensor, upstream: Tensor, cfg: Dict[str, int], *, autocast_bf16: bool,
                       want_taps: bool = False) -> Tuple[Tensor, StateDict, Dict[str, Tensor]]:
    """Forward + backward of the bare encoder (e.g. CLIP's vision tower, which has no loss of its own in the reference):
    the scalar is ``sum(out * upstream)``, i.e. ``upstream`` [B, out] is the gradient arriving at the encoder output."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    taps: Dict[str, Tensor] = {}
    with torch.autocast(x.device.type, dtype=torch.bfloat16, enabled=autocast_bf16):
        out = encoder_forward(params, x, cfg, taps if want_taps else None)
    (out.float() * upstream).sum().backward()
    return out.detach(), {k: v.grad for k, v in params.items()}, {k: v.detach() for k, v in taps.items()}


# -----------------------------------------------------------------------------------------------------------------
# text tower: TeTEncoder (nlp/encoder/transformer.py:16-99) as CLIP._init_text builds it (multimodal/clip.py:175-188)
# -----------------------------------------------------------------------------------------------------------------
def tet_config(name: str = "clip_text") -> Dict[str, int]:
    table = {
        # CLIP defaults (clip.py:47-62): width 512, 8 heads, 12 layers, context 77, causal mask, QuickGELU, eps 1e-5
        "clip_text": dict(latent_dim=512, context_length=77, num_layers=12, causal=True, activation="quick_gelu", eps=1e-5),
        "clip_text_tiny": dict(latent_dim=128, context_length=12, num_layers=2, causal=True, activation="quick_gelu", eps=1e-5),
        "clip_text_small": dict(latent_dim=256, context_length=77, num_layers=3, causal=True, activation="quick_gelu", eps=1e-5),
    }
    return dict(table[name])


def tet_state_dict_spec(cfg: Dict[str, int]) -> List[Tuple[str, Tuple[int, ...]]]:
    """Parameters of TeTEncoder in state_dict order (the persistent bool buffer ``attention_mask`` [T, T] comes first in
    the reference's state_dict when use_triu_attn_mask=True, transformer.py:42-48; it is not a parameter)."""
    d, t = cfg["latent_dim"], cfg["context_length"]
    ff = 4 * d
    spec: List[Tuple[str, Tuple[int, ...]]] = [("encoder.pos_encoding.pos_encoding", (1, t, d))]
    for i in range(cfg["num_layers"]):
        b = f"encoder.mixing_blocks.{i}."
        spec += [
            (b + "token_norm.weight", (d,)), (b + "token_norm.bias", (d,)),
            (b + "token_mixing.net.in_w", (3 * d, d)), (b + "token_mixing.net.qkv_bias", (3 * d,)),
            (b + "token_mixing.net.out_linear.linear.weight", (d, d)), (b + "token_mixing.net.out_linear.linear.bias", (d,)),
            (b + "channel_norm.weight", (d,)), (b + "channel_norm.bias", (d,)),
            (b + "channel_mixing.net.0.linear.weight", (ff, d)), (b + "channel_mixing.net.0.linear.bias", (ff,)),
            (b + "channel_mixing.net.3.linear.weight", (d, ff)), (b + "channel_mixing.net.3.linear.bias", (d,)),
        ]
    spec += [("encoder.head.norms.0.weight", (d,)), ("encoder.head.norms.0.bias", (d,))]
    return spec


def tet_init_state_dict(cfg: Dict[str, int], seed: int = 0) -> StateDict:
    g = torch.Generator().manual_seed(seed)
    sd: StateDict = {}
    for key, shape in tet_state_dict_spec(cfg):
        if key.endswith("norm.weight") or key.endswith("norms.0.weight"):
            t = torch.ones(shape) + 0.1 * torch.randn(shape, generator=g)
        elif key.endswith("bias"):
            t = 0.02 * torch.randn(shape, generator=g)
        elif key.endswith("pos_encoding"):
            t = 0.01 * torch.randn(shape, generator=g)  # clip.py:196
        else:
            t = torch.randn(shape, generator=g) * (shape[-1] ** -0.5) * 0.7  # clip.py:197-206 scales, roughly
        sd[key] = t.float()
    return sd


def tet_forward(sd: StateDict, net: Tensor, cfg: Dict[str, int]) -> Tensor:
    """TeTEncoder.forward (transformer.py:77-99) on an embedded sequence [B, T, D]: pre_process = + positional encoding
    (api.py:419-438, no head token), pre-norm blocks with the upper-triangular mask (:42-48 -> attentions.py:246-253),
    post_process = PreNorm(Identity): LayerNorm over every token (api.py:383-402)."""
    d = cfg["latent_dim"]
    heads = d // 64
    eps, act = cfg["eps"], cfg["activation"]
    net = net + sd["encoder.pos_encoding.pos_encoding"]
    for i in range(cfg["num_layers"]):
        b = f"encoder.mixing_blocks.{i}."
        t = F.layer_norm(net, (d,), sd[b + "token_norm.weight"], sd[b + "token_norm.bias"], eps)
        net = net + attention(sd, b + "token_mixing.net.", t, heads, causal=cfg["causal"])
        c = F.layer_norm(net, (d,), sd[b + "channel_norm.weight"], sd[b + "channel_norm.bias"], eps)
        net = net + feed_forward(sd, b + "channel_mixing.net.", c, act)
    return F.layer_norm(net, (d,), sd["encoder.head.norms.0.weight"], sd["encoder.head.norms.0.bias"], eps)


def tet_train_step(sd: StateDict, x: Tensor, upstream: Tensor, cfg: Dict[str, int], *, autocast_bf16: bool):
    """Forward + backward with ``upstream`` [B, T, D] as the gradient of the output; returns (out, dx, grads)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    xin = x.detach().clone().requires_grad_(True)
    with torch.autocast(x.device.type, dtype=torch.bfloat16, enabled=autocast_bf16):
        out = tet_forward(params, xin, cfg)
    (out.float() * upstream).sum().backward()
    return out.detach(), xin.grad, {k: v.grad for k, v in params.items()}
