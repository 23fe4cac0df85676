"""CPU tests (no GPU, no compute calls into the CUDA library): C-ABI export table, registry / module surface,
parameter arena, data-parallel bucket layout and the world_size=2 gloo run of the bucket reducer."""
import json
import os
import re
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import cflearn_b200  # noqa: E402
from cflearn_b200 import _cabi, dp, registry, vit  # noqa: E402


def _header_functions():
    text = open(os.path.join(ROOT, "include", "b200_cflearn.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    assert _cabi.available(), _cabi.load_error()
    declared = _header_functions()
    assert declared, "no declarations parsed from include/b200_cflearn.h"
    exported = set(_cabi.exported_symbols())
    assert set(declared) == set(_cabi.SIGNATURES), (set(declared) ^ set(_cabi.SIGNATURES))
    assert not [n for n in declared if n not in exported]
    assert _cabi.lib().b200_abi_version() == 2
    assert _cabi.lib().b200_gemm_pick_splits(768, 768, 50432) >= 1


def test_product_path_fails_loudly_without_gpu():
    m = registry.build_module("cv_clf", config=dict(in_channels=3, num_classes=8, img_size=32, latent_dim=64, encoder="vit",
                                                    encoder_config=dict(patch_size=16, num_layers=1)))
    with pytest.raises(cflearn_b200.B200Error):
        m(torch.randn(2, 3, 32, 32))  # CPU tensor: no fallback
    with pytest.raises(cflearn_b200.B200Error):
        vit.cross_entropy(torch.zeros(2, 8), torch.zeros(2, dtype=torch.long))


def test_missing_library_is_reported(tmp_path, monkeypatch):
    monkeypatch.setattr(_cabi, "LIB_PATH", str(tmp_path / "libb200_cflearn.so"))
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "_load_error", None)
    assert not _cabi.available()
    with pytest.raises(_cabi.B200Error, match="no CPU fallback"):
        _cabi.lib()


def test_registry_semantics_mirror_reference():
    # unknown keys are dropped (safe_execute), config is deep-copied, kwargs override the config
    cfg = dict(img_size=32, patch_size=16, in_channels=3, latent_dim=128, num_layers=1, something_else=3)
    e = registry.build_module("encoders.vit", config=cfg, num_layers=2)
    assert e.geo.L == 2 and cfg["num_layers"] == 1
    assert registry.encoders.has("vit") and registry.encoders.get("vit_b200") is vit.ViTEncoderB200
    with pytest.raises(KeyError):
        registry.build_module("no_such_module")
    fake_reference_dict = {"encoders.vit": object, "cv_clf": object}
    replaced = registry.install_into(fake_reference_dict)
    assert fake_reference_dict["encoders.vit"] is vit.ViTEncoderB200 and replaced["cv_clf"] is object
    for bad in (dict(dropout=0.1), dict(drop_path_rate=0.1), dict(norm_type="batch"), dict(output_dim=100),
                dict(feedforward_kwargs={"activation": "geglu"}), dict(to_patches_config={"padding": 1}),
                dict(embedding_norm=torch.nn.BatchNorm1d(128))):
        with pytest.raises(NotImplementedError):
            registry.build_module("encoders.vit", config=dict(img_size=32, patch_size=16, in_channels=3, latent_dim=128, **bad))


def test_fcnn_surface_mirrors_reference():
    # registry name, constructor defaults (fcnn.py:29-31), state_dict keys (golden fixture from the real reference)
    g = torch.load(os.path.join(ROOT, "tests", "golden", "fcnn_reference.pt"))
    m = registry.build_module("fcnn", config=dict(input_dim=10, output_dim=1, not_a_kwarg=1))
    assert list(m.state_dict().keys()) == g["keys"] and m.hidden_units == [32, 32]
    assert [tuple(v.shape) for v in m.state_dict().values()] == [(32, 10), (32,), (32, 32), (32,), (1, 32), (1,)]
    assert registry.build_module("fcnn", input_dim=600, output_dim=3).hidden_units == [1024, 1024]
    assert list(registry.build_module("fcnn", input_dim=4, output_dim=2, bias=False).state_dict()) == [
        "net.0.linear.linear.weight", "net.1.linear.linear.weight", "net.2.weight"]
    for bad in (dict(batch_norm=True), dict(dropout=0.3), dict(activation="GELU"), dict(mapping_type="res"), dict(rank=4)):
        with pytest.raises(NotImplementedError):
            registry.build_module("fcnn", input_dim=10, output_dim=1, **bad)
    with pytest.raises(cflearn_b200.B200Error):
        m(torch.randn(4, 10))  # CPU tensor: no fallback


def test_clip_vision_tower_options_mirror_reference_keys():
    # the ViTEncoder that CLIP._init_vision builds (multimodal/clip.py:121-135): key names AND order of the real reference
    with open(os.path.join(ROOT, "tests", "golden", "clip_vision_tiny_keys.json")) as f:
        golden = json.load(f)
    m = registry.build_module("encoders.vit", config=dict(
        img_size=64, patch_size=32, in_channels=3, latent_dim=128, to_patches_config={"bias": False}, num_layers=2,
        norm_kwargs={"eps": 1e-5}, embedding_norm=torch.nn.LayerNorm(128, 1e-5), attention_kwargs={"num_heads": 2},
        feedforward_kwargs={"activation": "quick_gelu"}, norm_after_head=True, output_dim=64))
    assert [[k, list(v.shape)] for k, v in m.state_dict().items()] == golden["keys"]
    assert m.geo.quick_gelu and m.geo.emb_eps == 1e-5 and m.geo.eps == 1e-5 and not m.geo.conv_bias


def test_tet_encoder_surface_mirrors_reference():
    # key names / order pinned by oracle/make_golden.py::pin_tet against the real TeTEncoder (buffer first, then parameters)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import vit_oracle as vo

    cfg = vo.tet_config("clip_text_tiny")
    m = registry.build_module("tet", config=dict(latent_dim=128, context_length=12, use_triu_attn_mask=True, num_layers=2,
                                                 norm_kwargs={"eps": 1e-5}, attention_kwargs={"num_heads": 2},
                                                 feedforward_kwargs={"activation": "quick_gelu"}, head_pooler=None))
    sd = m.state_dict()
    assert list(sd)[0] == "attention_mask" and sd["attention_mask"].dtype == torch.bool
    assert torch.equal(sd["attention_mask"], torch.ones(12, 12, dtype=torch.bool).triu(1))
    assert [(k, tuple(v.shape)) for k, v in list(sd.items())[1:]] == vo.tet_state_dict_spec(cfg)
    assert m.geo.causal and m.geo.tokens and m.geo.quick_gelu
    for bad in (dict(head_pooler="mean"), dict(dropout=0.1), dict(norm_position="post_norm"), dict(attention_kwargs={"num_heads": 6})):
        with pytest.raises(NotImplementedError):
            registry.build_module("tet", latent_dim=128, context_length=12, **bad)
    with pytest.raises(NotImplementedError):
        m(torch.zeros(1, 12, 128), mask=torch.zeros(12, 12, dtype=torch.bool))


def test_clip_module_surface_mirrors_reference():
    # CLIP.__init__ keywords, state_dict keys / order (pinned to the real reference CLIP by oracle/make_golden_clip.py)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import clip_oracle as co

    cfg = co.clip_config("clip_tiny")
    v, t = cfg["vision"], cfg["text"]
    m = registry.build_module("clip", config=dict(
        img_size=v["img_size"], latent_dim=cfg["latent_dim"], vision_latent_dim=v["latent_dim"], vision_patch_size=v["patch_size"],
        vision_num_heads=2, vision_num_layers=v["num_layers"], vocab_size=cfg["vocab_size"], context_length=t["context_length"],
        text_latent_dim=t["latent_dim"], text_num_heads=2, text_num_layers=t["num_layers"]))
    keys = [(k, tuple(p.shape)) for k, p in m.state_dict().items()]
    i = [k for k, _ in keys].index("text_transformer.attention_mask")
    assert keys[i - 1][0] == "token_embedding.weight"  # the buffer sits where the reference has it
    assert [kv for kv in keys if kv[0] != "text_transformer.attention_mask"] == co.state_dict_spec(cfg)
    assert abs(m.logit_scale.item() - 2.6592600) < 1e-5 and m.token_embedding.padding_idx == 0
    for bad in (dict(use_text=False), dict(token_type_size=2), dict(text_dropout=0.1), dict(text_head_pooler="mean")):
        with pytest.raises(NotImplementedError):
            registry.build_module("clip", img_size=64, latent_dim=64, vision_latent_dim=128, text_latent_dim=128, text_num_heads=2,
                                  vision_num_heads=2, **bad)
    with pytest.raises(cflearn_b200.B200Error):
        m(torch.zeros(1, 3, v["img_size"], v["img_size"]), torch.ones(1, t["context_length"], dtype=torch.long))  # CPU: no fallback


def test_param_arena_views_and_state_dict_roundtrip():
    m = registry.build_module("cv_clf", config=dict(in_channels=3, num_classes=16, img_size=32, latent_dim=128, encoder="vit",
                                                    encoder_config=dict(patch_size=16, num_layers=2)))
    a = m.arena
    for k, p in m.named_parameters():
        assert p.data_ptr() == a.flat.data_ptr() + 4 * a.offsets[k]
        assert a.offsets[k] % 64 == 0
    sd = {k: torch.randn_like(v) for k, v in m.state_dict().items()}
    m.load_state_dict(sd)
    a.ensure()
    for k in sd:
        assert torch.equal(a.p(k), sd[k])  # load_state_dict copies in place: the arena sees the new values
    m2 = m.double().float()  # _apply() re-allocates every parameter: views are broken, ensure() must repair them
    m2.arena.ensure()
    for k, p in m2.named_parameters():
        assert p.data_ptr() == m2.arena.flat.data_ptr() + 4 * m2.arena.offsets[k]
        assert torch.equal(p, sd[k])
    # reference cv_clf checkpoints carry an "encoder." prefix on the encoder's keys
    ref_sd = {("encoder." + k if not k.startswith("head.linear") else k): v for k, v in sd.items()}
    m.load_reference_state_dict(ref_sd)


def test_bucket_layout_covers_arena_without_overlap():
    m = registry.build_module("cv_clf", config=dict(in_channels=3, num_classes=16, img_size=32, latent_dim=128, encoder="vit",
                                                    encoder_config=dict(patch_size=16, num_layers=3)))
    red = dp.GradBucketReducer(m.arena, 3)
    spans = sorted(red.buckets.values())
    assert spans[0][0] == 0 and spans[-1][1] == m.arena.total
    for (lo0, hi0), (lo1, hi1) in zip(spans, spans[1:]):
        assert hi0 == lo1
    assert list(dp.shard_indices(10, 1, 4)) == [1, 5, 9]


def _worker(rank, world, port, total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    assert dp.ddp_info() == (rank, world, rank)
    torch.manual_seed(0)
    m = registry.build_module("cv_clf", config=dict(in_channels=3, num_classes=16, img_size=32, latent_dim=128, encoder="vit",
                                                    encoder_config=dict(patch_size=16, num_layers=2)))
    with torch.no_grad():
        m.arena.flat.add_(rank)  # replicas differ until the broadcast
    dp.broadcast_parameters(m, src=0)
    red = dp.attach_reducer(m)
    g = m.arena.grad
    g.copy_(torch.arange(g.numel(), dtype=torch.float32) * (rank + 1))
    # the engine signals buckets in backward order: tail, blocks L-1..0, stem
    for key in ["tail", 1, 0, "stem"]:
        red.ready(key, g)
    red.finish()
    q.put((rank, m.arena.flat.sum().item(), g.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_gloo_bucket_allreduce_matches_mean():
    world, port = 2, 29577
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 0, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, s0, g0), (_, s1, g1) = sorted(results, key=lambda t: t[0])
    assert s0 == s1  # parameters identical after the broadcast
    expect = torch.arange(g0.numel(), dtype=torch.float32) * 1.5  # mean of 1x and 2x
    assert torch.equal(g0, g1) and torch.allclose(g0, expect)
