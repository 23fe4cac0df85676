"""Plug-in surface: the reference's module registry, mirrored name for name.

The reference builds every network from ``module_dict`` (cflearn/modules/common.py:30) through
``register_module`` (:33-34) / ``build_module`` (:37-53) and the prefixed sub-registries (``PrefixModules``, :56-83 --
e.g. ``encoders.vit`` at cflearn/modules/cv/encoder/transformer.py:16-17, ``cv_clf`` at
cflearn/modules/cv/classifier/vanilla.py:16).  This file keeps those semantics (deep-copied config, kwargs merged
over it, unknown keys dropped -- the behaviour of ``cftool.misc.safe_execute``) and offers ``install_into`` to drop
the B200 modules into the reference's own dict when ``cflearn`` is importable, so existing configs
(``module_name="cv_clf"``, ``encoder="vit"``) resolve to the new kernels without being edited.
"""
from __future__ import annotations

import inspect
import json
from typing import Any, Callable, Dict, Optional, Type, Union

import torch.nn as nn

from .clip import CLIPB200
from .fcnn import FCNNB200
from .vit import TeTEncoderB200, VanillaClassifierB200, ViTEncoderB200

module_dict: Dict[str, Type[nn.Module]] = {}


def register_module(name: str, *, allow_duplicate: bool = False) -> Callable[[Type[nn.Module]], Type[nn.Module]]:
    def _deco(cls: Type[nn.Module]) -> Type[nn.Module]:
        if name in module_dict and not allow_duplicate:
            raise ValueError(f"module '{name}' is already registered")
        module_dict[name] = cls
        return cls

    return _deco


def _safe_execute(fn: Callable, kwargs: Dict[str, Any]) -> Any:
    target = fn.__init__ if inspect.isclass(fn) else fn
    sig = inspect.signature(target)
    if any(p.kind is inspect.Parameter.VAR_KEYWORD for p in sig.parameters.values()):
        return fn(**kwargs)
    return fn(**{k: v for k, v in kwargs.items() if k in sig.parameters})


def _copy_containers(d: Any) -> Any:
    """``cftool.misc.shallow_copy_dict``: nested dict / list containers are copied, leaves shared
    (an ``nn.Module`` passed as ``embedding_norm`` must stay the same object)."""
    if isinstance(d, dict):
        return {k: _copy_containers(v) for k, v in d.items()}
    if isinstance(d, list):
        return [_copy_containers(v) for v in d]
    return d


def _update_dict(src: Dict[str, Any], tgt: Dict[str, Any]) -> Dict[str, Any]:
    """``cftool.misc.update_dict``: merge ``src`` INTO ``tgt`` recursively (src wins on leaves), return ``tgt``."""
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(tgt.get(k), dict):
            _update_dict(v, tgt[k])
        else:
            tgt[k] = v
    return tgt


def build_module(name: str, *, config: Optional[Union[str, Dict[str, Any]]] = None, **kwargs: Any) -> nn.Module:
    """Same contract as cflearn/modules/common.py:37-53: ``config`` may be a dict or a JSON path; keyword arguments are
    merged over it RECURSIVELY (``update_dict``), so ``encoder_config={"num_layers": 2}`` overrides one nested key and
    keeps the others."""
    if config is None:
        kw: Dict[str, Any] = {}
    elif isinstance(config, dict):
        kw = _copy_containers(config)
    else:
        with open(config, "r") as f:
            kw = json.load(f)
    _update_dict(_copy_containers(kwargs), kw)
    if name not in module_dict:
        raise KeyError(f"module '{name}' is not registered (available: {sorted(module_dict)})")
    return _safe_execute(module_dict[name], kw)


class PrefixModules:
    """cflearn/modules/common.py:56-83: a view of ``module_dict`` whose keys carry ``"<prefix>."``."""

    def __init__(self, prefix: str) -> None:
        self.prefix = prefix

    def prefixed(self, name: str) -> str:
        return f"{self.prefix}.{name}"

    def has(self, name: str) -> bool:
        return self.prefixed(name) in module_dict

    def get(self, name: str) -> Optional[Type[nn.Module]]:
        return module_dict.get(self.prefixed(name))

    def register(self, name: str, **kwargs: Any) -> Callable:
        return register_module(self.prefixed(name), **kwargs)

    def build(self, name: str, *, config: Optional[Union[str, Dict[str, Any]]] = None, **kwargs: Any) -> nn.Module:
        return build_module(self.prefixed(name), config=config, **kwargs)


encoders = PrefixModules("encoders")
register_encoder = encoders.register
build_encoder = encoders.build

# the B200 modules, under new names and under the reference's names
register_module("encoders.vit_b200")(ViTEncoderB200)
register_module("encoders.vit")(ViTEncoderB200)
register_module("cv_clf_b200")(VanillaClassifierB200)
register_module("cv_clf")(VanillaClassifierB200)
register_module("fcnn_b200")(FCNNB200)
register_module("tet_b200")(TeTEncoderB200)
register_module("tet")(TeTEncoderB200)  # cflearn/modules/nlp/encoder/transformer.py:16
register_module("clip_b200")(CLIPB200)
register_module("clip")(CLIPB200)  # cflearn/modules/multimodal/clip.py:21
register_module("fcnn")(FCNNB200)  # cflearn/modules/ml/fcnn.py:12


def install_into(reference_module_dict: Dict[str, Any], *, override: bool = True) -> Dict[str, Any]:
    """Assign the B200 classes into the reference's ``cflearn.modules.common.module_dict``.

    ``register_core``'s duplicate policy is unverified, so we assign directly.  Returns the entries
    that were replaced so a caller can restore them."""
    replaced = {}
    for name in ("encoders.vit_b200", "cv_clf_b200", "fcnn_b200", "tet_b200", "clip_b200") + (("encoders.vit", "cv_clf", "fcnn", "tet", "clip") if override else ()):
        if name in reference_module_dict:
            replaced[name] = reference_module_dict[name]
        reference_module_dict[name] = module_dict[name]
    return replaced
