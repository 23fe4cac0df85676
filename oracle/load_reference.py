"""TEST INFRASTRUCTURE ONLY -- imports the real reference's ``cflearn.modules`` from /root/reference.

Used in THIS container to (a) pin ``oracle/vit_oracle.py`` (our CPU restatement) against the reference's own code
and (b) generate the golden fixtures under ``tests/golden/`` (see ``oracle/make_golden.py``).  /root/reference does
not exist on the GPU box, so nothing that runs there may import this module.

``import cflearn`` itself needs ``cftool`` (carefree-toolkit) and ``accelerate``, which are not installed and cannot
be (no network).  Following SURVEY.md section 8(c) / Appendix C we pre-seed ``sys.modules`` with
  * a bare ``cflearn`` package whose ``__path__`` points at /root/reference/cflearn (skips cflearn/__init__.py:1-31),
  * stub ``cftool.*`` / ``accelerate`` modules: real implementations of the handful of helpers that take part in
    module construction (``shallow_copy_dict``, ``update_dict``, ``safe_execute``, ``register_core``,
    ``WithRegister``, ``squeeze``, ``l2_normalize``) and distinct empty placeholder classes for everything else.
No reference source is copied; the reference files are imported where they lie.
"""
from __future__ import annotations

import inspect
import os
import sys
import types
from typing import Any, Callable, Dict, Optional

REFERENCE_ROOT = os.environ.get("CFLEARN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "cflearn", "modules"))


# ---------------------------------------------------------------------------------------------------------------
# the few cftool helpers with behaviour (spec: SURVEY.md Appendix C)
# ---------------------------------------------------------------------------------------------------------------
def shallow_copy_dict(d: Any) -> Any:
    if isinstance(d, dict):
        return {k: shallow_copy_dict(v) for k, v in d.items()}
    if isinstance(d, list):
        return [shallow_copy_dict(v) for v in d]
    return d


def update_dict(src: dict, tgt: dict) -> dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(tgt.get(k), dict):
            update_dict(v, tgt[k])
        else:
            tgt[k] = v
    return tgt


def safe_execute(fn: Callable, kwargs: Dict[str, Any]) -> Any:
    target = fn.__init__ if inspect.isclass(fn) else fn
    sig = inspect.signature(target)
    if any(p.kind is inspect.Parameter.VAR_KEYWORD for p in sig.parameters.values()):
        return fn(**kwargs)
    return fn(**{k: v for k, v in kwargs.items() if k in sig.parameters})


def register_core(name: str, registry: dict, *, before_register: Optional[Callable] = None,
                  after_register: Optional[Callable] = None, allow_duplicate: bool = False) -> Callable:
    def _deco(cls: Any) -> Any:
        if before_register is not None:
            before_register(cls)
        if name not in registry or allow_duplicate:
            registry[name] = cls
        if after_register is not None:
            after_register(cls)
        return cls

    return _deco


class WithRegister:
    d: Dict[str, Any]
    __identifier__: str

    def __class_getitem__(cls, item: Any) -> Any:
        return cls

    @classmethod
    def get(cls, name: str) -> Any:
        return cls.d[name]

    @classmethod
    def has(cls, name: str) -> bool:
        return name in cls.d

    @classmethod
    def make(cls, name: str, config: Dict[str, Any], *, ensure_safe: bool = False) -> Any:
        return safe_execute(cls.get(name), config)

    @classmethod
    def register(cls, name: str, **kwargs: Any) -> Callable:
        def before(cls_: Any) -> None:
            cls_.__identifier__ = name

        return register_core(name, cls.d, before_register=before)

    @classmethod
    def check_subclass(cls, name: str) -> bool:
        return issubclass(cls.d[name], cls)


def check_requires(fn: Callable, name: str, strict: bool = True) -> bool:
    """cftool.misc.check_requires as its call sites use it (cflearn/toolkit.py:1616-1618 ``scheduler_requires_metric``):
    does ``fn`` take a parameter called ``name``?  (strict: as a named parameter, not through **kwargs)"""
    params = inspect.signature(fn).parameters
    if name in params:
        return True
    return (not strict) and any(p.kind is inspect.Parameter.VAR_KEYWORD for p in params.values())


def squeeze(t: Any) -> Any:
    return t.squeeze()


def l2_normalize(t: Any) -> Any:
    return t / t.norm(dim=-1, keepdim=True)


def _noop(*args: Any, **kwargs: Any) -> None:
    return None


# ---------------------------------------------------------------------------------------------------------------
# stub modules
# ---------------------------------------------------------------------------------------------------------------
class _StubModule(types.ModuleType):
    """Module whose unknown attributes resolve to DISTINCT, subscriptable, empty placeholder classes."""

    def __getattr__(self, name: str) -> Any:
        if name.startswith("__"):
            raise AttributeError(name)

        def _register(cls, *a: Any, **k: Any) -> Callable:
            return lambda c: c

        def _getitem(cls, item: Any) -> Any:
            return cls

        placeholder = type(name, (), {"register": classmethod(_register), "__class_getitem__": classmethod(_getitem),
                                      "__module__": self.__name__})
        setattr(self, name, placeholder)
        return placeholder


def _install_stubs() -> None:
    real = {
        "shallow_copy_dict": shallow_copy_dict, "update_dict": update_dict, "safe_execute": safe_execute,
        "register_core": register_core, "WithRegister": WithRegister, "squeeze": squeeze, "l2_normalize": l2_normalize,
        "check_requires": check_requires,
        "print_info": _noop, "print_warning": _noop, "print_error": _noop,
    }
    names = ["cftool", "cftool.misc", "cftool.array", "cftool.types", "cftool.cv", "cftool.pipeline", "cftool.dist",
             "cftool.data_structures", "cftool.ml", "cftool.ml.utils", "cftool.web", "cftool.constants",
             "accelerate", "accelerate.utils"]
    for n in names:
        if n in sys.modules and not isinstance(sys.modules[n], _StubModule):
            continue  # a real install is present: use it
        mod = _StubModule(n)
        mod.__path__ = []  # type: ignore[attr-defined]
        for k, v in real.items():
            setattr(mod, k, v)
        sys.modules[n] = mod
    # typing aliases used in annotations at import time
    import numpy as np
    import torch

    t = sys.modules["cftool.types"]
    t.tensor_dict_type = Dict[str, Any]
    t.np_dict_type = Dict[str, Any]
    t.arr_type = Any
    t.TNumberPair = Any
    t.general_config_type = Any
    a = sys.modules["cftool.array"]
    a.to_torch = lambda x: torch.from_numpy(x) if isinstance(x, np.ndarray) else x
    a.to_numpy = lambda x: x.detach().cpu().numpy()
    a.squeeze = squeeze
    a.l2_normalize = l2_normalize
    a.tensor_dict_type = Dict[str, Any]


_loaded = False


def load_reference_modules() -> types.ModuleType:
    """Returns the reference's ``cflearn.modules`` package (imported from /root/reference, unmodified)."""
    global _loaded
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    if not _loaded:
        _install_stubs()
        if "cflearn" not in sys.modules:
            pkg = types.ModuleType("cflearn")
            pkg.__path__ = [os.path.join(REFERENCE_ROOT, "cflearn")]  # type: ignore[attr-defined]
            sys.modules["cflearn"] = pkg
        _loaded = True
    import importlib

    return importlib.import_module("cflearn.modules")


def load_modules() -> types.SimpleNamespace:
    """The reference's registry entry points, imported where they lie: ``module_dict`` / ``build_module``
    (cflearn/modules/common.py:30-53) and ``build_encoder`` (cflearn/modules/cv/common.py:286-292)."""
    load_reference_modules()
    import importlib

    common = importlib.import_module("cflearn.modules.common")
    cv_common = importlib.import_module("cflearn.modules.cv.common")
    return types.SimpleNamespace(module_dict=common.module_dict, build_module=common.build_module,
                                 build_encoder=cv_common.build_encoder, common=common)


if __name__ == "__main__":
    m = load_reference_modules()
    from cflearn.modules.common import module_dict

    print(f"reference cflearn.modules imported; {len(module_dict)} registered modules")
