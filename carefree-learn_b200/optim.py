"""Fused Adam over the flat parameter / gradient arenas: one kernel per step instead of one per tensor.

Mirrors ``torch.optim.Adam`` (the reference's default optimizer ``"adam"``, cflearn/optimizers.py:29-32, built by
BuildOptimizersBlock at cflearn/pipeline/blocks/basic.py:385-558 and stepped in cflearn/schema.py:983-984), so a
training run keeps the same update rule; ``zero_grad`` is a no-op because backward overwrites the gradient arena.
"""
from __future__ import annotations

from typing import Any, Dict

import torch

from . import ops
from ._cabi import call


class ArenaAdam:
    def __init__(self, module: Any, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        self.module = module
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        self.exp_avg = None
        self.exp_avg_sq = None

    def _state(self) -> None:
        arena = self.module.arena
        arena.ensure()
        if self.exp_avg is None or self.exp_avg.device != arena.flat.device:
            self.exp_avg = torch.zeros_like(arena.flat)
            self.exp_avg_sq = torch.zeros_like(arena.flat)

    def step(self) -> None:
        self._state()
        arena = self.module.arena
        self.step_count += 1
        call("b200_adam_step", arena.flat.data_ptr(), arena.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
             arena.total, float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
             self.step_count, ops._stream())

    def zero_grad(self, set_to_none: bool = True) -> None:
        if set_to_none:
            for p in self.module.parameters():
                p.grad = None

    def state_dict(self) -> Dict[str, Any]:
        return {"step": self.step_count, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "lr": self.lr,
                "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}
