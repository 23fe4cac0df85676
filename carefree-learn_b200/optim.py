"""Fused Adam over the flat parameter / gradient arenas, and a CUDA-graph wrapper for the whole training step.

``ArenaAdam`` mirrors ``torch.optim.Adam`` (the reference's default optimizer ``"adam"``, cflearn/optimizers.py:29-32,
built by BuildOptimizersBlock at cflearn/pipeline/blocks/basic.py:385-558 and stepped in cflearn/schema.py:983-984):
one kernel per step instead of one per tensor.  With ``capturable=True`` the step counter lives on the device so
the update can be replayed from a CUDA graph.

``GraphedTrainStep`` captures ``zero_grad -> forward -> cross-entropy -> backward (-> bucketed all-reduce) -> Adam``
(IDLModel.train, cflearn/schema.py:1174-1294, minus its per-step host work) into ONE CUDA graph: ~410 kernel launches
per ViT-B/16 step stop costing host time, so the GPU is never waiting on Python.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

import torch
from torch import Tensor

from . import ops
from ._cabi import call


class ArenaAdam:
    def __init__(self, module: Any, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 capturable: bool = False):
        self.module = module
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.capturable = capturable
        self.step_count = 0
        self.step_dev: Optional[Tensor] = None
        self.exp_avg: Optional[Tensor] = None
        self.exp_avg_sq: Optional[Tensor] = None

    def _state(self) -> None:
        arena = self.module.arena
        arena.ensure()
        if self.exp_avg is None or self.exp_avg.device != arena.flat.device:
            self.exp_avg = torch.zeros_like(arena.flat)
            self.exp_avg_sq = torch.zeros_like(arena.flat)
            if self.capturable:
                self.step_dev = torch.full((1,), self.step_count, dtype=torch.int32, device=arena.flat.device)

    def step(self) -> None:
        self._state()
        arena = self.module.arena
        self.step_count += 1
        call("b200_adam_step", arena.flat.data_ptr(), arena.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
             arena.total, float(self.lr), float(self.betas[0]), float(self.betas[1]), float(self.eps), float(self.weight_decay),
             self.step_count, None if self.step_dev is None else self.step_dev.data_ptr(), ops._stream())

    def zero_grad(self, set_to_none: bool = True) -> None:
        if set_to_none:
            for p in self.module.parameters():
                p.grad = None

    def state_dict(self) -> Dict[str, Any]:
        step = int(self.step_dev.item()) if self.step_dev is not None else self.step_count
        return {"step": step, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq, "lr": self.lr,
                "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay}


class GraphedTrainStep:
    """Whole training step of a ``VanillaClassifierB200`` as one CUDA graph with static input / loss buffers.

    ``step(x, labels)`` copies the batch into the static buffers (device->device or pinned-host->device, on the
    current stream), replays the graph and returns the static loss tensor (fp32 scalar, on the device)."""

    def __init__(self, model: Any, optimizer: ArenaAdam, batch: int, warmup: int = 2, process_group: Any = None):
        if not optimizer.capturable:
            raise ValueError("GraphedTrainStep needs ArenaAdam(capturable=True)")
        import torch.distributed as dist

        # data parallel: the graph holds zero_grad + forward + loss + backward; the gradient mean (one NCCL all-reduce of
        # the flat arena) and Adam run right after each replay.  (Capturing NCCL inside the graph hung on this stack.)
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.group = process_group
        if self.world > 1 and model.engine.reducer is not None:
            raise ValueError("detach the bucket reducer (model.engine.reducer = None) before graphing the DP step")
        g = model.geo
        dev = model.arena.flat.device
        self.model, self.optimizer = model, optimizer
        self.x = torch.zeros((batch, g.cin, g.img, g.img), dtype=torch.float32, device=dev)
        self.labels = torch.zeros((batch, 1), dtype=torch.int64, device=dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=dev)
        # warm up on a side stream (allocator pools, workspaces, lazy function attributes), then capture
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self._eager()
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        from . import _cabi

        self.graph = torch.cuda.CUDAGraph()
        n0 = _cabi.launch_count()
        with torch.cuda.graph(self.graph):
            self._eager()
        self.launches_per_replay = _cabi.launch_count() - n0  # b200 kernels inside one replay of the graph

    def _eager(self) -> None:
        self.optimizer.zero_grad()
        loss = self.model.train_step(self.x, self.labels)
        if self.world == 1:
            self.optimizer.step()
        self.loss.copy_(loss)

    def step(self, x: Tensor, labels: Tensor) -> Tensor:
        self.x.copy_(x, non_blocking=True)
        self.labels.copy_(labels.reshape(self.labels.shape), non_blocking=True)
        self.graph.replay()
        if self.world > 1:
            import torch.distributed as dist

            dist.all_reduce(self.model.arena.grad, op=dist.ReduceOp.AVG, group=self.group)
            self.optimizer.step()
        return self.loss
