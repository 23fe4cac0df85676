"""Data-parallel gradient exchange: bucketed NCCL all-reduce of the flat gradient arena, overlapped with backward.

Replaces what ``accelerator.prepare`` sets up in the reference (cflearn/trainer.py:226-229,266-273 -> torch DDP ->
NCCL; launched one process per GPU by cflearn/api/api.py:269-294).  Rank / world size come from the same environment
variables the reference reads (cflearn/toolkit.py:1917-1922: RANK, WORLD_SIZE, LOCAL_RANK).

The engine's backward produces gradients block by block, last block first.  The arena is laid out in forward order,
so each transformer block is one contiguous bucket (~28 MB fp32 for ViT-B/16, DDP's default is 25 MB).  As soon as a
block's last wgrad kernel has been queued, ``ready(i)`` records an event on the compute stream and launches
``all_reduce(AVG)`` for that slice on a dedicated communication stream; ``finish()`` makes the compute stream wait for
the last bucket.  Gradient averaging therefore equals the gradient of the global-batch mean loss (SURVEY.md
finding 3: the reference defines multi-GPU parity only mathematically).
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Tuple, Union

import torch
import torch.distributed as dist

BucketKey = Union[int, str]


def ddp_info() -> Optional[Tuple[int, int, int]]:
    """(rank, world_size, local_rank) or None -- same contract as cflearn/toolkit.py:1902-1922 get_ddp_info."""
    if "RANK" not in os.environ or "WORLD_SIZE" not in os.environ:
        return None
    return int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ.get("LOCAL_RANK", 0))


def shard_indices(n: int, rank: int, world: int) -> range:
    """DistributedSampler's rule without shuffling (cflearn/data/pytorch/api.py:62-71): rank r takes r, r+W, ..."""
    return range(rank, n, world)


class GradBucketReducer:
    def __init__(self, arena, num_layers: int, process_group=None):
        self.arena = arena
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_initialized() else 1
        self.buckets: Dict[BucketKey, Tuple[int, int]] = {}
        keys = [k for k, _ in arena.spec]
        offs = arena.offsets

        def span(pred) -> Tuple[int, int]:
            ks = [k for k in keys if pred(k)]
            lo = offs[ks[0]]
            last = ks[-1]
            hi = offs[last]
            n = 1
            for s in arena.shapes[last]:
                n *= s
            return lo, hi + (n + 63) // 64 * 64

        self.buckets["stem"] = span(lambda k: k.startswith("to_patches") or k.startswith("encoder.embedding_norm.")
                                    or k in ("encoder.head_token", "encoder.pos_encoding.pos_encoding"))
        for i in range(num_layers):
            self.buckets[i] = span(lambda k, i=i: k.startswith(f"encoder.mixing_blocks.{i}."))
        self.buckets["tail"] = span(lambda k: k.startswith("encoder.head.") or k.startswith("encoder.head_norm.") or k.startswith("head.linear"))
        if "output_projection" in offs:  # a parameter of the encoder itself: first in state_dict order, written with the tail
            self.buckets["proj"] = span(lambda k: k == "output_projection")
        self.comm_stream: Optional[torch.cuda.Stream] = None
        self._pending: List = []
        self.use_backend_avg = True

    def _stream(self) -> torch.cuda.Stream:
        if self.comm_stream is None:
            self.comm_stream = torch.cuda.Stream()
        return self.comm_stream

    def ready(self, key: BucketKey, grad: Optional[torch.Tensor] = None) -> None:
        """Bucket ``key`` of the gradient arena ``grad`` is complete on the current stream: start reducing it."""
        if self.world == 1:
            return
        if key == "tail" and "proj" in self.buckets:
            self.ready("proj", grad)
        lo, hi = self.buckets[key]
        if grad is None:
            grad = self.arena.grad
        if not grad.is_cuda:  # gloo path (host-side logic tests): same protocol, SUM then divide in finish()
            work = dist.all_reduce(grad[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self._pending.append((work, grad[lo:hi], True))
            return
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        cs = self._stream()
        cs.wait_event(ev)
        with torch.cuda.stream(cs):
            work = dist.all_reduce(grad[lo:hi], op=dist.ReduceOp.AVG, group=self.group, async_op=True)
            self._pending.append((work, None, False))

    def finish(self) -> None:
        if self.world == 1:
            return
        on_host = False
        for work, view, host in self._pending:
            work.wait()  # on CUDA this only orders streams, it does not block the host
            if host:
                view.div_(self.world)
                on_host = True
        self._pending.clear()
        if not on_host:
            torch.cuda.current_stream().wait_stream(self._stream())


def allreduce_flat_cpu(grad: torch.Tensor, buckets: List[Tuple[int, int]], group=None) -> None:
    """Host-side (gloo) version of the bucket protocol, used by the world_size=2 CPU tests."""
    world = dist.get_world_size(group)
    works = [dist.all_reduce(grad[lo:hi], op=dist.ReduceOp.SUM, group=group, async_op=True) for lo, hi in buckets]
    for w in works:
        w.wait()
    grad.div_(world)


def attach_reducer(module, process_group=None) -> GradBucketReducer:
    """Enable overlapped gradient averaging for a ViTEncoderB200 / VanillaClassifierB200."""
    module.arena.ensure()
    red = GradBucketReducer(module.arena, module.geo.L, process_group)
    module.engine.reducer = red
    return red


def broadcast_parameters(module, src: int = 0, process_group=None) -> None:
    """Identical replicas at start (what DDP's constructor does): one broadcast of the flat fp32 arena."""
    module.arena.ensure()
    dist.broadcast(module.arena.flat, src=src, group=process_group)
