"""TEST INFRASTRUCTURE ONLY -- plain-PyTorch (fp32, CPU) restatement of the reference's FCNN training step
(BASELINE.json configs[0]: "FCNN 2-hidden-layer on synthetic 10-feature tabular, CPU, batch 128").

Never imported by the product package; only ``tests/``, ``__graft_entry__`` and ``bench.py``'s CPU legs may use it.
Pinned bit-for-bit against the reference's own ``FCNN`` / ``MAELoss`` / ``MSELoss`` code by ``oracle/make_golden_fcnn.py``
(fixture: ``tests/golden/fcnn_reference.pt``).  Paths below are relative to /root/reference/cflearn/.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

StateDict = Dict[str, Tensor]


def hidden_units(input_dim: int, hidden: Optional[Sequence[int]] = None) -> List[int]:
    """modules/ml/fcnn.py:29-31: two hidden layers of max(32, min(1024, 2 * input_dim)) units by default."""
    if hidden is not None:
        return list(hidden)
    return 2 * [max(32, min(1024, 2 * input_dim))]


def state_dict_spec(input_dim: int, output_dim: int, hidden: Optional[Sequence[int]] = None) -> List[Tuple[str, Tuple[int, ...]]]:
    """Keys of FCNN.net = Sequential(Mapping..., nn.Linear) (fcnn.py:36-56): a Mapping holds the hijack-able ``Linear``
    wrapper (core/mappings.py:54-62 -> core/customs.py:41-100), hence ``net.{i}.linear.linear.*``; the output layer is a
    bare ``nn.Linear`` (fcnn.py:54)."""
    dims = [input_dim] + hidden_units(input_dim, hidden)
    out: List[Tuple[str, Tuple[int, ...]]] = []
    for i in range(len(dims) - 1):
        out += [(f"net.{i}.linear.linear.weight", (dims[i + 1], dims[i])), (f"net.{i}.linear.linear.bias", (dims[i + 1],))]
    n = len(dims) - 1
    out += [(f"net.{n}.weight", (output_dim, dims[-1])), (f"net.{n}.bias", (output_dim,))]
    return out


def init_state_dict(input_dim: int, output_dim: int, hidden: Optional[Sequence[int]] = None, seed: int = 0) -> StateDict:
    """Synthetic weights for parity runs (parity is checked with injected identical weights)."""
    g = torch.Generator().manual_seed(seed)
    sd: StateDict = {}
    for key, shape in state_dict_spec(input_dim, output_dim, hidden):
        scale = 0.1 if key.endswith("bias") else 1.0 / float(np.sqrt(shape[-1]))
        sd[key] = torch.randn(shape, generator=g) * scale
    return sd


def toy_data(n: int = 1000, input_dim: int = 10, seed: int = 123) -> Tuple[Tensor, Tensor]:
    """examples/ml/simple/toy.py:12-16: x ~ U[0,1)^{n x 10}, y = x w * 100 (numpy legacy RNG seeded by seed_everything)."""
    rs = np.random.RandomState(seed)
    x = rs.random_sample([n, input_dim])
    w = rs.random_sample([input_dim, 1])
    y = x.dot(w) * 100.0
    return torch.from_numpy(x).float(), torch.from_numpy(y).float()


def forward(sd: StateDict, x: Tensor) -> Tensor:
    """fcnn.py:58-59 over Mapping.forward (mappings.py:74-83: linear -> ReLU; no BN / dropout with FCNN's defaults,
    fcnn.py:23-24) and the final nn.Linear."""
    n = len(sd) // 2 - 1
    net = x
    for i in range(n):
        net = F.relu(F.linear(net, sd[f"net.{i}.linear.linear.weight"], sd[f"net.{i}.linear.linear.bias"]))
    return F.linear(net, sd[f"net.{n}.weight"], sd[f"net.{n}.bias"])


def multi_task_loss(pred: Tensor, y: Tensor, w_mae: float = 1.0, w_mse: float = 1.0) -> Tuple[Tensor, Tensor, Tensor]:
    """loss_name="multi_task", loss_names=["mae", "mse"] (toy.py:19-20): losses/basic.py:45-48 (l1, reduction none) and
    :58-61 (mse, reduction none), each reduced by ILoss._reduce = mean (schema.py:767-771), merged with unit weights
    (losses/common.py:72-79,84-88)."""
    mae = F.l1_loss(pred, y, reduction="none").mean()
    mse = F.mse_loss(pred, y, reduction="none").mean()
    return mae * w_mae + mse * w_mse, mae, mse


def train_step(sd: StateDict, x: Tensor, y: Tensor) -> Tuple[Tensor, Tensor, Dict[str, Tensor]]:
    """One forward + loss + backward (schema.py:1174-1294 minus the optimizer): returns (loss, predictions, grads)."""
    params = {k: v.detach().clone().requires_grad_(True) for k, v in sd.items()}
    pred = forward(params, x)
    loss, _, _ = multi_task_loss(pred, y)
    loss.backward()
    return loss.detach(), pred.detach(), {k: p.grad for k, p in params.items()}
