"""TEST INFRASTRUCTURE ONLY -- pins ``oracle/fcnn_oracle.py`` against the real reference and writes
tests/golden/fcnn_reference.pt.  Run in the build container (needs /root/reference):  python oracle/make_golden_fcnn.py

Builds the reference's own ``FCNN(10, 1)`` (modules/ml/fcnn.py) and ``MAELoss`` / ``MSELoss`` (losses/basic.py), imported
unmodified through oracle/load_reference.py, loads the oracle's synthetic weights with ``load_state_dict(strict=True)``
(which also checks keys and shapes), and asserts predictions, loss and every gradient are BIT-IDENTICAL to the oracle's
on the first 128-row batch of the toy data set of examples/ml/simple/toy.py.
"""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import fcnn_oracle as fo  # noqa: E402
from load_reference import load_reference_modules  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main() -> None:
    load_reference_modules()
    from cflearn.modules.ml.fcnn import FCNN
    from cflearn.losses.basic import MAELoss, MSELoss

    torch.set_num_threads(1)
    x_all, y_all = fo.toy_data()
    x, y = x_all[:128], y_all[:128]
    sd = fo.init_state_dict(10, 1, seed=0)
    ref = FCNN(10, 1)
    assert [tuple(v.shape) for v in ref.state_dict().values()] == [s for _, s in fo.state_dict_spec(10, 1)]
    assert list(ref.state_dict().keys()) == [k for k, _ in fo.state_dict_spec(10, 1)]
    ref.load_state_dict(sd, strict=True)
    ref.train()
    pred = ref(x)
    mae_fn, mse_fn = MAELoss(), MSELoss()
    mae = mae_fn._reduce(mae_fn(pred, y))
    mse = mse_fn._reduce(mse_fn(pred, y))
    loss = mae * 1.0 + mse * 1.0  # MultiTaskLoss._merge with unit weights (losses/common.py:72-79)
    loss.backward()
    r_grads = {k: p.grad for k, p in ref.named_parameters()}

    o_loss, o_pred, o_grads = fo.train_step(sd, x, y)
    assert torch.equal(o_pred, pred.detach()), "predictions differ from the reference"
    assert torch.equal(o_loss, loss.detach()), "loss differs from the reference"
    for k in r_grads:
        assert torch.equal(o_grads[k], r_grads[k]), f"grad {k} differs from the reference"
    print(f"fcnn oracle == reference (bit-exact): loss {loss.item():.6f}, {len(r_grads)} gradients")
    os.makedirs(GOLDEN, exist_ok=True)
    torch.save({"x": x, "y": y, "weights_seed": 0, "pred": pred.detach(), "loss": loss.detach(), "mae": mae.detach(),
                "mse": mse.detach(), "grads": {k: v.clone() for k, v in r_grads.items()},
                "keys": [k for k, _ in fo.state_dict_spec(10, 1)]}, os.path.join(GOLDEN, "fcnn_reference.pt"))
    print("wrote tests/golden/fcnn_reference.pt")


if __name__ == "__main__":
    main()
