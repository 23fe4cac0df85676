"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's per-item input pipeline for image batches (SURVEY.md 8f
row N4), the host-side work the fused ``b200_patch_im2col_u8`` kernel replaces.  Never imported by the product package.

  static_normalize     x.astype(float64) / division                    cflearn/data/blocks/cv/normalize.py:11-24
  imagenet_normalize   (x - mean) / std   (float64, per channel)        cflearn/data/blocks/cv/normalize.py:47-67
  affine_normalize     (x - center) / scale                             cflearn/data/blocks/cv/normalize.py:27-44
  hwc_to_chw           transpose([2, 0, 1]) + ascontiguousarray         cflearn/data/blocks/cv/hwc_to_chw.py:9-15
  TensorBatcher        np_batch_to_tensor + to_device                   cflearn/data/utils.py:255-283, toolkit.py:1182-1206

Parity UNPINNED: ``cflearn.data`` does not import through the loader shim (it needs ``cftool`` / ``OPT`` state that is not in
the container), and ``cftool.array.to_torch`` -- assumed to produce float32 tensors, as SURVEY.md N4 states -- is absent.
The arithmetic above is two numpy expressions; the restatement is checked against hand-computed values in tests/test_oracle.py.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def input_pipeline(x_u8_hwc: np.ndarray, division: float = 255.0, mean: Optional[Sequence[float]] = None,
                   std: Optional[Sequence[float]] = None) -> torch.Tensor:
    """uint8 [B, H, W, C] -> float32 tensor [B, C, H, W], op for op as the runtime blocks do it per item."""
    items = []
    for item in x_u8_hwc:
        inp = item.astype(np.float64) / division                       # static_normalize
        if mean is not None or std is not None:
            m = np.asarray(mean if mean is not None else [0.0] * item.shape[-1], dtype=np.float64)
            s = np.asarray(std if std is not None else [1.0] * item.shape[-1], dtype=np.float64)
            inp = (inp.astype(np.float64) - m) / s                     # imagenet_normalize / affine_normalize
        inp = np.ascontiguousarray(inp.transpose([2, 0, 1]))           # hwc_to_chw
        items.append(inp)
    return torch.from_numpy(np.stack(items, axis=0).astype(np.float32))  # to_torch: float32
