"""carefree-learn_b200: B200-native (sm_100a) kernels for carefree-learn's data-parallel training step.

Only the hot path named by BASELINE.json lives here: the forward/backward of the ``cflearn.modules`` transformer
block stack (patch-embed stem, LayerNorm, packed-QKV attention, FeedForward, classifier head, cross-entropy) and
the bucketed gradient all-reduce that replaces the ``accelerate`` DDP wrap.  Compute is hand-written CUDA behind the
C-ABI in ``include/b200_cflearn.h``; PyTorch provides device memory, streams and ``torch.distributed`` only.

The directory name contains a hyphen, so import it through the ``cflearn_b200`` shim at the repository root.
"""
from . import _cabi
from ._cabi import B200Error, available, load_error

__all__ = ["_cabi", "B200Error", "available", "load_error"]
__version__ = "0.1.0"
