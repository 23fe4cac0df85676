"""Tensor-level wrappers over the C-ABI: PyTorch supplies device memory and the current stream, nothing else.

Every function launches hand-written sm_100a kernels from ``libb200_cflearn.so`` on the *current* CUDA stream
(autograd's backward thread included) and raises ``B200Error`` on any failure.  No function here has a CPU or
ATen fallback.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch
from torch import Tensor

from . import _cabi
from ._cabi import B200Error, call

EPI_BIAS_BF16 = 0
EPI_BIAS_GELU_BF16 = 1
EPI_BIAS_RESID_F32 = 2
EPI_DGELU_BF16 = 3
EPI_PARTIAL_F32 = 4
EPI_BIAS_QGELU_BF16 = 5  # QuickGELU (CLIP towers) flavours of 1 and 3
EPI_DQGELU_BF16 = 6

_MAX_PARTS = 1024


def _ptr(t: Optional[Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts: Optional[Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise B200Error("b200 kernels need CUDA tensors (there is no CPU fallback)")


class Workspace:
    """Per-device scratch for split-K partials and column-sum partials (grown on demand, never freed)."""

    def __init__(self) -> None:
        self._bufs = {}

    def get(self, device: torch.device, nfloats: int, tag: str = "ws") -> Tensor:
        key = (device.index, tag)
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nfloats:
            buf = torch.empty(max(nfloats, 1 << 20), dtype=torch.float32, device=device)
            self._bufs[key] = buf
        return buf


WORKSPACE = Workspace()


# ----------------------------------------------------------------------------------------------------------------
# GEMM
# ----------------------------------------------------------------------------------------------------------------
def gemm(
    a: Tensor,
    b: Tensor,
    *,
    a_mn_major: bool = False,
    b_mn_major: bool = False,
    epilogue: int = EPI_BIAS_BF16,
    bias: Optional[Tensor] = None,
    out0: Optional[Tensor] = None,
    out1: Optional[Tensor] = None,
    aux: Optional[Tensor] = None,
    splits: int = 1,
    max_ctas: int = 0,
) -> Tensor:
    """C[M,N] = A . B^T on tcgen05.  ``a``: [M,K] (or [K,M] if a_mn_major); ``b``: [N,K] (or [K,N])."""
    _need_cuda(a, b, bias, out0, out1, aux)
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        raise B200Error("gemm operands must be bf16")
    if a.dim() != 2 or b.dim() != 2 or a.stride(1) != 1 or b.stride(1) != 1:
        raise B200Error("gemm operands must be 2-D with unit inner stride")
    if a_mn_major:
        K, M = a.shape
    else:
        M, K = a.shape
    if b_mn_major:
        Kb, N = b.shape
    else:
        N, Kb = b.shape
    if K != Kb:
        raise B200Error(f"gemm reduction mismatch: {K} vs {Kb}")
    f32_out = epilogue in (EPI_BIAS_RESID_F32, EPI_PARTIAL_F32)
    if out0 is None:
        if epilogue == EPI_PARTIAL_F32:
            out0 = torch.empty((splits, M, N), dtype=torch.float32, device=a.device)
        else:  # rows stay 16-byte aligned for TMA: pad the leading dimension, hand back the [M, N] view
            npad = (N + 7) // 8 * 8
            out0 = torch.empty((M, npad), dtype=torch.float32 if f32_out else torch.bfloat16, device=a.device)[:, :N]
    if out0.dtype != (torch.float32 if f32_out else torch.bfloat16):
        raise B200Error("gemm: out0 dtype does not match the epilogue")
    ldo = out0.stride(-2)
    call(
        "b200_gemm_bf16",
        a.data_ptr(), a.stride(0), int(a_mn_major),
        b.data_ptr(), b.stride(0), int(b_mn_major),
        M, N, K, epilogue, _ptr(bias), out0.data_ptr(), _ptr(out1), _ptr(aux), ldo, splits, max_ctas, _stream(),
    )
    return out0


_NUM_SMS: Dict[int, int] = {}


def num_sms() -> int:
    dev = torch.cuda.current_device()
    if dev not in _NUM_SMS:
        _NUM_SMS[dev] = torch.cuda.get_device_properties(dev).multi_processor_count
    return _NUM_SMS[dev]


def pick_splits(M: int, N: int, K: int) -> int:
    return int(_cabi.lib().b200_gemm_pick_splits(M, N, K))


def wgrad(dy: Tensor, x: Tensor, out: Tensor, *, round_bf16: bool = True, accumulate: bool = False) -> Tensor:
    """out[N_out, K_in] (fp32) (+)= dy[M, N_out]^T . x[M, K_in]   (split-K over the M tokens, then reduce)."""
    M, n_out = dy.shape
    k_in = x.shape[1]
    splits = pick_splits(n_out, k_in, M)
    part = WORKSPACE.get(dy.device, splits * n_out * k_in, "splitk")[: splits * n_out * k_in].view(splits, n_out, k_in)
    gemm(dy, x, a_mn_major=True, b_mn_major=True, epilogue=EPI_PARTIAL_F32, out0=part, splits=splits)
    call("b200_splitk_reduce", part.data_ptr(), splits, n_out * k_in, out.data_ptr(), int(round_bf16), int(accumulate), _stream())
    return out


def colsum(x: Tensor, out: Tensor, *, round_bf16: bool = True, accumulate: bool = False) -> Tensor:
    """out[cols] (fp32) (+)= sum over rows of x[rows, cols] (bf16): bias gradients."""
    _need_cuda(x, out)
    rows, cols = x.shape
    part = WORKSPACE.get(x.device, _MAX_PARTS * cols, "colsum")
    nparts = ctypes.c_int(0)
    call("b200_colsum_bf16", x.data_ptr(), x.stride(0), rows, cols, part.data_ptr(), _MAX_PARTS, ctypes.byref(nparts), _stream())
    call("b200_colsum_finish", part.data_ptr(), cols, nparts.value, cols, out.data_ptr(), int(round_bf16), int(accumulate), _stream())
    return out


# ----------------------------------------------------------------------------------------------------------------
# LayerNorm
# ----------------------------------------------------------------------------------------------------------------
def layernorm_fwd(x: Tensor, gamma: Tensor, beta: Tensor, eps: float, *, rows: int, dim: int, ld_x: int,
                  y_f32: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda(x, gamma, beta)
    y = torch.empty((rows, dim), dtype=torch.bfloat16, device=x.device)
    mean = torch.empty(rows, dtype=torch.float32, device=x.device)
    rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
    call("b200_layernorm_fwd", x.data_ptr(), ld_x, gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), _ptr(y_f32),
         mean.data_ptr(), rstd.data_ptr(), rows, dim, float(eps), _stream())
    return y, mean, rstd


def layernorm_bwd(
    dy: Tensor, x: Tensor, gamma: Tensor, mean: Tensor, rstd: Tensor, *, rows: int, dim: int, ld_x: int,
    dres: Optional[Tensor], dx_out: Tensor, ld_dx: int, dx_bf16: Optional[Tensor],
    dgamma: Tensor, dbeta: Tensor, accumulate: bool = False, dx_colsum: Optional[Tensor] = None,
) -> None:
    """LayerNorm backward.  ``dx_colsum`` (fp32 [dim]): also receives the bf16-rounded column sums of ``dx_bf16``,
    i.e. the bias gradient of the Linear layer whose dY this dx is -- fused, no extra pass over dx."""
    nw = 3 if dx_colsum is not None else 2
    part = WORKSPACE.get(x.device, 3 * _MAX_PARTS * dim, "lnbwd")
    nparts = ctypes.c_int(0)
    call(
        "b200_layernorm_bwd", dy.data_ptr(), int(dy.dtype == torch.float32), x.data_ptr(), ld_x, gamma.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
        _ptr(dres), dx_out.data_ptr(), ld_dx, _ptr(dx_bf16), part.data_ptr(), _MAX_PARTS,
        ctypes.byref(nparts), rows, dim, int(dx_colsum is not None), _stream(),
    )
    adjacent = dbeta.data_ptr() == dgamma.data_ptr() + 4 * dim  # adjacent in the gradient arena: one reduction for both
    if adjacent and dx_colsum is not None:
        call("b200_colsum_finish2", part.data_ptr(), nw * dim, nparts.value, 2 * dim, dgamma.data_ptr(), 0,
             dim, dx_colsum.data_ptr(), 1, int(accumulate), _stream())
        return
    if adjacent:
        call("b200_colsum_finish", part.data_ptr(), nw * dim, nparts.value, 2 * dim, dgamma.data_ptr(), 0, int(accumulate), _stream())
    else:
        call("b200_colsum_finish", part.data_ptr(), nw * dim, nparts.value, dim, dgamma.data_ptr(), 0, int(accumulate), _stream())
        call("b200_colsum_finish", part.data_ptr() + 4 * dim, nw * dim, nparts.value, dim, dbeta.data_ptr(), 0, int(accumulate), _stream())
    if dx_colsum is not None:
        call("b200_colsum_finish", part.data_ptr() + 8 * dim, nw * dim, nparts.value, dim, dx_colsum.data_ptr(), 1, int(accumulate), _stream())


# ----------------------------------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------------------------------
def attention_fwd(qkv: Tensor, B: int, T: int, H: int, *, scale: Optional[float] = None, causal: bool = False) -> Tuple[Tensor, Tensor]:
    _need_cuda(qkv)
    D = H * 64
    if qkv.dtype != torch.bfloat16 or qkv.numel() != B * T * 3 * D or not qkv.is_contiguous():
        raise B200Error("attention_fwd: qkv must be contiguous bf16 [B, T, 3*H*64]")
    out = torch.empty((B * T, D), dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty((B, H, T), dtype=torch.float32, device=qkv.device)
    sc = float(scale) if scale is not None else 0.125
    call("b200_attention_fwd", qkv.data_ptr(), out.data_ptr(), lse.data_ptr(), B, T, H, 64, sc, int(causal), _stream())
    return out, lse


def attention_bwd(qkv: Tensor, out: Tensor, dout: Tensor, lse: Tensor, B: int, T: int, H: int, *, scale: Optional[float] = None, causal: bool = False, dqkv: Optional[Tensor] = None,
                  dbias: Optional[Tensor] = None) -> Tensor:
    """dqkv (bf16 [B*T, 3*H*64]).  ``dbias`` (fp32 [3*H*64]): also receives the bf16-rounded column sums of dqkv -- the
    gradient of the packed qkv bias -- accumulated inside the kernel (no separate pass over dqkv)."""
    _need_cuda(qkv, out, dout, lse)
    D = H * 64
    if dqkv is None:
        dqkv = torch.empty((B * T, 3 * D), dtype=torch.bfloat16, device=qkv.device)
    sc = float(scale) if scale is not None else 0.125
    part = WORKSPACE.get(qkv.device, B * 3 * D, "attn_dbias") if dbias is not None else None
    call("b200_attention_bwd", qkv.data_ptr(), out.data_ptr(), dout.data_ptr(), lse.data_ptr(), dqkv.data_ptr(), B, T, H, 64, sc, int(causal),
         _ptr(part), _stream())
    if dbias is not None:
        call("b200_colsum_finish", part.data_ptr(), 3 * D, B, 3 * D, dbias.data_ptr(), 1, 0, _stream())
    return dqkv


# ----------------------------------------------------------------------------------------------------------------
# patch embedding glue, loss, casts
# ----------------------------------------------------------------------------------------------------------------
def patch_im2col(x: Tensor, patch: int) -> Tensor:
    _need_cuda(x)
    B, C, Hh, Ww = x.shape
    if Hh != Ww or x.dtype != torch.float32 or not x.is_contiguous():
        raise B200Error("patch_im2col: need contiguous fp32 [B, C, S, S]")
    side = Hh // patch
    cols = torch.empty((B * side * side, C * patch * patch), dtype=torch.bfloat16, device=x.device)
    call("b200_patch_im2col", x.data_ptr(), cols.data_ptr(), B, C, Hh, patch, _stream())
    return cols


def patch_im2col_u8(x: Tensor, patch: int, division: float = 255.0, mean=None, std=None) -> Tensor:
    """uint8 [B, S, S, C] (HWC) -> normalised bf16 im2col matrix: the reference's host-side input pipeline fused into the stem."""
    _need_cuda(x)
    if x.dtype != torch.uint8 or x.dim() != 4 or not x.is_contiguous() or x.shape[1] != x.shape[2]:
        raise B200Error("patch_im2col_u8: need contiguous uint8 [B, S, S, C]")
    B, S, _, C = x.shape
    side = S // patch
    cols = torch.empty((B * side * side, C * patch * patch), dtype=torch.bfloat16, device=x.device)
    dbl = ctypes.c_double * C
    m = dbl(*[float(v) for v in mean]) if mean is not None else None
    sd = dbl(*[float(v) for v in std]) if std is not None else None
    call("b200_patch_im2col_u8", x.data_ptr(), cols.data_ptr(), B, C, S, patch, float(division), m, sd, _stream())
    return cols


def add_pos(x: Tensor, pos: Tensor, B: int, T: int, D: int) -> Tensor:
    """net[b, t, :] = x[b, t, :] + pos[t, :] (fp32): the text tower's input stage."""
    _need_cuda(x, pos)
    net = torch.empty((B, T, D), dtype=torch.float32, device=x.device)
    call("b200_add_pos", x.data_ptr(), pos.data_ptr(), net.data_ptr(), B, T, D, _stream())
    return net


def add_pos_bwd(dnet: Tensor, dpos: Tensor, B: int, T: int, D: int, accumulate: bool = False) -> None:
    call("b200_add_pos_bwd", dnet.data_ptr(), dpos.data_ptr(), B, T, D, int(accumulate), _stream())


def assemble_tokens(patch: Tensor, cls: Tensor, pos: Tensor, B: int, np_: int, D: int, conv_bias: Optional[Tensor] = None) -> Tensor:
    """net[b, 0] = cls + pos[0]; net[b, t] = float(patch[b, t-1]) + pos[t].  ``conv_bias`` (bf16 [D]): the patch rows first become
    bf16(patch + bias) -- the separate bf16 bias add eager's convolution performs."""
    net = torch.empty((B, np_ + 1, D), dtype=torch.float32, device=patch.device)
    call("b200_assemble_tokens", patch.data_ptr(), cls.data_ptr(), pos.data_ptr(), net.data_ptr(), B, np_, D, _ptr(conv_bias), _stream())
    return net


def assemble_tokens_bwd(dnet: Tensor, dpos: Tensor, dcls: Tensor, B: int, np_: int, D: int, accumulate: bool = False) -> Tensor:
    dpatch = torch.empty((B * np_, D), dtype=torch.bfloat16, device=dnet.device)
    call("b200_assemble_tokens_bwd", dnet.data_ptr(), dpatch.data_ptr(), dpos.data_ptr(), dcls.data_ptr(), B, np_, D, int(accumulate), _stream())
    return dpatch


def softmax_xent(logits: Tensor, labels: Tensor, *, grad_scale: float = 1.0, need_grad: bool = True,
                 grad_scale_dev: Optional[Tensor] = None) -> Tuple[Tensor, Tensor, Optional[Tensor], Tensor]:
    """Returns (loss_mean[1], loss_rows[B], dlogits bf16 [B,C] or None, bad_label_flag int32[1])."""
    _need_cuda(logits, labels)
    Bn, C = logits.shape
    if logits.dtype != torch.bfloat16 or logits.stride(1) != 1:
        raise B200Error("softmax_xent: logits must be bf16 with unit inner stride")
    if labels.dtype != torch.int64 or labels.numel() != Bn or not labels.is_contiguous():
        raise B200Error("softmax_xent: labels must be contiguous int64 with one entry per row")
    loss_rows = torch.empty(Bn, dtype=torch.float32, device=logits.device)
    loss_mean = torch.empty(1, dtype=torch.float32, device=logits.device)
    # dlogits shares the logits' row stride (a multiple of 8 elements so the head dgrad / wgrad can TMA it)
    dlogits = torch.empty_strided((Bn, C), (logits.stride(0), 1), dtype=torch.bfloat16, device=logits.device) if need_grad else None
    bad = torch.zeros(1, dtype=torch.int32, device=logits.device)
    call(
        "b200_softmax_xent_fwd_bwd", logits.data_ptr(), logits.stride(0), labels.data_ptr(), loss_rows.data_ptr(),
        loss_mean.data_ptr(), _ptr(dlogits), bad.data_ptr(), Bn, C, float(grad_scale), _ptr(grad_scale_dev), _stream(),
    )
    return loss_mean, loss_rows, dlogits, bad


def cast_bf16(src: Tensor, dst: Optional[Tensor] = None) -> Tensor:
    _need_cuda(src)
    if src.dtype != torch.float32 or not src.is_contiguous():
        raise B200Error("cast_bf16: need contiguous fp32")
    if dst is None:
        dst = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    call("b200_cast_f32_to_bf16", src.data_ptr(), dst.data_ptr(), src.numel(), _stream())
    return dst


def fill_f32(dst: Tensor, value: float) -> Tensor:
    call("b200_fill_f32", dst.data_ptr(), float(value), dst.numel(), _stream())
    return dst


# ----------------------------------------------------------------------------------------------------------------
# GroupNorm(32) + SiLU on channels-last activations (SD-v1.5 UNet)
# ----------------------------------------------------------------------------------------------------------------
def groupnorm_silu_fwd(x: Tensor, gamma: Tensor, beta: Tensor, eps: float, *, silu: bool = True) -> Tuple[Tensor, Tensor, Tensor]:
    """x: bf16 [B, HW, C] (channels last).  Returns (y bf16 [B, HW, C], mean f32 [B, 32], rstd f32 [B, 32])."""
    _need_cuda(x, gamma, beta)
    if x.dtype != torch.bfloat16 or x.dim() != 3 or not x.is_contiguous():
        raise B200Error("groupnorm_silu_fwd: need contiguous bf16 [B, HW, C]")
    B, HW, C = x.shape
    y = torch.empty_like(x)
    mean = torch.empty((B, 32), dtype=torch.float32, device=x.device)
    rstd = torch.empty((B, 32), dtype=torch.float32, device=x.device)
    ws = WORKSPACE.get(x.device, int(_cabi.lib().b200_groupnorm_workspace_floats(B, HW, C)), "groupnorm")
    call("b200_groupnorm_silu_fwd", x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
         ws.data_ptr(), B, HW, C, float(eps), int(silu), _stream())
    return y, mean, rstd


def groupnorm_silu_bwd(x: Tensor, dy: Tensor, gamma: Tensor, beta: Tensor, mean: Tensor, rstd: Tensor, *, silu: bool = True,
                       dgamma: Optional[Tensor] = None, dbeta: Optional[Tensor] = None, accumulate: bool = False) -> Tuple[Tensor, Tensor, Tensor]:
    _need_cuda(x, dy, gamma, beta, mean, rstd)
    B, HW, C = x.shape
    if dy.dtype != torch.bfloat16 or dy.shape != x.shape or not dy.is_contiguous():
        raise B200Error("groupnorm_silu_bwd: dy must be contiguous bf16 with x's shape")
    dx = torch.empty_like(x)
    if dgamma is None:
        dgamma = torch.empty(C, dtype=torch.float32, device=x.device)
        dbeta = torch.empty(C, dtype=torch.float32, device=x.device)
        accumulate = False
    ws = WORKSPACE.get(x.device, int(_cabi.lib().b200_groupnorm_workspace_floats(B, HW, C)), "groupnorm")
    call("b200_groupnorm_silu_bwd", x.data_ptr(), dy.data_ptr(), gamma.data_ptr(), beta.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
         dx.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), ws.data_ptr(), B, HW, C, int(silu), int(accumulate), _stream())
    return dx, dgamma, dbeta


def pack_conv3x3_weight(w: Tensor) -> Tensor:
    """[Cout, Cin, 3, 3] (the reference's Conv2d.weight) -> bf16 [Cout, 9 * Cin], tap-major (k = (ky * 3 + kx) * Cin + ci)."""
    co, ci, kh, kw = w.shape
    if kh != 3 or kw != 3:
        raise B200Error("pack_conv3x3_weight: 3x3 kernels only")
    return w.permute(0, 2, 3, 1).reshape(co, 9 * ci).to(torch.bfloat16).contiguous()


def pack_conv3x3_weight_dgrad(w: Tensor) -> Tensor:
    """Weights of the INPUT-gradient convolution: [Cin, 9 * Cout] with the taps flipped (dx = conv3x3(dy, flip(w)^T))."""
    return pack_conv3x3_weight(w.flip(2, 3).transpose(0, 1))


def conv3x3_wgrad(dy: Tensor, x: Tensor, out: Optional[Tensor] = None, *, round_bf16: bool = True, accumulate: bool = False) -> Tensor:
    """Weight gradient of ``conv3x3``: dy bf16 [B, H, W, Cout], x bf16 [B, H, W, Cin] -> fp32 [Cout, 9 * Cin] in the PACKED layout
    of ``pack_conv3x3_weight`` (``unpack_conv3x3_weight`` gives the reference's [Cout, Cin, 3, 3]).  One split-K implicit GEMM over
    the B * H * W pixels plus the split-K reduction; ``round_bf16`` rounds like eager's bf16 conv backward does."""
    _need_cuda(dy, x, out)
    if dy.dtype != torch.bfloat16 or x.dtype != torch.bfloat16 or dy.dim() != 4 or x.dim() != 4 or not dy.is_contiguous() or not x.is_contiguous():
        raise B200Error("conv3x3_wgrad: need contiguous bf16 dy [B, H, W, Cout] and x [B, H, W, Cin]")
    B, H, W, Cin = x.shape
    Cout = dy.shape[3]
    if tuple(dy.shape[:3]) != (B, H, W):
        raise B200Error("conv3x3_wgrad: dy and x disagree on [B, H, W]")
    if out is None:
        out = torch.empty((Cout, 9 * Cin), dtype=torch.float32, device=x.device)
        accumulate = False
    n = Cout * 9 * Cin
    splits = pick_splits(Cout, 9 * Cin, B * H * W)
    part = WORKSPACE.get(x.device, splits * n, "splitk")[: splits * n]
    call("b200_conv3x3_wgrad_nhwc_bf16", dy.data_ptr(), Cout, x.data_ptr(), part.data_ptr(), B, H, W, Cin, Cout, splits, _stream())
    call("b200_splitk_reduce", part.data_ptr(), splits, n, out.data_ptr(), int(round_bf16), int(accumulate), _stream())
    return out


def unpack_conv3x3_weight(w_packed: Tensor, cin: int) -> Tensor:
    """[Cout, 9 * Cin] (tap-major) -> [Cout, Cin, 3, 3], the reference's Conv2d.weight layout."""
    co = w_packed.shape[0]
    return w_packed.view(co, 3, 3, cin).permute(0, 3, 1, 2).contiguous()


def conv3x3(x: Tensor, w_packed: Tensor, bias: Optional[Tensor] = None, *, epilogue: int = EPI_BIAS_BF16, aux: Optional[Tensor] = None) -> Tensor:
    """x: bf16 [B, H, W, Cin] channels-last; w_packed: bf16 [Cout, 9 * Cin] -> bf16 (or fp32 with the residual epilogue) [B, H, W, Cout]."""
    _need_cuda(x, w_packed, bias, aux)
    if x.dtype != torch.bfloat16 or x.dim() != 4 or not x.is_contiguous() or w_packed.dtype != torch.bfloat16 or not w_packed.is_contiguous():
        raise B200Error("conv3x3: need contiguous bf16 x [B, H, W, Cin] and w_packed [Cout, 9 * Cin]")
    B, H, W, Cin = x.shape
    Cout = w_packed.shape[0]
    if w_packed.shape[1] != 9 * Cin or Cout % 8 != 0:
        raise B200Error("conv3x3: weight shape does not match (need [Cout, 9 * Cin], Cout % 8 == 0)")
    f32 = epilogue == EPI_BIAS_RESID_F32
    out = torch.empty((B, H, W, Cout), dtype=torch.float32 if f32 else torch.bfloat16, device=x.device)
    call("b200_conv3x3_nhwc_bf16", x.data_ptr(), w_packed.data_ptr(), _ptr(bias), out.data_ptr(), _ptr(aux), Cout, B, H, W, Cin, Cout, epilogue, _stream())
    return out
