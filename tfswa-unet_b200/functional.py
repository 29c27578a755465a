"""Autograd-aware front of the C-ABI primitives.

Each primitive is one ``torch.autograd.Function`` whose forward and backward are C-ABI kernel calls
(``ops``); PyTorch only owns the graph bookkeeping and the memory.  Under ``torch.no_grad()`` (inference)
the raw ops are called directly.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _lib as L
from . import ops

Tensor = torch.Tensor

_KIND = {"conv3": 0, "down": 1, "up": 2}


def _needs_grad(*ts) -> bool:
    return torch.is_grad_enabled() and any(t is not None and torch.is_tensor(t) and t.requires_grad for t in ts)


def _no_backward(what: str):
    raise NotImplementedError(f"tfswa_unet_b200: backward of {what} is not available in this build")


# ------------------------------------------------------------------------------------------------
def row_stats(x: Tensor) -> Tensor:
    # statistics are treated as saved constants of the forward; the LN-hat backward in LinearFn
    # accounts for their dependence on x analytically.
    return ops.row_stats(x.detach())


class LinW:
    """Prepared weights of one (batched) token linear: fp32 master copy (nb,N,K) + bias (nb,N), and on demand the
    bf16 copy + per-output row sums the tcgen05 kernel consumes."""
    __slots__ = ("w", "b", "_tc")

    def __init__(self, w: Tensor, b: Optional[Tensor]):
        self.w = w.float().contiguous()
        self.b = None if b is None else b.float().contiguous()
        self._tc = None

    def tc(self):
        if self._tc is None:
            with torch.no_grad():
                wb = self.w.detach().to(torch.bfloat16).contiguous()
                self._tc = (wb, wb.float().sum(-1).contiguous())
        return self._tc


def _bf16_of(w: Tensor) -> Tensor:
    """bf16 copy of a prepared (cached, immutable) weight tensor.  The copy hangs on the tensor OBJECT (not on its
    address: the caching allocator hands a freed weight's address to the next rebuilt one) and is tied to its version."""
    # inference tensors (prepared under torch.inference_mode()) have no version counter and cannot be modified in place
    ver = -1 if torch.is_inference(w) else w._version
    hit = getattr(w, "_tfswa_bf16", None)
    if hit is None or hit[0] != ver:
        hit = (ver, w.detach().to(torch.bfloat16).contiguous())
        w._tfswa_bf16 = hit
    return hit[1]


USE_TC = True   # tcgen05 path for bf16 activations (set False to force the SIMT engine, e.g. for A/B tests)


def linear(x: Tensor, lw: LinW, *, prologue: int = 0, epilogue: int = 0,
           row_stats: Optional[Tensor] = None, r1: Optional[Tensor] = None, r2: Optional[Tensor] = None,
           want_col_stats: bool = False):
    w, bias = lw.w, lw.b
    if _needs_grad(x, w, bias, r1, r2):
        from .autograd import LinearFn
        return LinearFn.apply(x, w, bias, row_stats, r1, r2, prologue, epilogue, want_col_stats)
    if (USE_TC and x.dtype == torch.bfloat16 and not want_col_stats and prologue in (L.PRO_NONE, L.PRO_LNHAT)
            and w.shape[2] % 32 == 0 and w.shape[1] % 16 == 0):
        wb, wsum = lw.tc()
        return ops.linear_tc(x, wb, wsum, bias, prologue=prologue, epilogue=epilogue, row_stats=row_stats, r1=r1, r2=r2)
    if want_col_stats:
        stats = torch.zeros((2, w.shape[1]), dtype=torch.float32, device=x.device)
        if (USE_TC and x.dtype == torch.bfloat16 and prologue == L.PRO_NONE and x.shape[1] == 1
                and w.shape[2] % 32 == 0 and w.shape[1] % 16 == 0):
            pre = ops.linear_tc(x, lw.tc()[0], None, bias, col_stats=stats)
        else:
            pre = ops.linear(x, w, bias, prologue=prologue, row_stats=row_stats, col_stats=stats)
        return pre, stats
    return ops.linear(x, w, bias, prologue=prologue, epilogue=epilogue, row_stats=row_stats, r1=r1, r2=r2)


def gelu(u: Tensor) -> Tensor:
    """exact erf GELU of a dense token tensor (own autograd node when a gradient is needed)"""
    if _needs_grad(u):
        from .autograd import GeluFn
        return GeluFn.apply(u)
    return ops.affine_act(u, None, None, epilogue=L.EPI_GELU)


USE_FUSED_GELU_BWD = True   # training, bf16: GELU backward inside fc2's data-gradient GEMM (False: GeluFn + LinearFn, for A/B tests)


def gelu_linear(u: Tensor, lw: LinW, *, r1: Optional[Tensor] = None) -> Tensor:
    """GELU(u) W^T + b (+ r1): fc2 over the pre-activation u of fc1.  One autograd node whose backward fuses the GELU gradient
    into the data-gradient GEMM; without autograd (or with the switch off) the two-step form."""
    if USE_FUSED_GELU_BWD and USE_TC and u.dtype == torch.bfloat16 and _needs_grad(u, lw.w, lw.b, r1):
        from .autograd import GeluLinearFn
        return GeluLinearFn.apply(u, lw.w, lw.b, r1)
    return linear(gelu(u), lw, r1=r1)


USE_FUSED_TAIL = True   # one kernel for proj + residual + LN + MLP + residual at C in {32, 64, 128} (inference, bf16)
FUSED_TAIL_WIDTHS = (32, 64, 128)


def fused_tail_ok(att: Tensor, res: Tensor, proj: LinW, fc1: LinW, fc2: LinW) -> bool:
    C_ = att.shape[2]
    return (USE_TC and USE_FUSED_TAIL and att.dtype == torch.bfloat16 and C_ in FUSED_TAIL_WIDTHS and fc1.w.shape[1] == 4 * C_
            and proj.b is not None and fc1.b is not None and fc2.b is not None
            and not _needs_grad(att, res, proj.w, fc1.w, fc2.w))


def branch_tail(att: Tensor, res: Tensor, proj: LinW, fc1: LinW, fc2: LinW) -> Tensor:
    """res + proj(att) -> y;  y + fc2(GELU(fc1(LN_hat(y))))  in one kernel.  att (M,nb,C), res (M,1|nb,C)."""
    return ops.branch_tail_tc(att, res, proj.tc()[0], fc1.tc()[0], fc2.tc()[0], proj.b, fc1.b, fc2.b)


USE_FUSED_HEAD = True   # one kernel for input_proj + LN statistics + q|k|v GEMM at C in {32, 64} (inference, bf16)


def fused_head_ok(x: Tensor, inp: LinW, qkv: LinW) -> bool:
    C_ = x.shape[2]
    return (USE_TC and USE_FUSED_HEAD and x.dtype == torch.bfloat16 and x.shape[1] == 1 and C_ in (32, 64)
            and tuple(inp.w.shape) == (1, C_, C_) and tuple(qkv.w.shape) == (1, 9 * C_, C_)
            and inp.b is not None and qkv.b is not None and not _needs_grad(x, inp.w, qkv.w))


def block_head(x: Tensor, inp: LinW, qkv: LinW):
    """x1 = input_proj(x); qkv = qkv_linear(LN_hat(x1)) in one kernel -> (x1 (M,1,C), qkv (M,1,9C))."""
    return ops.block_head_tc(x, inp.tc()[0], qkv.tc()[0], inp.b, qkv.b)


def attention(qkv: Tensor, B: int, H: int, W: int, C: int, heads: int, geom: int, *, ws: int = 8, shift: int = 0,
              pad_kv: Optional[Tensor] = None, rel_bias: Optional[Tensor] = None, use_shift_mask: bool = False) -> Tensor:
    """qkv (M, 3C) view -> (M, C)."""
    if _needs_grad(qkv, pad_kv, rel_bias):
        from .autograd import AttentionFn
        return AttentionFn.apply(qkv, pad_kv, rel_bias, B, H, W, C, heads, geom, ws, shift, use_shift_mask)
    out = torch.empty((B * H * W, C), dtype=qkv.dtype, device=qkv.device)
    pk = None if pad_kv is None else pad_kv.contiguous()
    return ops.attention(qkv, out, B, H, W, C, heads, geom, ws=ws, shift=shift, pad_kv=pk, rel_bias=rel_bias,
                         use_shift_mask=use_shift_mask)


def attention3(qkv3: Tensor, B: int, H: int, W: int, C: int, heads: int, *, ws: int, shift: int, pad_kv: Tensor,
               rel_bias: Optional[Tensor] = None, use_shift_mask: bool = False) -> Tensor:
    """qkv3 (M, 3 branches, 3C) -> (M, 3, C): TSA, FSA and SW-MSA cores writing the concat buffer directly."""
    if _needs_grad(qkv3, pad_kv, rel_bias):
        from .autograd import Attention3Fn
        return Attention3Fn.apply(qkv3, pad_kv, rel_bias, B, H, W, C, heads, ws, shift, use_shift_mask)
    M = B * H * W
    att = torch.empty((M, 3, C), dtype=qkv3.dtype, device=qkv3.device)
    pk = pad_kv.contiguous()
    ops.attention(qkv3[:, 0, :], att[:, 0, :], B, H, W, C, heads, L.GEOM_TSA)
    ops.attention(qkv3[:, 1, :], att[:, 1, :], B, H, W, C, heads, L.GEOM_FSA)
    ops.attention(qkv3[:, 2, :], att[:, 2, :], B, H, W, C, heads, L.GEOM_SWA, ws=ws, shift=shift, pad_kv=pk,
                  rel_bias=rel_bias, use_shift_mask=use_shift_mask)
    return att


def bn_finalize(stats: Tensor, count: int, bn: nn.BatchNorm2d):
    """Train-mode BatchNorm2d: (scale, shift) from batch statistics + in-place running-stat update."""
    track = bn.track_running_stats and bn.running_mean is not None
    if track:
        bn.num_batches_tracked.add_(1)
        momentum = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked.item())
    else:
        momentum = 0.0
    if _needs_grad(stats, bn.weight, bn.bias):
        # per-channel bookkeeping on C-element vectors (host-side plumbing, differentiable): autograd carries the
        # gradient back to gamma/beta and, through `stats`, to the conv output (tfswa_act_bwd mode 1)
        mean = stats[0] / count
        var = (stats[1] / count - mean * mean).clamp_min(0.0)
        scale = bn.weight.float() * torch.rsqrt(var + bn.eps)
        shift = bn.bias.float() - mean * scale
        if track:
            with torch.no_grad():
                unbias = count / max(count - 1, 1)
                bn.running_mean.mul_(1 - momentum).add_(momentum * mean.detach())
                bn.running_var.mul_(1 - momentum).add_(momentum * unbias * var.detach())
        return scale, shift
    sc, sh, _ = ops.bn_finalize(stats, count, bn.weight.detach().float().contiguous(), bn.bias.detach().float().contiguous(),
                                bn.running_mean if track else None, bn.running_var if track else None, momentum, bn.eps)
    return sc, sh


def affine_act(v: Tensor, scale: Optional[Tensor], shift: Optional[Tensor], *, epilogue: int = 0,
               r1: Optional[Tensor] = None, r2: Optional[Tensor] = None) -> Tensor:
    if _needs_grad(v, scale, shift, r1, r2):
        from .autograd import AffineActFn
        return AffineActFn.apply(v, scale, shift, r1, r2, epilogue)
    return ops.affine_act(v, scale, shift, epilogue=epilogue, r1=r1, r2=r2)


def conv(x: Tensor, w_oihw: Tensor, wl: Tensor, b: Tensor, kind: str, out_hw, dtype: torch.dtype, *, epilogue: int = 0,
         want_col_stats: bool = False):
    """w_oihw: Cout-first (Cout,Cin,kh,kw) weight (autograd path); wl: the same weight in the kernel's layout."""
    if _needs_grad(x, w_oihw, b):
        from .autograd import ConvFn
        return ConvFn.apply(x, w_oihw, b, kind, tuple(out_hw), dtype, epilogue, want_col_stats)
    if (USE_TC and kind != "stem" and x.dtype == torch.bfloat16 and x.shape[1] % 32 == 0
            and b.shape[0] % 16 == 0 and b.shape[0] <= 256 and (not want_col_stats or epilogue == L.EPI_NONE)):
        if want_col_stats:
            stats = torch.zeros((2, b.shape[0]), dtype=torch.float32, device=x.device)
            return ops.conv_tc(x, _bf16_of(wl), b, _KIND[kind], out_hw, col_stats=stats), stats
        return ops.conv_tc(x, _bf16_of(wl), b, _KIND[kind], out_hw, epilogue=epilogue)
    stats = torch.zeros((2, b.shape[0]), dtype=torch.float32, device=x.device) if want_col_stats else None
    if kind == "stem":
        y = ops.stem(x, wl, b, dtype, epilogue=epilogue, col_stats=stats)
    else:
        y = ops.conv(x, wl, b, _KIND[kind], out_hw, epilogue=epilogue, col_stats=stats)
    return (y, stats) if want_col_stats else y


def head_tail(v: Tensor, w3: Tensor, b3: Tensor, scale: Optional[Tensor], shift: Optional[Tensor], want_logits: bool = False):
    if _needs_grad(v, w3, b3, scale, shift):
        from .autograd import HeadTailFn
        return HeadTailFn.apply(v, w3, b3, scale, shift, want_logits)
    return ops.head_tail(v, w3, b3, scale, shift, want_logits=want_logits)


def bilinear(x: Tensor, out_hw) -> Tensor:
    if _needs_grad(x):
        from .autograd import BilinearFn
        return BilinearFn.apply(x, tuple(out_hw))
    return ops.bilinear(x, out_hw)
