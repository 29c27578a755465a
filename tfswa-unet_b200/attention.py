"""Host-side mirror of the reference's ``src/models/attention.py`` (same class names, constructor and
forward signatures, parameter/buffer tree), executing on the sm_100a kernels of libtfswa_b200.

``nn.LayerNorm`` / ``nn.Linear`` / ``nn.Dropout`` / ``nn.GELU`` sub-modules are kept as *parameter
containers* so ``state_dict()`` keys, ``optim.AdamW(model.parameters())`` and tools that walk the module
tree (reference gradient_checkpoint.py:44-69, quantization.py:53-68) behave as with the reference; their
``forward`` methods are never called.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import nn

from . import _lib as L
from . import engine as E
from . import functional as Fn
from .config import work_dtype

Tensor = torch.Tensor


def _no_dropout(p: float, training: bool) -> None:
    if p > 0.0 and training:
        raise NotImplementedError("dropout > 0 in training mode is not implemented by the B200 kernels "
                                  "(no reference caller sets it; refusing rather than falling back)")


class ScaledDotProductAttention(nn.Module):
    """Unused helper of the reference (attention.py:12-31; scale = dim**-0.5), kept for import compatibility."""

    def __init__(self, dim: int) -> None:
        super().__init__()
        self.scale = dim ** -0.5

    def forward(self, q: Tensor, k: Tensor, v: Tensor, mask: Optional[Tensor] = None):
        raise NotImplementedError("ScaledDotProductAttention is an unused helper in the reference; "
                                  "use MultiHeadAttention / TSA / FSA / ShiftedWindowAttention")


class MultiHeadAttention(nn.Module):
    """attention.py:34-90.  x: (R, N, C) -> (R, N, C); each of the R rows is one attention sequence."""

    def __init__(self, dim: int, num_heads: int, dropout: float = 0.0) -> None:
        super().__init__()
        assert dim % num_heads == 0, f"dim {dim} must be divisible by num_heads {num_heads}"
        self.dim = dim
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.scale = self.head_dim ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=False)
        self.proj = nn.Linear(dim, dim)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x: Tensor, mask: Optional[Tensor] = None, *, return_attn_weights: bool = False):
        if mask is not None:
            raise NotImplementedError("MultiHeadAttention(mask=...) is never used by the reference's callers")
        if return_attn_weights:
            raise NotImplementedError("return_attn_weights=True would materialise the NxN weights the kernels avoid")
        _no_dropout(self.dropout.p, self.training)
        R, N, C = x.shape
        dt = work_dtype()
        if not x.is_cuda:
            raise RuntimeError("tfswa_unet_b200 modules run on CUDA (sm_100a) tensors only - there is no CPU fallback")
        xt = x.to(dt).reshape(R * N, 1, C)
        qkv = Fn.linear(xt, Fn.LinW(self.qkv.weight[None], None))
        att = Fn.attention(qkv[:, 0, :], R, 1, N, C, self.num_heads, L.GEOM_FSA)      # rows = (b, h=0), sequence along W=N
        y = Fn.linear(att[:, None, :], Fn.LinW(self.proj.weight[None], self.proj.bias[None]))
        return y.reshape(R, N, C).to(x.dtype)


class _Branch(nn.Module):
    """Shared body of TSA / FSA / SW-MSA: x + MHA(LN1(x)), then x + MLP(LN2(x)) over a token grouping."""

    _geom = -1

    def _build(self, dim: int, num_heads: int, dropout: float, mlp_ratio: float) -> None:
        self.norm1 = nn.LayerNorm(dim)
        self.attn = MultiHeadAttention(dim, num_heads, dropout)
        self.norm2 = nn.LayerNorm(dim)
        hidden = int(dim * mlp_ratio)
        self.mlp = nn.Sequential(nn.Linear(dim, hidden), nn.GELU(), nn.Dropout(dropout), nn.Linear(hidden, dim), nn.Dropout(dropout))

    def _run(self, x: Tensor, ws: int = 8, shift: int = 0) -> Tensor:
        _no_dropout(self.attn.dropout.p, self.training)
        dt = work_dtype()
        native = E.is_native(x, dt)
        xi = x if native else E.to_native(x, dt)
        p = E.cached_prep(self, "branch", lambda: E.prep_single_branch(self), self.training)
        y = E.branch_forward(xi, p, self._geom, self.num_heads, ws=ws, shift=shift,
                             rel_bias=getattr(self, "rel_bias", None), use_shift_mask=getattr(self, "use_shift_mask", False))
        return y if native else E.from_native(y, x)


class TemporalSequenceAttention(_Branch):
    """attention.py:93-164.  Attention along dim 2 of (B,C,T,F), one sequence per (b, f).
    ``attn_chunk_size`` is kept as an attribute only: the reference's 16-row chunk loop (:147-153) is
    arithmetically a no-op and the streaming-softmax kernel never materialises the scores it was bounding."""
    _geom = L.GEOM_TSA

    def __init__(self, dim: int, num_heads: int = 8, dropout: float = 0.0, mlp_ratio: float = 4.0,
                 attn_chunk_size: Optional[int] = 16) -> None:
        super().__init__()
        self.dim, self.num_heads, self.attn_chunk_size = dim, num_heads, attn_chunk_size
        self._build(dim, num_heads, dropout, mlp_ratio)

    def forward(self, x: Tensor) -> Tensor:
        return self._run(x)


class FrequencySequenceAttention(_Branch):
    """attention.py:167-238.  Attention along dim 3 of (B,C,T,F), one sequence per (b, t)."""
    _geom = L.GEOM_FSA

    def __init__(self, dim: int, num_heads: int = 8, dropout: float = 0.0, mlp_ratio: float = 4.0,
                 attn_chunk_size: Optional[int] = 16) -> None:
        super().__init__()
        self.dim, self.num_heads, self.attn_chunk_size = dim, num_heads, attn_chunk_size
        self._build(dim, num_heads, dropout, mlp_ratio)

    def forward(self, x: Tensor) -> Tensor:
        return self._run(x)


def window_partition(x: Tensor, window_size: int) -> Tensor:
    """attention.py:241-256: (B,C,H,W) -> (B*nWin, ws, ws, C).  Layout utility (not on the kernel path:
    the attention kernel addresses windows in place); provided for API compatibility."""
    B, C, H, W = x.shape
    t = x.reshape(B, C, H // window_size, window_size, W // window_size, window_size)
    return t.permute(0, 2, 4, 3, 5, 1).reshape(-1, window_size, window_size, C)


def window_reverse(windows: Tensor, window_size: int, H: int, W: int) -> Tensor:
    """attention.py:259-277: inverse of :func:`window_partition`."""
    C = windows.shape[-1]
    B = windows.shape[0] // ((H // window_size) * (W // window_size))
    t = windows.reshape(B, H // window_size, W // window_size, window_size, window_size, C)
    return t.permute(0, 5, 1, 3, 2, 4).reshape(B, C, H, W)


def _reference_mask_buffer(ws: int, shift: int) -> Tensor:
    """The (64, ws*ws, ws*ws) {0,-100} buffer the reference registers for a fixed (8ws x 8ws) map
    (attention.py:318-343).  It never enters the reference's arithmetic (:380-382); kept for state_dict parity."""
    Hm = ws * 8
    edges = (0, Hm - ws, Hm - shift, Hm)
    img = torch.zeros(Hm, Hm)
    r = 0
    for i in range(3):
        for j in range(3):
            img[edges[i]:edges[i + 1], edges[j]:edges[j + 1]] = r
            r += 1
    ids = window_partition(img[None, None], ws).reshape(-1, ws * ws)
    diff = ids[:, None, :] - ids[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


class ShiftedWindowAttention(_Branch):
    """attention.py:280-403.  Reference behaviour (default): NO shift mask and NO relative-position bias.
    ``use_shift_mask=True`` / ``rel_bias`` (heads, ws*ws, ws*ws) enable the Swin features inside the kernel."""
    _geom = L.GEOM_SWA

    def __init__(self, dim: int, window_size: int, num_heads: int, shift_size: int = 0, dropout: float = 0.0,
                 mlp_ratio: float = 4.0) -> None:
        super().__init__()
        self.dim, self.window_size, self.num_heads, self.shift_size = dim, window_size, num_heads, shift_size
        self._build(dim, num_heads, dropout, mlp_ratio)
        self.use_shift_mask = False
        self.rel_bias = None
        if shift_size > 0:
            self.register_buffer("attn_mask", _reference_mask_buffer(window_size, shift_size))
        else:
            self.attn_mask = None

    def forward(self, x: Tensor) -> Tensor:
        return self._run(x, ws=self.window_size, shift=self.shift_size)
