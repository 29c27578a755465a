"""Host-side mirror of the reference's ``src/models/blocks.py`` on the sm_100a kernels."""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import engine as E
from .attention import FrequencySequenceAttention, ShiftedWindowAttention, TemporalSequenceAttention
from .config import work_dtype

Tensor = torch.Tensor


class TFSWABlock(nn.Module):
    """blocks.py:16-148: input_proj (1x1 conv + BN) -> TSA | FSA | SW-MSA on the same tensor -> concat ->
    fusion (1x1 conv + BN + GELU) -> + identity (+ skip).  The class name and an assignable instance
    ``forward`` are part of the contract (reference gradient_checkpoint.py:44-69)."""

    def __init__(self, in_channels: int, out_channels: int, window_size: int, shift_size: int, num_heads: int,
                 dropout: float = 0.0, mlp_ratio: float = 4.0) -> None:
        super().__init__()
        if in_channels != out_channels:
            raise NotImplementedError("TFSWABlock with in_channels != out_channels (skip_proj, blocks.py:92-94) is never "
                                      "constructed by TFSWAUNet and is not implemented")
        self.in_channels, self.out_channels = in_channels, out_channels
        self.window_size, self.shift_size, self.num_heads = window_size, shift_size, num_heads
        self.input_proj = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=1), nn.BatchNorm2d(out_channels))
        self.tsa = TemporalSequenceAttention(dim=out_channels, num_heads=num_heads, dropout=dropout, mlp_ratio=mlp_ratio)
        self.fsa = FrequencySequenceAttention(dim=out_channels, num_heads=num_heads, dropout=dropout, mlp_ratio=mlp_ratio)
        self.swa = ShiftedWindowAttention(dim=out_channels, window_size=window_size, num_heads=num_heads,
                                          shift_size=shift_size, dropout=dropout, mlp_ratio=mlp_ratio)
        self.fusion = nn.Sequential(nn.Conv2d(out_channels * 3, out_channels, kernel_size=1), nn.BatchNorm2d(out_channels), nn.GELU())
        self.skip_proj = None

    def forward(self, x: Tensor, skip: Optional[Tensor] = None) -> Tensor:
        if self.training and self.tsa.attn.dropout.p > 0:
            raise NotImplementedError("dropout > 0 in training mode is not implemented by the B200 kernels")
        dt = work_dtype()
        native = E.is_native(x, dt)
        xi = x if native else E.to_native(x, dt)
        if skip is not None:
            if skip.shape != x.shape:
                # blocks.py:136-145 resizes and, on a channel mismatch, creates a randomly initialised conv per call;
                # unreachable from TFSWAUNet and not a function of the parameters -> refuse.
                raise ValueError("skip must have the block's output shape")
            skip = skip if E.is_native(skip, dt) else E.to_native(skip, dt)
        p = E.cached_prep(self, "block", lambda: E.prep_block(self, self.training), self.training)
        y = E.block_forward(self, xi, skip, p)
        return y if native else E.from_native(y, x)


class _Resample(nn.Module):
    _kind = ""
    _seq = ""

    def _run(self, x: Tensor, out_hw) -> Tensor:
        dt = work_dtype()
        native = E.is_native(x, dt)
        xi = x if native else E.to_native(x, dt)
        seq = getattr(self, self._seq)
        prep = E.cached_prep(self, "conv", lambda: E.prep_conv_bn(seq[0], seq[1], self._kind, self.training), self.training)
        y = E.conv_bn_gelu(xi, seq[1], self._kind, prep, self.training, out_hw, dt)
        return y if native else E.from_native(y, x)


class DownsampleBlock(_Resample):
    """blocks.py:151-163: Conv2d(k=4, s=2, p=1) + BN + GELU."""
    _kind, _seq = "down", "downsample"

    def __init__(self, in_channels: int, out_channels: int) -> None:
        super().__init__()
        self.downsample = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=4, stride=2, padding=1),
                                        nn.BatchNorm2d(out_channels), nn.GELU())

    def forward(self, x: Tensor) -> Tensor:
        H, W = x.shape[2:]
        return self._run(x, ((H - 2) // 2 + 1, (W - 2) // 2 + 1))


class UpsampleBlock(_Resample):
    """blocks.py:166-178: ConvTranspose2d(k=4, s=2, p=1) + BN + GELU (output exactly 2x the input)."""
    _kind, _seq = "up", "upsample"

    def __init__(self, in_channels: int, out_channels: int) -> None:
        super().__init__()
        self.upsample = nn.Sequential(nn.ConvTranspose2d(in_channels, out_channels, kernel_size=4, stride=2, padding=1),
                                      nn.BatchNorm2d(out_channels), nn.GELU())

    def forward(self, x: Tensor) -> Tensor:
        H, W = x.shape[2:]
        return self._run(x, (2 * H, 2 * W))
