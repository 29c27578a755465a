"""state_dict layout helpers: the parameter/buffer tree of each mirrored module."""
from __future__ import annotations

from typing import Dict

import torch


def empty_state(kind: str, C: int, shift: int = 0, cout: int = 0, cin: int = 2, gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """A fresh CPU ``state_dict`` (cloned tensors, reference key order) of the mirrored module ``kind``."""
    from .attention import FrequencySequenceAttention, ShiftedWindowAttention, TemporalSequenceAttention
    from .blocks import DownsampleBlock, TFSWABlock, UpsampleBlock
    from .tfswa_unet import TFSWAUNet
    if kind == "tsa":
        m = TemporalSequenceAttention(C, 8)
    elif kind == "fsa":
        m = FrequencySequenceAttention(C, 8)
    elif kind == "swa":
        m = ShiftedWindowAttention(C, 8, 8, shift)
    elif kind == "block":
        m = TFSWABlock(C, C, 8, shift, 8)
    elif kind == "down":
        m = DownsampleBlock(C, cout)
    elif kind == "up":
        m = UpsampleBlock(C, cout)
    elif kind == "unet":
        m = TFSWAUNet(cin, cout or 2, [2, 2, 6, 2], [32, 64, 128, 256], 8, 4, 8)
    else:
        raise ValueError(kind)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}
