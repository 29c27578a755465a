"""Algorithmic work model of the path: the FLOP counts `bench.py` and `tools/block_bench.py` divide by
measured time.  Counts multiply-adds of the reference's op sequence as 2 FLOPs (attention.py:70-86, :121-128,
blocks.py:53-56, :85-89, :156-175, tfswa_unet.py:58-62, :139-145); validated against torch's FlopCounterMode on the
reference."""
from __future__ import annotations


def block_flops(B: int, C: int, H: int, W: int, ws: int = 8) -> int:
    """forward FLOPs of one TFSWABlock: input_proj + 3 x [qkv 6 + proj 2 + MLP 16] tokens C^2 + attention + fusion conv"""
    M = B * H * W
    Hp, Wp = H + (-H) % ws, W + (-W) % ws
    Mp = B * Hp * Wp
    return (2 * M * C * C + 24 * C * C * (2 * M + Mp) + 4 * C * (M * H + M * W + ws * ws * Mp) + 6 * M * C * C)


def model_flops(B: int, Cin: int, Cout: int, H: int, W: int, depths=(2, 2, 6, 2), dims=(32, 64, 128, 256)) -> int:
    """forward FLOPs of the whole TFSWAUNet"""
    total = 2 * B * H * W * Cin * dims[0] * 49                                # stem 7x7
    sizes = [(H, W)]
    for _ in range(3):
        h, w = sizes[-1]
        sizes.append(((h - 2) // 2 + 1, (w - 2) // 2 + 1))
    for s in range(4):
        h, w = sizes[s]
        total += depths[s] * (2 if s < 3 else 1) * block_flops(B, dims[s], h, w)
    for s in range(3):
        ho, wo = sizes[s + 1]
        total += 2 * B * ho * wo * dims[s] * dims[s + 1] * 16                 # down 4x4 / stride 2
        total += 2 * B * (2 * ho) * (2 * wo) * dims[s + 1] * dims[s] * 4      # up (4 taps per output pixel)
    total += 2 * B * H * W * (dims[0] * dims[0] * 9 + dims[0] * Cout)         # head 3x3 + 1x1
    return total
