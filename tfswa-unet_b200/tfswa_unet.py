"""Host-side mirror of the reference's ``src/models/tfswa_unet.py``: same constructor, attribute names
(hence the same 1103 ``state_dict`` keys), ``forward`` contract ((B,Cin,T,F) fp32 -> sigmoid masks
(B,Cout,T,F) fp32), ``get_num_parameters`` and ``get_model_info``."""
from __future__ import annotations

from typing import List

import torch
from torch import nn

from . import _lib as L
from . import engine as E
from . import functional as Fn
from .blocks import DownsampleBlock, TFSWABlock, UpsampleBlock
from .config import work_dtype

Tensor = torch.Tensor


class TFSWAUNet(nn.Module):
    """tfswa_unet.py:11-245."""

    def __init__(self, in_channels: int, out_channels: int, depths: List[int], dims: List[int], window_size: int,
                 shift_size: int, num_heads: int, dropout: float = 0.0, mlp_ratio: float = 4.0) -> None:
        super().__init__()
        assert len(depths) == len(dims), "depths and dims must have the same length"
        assert len(depths) == 4, "Expected 4 stages (3 encoder + 1 bottleneck)"
        self.in_channels, self.out_channels = in_channels, out_channels
        self.depths, self.dims = depths, dims
        self.num_stages = len(depths)

        def stage(dim: int, n: int) -> nn.ModuleList:
            # even blocks: W-MSA (shift 0); odd blocks: SW-MSA (shift_size)      tfswa_unet.py:73,96,123
            return nn.ModuleList([TFSWABlock(dim, dim, window_size, 0 if i % 2 == 0 else shift_size, num_heads,
                                             dropout=dropout, mlp_ratio=mlp_ratio) for i in range(n)])

        self.stem = nn.Sequential(nn.Conv2d(in_channels, dims[0], kernel_size=7, stride=1, padding=3),
                                  nn.BatchNorm2d(dims[0]), nn.GELU())
        self.encoder_stages = nn.ModuleList()
        self.downsample_layers = nn.ModuleList()
        for s in range(self.num_stages - 1):
            self.encoder_stages.append(stage(dims[s], depths[s]))
            self.downsample_layers.append(DownsampleBlock(dims[s], dims[s + 1]))
        self.bottleneck = stage(dims[-1], depths[-1])
        self.upsample_layers = nn.ModuleList()
        self.decoder_stages = nn.ModuleList()
        for s in range(self.num_stages - 2, -1, -1):
            self.upsample_layers.append(UpsampleBlock(dims[s + 1], dims[s]))
            self.decoder_stages.append(stage(dims[s], depths[s]))
        self.output_head = nn.Sequential(nn.Conv2d(dims[0], dims[0], kernel_size=3, padding=1), nn.BatchNorm2d(dims[0]),
                                         nn.GELU(), nn.Conv2d(dims[0], out_channels, kernel_size=1), nn.Sigmoid())
        self._init_weights()

    def _init_weights(self) -> None:
        """tfswa_unet.py:149-162."""
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.Linear):
                nn.init.trunc_normal_(m.weight, std=0.02)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    # ------------------------------------------------------------------------------------------
    def forward(self, x: Tensor, *, return_logits: bool = False) -> Tensor:
        """tfswa_unet.py:164-229.  ``return_logits`` (extension) also returns the pre-sigmoid logits."""
        if not x.is_cuda:
            raise RuntimeError("tfswa_unet_b200.TFSWAUNet runs on CUDA (sm_100a) tensors only - there is no CPU fallback")
        dt = work_dtype()
        training = self.training
        xin = x.float().contiguous()
        # stem: NCHW fp32 -> native NHWC activations
        prep = E.cached_prep(self.stem, "conv", lambda: E.prep_conv_bn(self.stem[0], self.stem[1], "stem", training), training)
        h = E.conv_bn_gelu(xin, self.stem[1], "stem", prep, training, tuple(x.shape[2:]), dt)
        skips = []
        for blocks, down in zip(self.encoder_stages, self.downsample_layers):
            for blk in blocks:
                h = blk(h)
            skips.append(h)
            h = down(h)
        for blk in self.bottleneck:
            h = blk(h)
        for j, (up, blocks) in enumerate(zip(self.upsample_layers, self.decoder_stages)):
            h = up(h)
            skip = skips[-(j + 1)]
            if h.shape[2:] != skip.shape[2:]:
                h = Fn.bilinear(h, tuple(skip.shape[2:]))                       # tfswa_unet.py:210-216
            for i, blk in enumerate(blocks):
                h = blk(h, skip=skip) if i == 0 else blk(h)                      # tfswa_unet.py:221-224
        # head: conv3x3 (+BN folded in eval) -> GELU -> conv1x1 -> sigmoid, written as NCHW fp32
        conv0, bn, conv3 = self.output_head[0], self.output_head[1], self.output_head[3]
        w, wl, b = E.cached_prep(self.output_head, "conv", lambda: E.prep_conv_bn(conv0, bn, "conv3", training), training)
        w3 = conv3.weight.reshape(self.out_channels, -1).float().contiguous()
        b3 = conv3.bias.float().contiguous()
        if training:
            pre, stats = Fn.conv(h, w, wl, b, "conv3", tuple(h.shape[2:]), dt, want_col_stats=True)
            sc, sh = E._bn_train(pre.shape[0] * pre.shape[2] * pre.shape[3], stats, bn)
            out = Fn.head_tail(pre, w3, b3, sc, sh, want_logits=return_logits)
        else:
            v = Fn.conv(h, w, wl, b, "conv3", tuple(h.shape[2:]), dt)
            out = Fn.head_tail(v, w3, b3, None, None, want_logits=return_logits)
        return out

    def get_num_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def get_model_info(self) -> dict:
        return {"architecture": "TFSWA-UNet", "in_channels": self.in_channels, "out_channels": self.out_channels,
                "depths": self.depths, "dims": self.dims, "num_parameters": self.get_num_parameters(),
                "num_stages": self.num_stages}
