"""Composition of the C-ABI kernels into the TFSWA-UNet hot path (host-side orchestration only).

Algebraic restructuring relative to the reference's eager op sequence (all exact up to rounding):
  * activations stay NHWC, so TSA/FSA permutes (attention.py:143,162,217,236) and the SW-MSA
    pad/roll/partition chain (attention.py:358-375,390-401) become index maps inside the attention kernel;
  * LayerNorm is split into its statistics (row_stats) and its affine part, and the affine part is folded
    into the following Linear (W' = W diag(gamma), b' = W beta + b).  The three branches normalise the SAME
    tensor with different (gamma, beta), so one (M,C)x(C,9C) GEMM produces q|k|v of all three branches;
  * a zero-padded SW-MSA token has LN(0) = beta, i.e. its k|v equal the folded qkv bias (pad_kv);
  * eval-mode BatchNorm is folded into the preceding conv; train-mode BatchNorm runs as
    conv(+column sums) -> bn_finalize -> affine_act;
  * torch.cat([tsa, fsa, swa]) (blocks.py:123) is never materialised as a copy: the three branch outputs
    are written straight into one (M, 3C) buffer that the fusion GEMM reads with K = 3C.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from . import _lib as L
from . import functional as Fn

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------
# native layout helpers
# ----------------------------------------------------------------------------------------------
def is_native(x: Tensor, dtype: torch.dtype) -> bool:
    return (x.is_cuda and x.dim() == 4 and x.dtype == dtype and x.stride(1) == 1
            and x.is_contiguous(memory_format=torch.channels_last))


def to_native(x: Tensor, dtype: torch.dtype) -> Tensor:
    """Entry conversion for sub-module drop-in use (a bare TSA(x) on an NCHW fp32 tensor).  Host plumbing:
    whole-model calls never take this path (stem/head kernels convert in flight)."""
    if not x.is_cuda:
        raise RuntimeError("tfswa_unet_b200 modules run on CUDA (sm_100a) tensors only - there is no CPU fallback")
    return x.to(dtype).contiguous(memory_format=torch.channels_last)


def from_native(y: Tensor, like: Tensor) -> Tensor:
    return y.to(like.dtype).contiguous()


def tokens(x: Tensor) -> Tensor:
    """native (B,C,H,W) -> (M, C) view."""
    B, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(B * H * W, C)


def untokens(t: Tensor, B: int, H: int, W: int) -> Tensor:
    return t.reshape(B, H, W, -1).permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------------------------
# weight preparation (differentiable torch ops on tiny tensors; cached per parameter version for inference)
# ----------------------------------------------------------------------------------------------
def _bn_fold(w2d: Tensor, b: Tensor, bn: nn.BatchNorm2d):
    s = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
    return w2d * s[:, None], (b - bn.running_mean) * s + bn.bias


def _versions(mod: nn.Module):
    return tuple((id(t), t._version) for t in list(mod.parameters()) + list(mod.buffers()))


def cached_prep(mod: nn.Module, name: str, builder, training: bool):
    """Cache prepared weights while no gradient is required; rebuild when any parameter/buffer changed."""
    grad = torch.is_grad_enabled() and any(p.requires_grad for p in mod.parameters())
    if grad:
        return builder()
    dev = next(mod.parameters()).device
    key = (name, training, dev, _versions(mod))
    cache = mod.__dict__.setdefault("_tfswa_prep", {})
    hit = cache.get(name)
    if hit is not None and hit[0] == key:
        return hit[1]
    with torch.no_grad():
        val = builder()
    cache[name] = (key, val)
    return val


def prep_branch(br: nn.Module):
    """LN affine folded into qkv / fc1 (see module docstring)."""
    g1, b1 = br.norm1.weight, br.norm1.bias
    g2, b2 = br.norm2.weight, br.norm2.bias
    wqkv = br.attn.qkv.weight
    w1, bb1 = br.mlp[0].weight, br.mlp[0].bias
    return {
        "wqkv": wqkv * g1[None, :], "bqkv": wqkv @ b1,
        "wp": br.attn.proj.weight, "bp": br.attn.proj.bias,
        "w1": w1 * g2[None, :], "b1": w1 @ b2 + bb1,
        "w2": br.mlp[3].weight, "b2": br.mlp[3].bias,
    }


def prep_block(blk: nn.Module, training: bool):
    C = blk.out_channels
    brs = [prep_branch(b) for b in (blk.tsa, blk.fsa, blk.swa)]
    wi, bi = blk.input_proj[0].weight.reshape(C, -1), blk.input_proj[0].bias
    wf, bf = blk.fusion[0].weight.reshape(C, -1), blk.fusion[0].bias
    if not training:
        wi, bi = _bn_fold(wi, bi, blk.input_proj[1])
        wf, bf = _bn_fold(wf, bf, blk.fusion[1])
    st = lambda k: torch.stack([b[k] for b in brs])
    return {
        "in": Fn.LinW(wi[None], bi[None]),
        "qkv": Fn.LinW(torch.cat([b["wqkv"] for b in brs], 0)[None], torch.cat([b["bqkv"] for b in brs], 0)[None]),
        "proj": Fn.LinW(st("wp"), st("bp")),
        "fc1": Fn.LinW(st("w1"), st("b1")),
        "fc2": Fn.LinW(st("w2"), st("b2")),
        "fuse": Fn.LinW(wf[None], bf[None]),
    }


def prep_single_branch(br: nn.Module):
    p = prep_branch(br)
    return {"qkv": Fn.LinW(p["wqkv"][None], p["bqkv"][None]), "proj": Fn.LinW(p["wp"][None], p["bp"][None]),
            "fc1": Fn.LinW(p["w1"][None], p["b1"][None]), "fc2": Fn.LinW(p["w2"][None], p["b2"][None])}


def _mlp(y: Tensor, fc1, fc2) -> Tensor:
    """y + fc2(GELU(fc1(LN_hat(y))))  (attention.py:121-128,159).  Without autograd the GELU runs in fc1's epilogue
    (hidden stored post-activation); with autograd the pre-activation is what is stored, and fc2 applies GELU on load."""
    st2 = Fn.row_stats(y)
    grad = torch.is_grad_enabled() and (y.requires_grad or fc1.w.requires_grad)
    if grad:
        u = Fn.linear(y, fc1, prologue=L.PRO_LNHAT, row_stats=st2)
        if Fn.USE_TC and y.dtype == torch.bfloat16:
            # bf16: GELU as its own node so that fc2 (forward, dgrad, wgrad) stays on the tensor-core kernels
            return Fn.gelu_linear(u, fc2, r1=y)
        return Fn.linear(u, fc2, prologue=L.PRO_GELU, r1=y)
    h = Fn.linear(y, fc1, prologue=L.PRO_LNHAT, epilogue=L.EPI_GELU, row_stats=st2)
    return Fn.linear(h, fc2, r1=y)


def _tail(att: Tensor, res: Tensor, p: dict) -> Tensor:
    """res + proj(att), then the MLP with its residual (attention.py:86,146,159).  Narrow stages run it as one fused
    tensor-core kernel at inference; otherwise (and under autograd) it is three linears + a statistics pass."""
    if Fn.fused_tail_ok(att, res, p["proj"], p["fc1"], p["fc2"]):
        return Fn.branch_tail(att, res, p["proj"], p["fc1"], p["fc2"])
    y = Fn.linear(att, p["proj"], r1=res)
    return _mlp(y, p["fc1"], p["fc2"])


def _bn_train(count: int, stats: Tensor, bn: nn.BatchNorm2d):
    """Train-mode BatchNorm2d bookkeeping on accumulated column sums -> per-channel (scale, shift)."""
    return Fn.bn_finalize(stats, count, bn)


# ----------------------------------------------------------------------------------------------
# one pre-LN transformer branch on tokens (TSA / FSA / SW-MSA differ only in `geom`)
#   attention.py:146-159 / :220-233 / :378-387
# ----------------------------------------------------------------------------------------------
def branch_forward(x: Tensor, p: dict, geom: int, heads: int, ws: int = 8, shift: int = 0,
                   rel_bias: Optional[Tensor] = None, use_shift_mask: bool = False) -> Tensor:
    B, C, H, W = x.shape
    M = B * H * W
    xt = tokens(x)[:, None, :]                                              # (M,1,C)
    st1 = Fn.row_stats(xt)
    qkv = Fn.linear(xt, p["qkv"], prologue=L.PRO_LNHAT, row_stats=st1)                  # (M,1,3C)
    att = Fn.attention(qkv[:, 0, :], B, H, W, C, heads, geom, ws=ws, shift=shift,
                       pad_kv=p["qkv"].b[0, C:], rel_bias=rel_bias, use_shift_mask=use_shift_mask)   # (M,C)
    z = _tail(att[:, None, :], xt, p)
    return untokens(z[:, 0, :], B, H, W)


# ----------------------------------------------------------------------------------------------
# TFSWABlock (blocks.py:96-148)
# ----------------------------------------------------------------------------------------------
def block_forward(blk: nn.Module, x: Tensor, skip: Optional[Tensor], p: dict) -> Tensor:
    B, C, H, W = x.shape
    M = B * H * W
    training = blk.training
    xt = tokens(x)[:, None, :]
    # input_proj: 1x1 conv + BN (no activation)                                   blocks.py:53-56,115
    if not training and Fn.fused_head_ok(xt, p["in"], p["qkv"]):
        x1, qkv = Fn.block_head(xt, p["in"], p["qkv"])                                   # one kernel at C = 32 / 64
    else:
        if training:
            pre, stats = Fn.linear(xt, p["in"], want_col_stats=True)
            sc, sh = _bn_train(M, stats, blk.input_proj[1])
            x1 = Fn.affine_act(pre, sc, sh)
        else:
            x1 = Fn.linear(xt, p["in"])
        # LN statistics once, q|k|v of all three branches in one GEMM
        st1 = Fn.row_stats(x1)
        qkv = Fn.linear(x1, p["qkv"], prologue=L.PRO_LNHAT, row_stats=st1)              # (M,1,9C)
    qkv3 = qkv.view(M, 3, 3 * C)
    b9 = p["qkv"].b.view(3, 3 * C)
    att = Fn.attention3(qkv3, B, H, W, C, blk.num_heads, ws=blk.window_size, shift=blk.shift_size,
                        pad_kv=b9[2, C:], use_shift_mask=getattr(blk.swa, "use_shift_mask", False),
                        rel_bias=getattr(blk.swa, "rel_bias", None))                      # (M,3,C)
    z = _tail(att, x1, p)                                                                # (M,3,C) == cat along C
    zc = z.view(M, 1, 3 * C)
    skt = None if skip is None else tokens(skip)[:, None, :]
    # fusion: 1x1 conv (3C->C) + BN + GELU, + identity (+ skip)                    blocks.py:85-89,123-146
    if training:
        pre, stats = Fn.linear(zc, p["fuse"], want_col_stats=True)
        sc, sh = _bn_train(M, stats, blk.fusion[1])
        out = Fn.affine_act(pre, sc, sh, epilogue=L.EPI_GELU, r1=xt, r2=skt)
    else:
        out = Fn.linear(zc, p["fuse"], epilogue=L.EPI_GELU, r1=xt, r2=skt)
    return untokens(out[:, 0, :], B, H, W)


# ----------------------------------------------------------------------------------------------
# stem / down / up / head   (tfswa_unet.py:58-62,139-145; blocks.py:156-160,171-175)
# ----------------------------------------------------------------------------------------------
def prep_conv_bn(conv: nn.Module, bn: nn.BatchNorm2d, kind: str, training: bool):
    """-> (w_oihw Cout-first, wl kernel layout, bias); eval mode folds the BatchNorm into weight and bias."""
    from .autograd import conv_layout
    w, b = conv.weight, conv.bias
    if kind == "up":            # ConvTranspose2d weight (Cin, Cout, 4, 4) -> Cout-first
        w = w.permute(1, 0, 2, 3)
    if not training:
        s = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
        w = w * s[:, None, None, None]
        b = (b - bn.running_mean) * s + bn.bias
    w = w.float()
    return w, conv_layout(w.detach(), kind), b.float().contiguous()


def conv_bn_gelu(x: Tensor, bn: nn.BatchNorm2d, kind: str, prep, training: bool, out_hw, dtype: torch.dtype) -> Tensor:
    w, wl, b = prep
    if training:
        pre, stats = Fn.conv(x, w, wl, b, kind, out_hw, dtype, want_col_stats=True)
        sc, sh = _bn_train(pre.shape[0] * pre.shape[2] * pre.shape[3], stats, bn)
        return Fn.affine_act(pre, sc, sh, epilogue=L.EPI_GELU)
    return Fn.conv(x, w, wl, b, kind, out_hw, dtype, epilogue=L.EPI_GELU)
