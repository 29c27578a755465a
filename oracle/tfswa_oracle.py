"""Functional fp32/fp64 restatement of the TFSWA-UNet hot path (CPU oracle).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

This file restates, as pure functions over a ``{state_dict key: tensor}``
mapping, the arithmetic of the reference modules.  It is *not* a copy of the
reference: there are no ``nn.Module``s, the three attention branches share one
``_encoder_layer`` helper, and the token regrouping is written with einops-free
reshapes of an NHWC view.  Every function cites the reference lines it follows
(paths relative to ``/root/reference``).

Pinning: the reference's own tests hold **no** numerical fixtures for this
path, so the oracle is pinned against outputs of the
live reference generated in the build container by
``tests/golden/make_golden.py`` and committed under ``tests/golden/``
(``tests/test_oracle_golden.py`` checks them on every CPU run).

All arithmetic is plain ``torch`` on whatever dtype/device the inputs have
(use float64 on CPU for the tightest check).
"""
from __future__ import annotations

import math
from typing import Dict, Mapping, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Mapping[str, Tensor]

LN_EPS = 1e-5  # nn.LayerNorm default, attention.py:113,115
BN_EPS = 1e-5  # nn.BatchNorm2d default, blocks.py:55
BN_MOMENTUM = 0.1


def _sub(p: Params, prefix: str) -> Dict[str, Tensor]:
    """View of ``p`` restricted to keys under ``prefix`` (prefix stripped)."""
    n = len(prefix)
    return {k[n:]: v for k, v in p.items() if k.startswith(prefix)}


# --------------------------------------------------------------------------
# a1: MultiHeadAttention.forward            attention.py:50-90
# --------------------------------------------------------------------------
def mha(x: Tensor, p: Params, num_heads: int, bias: Optional[Tensor] = None) -> Tensor:
    """x: (R, N, C).  qkv has no bias (attention.py:46), proj has (attention.py:47).

    ``bias`` (optional, (R or 1, heads or 1, N, N)) is added to the scaled
    scores before the softmax; the reference never passes one
    (attention.py:380-382), it exists for the flag-gated mask/bias feature.
    """
    R, N, C = x.shape
    d = C // num_heads
    qkv = x @ p["qkv.weight"].t()                                  # attention.py:70
    qkv = qkv.view(R, N, 3, num_heads, d)
    q = qkv[:, :, 0].transpose(1, 2)                               # (R, h, N, d) attention.py:71-72
    k = qkv[:, :, 1].transpose(1, 2)
    v = qkv[:, :, 2].transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (d ** -0.5)                    # attention.py:44,75
    if bias is not None:
        s = s + bias
    w = torch.softmax(s, dim=-1)                                   # attention.py:80
    o = (w @ v).transpose(1, 2).reshape(R, N, C)                   # attention.py:84-85
    return o @ p["proj.weight"].t() + p["proj.bias"]               # attention.py:86


# --------------------------------------------------------------------------
# shared pre-LN transformer layer used by TSA / FSA / SWA
#   attention.py:146-159 (TSA), :220-233 (FSA), :378-387 (SWA)
# --------------------------------------------------------------------------
def _encoder_layer(tok: Tensor, p: Params, num_heads: int, bias: Optional[Tensor] = None) -> Tensor:
    C = tok.shape[-1]
    n1 = F.layer_norm(tok, (C,), p["norm1.weight"], p["norm1.bias"], LN_EPS)
    tok = tok + mha(n1, _sub(p, "attn."), num_heads, bias)
    n2 = F.layer_norm(tok, (C,), p["norm2.weight"], p["norm2.bias"], LN_EPS)
    h = F.gelu(n2 @ p["mlp.0.weight"].t() + p["mlp.0.bias"])       # exact erf GELU, attention.py:121-124
    return tok + (h @ p["mlp.3.weight"].t() + p["mlp.3.bias"])      # attention.py:126,159


# a2: TemporalSequenceAttention.forward     attention.py:130-164
def tsa(x: Tensor, p: Params, num_heads: int = 8) -> Tensor:
    """Attention along dim 2 of (B,C,H,W); one sequence per (b, w).  The 16-row
    chunk loop of attention.py:147-153 is arithmetically a no-op and is not restated."""
    B, C, H, W = x.shape
    tok = x.permute(0, 3, 2, 1).reshape(B * W, H, C)
    out = _encoder_layer(tok, p, num_heads)
    return out.view(B, W, H, C).permute(0, 3, 2, 1)


# a3: FrequencySequenceAttention.forward    attention.py:204-238
def fsa(x: Tensor, p: Params, num_heads: int = 8) -> Tensor:
    """Attention along dim 3 of (B,C,H,W); one sequence per (b, h)."""
    B, C, H, W = x.shape
    tok = x.permute(0, 2, 3, 1).reshape(B * H, W, C)
    out = _encoder_layer(tok, p, num_heads)
    return out.view(B, H, W, C).permute(0, 3, 1, 2)


# a4: window_partition / window_reverse     attention.py:241-277
def window_partition(x: Tensor, ws: int) -> Tensor:
    """(B,C,H,W) -> (B*nH*nW, ws, ws, C), windows ordered (b, wh, ww)."""
    B, C, H, W = x.shape
    t = x.reshape(B, C, H // ws, ws, W // ws, ws)
    return t.permute(0, 2, 4, 3, 5, 1).reshape(-1, ws, ws, C)


def window_reverse(win: Tensor, ws: int, H: int, W: int) -> Tensor:
    """Inverse of :func:`window_partition`."""
    C = win.shape[-1]
    B = win.shape[0] // ((H // ws) * (W // ws))
    t = win.reshape(B, H // ws, W // ws, ws, ws, C)
    return t.permute(0, 5, 1, 3, 2, 4).reshape(B, C, H, W)


def swin_shift_mask(Hp: int, Wp: int, ws: int, shift: int, dtype=torch.float32) -> Tensor:
    """Swin-style additive mask for an (Hp, Wp) padded map: (nWin, ws*ws, ws*ws) of {0,-100}.

    Independent oracle for the *flag-gated* mask feature (default OFF: the
    reference builds such a mask for a fixed 64x64 map at attention.py:318-343
    but never applies it, attention.py:380-382)."""
    img = torch.zeros(Hp, Wp, dtype=dtype)
    edges_h = (0, Hp - ws, Hp - shift, Hp)
    edges_w = (0, Wp - ws, Wp - shift, Wp)
    region = 0
    for i in range(3):
        for j in range(3):
            img[edges_h[i]:edges_h[i + 1], edges_w[j]:edges_w[j + 1]] = region
            region += 1
    ids = window_partition(img[None, None], ws).reshape(-1, ws * ws)
    diff = ids[:, None, :] - ids[:, :, None]
    return torch.where(diff != 0, torch.full_like(diff, -100.0), torch.zeros_like(diff))


def reference_attn_mask_buffer(ws: int, shift: int) -> Tensor:
    """The (64,64,64) ``attn_mask`` *buffer* the reference registers when shift>0
    (attention.py:318-343): a Swin mask for a fixed (8*ws, 8*ws) map.  Needed only
    for state_dict key/shape/value compatibility - it never enters the arithmetic."""
    return swin_shift_mask(ws * 8, ws * 8, ws, shift)


# a5: ShiftedWindowAttention.forward        attention.py:347-403
def swa(x: Tensor, p: Params, ws: int = 8, shift: int = 0, num_heads: int = 8,
        use_mask: bool = False, rel_bias: Optional[Tensor] = None) -> Tensor:
    """Zero-pad to a multiple of ``ws`` (pad tokens are *real* keys: LN(0)=beta),
    roll by (-shift,-shift), attend inside ws x ws windows with NO mask and NO
    relative-position bias (reference behaviour), un-roll, crop.

    ``use_mask`` / ``rel_bias`` ((heads, ws*ws, ws*ws)) are the optional,
    default-off Swin features requested by north_star."""
    B, C, H, W = x.shape
    ph, pw = (-H) % ws, (-W) % ws
    xp = F.pad(x, (0, pw, 0, ph))                                   # attention.py:358-365
    Hp, Wp = H + ph, W + pw
    if shift > 0:
        xp = torch.roll(xp, shifts=(-shift, -shift), dims=(2, 3))   # attention.py:368-371
    tok = window_partition(xp, ws).reshape(-1, ws * ws, C)          # attention.py:374-375
    bias = None
    if use_mask and shift > 0:
        m = swin_shift_mask(Hp, Wp, ws, shift, x.dtype).to(x.device)   # (nWin, N, N)
        bias = m.repeat(B, 1, 1)[:, None]                            # windows are (b, wh, ww)-ordered
    if rel_bias is not None:
        rb = rel_bias[None].to(x.dtype)
        bias = rb if bias is None else bias + rb
    out = _encoder_layer(tok, p, num_heads, bias)
    xp = window_reverse(out.reshape(-1, ws, ws, C), ws, Hp, Wp)     # attention.py:390-391
    if shift > 0:
        xp = torch.roll(xp, shifts=(shift, shift), dims=(2, 3))     # attention.py:394-397
    return xp[:, :, :H, :W]                                         # attention.py:400-401


# --------------------------------------------------------------------------
# BatchNorm2d as used everywhere on the path (eps 1e-5, momentum 0.1)
# --------------------------------------------------------------------------
def batch_norm(x: Tensor, p: Params, training: bool, new_stats: Optional[dict] = None,
               prefix: str = "") -> Tensor:
    """Eval: running stats.  Train: biased batch variance to normalise, unbiased for the
    running update (torch semantics); updated running stats are written to ``new_stats``."""
    w, b = p["weight"], p["bias"]
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if new_stats is not None:
            n = x.numel() // x.shape[1]
            new_stats[prefix + "running_mean"] = (1 - BN_MOMENTUM) * p["running_mean"] + BN_MOMENTUM * mean.detach()
            new_stats[prefix + "running_var"] = (1 - BN_MOMENTUM) * p["running_var"] + BN_MOMENTUM * var.detach() * n / max(n - 1, 1)
            new_stats[prefix + "num_batches_tracked"] = p["num_batches_tracked"] + 1
    else:
        mean, var = p["running_mean"], p["running_var"]
    scale = w * torch.rsqrt(var + BN_EPS)
    return x * scale[None, :, None, None] + (b - mean * scale)[None, :, None, None]


# a6: TFSWABlock.forward                     blocks.py:96-148
def tfswa_block(x: Tensor, p: Params, shift: int, ws: int = 8, num_heads: int = 8,
                skip: Optional[Tensor] = None, training: bool = False,
                new_stats: Optional[dict] = None, prefix: str = "",
                use_mask: bool = False) -> Tensor:
    identity = x                                                    # blocks.py:112
    y = F.conv2d(x, p["input_proj.0.weight"], p["input_proj.0.bias"])
    y = batch_norm(y, _sub(p, "input_proj.1."), training, new_stats, prefix + "input_proj.1.")  # blocks.py:115
    t = tsa(y, _sub(p, "tsa."), num_heads)                          # blocks.py:118-120
    f = fsa(y, _sub(p, "fsa."), num_heads)
    s = swa(y, _sub(p, "swa."), ws, shift, num_heads, use_mask=use_mask)
    z = F.conv2d(torch.cat([t, f, s], dim=1), p["fusion.0.weight"], p["fusion.0.bias"])  # blocks.py:123-126
    z = F.gelu(batch_norm(z, _sub(p, "fusion.1."), training, new_stats, prefix + "fusion.1."))
    z = z + identity                                                # blocks.py:129-131 (skip_proj is None: in==out)
    if skip is not None:
        if skip.shape != z.shape:
            # blocks.py:136-145 would resize / create a random conv per call; unreachable from
            # TFSWAUNet and not a defined function of the parameters -> refuse.
            raise ValueError("skip must have the block's output shape")
        z = z + skip                                                # blocks.py:146
    return z


# a7: DownsampleBlock                        blocks.py:156-163
def downsample(x: Tensor, p: Params, training: bool = False, new_stats=None, prefix: str = "") -> Tensor:
    y = F.conv2d(x, p["downsample.0.weight"], p["downsample.0.bias"], stride=2, padding=1)
    return F.gelu(batch_norm(y, _sub(p, "downsample.1."), training, new_stats, prefix + "downsample.1."))


# a8: UpsampleBlock                          blocks.py:171-178
def upsample(x: Tensor, p: Params, training: bool = False, new_stats=None, prefix: str = "") -> Tensor:
    y = F.conv_transpose2d(x, p["upsample.0.weight"], p["upsample.0.bias"], stride=2, padding=1)
    return F.gelu(batch_norm(y, _sub(p, "upsample.1."), training, new_stats, prefix + "upsample.1."))


# a11: stem / output head                    tfswa_unet.py:58-62, :139-145
def stem(x: Tensor, p: Params, training: bool = False, new_stats=None, prefix: str = "") -> Tensor:
    y = F.conv2d(x, p["0.weight"], p["0.bias"], padding=3)
    return F.gelu(batch_norm(y, _sub(p, "1."), training, new_stats, prefix + "1."))


def head_logits(x: Tensor, p: Params, training: bool = False, new_stats=None, prefix: str = "") -> Tensor:
    y = F.conv2d(x, p["0.weight"], p["0.bias"], padding=1)
    y = F.gelu(batch_norm(y, _sub(p, "1."), training, new_stats, prefix + "1."))
    return F.conv2d(y, p["3.weight"], p["3.bias"])


# a10: TFSWAUNet.forward                     tfswa_unet.py:164-229
def unet_forward(x: Tensor, p: Params, depths: Sequence[int] = (2, 2, 6, 2), ws: int = 8,
                 shift_size: int = 4, num_heads: int = 8, training: bool = False,
                 new_stats: Optional[dict] = None, return_logits: bool = False,
                 taps: Optional[dict] = None) -> Tensor:
    """(B,Cin,T,F) -> sigmoid masks (B,Cout,T,F).  ``taps`` (optional dict) receives
    intermediate tensors keyed by the module path that produced them."""
    def tap(name, t):
        if taps is not None:
            taps[name] = t.detach()
        return t

    def run_blocks(x, prefix, n, skip=None):
        for i in range(n):
            shift = 0 if i % 2 == 0 else shift_size               # tfswa_unet.py:73,96,123
            bp = f"{prefix}{i}."
            x = tfswa_block(x, _sub(p, bp), shift, ws, num_heads,
                            skip=skip if i == 0 else None,          # tfswa_unet.py:221-224
                            training=training, new_stats=new_stats, prefix=bp)
            tap(bp[:-1], x)
        return x

    x = tap("stem", stem(x, _sub(p, "stem."), training, new_stats, "stem."))          # :177
    skips = []
    n_enc = len(depths) - 1
    for s in range(n_enc):                                                              # :183-193
        x = run_blocks(x, f"encoder_stages.{s}.", depths[s])
        skips.append(x)
        dp = f"downsample_layers.{s}."
        x = tap(dp[:-1], downsample(x, _sub(p, dp), training, new_stats, dp))
    x = run_blocks(x, "bottleneck.", depths[-1])                                        # :196-197
    for j in range(n_enc):                                                              # :200-224
        up = f"upsample_layers.{j}."
        x = upsample(x, _sub(p, up), training, new_stats, up)
        skip = skips[n_enc - 1 - j]
        if x.shape[2:] != skip.shape[2:]:
            x = F.interpolate(x, size=skip.shape[2:], mode="bilinear", align_corners=False)  # :210-216
        tap(up[:-1], x)
        x = run_blocks(x, f"decoder_stages.{j}.", depths[n_enc - 1 - j], skip=skip)
    logits = tap("logits", head_logits(x, _sub(p, "output_head."), training, new_stats, "output_head."))
    return logits if return_logits else torch.sigmoid(logits)                           # :144,227


# --------------------------------------------------------------------------
# deterministic parameter recipes (shared by golden generation and tests)
# --------------------------------------------------------------------------
def randomize_state_(state: Dict[str, Tensor], seed: int, gain: float = 1.0) -> Dict[str, Tensor]:
    """Fill ``state`` in place (sorted key order, CPU generator) with a recipe that
    exercises every affine/bias term (SURVEY 8c oracle hygiene): non-trivial LN/BN
    affine, non-zero biases, non-identity BN running statistics, and weight scales
    that keep activations O(1) through 22 blocks in eval mode."""
    g = torch.Generator().manual_seed(seed)
    for k in sorted(state.keys()):
        t = state[k]
        if k.endswith("num_batches_tracked"):
            t.fill_(3)
            continue
        if k.endswith("attn_mask"):
            continue
        r = torch.randn(t.shape, generator=g, dtype=torch.float32)
        leaf = k.split(".")[-1]
        is_norm = (".norm1." in k or ".norm2." in k or
                   (t.dim() == 1 and (k.endswith(".1.weight") or k.endswith(".1.bias"))))
        if leaf == "running_mean":
            v = 0.1 * r
        elif leaf == "running_var":
            v = 0.75 + 0.5 * torch.rand(t.shape, generator=g)
        elif leaf == "weight" and t.dim() == 1:          # LN / BN gamma
            v = 1.0 + 0.2 * r
        elif leaf == "bias" and is_norm:                  # LN / BN beta
            v = 0.1 * r
        elif leaf == "bias":
            v = 0.05 * r
        elif leaf == "weight":
            if "upsample.0" in k:                         # ConvTranspose2d weight (Cin,Cout,kh,kw)
                fan_in = t.shape[0] * t.shape[2] * t.shape[3] / 4.0
            else:
                fan_in = t[0].numel()
            v = gain * r / math.sqrt(fan_in)
        else:
            v = r
        t.copy_(v.to(t.dtype))
    return state


def count_block_flops(B: int, C: int, H: int, W: int, ws: int = 8) -> int:
    """Forward FLOPs of one TFSWABlock (SURVEY 8d / BASELINE.md section 3)."""
    M = B * H * W
    Hp, Wp = H + (-H) % ws, W + (-W) % ws
    Mp = B * Hp * Wp
    return (2 * M * C * C + 24 * C * C * (2 * M + Mp)
            + 4 * C * (M * H + M * W + ws * ws * Mp) + 6 * M * C * C)


def count_model_flops(B: int, Cin: int, Cout: int, H: int, W: int,
                      depths=(2, 2, 6, 2), dims=(32, 64, 128, 256)) -> int:
    """Forward FLOPs of the whole TFSWAUNet (validated against FlopCounterMode in SURVEY A.2)."""
    total = 2 * B * H * W * Cin * dims[0] * 49
    sizes = [(H, W)]
    for _ in range(3):
        h, w = sizes[-1]
        sizes.append(((h - 2) // 2 + 1, (w - 2) // 2 + 1))
    for s in range(4):
        h, w = sizes[s]
        nblk = depths[s] * (2 if s < 3 else 1)
        total += nblk * count_block_flops(B, dims[s], h, w)
    for s in range(3):
        ho, wo = sizes[s + 1]
        total += 2 * B * ho * wo * dims[s] * dims[s + 1] * 16          # down
        total += 2 * B * (2 * ho) * (2 * wo) * dims[s + 1] * dims[s] * 4  # up (output = 2x input)
    total += 2 * B * H * W * (dims[0] * dims[0] * 9 + dims[0] * Cout)
    return total
