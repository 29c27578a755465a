"""GPU (-m gpu): SW-MSA window attention (tfswa_attn_win_tc_fwd: tcgen05 + TMA for the interior windows at head_dim 4 / 8,
warp-level MMAs for the pad / wrap-around fringe and head_dim 16 / 32) against the fp32-math SIMT kernel on the same bf16
q|k|v AND against a torch fp32 restatement of attention.py:358-401, for every head_dim of the model, padded windows and
cyclic shift."""
import pytest
import torch

from helpers import seeded

pytestmark = pytest.mark.gpu


def _ref_windows(qkv, B, H, W, C, heads, ws, shift, pad_kv):
    """torch fp32 restatement of attention.py:358-401 on q|k|v tokens: zero-padded tokens are real keys whose k|v are
    `pad_kv` (= the folded qkv bias of a LayerNorm'ed zero token), roll(-shift), 8x8 windows, softmax(QK^T/sqrt(d)) V,
    reverse, roll back, crop."""
    d = C // heads
    Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
    t = qkv.float().view(B, H, W, 3 * C)
    full = torch.zeros((B, Hp, Wp, 3 * C), device=qkv.device)
    full[..., C:] = pad_kv.view(1, 1, 1, 2 * C)
    full[:, :H, :W] = t
    if shift:
        full = torch.roll(full, shifts=(-shift, -shift), dims=(1, 2))
    win = full.view(B, Hp // ws, ws, Wp // ws, ws, 3 * C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, 3, heads, d)
    q, k, v = win[:, :, 0].transpose(1, 2), win[:, :, 1].transpose(1, 2), win[:, :, 2].transpose(1, 2)     # (nW, h, 64, d)
    o = torch.softmax((q @ k.transpose(-1, -2)) * d ** -0.5, -1) @ v
    o = o.transpose(1, 2).reshape(B, Hp // ws, Wp // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, C)
    if shift:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    return o[:, :H, :W].reshape(B * H * W, C)


@pytest.mark.parametrize("B,H,W,C,shift", [
    (1, 8, 8, 32, 0), (1, 8, 8, 32, 4),               # one window
    (2, 13, 21, 32, 4), (1, 17, 9, 32, 0),            # padded windows (pad tokens are real keys = pad_kv)
    (1, 24, 16, 64, 4), (2, 11, 30, 64, 0),           # head_dim 8
    (1, 16, 24, 128, 4), (1, 9, 17, 128, 0),          # head_dim 16, two channel slabs
    (1, 16, 8, 256, 4), (2, 13, 9, 256, 0),           # head_dim 32
    (1, 40, 16, 64, 4),                               # heads = 16 -> head_dim 4 with a 64-channel slab
    # shapes with many interior windows (tcgen05 + TMA path): > 8 items per persistent CTA (the published-item ring wraps),
    # odd interior widths (head_dim 8 pairs windows: the last pair has one), no fringe at all, fringe only on one side
    (2, 200, 264, 32, 4), (1, 129, 101, 32, 0), (1, 64, 64, 32, 0), (1, 68, 33, 32, 4),
    (2, 100, 93, 64, 4), (1, 259, 131, 64, 0), (1, 48, 24, 64, 0), (3, 28, 20, 64, 4),
    (1, 96, 72, 64, 4),                               # heads = 16 -> head_dim 4, four quads
])
def test_window_attention_matches_simt(B, H, W, C, shift):
    from tfswa_unet_b200 import ops, _lib
    heads = 16 if (C, H) in ((64, 40), (64, 96)) else 8
    tc_before = _lib.lib().tfswa_attn_win_tc_interior_launches()
    M = B * H * W
    big = seeded((M, 9 * C), 31, 1.5).cuda().to(torch.bfloat16)
    qkv = big[:, 6 * C:9 * C]                                     # the SW-MSA slab of a fused 9C-wide qkv buffer
    pad_kv = seeded((2 * C,), 32, 0.7).cuda().float().contiguous()
    out_tc = torch.empty((M, 3, C), dtype=torch.bfloat16, device="cuda")[:, 2, :]
    lse_tc = torch.empty((M, heads), dtype=torch.float32, device="cuda")
    out_s = torch.empty((M, C), dtype=torch.bfloat16, device="cuda")
    lse_s = torch.empty((M, heads), dtype=torch.float32, device="cuda")
    ops.USE_TC_ATTENTION = True
    n0 = ops.LAUNCHES
    ops.enable_timing(True)
    try:
        ops.attention(qkv, out_tc, B, H, W, C, heads, 2, ws=8, shift=shift, pad_kv=pad_kv, lse=lse_tc)
        torch.cuda.synchronize()
        tags = list(ops.collect_timing().keys())
    finally:
        ops.enable_timing(False)
    assert tags and tags[0].startswith("attn_tc[swa"), tags     # the tensor-core window kernel really ran
    interior = ((H - shift) // 8) * ((W - shift) // 8)
    want_tc = 1 if (C // heads in (4, 8) and interior > 0) else 0   # ... and its tcgen05 form wherever an interior window exists
    assert _lib.lib().tfswa_attn_win_tc_interior_launches() - tc_before == want_tc
    ops.USE_TC_ATTENTION = False
    try:
        ops.attention(qkv, out_s, B, H, W, C, heads, 2, ws=8, shift=shift, pad_kv=pad_kv, lse=lse_s)
    finally:
        ops.USE_TC_ATTENTION = True
    torch.cuda.synchronize()
    assert ops.LAUNCHES == n0 + 2
    scale = float(out_s.float().abs().max())
    assert torch.isfinite(out_tc.float()).all()
    err = float((out_tc.float() - out_s.float()).abs().max())
    assert err <= 2e-2 * scale, f"window tc vs simt: {err:.3e} (scale {scale:.3e})"
    assert float((lse_tc - lse_s).abs().max()) <= 3e-2, "log-sum-exp mismatch"
    # and against a torch fp32 restatement of the same op (not only CUDA against CUDA)
    ref = _ref_windows(qkv.contiguous(), B, H, W, C, heads, 8, shift, pad_kv)
    rscale = float(ref.abs().max())
    e_tc = float((out_tc.float() - ref).abs().max())
    e_s = float((out_s.float() - ref).abs().max())
    assert e_tc <= 2e-2 * rscale, f"window tc vs torch fp32: {e_tc:.3e} (scale {rscale:.3e})"
    assert e_s <= 1e-2 * rscale, f"window simt vs torch fp32: {e_s:.3e}"


def test_window_features_stay_on_simt():
    """mask / relative bias are default-off features of the reference: they are served by tfswa_attn_fwd"""
    from tfswa_unet_b200 import ops
    B, H, W, C = 1, 16, 16, 32
    qkv = seeded((B * H * W, 3 * C), 5, 1.0).cuda().to(torch.bfloat16)
    out = torch.empty((B * H * W, C), dtype=torch.bfloat16, device="cuda")
    ops.enable_timing(True)
    try:
        ops.attention(qkv, out, B, H, W, C, 8, 2, ws=8, shift=4, use_shift_mask=True)
        torch.cuda.synchronize()
        tags = list(ops.collect_timing().keys())
    finally:
        ops.enable_timing(False)
    assert tags[0].startswith("attn[swa")


@pytest.mark.parametrize("B,H,W,C,shift,heads", [
    (1, 8, 8, 32, 0, 8), (2, 13, 21, 32, 4, 8),       # head_dim 4: one window; padded + shifted windows
    (1, 24, 16, 64, 4, 8), (2, 11, 30, 64, 0, 8),     # head_dim 8
    (1, 16, 24, 128, 4, 8), (1, 9, 17, 128, 0, 8),    # head_dim 16: two 32-row CTAs per window
    (1, 40, 16, 64, 4, 16),                           # two 8-head slabs
])
def test_window_attention_backward_mma_matches_simt(B, H, W, C, shift, heads):
    """tfswa_attn_bwd on SW-MSA windows: the warp-MMA kernels (default for bf16) against the CUDA-core kernels
    (TFSWA_ATTN_BWD_SIMT=1) on the same q|k|v, output gradient and saved log-sum-exp; pad-token gradients included."""
    import os
    from tfswa_unet_b200 import ops
    M = B * H * W
    qkv = seeded((M, 3 * C), 41, 1.2).cuda().to(torch.bfloat16)
    dout = seeded((M, C), 42, 1.0).cuda().to(torch.bfloat16)
    pad_kv = seeded((2 * C,), 43, 0.7).cuda().float().contiguous()
    out = torch.empty((M, C), dtype=torch.bfloat16, device="cuda")
    lse = torch.empty((M, heads), dtype=torch.float32, device="cuda")
    ops.attention(qkv, out, B, H, W, C, heads, 2, ws=8, shift=shift, pad_kv=pad_kv, lse=lse)
    res = {}
    for mode in ("mma", "simt"):
        dqkv = torch.full((M, 3 * C), float("nan"), dtype=torch.bfloat16, device="cuda")
        dsum = torch.zeros((M, heads), dtype=torch.float32, device="cuda")
        dpad = torch.zeros((2 * C,), dtype=torch.float32, device="cuda")
        if mode == "simt":
            os.environ["TFSWA_ATTN_BWD_SIMT"] = "1"
        try:
            ops.attention_bwd(qkv, out, lse, dout, dqkv, dsum, B, H, W, C, heads, 2, ws=8, shift=shift, pad_kv=pad_kv, dpad=dpad)
            torch.cuda.synchronize()
        finally:
            os.environ.pop("TFSWA_ATTN_BWD_SIMT", None)
        res[mode] = (dqkv.float(), dsum, dpad)
    (a, sa, pa), (b, sb, pb) = res["mma"], res["simt"]
    assert torch.isfinite(a).all(), "every token's dq|dk|dv must be written"
    for name, lo in (("dq", 0), ("dk", C), ("dv", 2 * C)):
        x, y = a[:, lo:lo + C], b[:, lo:lo + C]
        rel = float((x - y).norm() / y.norm())
        assert rel <= 2e-2, f"{name}: mma vs simt rel-L2 {rel:.3e}"          # P, dS rounded to bf16 as MMA operands
    assert float((sa - sb).abs().max()) <= 1e-3 * float(sb.abs().max()) + 1e-5
    if (H % 8) or (W % 8):
        assert float(pb.abs().max()) > 0
        assert float((pa - pb).norm() / pb.norm()) <= 2e-2, "pad-token gradient"
    else:
        assert float(pa.abs().max()) == 0.0
