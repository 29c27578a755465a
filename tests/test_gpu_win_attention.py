"""GPU (-m gpu): register-resident SW-MSA window attention (tfswa_attn_win_tc_fwd, warp-level MMAs) against the fp32-math
SIMT kernel on the same bf16 q|k|v, for every head_dim of the model, padded windows and cyclic shift."""
import pytest
import torch

from helpers import seeded

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,H,W,C,shift", [
    (1, 8, 8, 32, 0), (1, 8, 8, 32, 4),               # one window
    (2, 13, 21, 32, 4), (1, 17, 9, 32, 0),            # padded windows (pad tokens are real keys = pad_kv)
    (1, 24, 16, 64, 4), (2, 11, 30, 64, 0),           # head_dim 8
    (1, 16, 24, 128, 4), (1, 9, 17, 128, 0),          # head_dim 16, two channel slabs
    (1, 16, 8, 256, 4), (2, 13, 9, 256, 0),           # head_dim 32
    (1, 40, 16, 64, 4),                               # heads = 16 -> head_dim 4 with a 64-channel slab
])
def test_window_attention_matches_simt(B, H, W, C, shift):
    from tfswa_unet_b200 import ops
    heads = 16 if (C, H) == (64, 40) else 8
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
