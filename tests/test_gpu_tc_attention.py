"""GPU (-m gpu): tcgen05 axial attention (head_dim 4 / 8, bf16) against the fp32-math SIMT kernel on the same bf16
q|k|v, and against a torch fp32 reference of the same op."""
import pytest
import torch

from helpers import seeded

pytestmark = pytest.mark.gpu


def _ref(qkv, B, H, W, C, heads, geom):
    d = C // heads
    t = qkv.float().view(B, H, W, 3, heads, d)
    if geom == 0:      # TSA: sequences along H
        t = t.permute(0, 2, 3, 4, 1, 5)          # (B, W, 3, h, H, d)
    else:
        t = t.permute(0, 1, 3, 4, 2, 5)          # (B, H, 3, h, W, d)
    q, k, v = t[:, :, 0], t[:, :, 1], t[:, :, 2]
    s = (q @ k.transpose(-1, -2)) * d ** -0.5
    o = torch.softmax(s, -1) @ v                 # (B, R, h, N, d)
    if geom == 0:
        o = o.permute(0, 3, 1, 2, 4)             # (B, H, W, h, d)
    else:
        o = o.permute(0, 1, 3, 2, 4)
    return o.reshape(B * H * W, C)


@pytest.mark.parametrize("B,H,W,C,geom", [
    (1, 37, 5, 32, 0), (1, 5, 37, 32, 1),            # one ragged tile
    (2, 129, 3, 32, 0), (1, 3, 300, 32, 1),          # several query tiles, key tail
    (1, 1025, 2, 32, 0), (1, 2, 517, 32, 1),         # the C3 stage-1 sequence lengths
    (1, 64, 4, 64, 0), (2, 4, 258, 64, 1),           # head_dim 8
    (1, 512, 2, 64, 0), (1, 128, 3, 32, 0), (1, 2, 33, 64, 1),
    (1, 256, 3, 128, 0), (2, 3, 200, 128, 1), (1, 140, 2, 128, 0), (1, 2, 300, 128, 1),   # head_dim 16 (one head per CTA)
    (1, 257, 2, 128, 0), (1, 2, 130, 128, 1), (2, 3, 129, 128, 1), (1, 259, 2, 64, 0), (1, 2, 513, 64, 1),   # 1-3 keys past the last full tile (C3 stage 3: 129)
])
def test_tc_attention_matches_simt_and_reference(B, H, W, C, geom):
    from tfswa_unet_b200 import ops
    heads = 8
    M = B * H * W
    big = seeded((M, 9 * C), 11, 1.5).cuda().to(torch.bfloat16)     # slab view: q|k|v of "branch 1" inside a 9C-wide buffer
    qkv = big[:, 3 * C:6 * C]
    out_tc = torch.empty((M, 3, C), dtype=torch.bfloat16, device="cuda")[:, 1, :]      # strided output slab
    lse_tc = torch.empty((M, heads), dtype=torch.float32, device="cuda")
    ops.USE_TC_ATTENTION = True
    ops.attention(qkv, out_tc, B, H, W, C, heads, geom, lse=lse_tc)
    out_s = torch.empty((M, C), dtype=torch.bfloat16, device="cuda")
    lse_s = torch.empty((M, heads), dtype=torch.float32, device="cuda")
    ops.USE_TC_ATTENTION = False
    try:
        ops.attention(qkv, out_s, B, H, W, C, heads, geom, lse=lse_s)
    finally:
        ops.USE_TC_ATTENTION = True
    torch.cuda.synchronize()
    ref = _ref(qkv.contiguous(), B, H, W, C, heads, geom)
    scale = float(ref.abs().max())
    e_ref = float((out_tc.float() - ref).abs().max())
    e_simt = float((out_tc.float() - out_s.float()).abs().max())
    assert torch.isfinite(out_tc.float()).all()
    assert e_ref <= 2e-2 * scale, f"tc vs fp32 reference: {e_ref:.3e} (scale {scale:.3e})"
    assert e_simt <= 2e-2 * scale, f"tc vs simt: {e_simt:.3e}"
    assert float((lse_tc - lse_s).abs().max()) <= 3e-2, "log-sum-exp mismatch"


@pytest.mark.parametrize("B,H,W,C,geom", [(1, 300, 2, 32, 0), (1, 2, 517, 32, 1), (1, 3, 200, 64, 1), (1, 2, 260, 128, 1)])
def test_tc_attention_exact_two_pass_path(B, H, W, C, geom):
    """the in-kernel fallback (exact row maxima) selected explicitly must agree with the bounded fast path"""
    from tfswa_unet_b200 import ops
    M = B * H * W
    qkv = seeded((M, 3 * C), 21, 1.5).cuda().to(torch.bfloat16)
    a = torch.empty((M, C), dtype=torch.bfloat16, device="cuda")
    b = torch.empty_like(a)
    la = torch.empty((M, 8), dtype=torch.float32, device="cuda")
    lb = torch.empty_like(la)
    ops.attention(qkv, a, B, H, W, C, 8, geom, lse=la)
    ops.attention(qkv, b, B, H, W, C, 8, geom, lse=lb, force_exact=True)      # test hook: the exact two-pass path
    torch.cuda.synchronize()
    ref = _ref(qkv, B, H, W, C, 8, geom)
    scale = float(ref.abs().max())
    assert float((b.float() - ref).abs().max()) <= 2e-2 * scale
    assert float((a.float() - b.float()).abs().max()) <= 2e-2 * scale
    assert float((la - lb).abs().max()) <= 3e-2


def test_tc_attention_loose_bound_falls_back_to_exact():
    """q = (a,a,0,0), k = +-(b,-b,0,0): every score is 0 but the per-channel bound is 2ab*scale*log2e >> 126,
    so every exponential underflows on the fast path and the CTA must redo the rows with the exact maximum"""
    from tfswa_unet_b200 import ops
    B, H, W, C, heads = 1, 70, 1, 32, 8
    M = B * H * W
    qkv = torch.zeros((M, 3 * C), device="cuda")
    for h in range(heads):
        qkv[:, 4 * h] = 16.0
        qkv[:, 4 * h + 1] = 16.0
        sign = torch.where(torch.arange(M, device="cuda") % 2 == 0, 1.0, -1.0)
        qkv[:, C + 4 * h] = 16.0 * sign
        qkv[:, C + 4 * h + 1] = -16.0 * sign
    qkv[:, 2 * C:] = seeded((M, C), 5).cuda()
    qkv = qkv.to(torch.bfloat16)
    out = torch.empty((M, C), dtype=torch.bfloat16, device="cuda")
    ops.attention(qkv, out, B, H, W, C, heads, 0)
    torch.cuda.synchronize()
    ref = qkv[:, 2 * C:].float().mean(0, keepdim=True).expand(M, C)     # all scores equal -> uniform attention
    assert torch.isfinite(out.float()).all()
    assert float((out.float() - ref).abs().max()) <= 2e-2 * float(ref.abs().max()) + 1e-2


def test_tc_attention_large_negative_scores_do_not_overflow():
    """all true scores << 0 while absent (zero) keys of the tail tile score 0: the tail clamp must keep P finite"""
    from tfswa_unet_b200 import ops
    B, H, W, C, heads = 1, 33, 1, 32, 8
    M = B * H * W
    qkv = torch.zeros((M, 3 * C), device="cuda")
    qkv[:, :C] = 6.0
    qkv[:, C:2 * C] = -6.0            # q.k = -144 per head -> scaled -72
    qkv[:, 2 * C:] = seeded((M, C), 3).cuda()
    qkv = qkv.to(torch.bfloat16)
    out = torch.empty((M, C), dtype=torch.bfloat16, device="cuda")
    ops.attention(qkv, out, B, H, W, C, heads, 0)
    torch.cuda.synchronize()
    ref = qkv[:, 2 * C:].float().mean(0, keepdim=True).expand(M, C)     # uniform attention
    assert torch.isfinite(out.float()).all()
    assert float((out.float() - ref).abs().max()) <= 2e-2 * float(ref.abs().max()) + 1e-2
