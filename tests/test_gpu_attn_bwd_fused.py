"""GPU (-m gpu): the one-pass head_dim-4 / 8 attention backward (attn_bwd_fused_kernel<D>: P computed once, dQ accumulated in
shared memory) against the two-kernel warp-MMA path (TFSWA_ATTN_BWD_FUSED=0), the CUDA-core path (TFSWA_ATTN_BWD_SIMT=1)
and torch autograd of the same op in fp32."""
import os

import pytest
import torch

from helpers import seeded

pytestmark = pytest.mark.gpu


def _torch_ref(qkv, dout, B, H, W, C, heads, geom):
    d = C // heads
    x = qkv.float().clone().requires_grad_(True)
    t = x.view(B, H, W, 3, heads, d)
    t = t.permute(0, 2, 3, 4, 1, 5) if geom == 0 else t.permute(0, 1, 3, 4, 2, 5)      # (B, R, 3, h, N, d)
    q, k, v = t[:, :, 0], t[:, :, 1], t[:, :, 2]
    o = torch.softmax((q @ k.transpose(-1, -2)) * d ** -0.5, -1) @ v                  # (B, R, h, N, d)
    o = o.permute(0, 3, 1, 2, 4) if geom == 0 else o.permute(0, 1, 3, 2, 4)
    o = o.reshape(B * H * W, C)
    o.backward(dout.float())
    return x.grad


@pytest.mark.parametrize("B,H,W,C,heads,geom", [
    (1, 37, 5, 32, 8, 0), (1, 5, 37, 32, 8, 1),            # one ragged query tile, ragged last key block
    (2, 129, 3, 32, 8, 0), (1, 3, 300, 32, 8, 1),          # several query tiles / key blocks
    (1, 64, 4, 32, 8, 0), (1, 4, 96, 32, 8, 1),            # exact multiples of the tile sizes
    (1, 1025, 2, 32, 8, 0), (1, 2, 517, 32, 8, 1),         # the C3 stage-1 sequence lengths
    (1, 70, 3, 64, 16, 0),                                 # two 8-head slabs
    (1, 37, 5, 64, 8, 0), (1, 5, 37, 64, 8, 1),            # head_dim 8: ragged tile / key block
    (2, 129, 3, 64, 8, 0), (1, 3, 300, 64, 8, 1),
    (1, 512, 2, 64, 8, 0), (1, 2, 258, 64, 8, 1),          # the C3 stage-2 sequence lengths
    (1, 70, 3, 128, 16, 1),                                # head_dim 8, two slabs
])
def test_fused_attention_backward_matches_other_paths(B, H, W, C, heads, geom):
    from tfswa_unet_b200 import ops
    M = B * H * W
    qkv = seeded((M, 3 * C), 81, 1.2).cuda().to(torch.bfloat16)
    dout = seeded((M, C), 82, 1.0).cuda().to(torch.bfloat16)
    out = torch.empty((M, C), dtype=torch.bfloat16, device="cuda")
    lse = torch.empty((M, heads), dtype=torch.float32, device="cuda")
    ops.attention(qkv, out, B, H, W, C, heads, geom, lse=lse)
    res = {}
    for mode, env in (("fused", {}), ("pair", {"TFSWA_ATTN_BWD_FUSED": "0"}), ("simt", {"TFSWA_ATTN_BWD_SIMT": "1"})):
        dqkv = torch.full((M, 3 * C), float("nan"), dtype=torch.bfloat16, device="cuda")
        dsum = torch.full((M, heads), float("nan"), dtype=torch.float32, device="cuda")
        os.environ.update(env)
        try:
            ops.attention_bwd(qkv, out, lse, dout, dqkv, dsum, B, H, W, C, heads, geom)
            torch.cuda.synchronize()
        finally:
            for k in env:
                os.environ.pop(k, None)
        res[mode] = (dqkv.float(), dsum)
    ref = _torch_ref(qkv, dout, B, H, W, C, heads, geom)
    f, fs = res["fused"]
    assert torch.isfinite(f).all(), "every token's dq|dk|dv must be written"
    assert float((fs - res["simt"][1]).abs().max()) <= 1e-3 * float(res["simt"][1].abs().max()) + 1e-5
    for name, lo in (("dq", 0), ("dk", C), ("dv", 2 * C)):
        x = f[:, lo:lo + C]
        for other in ("pair", "simt"):
            y = res[other][0][:, lo:lo + C]
            rel = float((x - y).norm() / y.norm())
            assert rel <= 2e-2, f"{name}: fused vs {other} rel-L2 {rel:.3e}"
        r = ref[:, lo:lo + C]
        rel = float((x - r).norm() / r.norm())
        assert rel <= 3e-2, f"{name}: fused vs torch fp32 autograd rel-L2 {rel:.3e}"     # forward O and lse are bf16-path values
