"""GPU (-m gpu): the tcgen05 + TMA weight-gradient kernel (wgrad_tc.cu, behind tfswa_linear_wgrad for bf16) against a torch
fp32 reference of the same contraction, for the (K, N, batch, prologue) shapes the model produces, strided slab views and
token counts that are not multiples of the 64-token stage."""
import pytest
import torch

from helpers import seeded

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("M,K,N,nb,ln", [
    (5000, 32, 288, 1, True),       # stage-1 q|k|v of one branch (N = 9 x 32: 32-column boxes, three n tiles, last one ragged)
    (4097, 32, 128, 3, True),       # stage-1 fc1, three branches, LayerNorm prologue
    (3000, 128, 32, 3, False),      # stage-1 fc2 (N < 128: zero-filled rows of the G tile)
    (2049, 96, 32, 1, False),       # stage-1 fusion conv (K = 3 x 32)
    (1500, 64, 64, 1, False),
    (1111, 128, 1152, 1, True),     # stage-3 q|k|v of three branches
    (900, 512, 128, 3, False),      # stage-3 fc2: two k tiles of 256
    (700, 1024, 256, 3, False),     # stage-4 fc2: four k tiles
    (640, 384, 128, 1, False),      # stage-3 fusion conv: two k tiles of 192
    (300, 256, 768, 1, True),
    (1554, 32, 96, 1, True),        # a single branch's q|k|v (tests/test_gpu_backward.py tsa_c32)
    (1554, 32, 32, 1, False),
])
@pytest.mark.parametrize("slab", [True, False])
def test_wgrad_tc_matches_torch(M, K, N, nb, ln, slab):
    from tfswa_unet_b200 import ops, _lib as L
    if slab:      # slab views: x lives inside a wider (M, nb, K + 32) buffer, g inside (M, nb, N + 64)
        xbig = seeded((M, nb, K + 32), 51, 1.0).cuda().to(torch.bfloat16)
        gbig = seeded((M, nb, N + 64), 52, 0.5).cuda().to(torch.bfloat16)
        x, g = xbig[:, :, :K], gbig[:, :, :N]
    else:
        x = seeded((M, nb, K), 51, 1.0).cuda().to(torch.bfloat16)
        g = seeded((M, nb, N), 52, 0.5).cuda().to(torch.bfloat16)
    stats = None
    xf = x.float()
    if ln:
        mean = xf.mean(-1)                                             # (M, nb)
        rstd = torch.rsqrt(xf.var(-1, unbiased=False) + 1e-5)
        stats = torch.stack([mean.t(), rstd.t()], -1).contiguous()     # (nb, M, 2): (mean, rstd) per token, as tfswa_row_stats writes them
        xf = ((xf - mean[..., None]) * rstd[..., None]).to(torch.bfloat16).float()     # the kernel feeds bf16 to the MMA
    dw, db = ops.linear_wgrad(x, g, prologue=L.PRO_LNHAT if ln else L.PRO_NONE, row_stats=stats, want_bias=True)
    torch.cuda.synchronize()
    gf = g.float()
    ref_dw = torch.einsum("mbn,mbk->bnk", gf, xf)
    ref_db = gf.sum(0)
    assert dw.shape == (nb, N, K) and db.shape == (nb, N)
    e_w = float((dw - ref_dw).abs().max()) / float(ref_dw.abs().max())
    e_b = float((db - ref_db).abs().max()) / float(ref_db.abs().max())
    assert e_w <= 2e-3, f"dW rel-max-err {e_w:.3e}"
    assert e_b <= 2e-3, f"dbias rel-max-err {e_b:.3e}"


def test_wgrad_tc_is_the_kernel_that_runs():
    """the launch really goes through the tcgen05 kernel (the entry point would silently fall back to the warp-MMA kernel
    for shapes it does not cover): wrong results from a forced layout mismatch are not what we want to find later"""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import torch, sys; sys.path.insert(0, %r)\n"
            "from tfswa_unet_b200 import ops\n"
            "x = torch.randn(4096, 1, 64, device='cuda').bfloat16(); g = torch.randn(4096, 1, 128, device='cuda').bfloat16()\n"
            "from torch.profiler import profile, ProfilerActivity\n"
            "with profile(activities=[ProfilerActivity.CUDA]) as prof:\n"
            "    ops.linear_wgrad(x, g); torch.cuda.synchronize()\n"
            "names = [e.key for e in prof.key_averages()]\n"
            "print('KERNELS', names)\n" % root)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "wgrad_tc_kernel" in r.stdout, r.stdout[-1000:]
