"""GPU (-m gpu): tcgen05/TMA token GEMM against the fp32 SIMT engine and a torch fp32 reference of the same op."""
import pytest
import torch

from helpers import seeded

pytestmark = pytest.mark.gpu


def _ref(x, w, b, st, epi, r1, r2, ln):
    xf = x.float()
    if ln:
        mean = xf.mean(-1, keepdim=True)
        var = xf.var(-1, unbiased=False, keepdim=True)
        xf = (xf - mean) * torch.rsqrt(var + 1e-5)
    y = torch.einsum("mbk,bnk->mbn", xf, w) + b[None]
    if epi:
        y = torch.nn.functional.gelu(y)
    if r1 is not None:
        y = y + r1.float()
    if r2 is not None:
        y = y + r2.float()
    return y


# (M, nb, K, N, ln, gelu, r1, r2): every (K, N) pair the model uses at its four stages, plus ragged M
CASES = [
    (300, 1, 32, 32, False, False, False, False),      # input_proj C=32
    (1000, 1, 32, 288, True, False, False, False),     # qkv x3 branches, C=32 (BN=144, 64B swizzle)
    (515, 3, 32, 32, False, False, True, False),       # proj + residual (broadcast r1)
    (515, 3, 32, 128, True, True, False, False),       # fc1 + GELU
    (515, 3, 128, 32, False, False, True, False),      # fc2 + residual
    (515, 1, 96, 32, False, True, True, True),         # fusion (K = 3C, 64B swizzle, 3 k-blocks) + 2 residuals
    (777, 1, 64, 576, True, False, False, False),      # C=64 qkv (BN=192)
    (260, 3, 64, 256, True, True, False, False),
    (260, 3, 256, 64, False, False, True, False),
    (129, 1, 128, 1152, True, False, False, False),    # C=128 qkv
    (129, 3, 128, 512, True, True, False, False),
    (129, 3, 512, 128, False, False, True, False),     # 8 k-blocks > 4 stages: ring wrap-around
    (129, 1, 384, 128, False, True, True, True),
    (200, 1, 256, 2304, True, False, False, False),    # C=256 qkv (BN=256, 2 stages)
    (200, 3, 256, 1024, True, True, False, False),
    (200, 3, 1024, 256, False, False, True, False),    # 16 k-blocks
    (200, 1, 768, 256, False, True, True, False),
    (128, 1, 256, 256, False, False, False, False),
    (1, 1, 32, 32, False, False, False, False),        # a single token
]


@pytest.mark.parametrize("M,nb,K,N,ln,gelu,use_r1,use_r2", CASES)
def test_tc_linear_matches_reference(M, nb, K, N, ln, gelu, use_r1, use_r2):
    from tfswa_unet_b200 import ops, _lib as L
    dev = "cuda"
    x = seeded((M, nb, K), 1, 1.0).add_(0.3).to(dev).to(torch.bfloat16)
    w = (seeded((nb, N, K), 2) / K ** 0.5).to(dev)
    wb = w.to(torch.bfloat16).contiguous()
    wsum = wb.float().sum(-1).contiguous()
    b = seeded((nb, N), 3, 0.1).to(dev)
    r1 = seeded((M, 1, N), 4).to(dev).to(torch.bfloat16) if use_r1 else None     # broadcast over nb
    r2 = seeded((M, nb, N), 5).to(dev).to(torch.bfloat16) if use_r2 else None
    st = ops.row_stats(x) if ln else None
    y = ops.linear_tc(x, wb, wsum, b, prologue=L.PRO_LNHAT if ln else 0, epilogue=L.EPI_GELU if gelu else 0, row_stats=st,
                      r1=r1, r2=r2)
    ref = _ref(x, wb.float(), b, st, gelu, r1, r2, ln)
    torch.cuda.synchronize()
    err = float((y.float() - ref).abs().max())
    scale = float(ref.abs().max())
    assert err <= 1e-2 * scale + 1e-3, f"tc_linear err {err:.3e} vs scale {scale:.3e}"
    # and against the SIMT kernel on the same bf16 inputs (independent implementation, same contract)
    y2 = ops.linear(x, wb.float().contiguous(), b, prologue=L.PRO_LNHAT if ln else 0, epilogue=L.EPI_GELU if gelu else 0,
                    row_stats=st, r1=r1, r2=r2)
    err2 = float((y.float() - y2.float()).abs().max())
    assert err2 <= 1e-2 * scale + 1e-3, f"tc vs simt {err2:.3e}"


# the PERSISTENT kernel is dispatched when every SM gets at least two tiles (148 SMs: >= 296 tiles of 128 rows x BN columns); wide
# erf-GELU epilogues (N > K) stay on the one-tile-per-CTA kernel.  (M, nb, K, N, ln, gelu, r1, r2)
PERSISTENT_CASES = [
    (40000, 1, 32, 32, False, False, False, False),     # BN = 32: three epilogue groups, one n-tile, interleaved walk
    (40001, 3, 32, 32, False, False, True, False),      # residual tile through TMA into the staging buffer, ragged last row block, batch 3
    (20000, 3, 128, 32, False, False, True, False),     # fc2 + residual: 2 k-blocks, bias rows reloaded per batch entry
    (30000, 1, 96, 32, False, True, True, True),        # fusion conv: GELU with N < K, both residual paths (r1 by TMA, r2 by loads)
    (160000, 1, 32, 288, True, False, False, False),    # BN = 96, three n-tiles, LayerNorm algebra, ROW-MAJOR walk (>= 8 row blocks per SM)
    (20000, 1, 64, 576, True, False, False, False),     # BN = 64 / 192-column split, interleaved walk with several n-tiles
    (8000, 1, 128, 1152, True, False, False, False),    # BN = 128: two groups at 96 registers, 9 n-tiles
    (9000, 3, 512, 128, False, False, True, False),     # 8 k-blocks per tile through a 4-stage ring
    (5000, 1, 256, 2304, True, False, False, False),    # 18 n-tiles, bias rows of 2304 columns in shared memory
]


@pytest.mark.parametrize("M,nb,K,N,ln,gelu,use_r1,use_r2", PERSISTENT_CASES)
def test_tc_linear_persistent_kernel_matches_reference(M, nb, K, N, ln, gelu, use_r1, use_r2):
    from tfswa_unet_b200 import ops, _lib as L
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(7)
    x = (torch.randn((M, nb, K), device=dev, generator=g) + 0.3).to(torch.bfloat16)
    w = torch.randn((nb, N, K), device=dev, generator=g) / K ** 0.5
    wb = w.to(torch.bfloat16).contiguous()
    wsum = wb.float().sum(-1).contiguous()
    b = 0.1 * torch.randn((nb, N), device=dev, generator=g)
    r1 = torch.randn((M, 1, N), device=dev, generator=g).to(torch.bfloat16) if use_r1 else None     # broadcast over nb
    r2 = torch.randn((M, nb, N), device=dev, generator=g).to(torch.bfloat16) if use_r2 else None
    st = ops.row_stats(x) if ln else None
    ops.reset_launch_count()
    y = ops.linear_tc(x, wb, wsum, b, prologue=L.PRO_LNHAT if ln else 0, epilogue=L.EPI_GELU if gelu else 0, row_stats=st,
                      r1=r1, r2=r2)
    ref = _ref(x, wb.float(), b, st, gelu, r1, r2, ln)
    torch.cuda.synchronize()
    # every row block and every column block: relative L2 per 4096-row band (a dropped or duplicated tile shows up as O(1))
    for m0 in range(0, M, 4096):
        a, r = y[m0:m0 + 4096].float(), ref[m0:m0 + 4096]
        assert float((a - r).norm() / r.norm()) <= 6e-3, m0
    err = float((y.float() - ref).abs().max())
    scale = float(ref.abs().max())
    assert err <= 1e-2 * scale + 1e-3, f"tc_linear(persistent) err {err:.3e} vs scale {scale:.3e}"


def test_tc_linear_persistent_is_the_kernel_that_runs():
    """large problems really go through tc_linear_persist_kernel (<3> for narrow tiles, <2> for 96-128-column tiles), and
    TFSWA_LINEAR_KERNEL=v1 really selects the one-tile-per-CTA kernel; both give the same bits (same arithmetic per element)"""
    import subprocess, sys, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import torch, sys; sys.path.insert(0, %r)\n"
            "from tfswa_unet_b200 import ops\n"
            "torch.manual_seed(0)\n"
            "x = torch.randn(40000, 1, 32, device='cuda').bfloat16()\n"
            "wa = (torch.randn(1, 32, 32, device='cuda') / 6).bfloat16(); wb = (torch.randn(1, 288, 32, device='cuda') / 6).bfloat16()\n"
            "r1 = torch.randn(40000, 1, 32, device='cuda').bfloat16()\n"
            "from torch.profiler import profile, ProfilerActivity\n"
            "with profile(activities=[ProfilerActivity.CUDA]) as prof:\n"
            "    ya = ops.linear_tc(x, wa, None, None, r1=r1); yb = ops.linear_tc(x, wb, None, None); torch.cuda.synchronize()\n"
            "print('KERNELS', [e.key for e in prof.key_averages()])\n"
            "print('SUMS', float(ya.float().sum()), float(yb.float().sum()), float(ya.float().abs().sum()), float(yb.float().abs().sum()))\n" % root)
    outs = {}
    for mode in ("", "v1"):
        env = dict(os.environ, TFSWA_LINEAR_KERNEL=mode)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[mode] = r.stdout
    assert "tc_linear_persist_kernel<3>" in outs[""] and "tc_linear_persist_kernel<2>" in outs[""], outs[""][-1500:]
    assert "tc_linear_persist_kernel" not in outs["v1"] and "tc_linear_kernel" in outs["v1"], outs["v1"][-1500:]
    sums = {m: [l for l in o.splitlines() if l.startswith("SUMS")][0] for m, o in outs.items()}
    assert sums[""] == sums["v1"], sums


def test_tc_linear_persistent_column_sums():
    """train-mode BatchNorm sums from the persistent kernel's staged tiles (M large enough for the persistent path)"""
    from tfswa_unet_b200 import ops
    dev = "cuda"
    M, K, N = 50000, 96, 32
    g = torch.Generator(device=dev).manual_seed(9)
    x = torch.randn((M, 1, K), device=dev, generator=g).to(torch.bfloat16)
    wb = (torch.randn((1, N, K), device=dev, generator=g) / K ** 0.5).to(torch.bfloat16)
    b = 0.1 * torch.randn((1, N), device=dev, generator=g)
    stats = torch.zeros((2, N), dtype=torch.float32, device=dev)
    y = ops.linear_tc(x, wb, None, b, col_stats=stats)
    torch.cuda.synchronize()
    yf = y.float().reshape(M, N)
    assert torch.allclose(stats[0], yf.sum(0), rtol=2e-4, atol=2e-2)
    assert torch.allclose(stats[1], (yf * yf).sum(0), rtol=2e-4, atol=2e-2)


def test_tc_linear_strided_slab_views():
    """x / y / r address column slabs of wider buffers (the (M,3,C) concat buffer and the (M,9C) qkv buffer)."""
    from tfswa_unet_b200 import ops
    dev = "cuda"
    M, C = 400, 64
    big = seeded((M, 3, C), 7).to(dev).to(torch.bfloat16)          # x[:, b, :] slabs, ld = 3C
    w = (seeded((3, C, C), 8) / 8).to(dev)
    wb = w.to(torch.bfloat16).contiguous()
    y = ops.linear_tc(big, wb, None, None)
    ref = torch.einsum("mbk,bnk->mbn", big.float(), wb.float())
    assert float((y.float() - ref).abs().max()) <= 1e-2 * float(ref.abs().max())
    # K = 3C read of the same buffer as one (M,1,3C) operand (the fusion conv)
    wf = (seeded((1, C, 3 * C), 9) / 14).to(dev).to(torch.bfloat16).contiguous()
    y2 = ops.linear_tc(big.view(M, 1, 3 * C), wf, None, None)
    ref2 = big.view(M, 3 * C).float() @ wf[0].float().t()
    assert float((y2[:, 0].float() - ref2).abs().max()) <= 1e-2 * float(ref2.abs().max())


@pytest.mark.parametrize("M,K,N", [(1000, 32, 32), (128 * 5 + 17, 96, 32), (4096 + 3, 384, 128), (700, 768, 256), (300, 64, 64)])
def test_tc_linear_column_statistics(M, K, N):
    """train-mode BatchNorm sums from the tcgen05 epilogue: sum and sum of squares of the STORED (bf16) output over the
    valid rows only (the last 128-row tile is ragged), every N tile."""
    from tfswa_unet_b200 import ops
    x = seeded((M, 1, K), 61, 1.0).cuda().to(torch.bfloat16)
    w = seeded((1, N, K), 62, K ** -0.5).cuda().to(torch.bfloat16)
    b = seeded((1, N), 63, 0.5).cuda().float()
    stats = torch.zeros((2, N), dtype=torch.float32, device="cuda")
    y = ops.linear_tc(x, w, None, b, col_stats=stats)
    torch.cuda.synchronize()
    ref = x[:, 0].float() @ w[0].float().t() + b[0]
    assert float((y[:, 0].float() - ref).abs().max()) <= 2e-2 * float(ref.abs().max())
    yf = y[:, 0].double()
    s1, s2 = yf.sum(0), (yf * yf).sum(0)
    assert float((stats[0].double() - s1).abs().max()) <= 1e-4 * float(yf.abs().sum(0).max()) + 1e-3
    assert float((stats[1].double() - s2).abs().max()) <= 1e-4 * float(s2.max())
    with pytest.raises(RuntimeError, match="col_stats"):
        ops.linear_tc(x, w, None, b, col_stats=torch.zeros((2, N), device="cuda"), epilogue=1)
