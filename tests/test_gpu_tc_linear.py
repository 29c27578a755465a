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
