"""Fused block head (tfswa_block_head_tc_fwd: input_proj + LayerNorm + q|k|v GEMM) against an fp32 torch restatement
of blocks.py:53-56,115 + attention.py:70,146 and against the unfused tensor-core sequence it replaces."""
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu


def _case(M, C, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    t = dict(x=rn(M, 1, C) + 0.2, wi=rn(1, C, C) / C ** 0.5, bi=rn(1, C) * 0.3,
             wq=rn(1, 9 * C, C) / C ** 0.5, bq=rn(1, 9 * C) * 0.2)
    return {k: v.cuda() for k, v in t.items()}


def _ref(t, eps=1e-5):
    bf = lambda x: x.to(torch.bfloat16).float()
    x1 = bf(bf(t["x"])[:, 0] @ bf(t["wi"])[0].T + t["bi"][0])          # x1 is stored (and re-read) as bf16
    mu = x1.mean(-1, keepdim=True)
    var = x1.var(-1, unbiased=False, keepdim=True)
    xh = (x1 - mu) * torch.rsqrt(var + eps)
    return x1, xh @ bf(t["wq"])[0].T + t["bq"][0]


def _run(t):
    from tfswa_unet_b200 import ops
    b = lambda x: x.to(torch.bfloat16).contiguous()
    return ops.block_head_tc(b(t["x"]), b(t["wi"]), b(t["wq"]), t["bi"].contiguous(), t["bq"].contiguous())


@pytest.mark.parametrize("C", [32, 64])
@pytest.mark.parametrize("M", [1, 127, 128, 1000, 128 * 700 + 5])
def test_head_vs_fp32(C, M):
    t = _case(M, C, seed=M % 97)
    x1, qkv = _run(t)
    r1, rq = _ref(t)
    assert x1.shape == (M, 1, C) and qkv.shape == (M, 1, 9 * C)
    assert torch.isfinite(qkv.float()).all()
    assert rel_l2(x1[:, 0].float(), r1) < 5e-3
    assert rel_l2(qkv[:, 0].float(), rq) < 8e-3
    if M > 128:                                   # every tile, not just the average
        nt = M // 128
        d = ((qkv[: nt * 128, 0].float() - rq[: nt * 128]) ** 2).reshape(nt, -1).sum(1).sqrt()
        n = (rq[: nt * 128] ** 2).reshape(nt, -1).sum(1).sqrt()
        assert (d / n).max() < 2e-2


@pytest.mark.parametrize("C", [32, 64])
def test_head_matches_unfused_sequence(C):
    from tfswa_unet_b200 import _lib as L
    from tfswa_unet_b200 import functional as Fn
    t = _case(3000, C, seed=4)
    x = t["x"].to(torch.bfloat16)
    inp, qk = Fn.LinW(t["wi"], t["bi"]), Fn.LinW(t["wq"], t["bq"])
    with torch.no_grad():
        assert Fn.fused_head_ok(x, inp, qk)
        x1f, qf = Fn.block_head(x, inp, qk)
        x1u = Fn.linear(x, inp)
        qu = Fn.linear(x1u, qk, prologue=L.PRO_LNHAT, row_stats=Fn.row_stats(x1u))
    assert rel_l2(x1f.float(), x1u.float()) < 3e-3
    assert rel_l2(qf.float(), qu.float()) < 8e-3


def test_head_rejects_other_widths():
    t = _case(64, 128)
    with pytest.raises(RuntimeError, match="not in"):
        _run(t)
