"""Fused branch tail (tfswa_branch_tail_tc_fwd) against an fp32 torch restatement of attention.py:86,121-128 with the
residuals of :146/:159, and against the unfused tensor-core sequence it replaces."""
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu


def _case(M, nb, C, res_nb, seed=0, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    hid = 4 * C
    t = dict(att=rn(M, nb, C) * scale, res=rn(M, res_nb, C) * scale + 0.3,
             wp=rn(nb, C, C) / C ** 0.5, bp=rn(nb, C) * 0.1,
             w1=rn(nb, hid, C) / C ** 0.5, b1=rn(nb, hid) * 0.1,
             w2=rn(nb, C, hid) / hid ** 0.5, b2=rn(nb, C) * 0.1)
    return {k: v.cuda() for k, v in t.items()}


def _ref(t, eps=1e-5):
    """fp32 math on the bf16-rounded inputs/weights the kernel sees"""
    bf = lambda x: x.to(torch.bfloat16).float()
    att, res = bf(t["att"]), bf(t["res"])
    y = torch.einsum("mbk,bnk->mbn", att, bf(t["wp"])) + t["bp"][None] + res
    mu = y.mean(-1, keepdim=True)
    var = y.var(-1, unbiased=False, keepdim=True)
    yh = (y - mu) * torch.rsqrt(var + eps)
    h = torch.nn.functional.gelu(torch.einsum("mbk,bnk->mbn", yh, bf(t["w1"])) + t["b1"][None])
    return y + torch.einsum("mbk,bnk->mbn", h, bf(t["w2"])) + t["b2"][None]


def _run(t):
    from tfswa_unet_b200 import ops
    b = lambda x: x.to(torch.bfloat16).contiguous()
    return ops.branch_tail_tc(b(t["att"]), b(t["res"]), b(t["wp"]), b(t["w1"]), b(t["w2"]),
                              t["bp"].contiguous(), t["b1"].contiguous(), t["b2"].contiguous())


@pytest.mark.parametrize("C", [32, 64, 128])
@pytest.mark.parametrize("M,nb,res_nb", [(1, 1, 1), (127, 3, 1), (128, 1, 1), (1000, 3, 1), (1000, 3, 3), (333, 2, 2)])
def test_tail_vs_fp32(C, M, nb, res_nb):
    t = _case(M, nb, C, res_nb, seed=M + nb)
    out = _run(t).float()
    ref = _ref(t)
    assert out.shape == ref.shape
    assert torch.isfinite(out).all()
    # bf16 operands between the three GEMMs (LN_hat(y), hidden) and a bf16 result
    assert rel_l2(out, ref) < 8e-3
    assert (out - ref).abs().max() < 0.08 * max(1.0, ref.abs().max().item() / 4)


@pytest.mark.parametrize("C", [32, 64, 128])
def test_tail_many_tiles_per_cta(C):
    """more tiles than the persistent grid holds: exercises the prefetch double buffer and barrier phases"""
    M = 128 * (1500 if C < 128 else 700) + 77
    t = _case(M, 3, C, 1, seed=5)
    out = _run(t).float()
    ref = _ref(t)
    assert rel_l2(out, ref) < 8e-3
    # every tile, not just the average: per-tile relative error
    nt = M // 128
    d = ((out[: nt * 128] - ref[: nt * 128]) ** 2).reshape(nt, -1).sum(1).sqrt()
    n = (ref[: nt * 128] ** 2).reshape(nt, -1).sum(1).sqrt()
    assert (d / n).max() < 2e-2


@pytest.mark.parametrize("C", [32, 64, 128])
def test_tail_matches_unfused_sequence(C):
    """same rounding points as proj -> row_stats -> fc1(GELU) -> fc2 on tfswa_linear_tc_fwd, except that y stays fp32"""
    from tfswa_unet_b200 import _lib as L
    from tfswa_unet_b200 import functional as Fn
    M, nb = 2000, 3
    t = _case(M, nb, C, 1, seed=9)
    att, res = t["att"].to(torch.bfloat16), t["res"].to(torch.bfloat16)
    proj, fc1, fc2 = Fn.LinW(t["wp"], t["bp"]), Fn.LinW(t["w1"], t["b1"]), Fn.LinW(t["w2"], t["b2"])
    with torch.no_grad():
        assert Fn.fused_tail_ok(att, res, proj, fc1, fc2)
        fused = Fn.branch_tail(att, res, proj, fc1, fc2).float()
        y = Fn.linear(att, proj, r1=res)
        st = Fn.row_stats(y)
        h = Fn.linear(y, fc1, prologue=L.PRO_LNHAT, epilogue=L.EPI_GELU, row_stats=st)
        unfused = Fn.linear(h, fc2, r1=y).float()
    assert rel_l2(fused, unfused) < 8e-3


def test_tail_rejects_other_widths():
    from tfswa_unet_b200 import ops
    t = _case(64, 1, 256, 1)
    with pytest.raises(RuntimeError, match="not in"):
        _run(t)
    t = _case(64, 1, 32, 1)
    t["w1"] = t["w1"][:, :64].contiguous(); t["b1"] = t["b1"][:, :64].contiguous(); t["w2"] = t["w2"][:, :, :64].contiguous()
    with pytest.raises(RuntimeError, match="4\\*C"):
        _run(t)


def test_fused_tail_not_used_under_autograd():
    from tfswa_unet_b200 import functional as Fn
    t = _case(64, 1, 32, 1)
    att = t["att"].to(torch.bfloat16).requires_grad_(True)
    proj, fc1, fc2 = Fn.LinW(t["wp"], t["bp"]), Fn.LinW(t["w1"], t["b1"]), Fn.LinW(t["w2"], t["b2"])
    assert not Fn.fused_tail_ok(att, t["res"].to(torch.bfloat16), proj, fc1, fc2)
