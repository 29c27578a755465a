"""GPU (-m gpu): gradients of the CUDA path (autograd.Functions over C-ABI kernels) against the golden gradients
produced by the live reference, for branches, blocks (eval + train-mode BatchNorm), down/up-sampling and the full model.

Tolerances: fp32 path max-abs <= 1e-3 * max|ref| per tensor (fp32 atomics reorder sums); bf16 path relative L2 <= 6e-2
per tensor (bf16 activations AND bf16 activation-gradients), with an absolute floor for analytically-zero gradients."""
import pytest
import torch

from oracle import tfswa_oracle as O
from helpers import seeded, assert_close, rel_l2, unpack_grad

pytestmark = pytest.mark.gpu


def _T():
    import tfswa_unet_b200 as T
    return T


def _build(kind, C, shift=0, cout=0, cin=2):
    T = _T()
    return {"tsa": lambda: T.TemporalSequenceAttention(C, 8), "fsa": lambda: T.FrequencySequenceAttention(C, 8),
            "swa": lambda: T.ShiftedWindowAttention(C, 8, 8, shift), "block": lambda: T.TFSWABlock(C, C, 8, shift, 8),
            "down": lambda: T.DownsampleBlock(C, cout), "up": lambda: T.UpsampleBlock(C, cout),
            "unet": lambda: T.TFSWAUNet(cin, cout, [2, 2, 6, 2], [32, 64, 128, 256], 8, 4, 8)}[kind]()


def _filled(kind, C, seed, gain=1.0, **kw):
    m = _build(kind, C, **kw)
    sd = m.state_dict()
    O.randomize_state_(sd, seed, gain)
    m.load_state_dict(sd)
    return m


def _fwd_bwd(m, x, seed, skip=None):
    x = x.clone().requires_grad_(True)
    y = m(x, skip=skip) if skip is not None else m(x)
    w = seeded(y.shape, seed + 7).cuda()
    (y.float() * w).sum().backward()
    torch.cuda.synchronize()
    return y, x.grad, {k: p.grad for k, p in m.named_parameters()}


def _cmp(name, got, ref_entry, precision, atol_scale):
    ref, stride = unpack_grad(ref_entry)
    g = got.detach().float().flatten()[::stride] if stride else got.detach().float()
    g, ref = g.cpu().double().reshape(-1), ref.double().reshape(-1)
    assert torch.isfinite(g).all(), f"{name}: non-finite gradient"
    if precision == "fp32":
        err, scale = float((g - ref).abs().max()), float(ref.abs().max())
        assert err <= 1e-3 * scale + 1e-4 * atol_scale, f"{name}: fp32 grad err {err:.3e} (max|ref| {scale:.3e})"
    else:
        err = float((g - ref).norm())
        assert err <= 6e-2 * float(ref.norm()) + 2e-3 * atol_scale * ref.numel() ** 0.5, \
            f"{name}: bf16 grad rel-L2 {err / max(float(ref.norm()), 1e-30):.3e}"


def _check_all(name, case, y, dx, grads, precision):
    if precision == "fp32":
        assert_close(name + ".y", y, case["y"], 2e-4)
    else:
        assert rel_l2(y.float(), case["y"]) <= 2e-2
    # scale for the absolute floor: typical gradient magnitude of this case
    atol_scale = float(unpack_grad(next(iter(case["grads"].values())))[0].abs().max()) if case["grads"] else 1.0
    atol_scale = max(atol_scale, float(case["dx"].abs().max()))
    _cmp(name + ".dx", dx, case["dx"], precision, atol_scale)
    assert set(grads) == set(case["grads"]), set(grads) ^ set(case["grads"])
    for k, g in grads.items():
        assert g is not None, f"{name}.{k}: no gradient"
        _cmp(f"{name}.{k}", g, case["grads"][k], precision, atol_scale)


BRANCH_CASES = [("tsa_c32", "tsa", 32, 0), ("fsa_c32", "fsa", 32, 0), ("swa_c32_s0", "swa", 32, 0),
                ("swa_c32_s4", "swa", 32, 4), ("tsa_c64", "tsa", 64, 0), ("fsa_c128", "fsa", 128, 0),
                ("swa_c256_s4", "swa", 256, 4)]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name,kind,C,shift", BRANCH_CASES)
def test_branch_backward_golden(golden, name, kind, C, shift, precision):
    T = _T()
    T.set_precision(precision)
    case = golden["cases"][name]
    m = _filled(kind, C, case["seed"], shift=shift).eval().cuda()
    y, dx, grads = _fwd_bwd(m, seeded(case["shape"], case["seed"] + 100).cuda(), case["seed"])
    _check_all(name, case, y, dx, grads, precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["block_c32_s0_eval", "block_c32_s4_skip_eval", "block_c32_s4_train", "block_c64_s4_skip_train"])
def test_block_backward_golden(golden, name, precision):
    T = _T()
    T.set_precision(precision)
    case = golden["cases"][name]
    m = _filled("block", case["shape"][1], case["seed"], shift=case["shift"]).train(case["train"]).cuda()
    skip = seeded(case["shape"], case["seed"] + 200).cuda() if case["with_skip"] else None
    y, dx, grads = _fwd_bwd(m, seeded(case["shape"], case["seed"] + 100).cuda(), case["seed"], skip=skip)
    _check_all(name, case, y, dx, grads, precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name,kind,cin,cout", [("down_32_64_eval", "down", 32, 64), ("down_64_128_train", "down", 64, 128),
                                                ("up_64_32_eval", "up", 64, 32), ("up_128_64_train", "up", 128, 64)])
def test_resample_backward_golden(golden, name, kind, cin, cout, precision):
    T = _T()
    T.set_precision(precision)
    case = golden["cases"][name]
    m = _filled(kind, cin, case["seed"], cout=cout).train(case["train"]).cuda()
    y, dx, grads = _fwd_bwd(m, seeded(case["shape"], case["seed"] + 100).cuda(), case["seed"])
    _check_all(name, case, y, dx, grads, precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["unet_65x41_eval", "unet_64x96_train"])
def test_unet_backward_golden(golden, name, precision):
    T = _T()
    T.set_precision(precision)
    case = golden["cases"][name]
    m = _filled("unet", 32, case["seed"], gain=case["gain"], cin=case["cin"], cout=case["cout"]).train(case["train"]).cuda()
    x = seeded(case["shape"], case["seed"] + 100).cuda().requires_grad_(True)
    y = m(x)
    w = seeded(y.shape, case["seed"] + 7).cuda()
    (y * w).sum().backward()
    torch.cuda.synchronize()
    grads = {k: p.grad for k, p in m.named_parameters()}
    assert set(grads) == set(case["grads"])
    if precision == "fp32":
        assert_close(name + ".y", y, case["y"], 2e-4)
        assert_close(name + ".dx", x.grad, case["dx"], 2e-3)
        worst = 0.0
        for k, g in grads.items():
            ref, stride = unpack_grad(case["grads"][k])
            gg = g.detach().float().flatten()[::stride].cpu() if stride else g.detach().float().cpu()
            err = float((gg.double().reshape(-1) - ref.double().reshape(-1)).abs().max())
            scale = float(ref.abs().max())
            worst = max(worst, err / (scale + 1e-3))
            assert err <= 3e-3 * scale + 3e-4, f"{name}.{k}: err {err:.3e} scale {scale:.3e}"
    else:
        assert float((y.cpu() - case["y"]).abs().max()) <= 2e-2
        # whole-gradient relative L2 over all 936 parameter tensors (per-tensor bf16 noise is dominated by tiny tensors)
        num = den = 0.0
        for k, g in grads.items():
            ref, stride = unpack_grad(case["grads"][k])
            gg = g.detach().float().flatten()[::stride].cpu() if stride else g.detach().float().cpu()
            num += float((gg.double().reshape(-1) - ref.double().reshape(-1)).pow(2).sum())
            den += float(ref.double().pow(2).sum())
        assert (num / den) ** 0.5 <= 8e-2, f"{name}: bf16 global grad rel-L2 {(num / den) ** 0.5:.3e}"
    if case["train"]:
        bufs = dict(m.named_buffers())
        for k, v in case["buffers"].items():
            assert_close(f"{name}.{k}", bufs[k].float(), v.float(), 2e-4 if precision == "fp32" else 3e-2,
                         atol=0.0 if precision == "fp32" else 2e-3)


@pytest.mark.parametrize("M,nb,C,r1_nb", [(3000, 3, 32, 1), (50000, 3, 32, 1), (40000, 1, 64, 1), (9000, 3, 128, 3)])
def test_fused_gelu_backward_matches_two_node_form_and_torch(M, nb, C, r1_nb):
    """fc2 over GELU(u): the one-node form whose data-gradient GEMM multiplies by gelu'(u) in its epilogue (TFSWA_EPI_MUL_DGELU;
    large M: persistent kernel with u arriving as the residual tile, small M: one-tile kernel) against the GeluFn + LinearFn pair
    and against torch autograd in fp32 on the same bf16-rounded inputs (attention.py:121-128)."""
    import torch.nn.functional as TF
    from tfswa_unet_b200 import functional as Fn
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(5)
    u0 = torch.randn((M, nb, 4 * C), device=dev, generator=g).to(torch.bfloat16)
    w0 = torch.randn((nb, C, 4 * C), device=dev, generator=g) / (4 * C) ** 0.5
    b0 = 0.1 * torch.randn((nb, C), device=dev, generator=g)
    r0 = torch.randn((M, r1_nb, C), device=dev, generator=g).to(torch.bfloat16)
    dy = torch.randn((M, nb, C), device=dev, generator=g).to(torch.bfloat16)
    res = {}
    for fused in (True, False):
        Fn.USE_FUSED_GELU_BWD = fused
        try:
            u = u0.clone().requires_grad_(True)
            w = w0.clone().requires_grad_(True)
            b = b0.clone().requires_grad_(True)
            r = r0.clone().requires_grad_(True)
            y = Fn.gelu_linear(u, Fn.LinW(w, b), r1=r)
            y.backward(dy)
            res[fused] = (y.detach().float(), u.grad.float(), w.grad.float(), b.grad.float(), r.grad.float())
        finally:
            Fn.USE_FUSED_GELU_BWD = True
    # torch reference (fp32 math on the bf16 inputs; h rounded to bf16 as both product forms store it)
    u = u0.float().requires_grad_(True)
    w = w0.to(torch.bfloat16).float().requires_grad_(True)
    b = b0.clone().requires_grad_(True)
    r = r0.float().requires_grad_(True)
    h = TF.gelu(u)
    y = torch.einsum("mbk,bnk->mbn", h, w) + b[None] + r
    y.backward(dy.float())
    ref = (y.detach(), u.grad, w.grad, b.grad, r.grad)
    names = ("y", "du", "dw", "db", "dr1")
    for i, n in enumerate(names):
        a, t, rr = res[True][i], res[False][i], ref[i]
        rel_pair = float((a - t).norm() / t.norm())
        rel_ref = float((a - rr).norm() / rr.norm())
        # du: one bf16 rounding in the fused form, two in the pair (dh is rounded before the GELU gradient multiplies it)
        assert rel_pair <= (6e-3 if n == "du" else 1e-5), (n, rel_pair)
        assert rel_ref <= 8e-3, (n, rel_ref)
    # the fused form is the more accurate one for du
    assert float((res[True][1] - ref[1]).norm()) <= 1.05 * float((res[False][1] - ref[1]).norm())
