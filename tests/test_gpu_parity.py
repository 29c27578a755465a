"""GPU (-m gpu): the CUDA path, called through the C ABI via the module mirror, against
(a) the golden vectors produced by the live reference and (b) the oracle on fresh seeded inputs.

Tolerances (parity is unpinned by the reference's own tests, so they are ours to state):
  fp32 path : max-abs error <= 2e-4 * max|ref| on outputs / logits
  bf16 path : relative L2 error <= 2e-2 on branch/block outputs and logits, masks max-abs <= 2e-2
"""
import pytest
import torch

from oracle import tfswa_oracle as O
from helpers import seeded, assert_close, rel_l2

pytestmark = pytest.mark.gpu

FP32_TOL = 2e-4
BF16_L2 = 2e-2


def _T():
    import tfswa_unet_b200 as T
    return T


def _build(kind, C, shift=0, cout=0, cin=2):
    T = _T()
    if kind == "tsa":
        return T.TemporalSequenceAttention(C, 8)
    if kind == "fsa":
        return T.FrequencySequenceAttention(C, 8)
    if kind == "swa":
        return T.ShiftedWindowAttention(C, 8, 8, shift)
    if kind == "block":
        return T.TFSWABlock(C, C, 8, shift, 8)
    if kind == "down":
        return T.DownsampleBlock(C, cout)
    if kind == "up":
        return T.UpsampleBlock(C, cout)
    return T.TFSWAUNet(cin, cout, [2, 2, 6, 2], [32, 64, 128, 256], 8, 4, 8)


def _filled(kind, C, seed, gain=1.0, **kw):
    m = _build(kind, C, **kw)
    sd = m.state_dict()
    O.randomize_state_(sd, seed, gain)
    m.load_state_dict(sd)
    return m, sd


def _check(name, got, ref, precision):
    if precision == "fp32":
        assert_close(name, got, ref, FP32_TOL)
    else:
        e = rel_l2(got.float(), ref)
        assert e <= BF16_L2, f"{name}: bf16 rel-L2 {e:.3e} > {BF16_L2}"


BRANCH_CASES = [("tsa_c32", "tsa", 32, 0), ("fsa_c32", "fsa", 32, 0), ("swa_c32_s0", "swa", 32, 0),
                ("swa_c32_s4", "swa", 32, 4), ("tsa_c64", "tsa", 64, 0), ("fsa_c128", "fsa", 128, 0),
                ("swa_c256_s4", "swa", 256, 4)]


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name,kind,C,shift", BRANCH_CASES)
def test_branch_forward_golden(golden, name, kind, C, shift, precision):
    T = _T()
    T.set_precision(precision)
    case = golden["cases"][name]
    m, _ = _filled(kind, C, case["seed"], shift=shift)
    m.eval().cuda()
    x = seeded(case["shape"], case["seed"] + 100).cuda()
    with torch.no_grad():
        y = m(x)
    assert y.shape == x.shape and y.dtype == x.dtype
    _check(name, y, case["y"], precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["block_c32_s0_eval", "block_c32_s4_skip_eval", "block_c32_s4_train", "block_c64_s4_skip_train"])
def test_block_forward_golden(golden, name, precision):
    T = _T()
    T.set_precision(precision)
    case = golden["cases"][name]
    C = case["shape"][1]
    m, _ = _filled("block", C, case["seed"], shift=case["shift"])
    m.train(case["train"]).cuda()
    x = seeded(case["shape"], case["seed"] + 100).cuda()
    skip = seeded(case["shape"], case["seed"] + 200).cuda() if case["with_skip"] else None
    with torch.no_grad():
        y = m(x, skip=skip) if skip is not None else m(x)
    _check(name, y, case["y"], precision)
    if case["train"]:
        bufs = dict(m.named_buffers())
        for k, v in case["buffers"].items():
            tol = 1e-4 if precision == "fp32" else 2e-2
            assert_close(f"{name}.{k}", bufs[k].float(), v.float(), tol, atol=1e-3 if precision == "bf16" else 0.0)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name,kind,cin,cout", [("down_32_64_eval", "down", 32, 64), ("down_64_128_train", "down", 64, 128),
                                                ("up_64_32_eval", "up", 64, 32), ("up_128_64_train", "up", 128, 64)])
def test_resample_forward_golden(golden, name, kind, cin, cout, precision):
    T = _T()
    T.set_precision(precision)
    case = golden["cases"][name]
    m, _ = _filled(kind, cin, case["seed"], cout=cout)
    m.train(case["train"]).cuda()
    x = seeded(case["shape"], case["seed"] + 100).cuda()
    with torch.no_grad():
        y = m(x)
    _check(name, y, case["y"], precision)


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
@pytest.mark.parametrize("name", ["unet_65x41_eval", "unet_64x96_train"])
def test_unet_forward_golden(golden, name, precision):
    T = _T()
    T.set_precision(precision)
    case = golden["cases"][name]
    m, _ = _filled("unet", 32, case["seed"], gain=case["gain"], cin=case["cin"], cout=case["cout"])
    m.train(case["train"]).cuda()
    x = seeded(case["shape"], case["seed"] + 100).cuda()
    with torch.no_grad():
        masks, logits = m(x, return_logits=True)
    assert masks.shape == case["y"].shape and masks.dtype == torch.float32
    if precision == "fp32":
        assert_close(name + ".logits", logits, case["logits"], 5e-4)
        assert_close(name + ".masks", masks, case["y"], 2e-4)
    else:
        e = rel_l2(logits, case["logits"])
        assert e <= 3e-2, f"{name}: bf16 logits rel-L2 {e:.3e}"
        assert float((masks.cpu() - case["y"]).abs().max()) <= 2e-2


_CONFIG0_REF = {}


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_unet_config0_shape_vs_oracle(precision):
    """BASELINE.json configs[0] (the reference's tests/test_model.py forward): TFSWA-UNet 15.4 M params, input
    (2,2,256,512) - the fp32 kernels AND the bf16 tensor-core path against the CPU oracle (fp32, computed once) on the
    same seeded weights and input."""
    T = _T()
    try:
        T.set_precision(precision)
        m, sd = _filled("unet", 32, 5, gain=0.7, cin=2, cout=2)
        m.eval().cuda()
        assert sum(p.numel() for p in m.parameters()) == 15404834
        x = seeded((2, 2, 256, 512), 6)
        if "ref" not in _CONFIG0_REF:
            with torch.no_grad():
                _CONFIG0_REF["ref"] = O.unet_forward(x, sd, return_logits=True)
        ref_logits = _CONFIG0_REF["ref"]
        ref_masks = torch.sigmoid(ref_logits)
        with torch.no_grad():
            masks, logits = m(x.cuda(), return_logits=True)
        assert masks.shape == (2, 2, 256, 512) and masks.dtype == torch.float32
        if precision == "fp32":
            assert_close("config0.masks", masks, ref_masks, 2e-4)
            assert_close("config0.logits", logits, ref_logits, 2e-4)
        else:
            e = rel_l2(logits.float(), ref_logits)
            assert e <= BF16_L2, f"config0 bf16 logits rel-L2 {e:.3e}"
            assert float((masks.cpu() - ref_masks).abs().max()) <= 2e-2
    finally:
        T.set_precision("bf16")


def test_gradient_checkpointing_over_tfswa_blocks():
    """the reference's enable_gradient_checkpointing (src/optimization/gradient_checkpoint.py:44-69) finds
    modules whose class name contains 'TFSWABlock', replaces their instance `forward` with
    torch.utils.checkpoint.checkpoint(original_forward, ..., use_reentrant=False) in training mode.  The same patch applied
    to the product model must give the same loss and gradients as the unpatched model (recompute = same kernels)."""
    from torch.utils.checkpoint import checkpoint
    T = _T()
    try:
        T.set_precision("fp32")
        x = seeded((1, 2, 40, 24), 91).cuda()
        wgt = seeded((1, 2, 40, 24), 92).cuda()
        res = {}
        for mode in ("plain", "checkpointed"):
            m, _ = _filled("unet", 32, 5, gain=0.7, cin=2, cout=2)
            m.train().cuda()
            patched = 0
            if mode == "checkpointed":
                for mod in m.modules():
                    if "TFSWABlock" in type(mod).__name__:
                        orig = mod.forward

                        def fwd(*a, _orig=orig, _mod=mod, **kw):
                            return checkpoint(_orig, *a, **kw, use_reentrant=False) if _mod.training else _orig(*a, **kw)
                        mod.forward = fwd
                        patched += 1
                assert patched == 22
            y = m(x)
            loss = (y * wgt).sum()
            loss.backward()
            torch.cuda.synchronize()
            res[mode] = (loss.detach(), {k: p.grad.detach().clone() for k, p in m.named_parameters()})
        (l0, g0), (l1, g1) = res["plain"], res["checkpointed"]
        assert abs(float(l0) - float(l1)) <= 1e-5 * abs(float(l0)) + 1e-6
        for k in g0:
            # atol: a conv bias that feeds a train-mode BatchNorm has an analytically zero gradient (fp32 noise ~1e-6)
            assert_close(f"ckpt.{k}", g1[k], g0[k], 1e-4, atol=3e-5)
    finally:
        T.set_precision("bf16")


@pytest.mark.parametrize("shape,C,shift", [((1, 32, 65, 41), 32, 4), ((2, 64, 33, 50), 64, 4), ((1, 128, 24, 40), 128, 0),
                                           ((1, 256, 16, 24), 256, 4), ((1, 32, 8, 8), 32, 4), ((1, 32, 130, 9), 32, 4)])
def test_block_vs_oracle_odd_sizes(shape, C, shift):
    """fresh seeds, odd sizes (padding + wrap-around windows, multi-tile sequences), fp32 path vs CPU oracle"""
    T = _T()
    T.set_precision("fp32")
    m, sd = _filled("block", C, 77, shift=shift)
    m.eval().cuda()
    x = seeded(shape, 78)
    skip = seeded(shape, 79)
    with torch.no_grad():
        y = m(x.cuda(), skip=skip.cuda())
        ref = O.tfswa_block(x, sd, shift, skip=skip)
    assert_close("block", y, ref, FP32_TOL)


def test_swin_mask_and_relbias_feature_vs_oracle():
    """flag-gated (default OFF) Swin shift mask + relative-position bias inside the window kernel"""
    T = _T()
    T.set_precision("fp32")
    m, sd = _filled("swa", 32, 91, shift=4)
    m.eval().cuda()
    x = seeded((2, 32, 20, 27), 92)
    rb = seeded((8, 64, 64), 93, 0.5)
    m.use_shift_mask = True
    m.rel_bias = rb.cuda().contiguous()
    with torch.no_grad():
        y = m(x.cuda())
        ref = O.swa(x, sd, 8, 4, 8, use_mask=True, rel_bias=rb)
    assert_close("swa+mask+bias", y, ref, FP32_TOL)
    # and the default path ignores the registered attn_mask buffer exactly like the reference
    m.use_shift_mask = False
    m.rel_bias = None
    m.attn_mask.fill_(123.0)
    with torch.no_grad():
        y2 = m(x.cuda())
    assert_close("swa default", y2, O.swa(x, sd, 8, 4, 8), FP32_TOL)


def test_mha_module_vs_oracle():
    T = _T()
    T.set_precision("fp32")
    mha = T.MultiHeadAttention(64, 8)
    sd = mha.state_dict()
    O.randomize_state_(sd, 5)
    mha.load_state_dict(sd)
    mha.cuda().eval()
    x = seeded((3, 50, 64), 6)
    with torch.no_grad():
        y = mha(x.cuda())
    assert_close("mha", y, O.mha(x, sd, 8), FP32_TOL)


def test_convert_shares_parameters_and_matches():
    T = _T()
    T.set_precision("fp32")
    a, sd = _filled("unet", 32, 3, gain=0.7, cin=2, cout=2)
    a.eval().cuda()
    b = T.convert(a)
    assert all(p is q for p, q in zip(a.parameters(), b.parameters()))
    x = seeded((1, 2, 32, 24), 4).cuda()
    with torch.no_grad():
        assert torch.equal(a(x), b(x))


def test_host_pipeline_matches_plain_forward():
    """HostPipeline (double-buffered H2D / D2H on copy streams around the eval forward) returns, for every step and in order,
    exactly what the plain forward returns - different inputs per step so that a buffer race would show."""
    import tfswa_unet_b200 as T
    torch.manual_seed(3)
    model = T.TFSWAUNet(2, 2, [1, 1, 1, 1], [32, 64, 128, 256], 8, 4, 8).cuda().eval()
    xs = [torch.randn(2, 2, 72, 40).pin_memory() for _ in range(5)]
    outs = [torch.empty(2, 2, 72, 40).pin_memory() for _ in range(5)]
    pipe = T.HostPipeline(model)
    for x, o in zip(xs, outs):
        pipe.step(x, o)
    pipe.flush()
    torch.cuda.synchronize()
    with torch.no_grad():
        for x, o in zip(xs, outs):
            assert torch.equal(o, model(x.cuda()).cpu())
