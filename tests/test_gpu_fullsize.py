"""GPU (-m gpu): the path at BASELINE.json's full sizes (6 s / n_fft 2048 segments -> 1025 x 517 maps), where the CPU oracle
would take minutes per case: size-independent properties instead of element-wise goldens.

  * sampled sequences of the full-size axial attention launches against a torch fp32 softmax of the same rows
  * equivariance: permuting the sequences of an axial attention launch permutes its output bit for bit
  * window attention at full size against the same op on a crop that contains whole windows
  * whole model, batch 8: masks finite and in [0,1], every sample equal to the same sample run alone
  * whole model: the bf16 path against the fp32 parity path (which the small-size tests pin to the oracle)
  * one training step at full size: finite loss, finite and non-trivial gradients for every parameter
"""
import pytest
import torch

from helpers import seeded

pytestmark = pytest.mark.gpu
H, W = 1025, 517
MODEL = dict(depths=[2, 2, 6, 2], dims=[32, 64, 128, 256], window_size=8, shift_size=4, num_heads=8)


def _ref_rows(qkv, rows, N, stride, base, C, heads):
    """torch fp32 attention of the given sequences: token index of element n of row r = base(r) + n*stride"""
    d = C // heads
    out = {}
    for r in rows:
        idx = base(r) + torch.arange(N, device=qkv.device) * stride
        t = qkv[idx].float().view(N, 3, heads, d)
        q, k, v = t[:, 0].transpose(0, 1), t[:, 1].transpose(0, 1), t[:, 2].transpose(0, 1)      # (h, N, d)
        o = torch.softmax((q @ k.transpose(-1, -2)) * d ** -0.5, -1) @ v
        out[r] = (idx, o.transpose(0, 1).reshape(N, C))
    return out


@pytest.mark.parametrize("geom,C,h,w", [(0, 32, H, W), (1, 32, H, W), (0, 64, 512, 258), (1, 64, 512, 258),
                                         (0, 128, 256, 129), (1, 128, 256, 129)])
def test_fullsize_axial_attention_sampled_sequences(geom, C, h, w):
    from tfswa_unet_b200 import ops
    B, heads = 2, 8
    M = B * h * w
    qkv = seeded((M, 3 * C), 900 + geom + C, 1.0).cuda().to(torch.bfloat16)
    out = torch.empty((M, C), dtype=torch.bfloat16, device="cuda")
    ops.attention(qkv, out, B, h, w, C, heads, geom)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    if geom == 0:       # TSA: sequence r = (b, column), along H
        N, stride, nrows = h, w, B * w
        base = lambda r: (r // w) * h * w + (r % w)
    else:               # FSA: sequence r = (b, row), along W
        N, stride, nrows = w, 1, B * h
        base = lambda r: r * w
    rows = [0, 1, nrows // 2, nrows - 1]
    for r, (idx, ref) in _ref_rows(qkv, rows, N, stride, base, C, heads).items():
        got = out[idx].float()
        rel = float((got - ref).norm() / ref.norm())
        assert rel <= 1.5e-2, f"geom {geom} C {C} sequence {r}: rel-L2 {rel:.3e}"          # bf16 P and output rounding


@pytest.mark.parametrize("geom", [0, 1])
def test_fullsize_axial_attention_is_equivariant_under_sequence_permutation(geom):
    from tfswa_unet_b200 import ops
    B, C, heads = 1, 32, 8
    M = B * H * W
    qkv = seeded((M, 3 * C), 910 + geom, 1.0).cuda().to(torch.bfloat16)
    out = torch.empty((M, C), dtype=torch.bfloat16, device="cuda")
    ops.attention(qkv, out, B, H, W, C, heads, geom)
    g = torch.Generator().manual_seed(5)
    x4 = qkv.view(B, H, W, 3 * C)
    if geom == 0:       # sequences are the W columns
        perm = torch.randperm(W, generator=g).cuda()
        qp = x4[:, :, perm].contiguous().view(M, 3 * C)
    else:               # sequences are the H rows
        perm = torch.randperm(H, generator=g).cuda()
        qp = x4[:, perm].contiguous().view(M, 3 * C)
    outp = torch.empty_like(out)
    ops.attention(qp, outp, B, H, W, C, heads, geom)
    torch.cuda.synchronize()
    o4, p4 = out.view(B, H, W, C), outp.view(B, H, W, C)
    expect = o4[:, :, perm] if geom == 0 else o4[:, perm]
    assert torch.equal(p4, expect), "a sequence's result must not depend on where it sits in the launch"


def test_fullsize_window_attention_matches_crop_of_whole_windows():
    """unshifted windows are independent: the top-left 512 x 256 crop (whole windows) must reproduce the same tokens"""
    from tfswa_unet_b200 import ops
    B, C, heads = 1, 32, 8
    qkv = seeded((B * H * W, 3 * C), 920, 1.0).cuda().to(torch.bfloat16)
    pad_kv = seeded((2 * C,), 921, 0.5).cuda().float().contiguous()
    out = torch.empty((B * H * W, C), dtype=torch.bfloat16, device="cuda")
    ops.attention(qkv, out, B, H, W, C, heads, 2, ws=8, shift=0, pad_kv=pad_kv)
    h2, w2 = 512, 256
    crop = qkv.view(B, H, W, 3 * C)[:, :h2, :w2].contiguous().view(B * h2 * w2, 3 * C)
    outc = torch.empty((B * h2 * w2, C), dtype=torch.bfloat16, device="cuda")
    ops.attention(crop, outc, B, h2, w2, C, heads, 2, ws=8, shift=0, pad_kv=pad_kv)
    torch.cuda.synchronize()
    assert torch.equal(out.view(B, H, W, C)[:, :h2, :w2].contiguous().view(-1, C), outc)
    assert torch.isfinite(out.float()).all()


def test_fullsize_model_batch8_samples_are_independent():
    import tfswa_unet_b200 as T
    T.set_precision("bf16")
    torch.manual_seed(0)
    model = T.TFSWAUNet(2, 2, **MODEL).eval().cuda()
    x = seeded((8, 2, H, W), 930, 1.0).cuda()
    with torch.no_grad():
        y = model(x)
        y3 = model(x[3:4].contiguous())
        y7 = model(x[7:8].contiguous())
    torch.cuda.synchronize()
    assert y.shape == x.shape and y.dtype == torch.float32
    assert torch.isfinite(y).all() and float(y.min()) >= 0.0 and float(y.max()) <= 1.0
    assert float(y.std()) > 1e-3, "degenerate masks"
    # eval-mode samples never meet (BatchNorm uses running statistics); tiles and CTAs are assigned differently at B=1
    assert float((y[3:4] - y3).abs().max()) <= 1e-6 and float((y[7:8] - y7).abs().max()) <= 1e-6


def test_fullsize_model_bf16_path_against_fp32_path():
    import tfswa_unet_b200 as T
    from oracle import tfswa_oracle as O
    try:
        torch.manual_seed(0)
        model = T.TFSWAUNet(2, 2, **MODEL)
        sd = model.state_dict()
        O.randomize_state_(sd, 31, 0.7)                      # non-trivial LN / BN affine, biases and running statistics
        model.load_state_dict(sd)
        model = model.eval().cuda()
        x = seeded((1, 2, H, W), 940, 1.0).cuda()
        res = {}
        for prec in ("fp32", "bf16"):
            T.set_precision(prec)
            with torch.no_grad():
                res[prec] = model(x, return_logits=True)
        torch.cuda.synchronize()
        (m32, l32), (m16, l16) = res["fp32"], res["bf16"]
        assert torch.isfinite(l32).all() and torch.isfinite(l16).all()
        rel = float((l16 - l32).norm() / l32.norm())
        assert rel <= 2e-2, f"full-size logits bf16 vs fp32 rel-L2 {rel:.3e}"
        assert float((m16 - m32).abs().max()) <= 4e-2
    finally:
        T.set_precision("bf16")


def test_fullsize_training_step_is_finite():
    import tfswa_unet_b200 as T
    from tfswa_unet_b200.train_step import TrainStep
    T.set_precision("bf16")
    torch.manual_seed(0)
    model = T.TFSWAUNet(4, 4, **MODEL).train().cuda()
    before = [p.detach().clone() for p in model.parameters()]
    step = TrainStep(model, lr=1e-3)
    x = seeded((1, 4, H, W), 950, 1.0).cuda()
    mix = seeded((1, H, W), 951, 1.0).abs().cuda()
    tg = [seeded((1, H, W), 952 + i, 1.0).abs().cuda() for i in range(2)]
    loss, norm = step(x, mix, tg)
    torch.cuda.synchronize()
    assert torch.isfinite(loss) and torch.isfinite(norm) and float(norm) > 0
    assert torch.isfinite(step.arena.flat_g).all() and torch.isfinite(step.arena.flat_p).all()
    zero = [n for (n, p) in model.named_parameters() if float(p.grad.abs().max()) == 0.0]
    # a conv bias that feeds a train-mode BatchNorm has an analytically zero gradient; nothing else may be dead
    assert all(".0.bias" in n or n.endswith("downsample.0.bias") or n.endswith("upsample.0.bias") for n in zero), zero
    stuck = [n for (n, p), a in zip(model.named_parameters(), before) if torch.equal(a, p.detach()) and n not in zero]
    assert not stuck, f"parameters with a gradient that did not take an AdamW step: {stuck[:5]}"


def test_fullsize_forward_is_bitwise_reproducible():
    """eval forward has no atomics: any run-to-run difference is a synchronisation bug (this test found a barrier race in the
    tcgen05 attention kernel that corrupted 32 rows of one CTA in about one full-size launch out of ten)"""
    import tfswa_unet_b200 as T
    T.set_precision("bf16")
    torch.manual_seed(0)
    model = T.TFSWAUNet(2, 2, **MODEL).eval().cuda()
    x = seeded((2, 2, H, W), 960, 1.0).cuda()
    with torch.no_grad():
        ref = model(x, return_logits=True)[1]
        for i in range(10):
            junk = torch.randn(32 << 20, device="cuda")                 # perturb allocator state, L2 and leftover shared memory
            got = model(x, return_logits=True)[1]
            del junk
            assert torch.equal(got, ref), f"run {i}: {int((got != ref).sum())} logits differ, max {float((got - ref).abs().max()):.3e}"


@pytest.mark.parametrize("geom,C,h,w", [(0, 32, H, W), (1, 32, H, W), (2, 32, H, W), (0, 64, 512, 258), (1, 128, 256, 129)])
def test_fullsize_attention_backward_is_bitwise_reproducible(geom, C, h, w):
    """dq|dk|dv are written without atomics: they must not change from launch to launch"""
    from tfswa_unet_b200 import ops
    B, heads = 2, 8
    M = B * h * w
    qkv = seeded((M, 3 * C), 970 + geom, 1.0).cuda().to(torch.bfloat16)
    dout = seeded((M, C), 971 + geom, 1.0).cuda().to(torch.bfloat16)
    pad_kv = seeded((2 * C,), 972, 0.5).cuda().float().contiguous() if geom == 2 else None
    out = torch.empty((M, C), dtype=torch.bfloat16, device="cuda")
    lse = torch.empty((M, heads), dtype=torch.float32, device="cuda")
    kw = dict(ws=8, shift=4, pad_kv=pad_kv) if geom == 2 else {}
    ops.attention(qkv, out, B, h, w, C, heads, geom, lse=lse, **kw)
    ref = None
    for i in range(6):
        dqkv = torch.empty((M, 3 * C), dtype=torch.bfloat16, device="cuda")
        dsum = torch.empty((M, heads), dtype=torch.float32, device="cuda")
        dpad = torch.zeros((2 * C,), dtype=torch.float32, device="cuda") if geom == 2 else None
        ops.attention_bwd(qkv, out, lse, dout, dqkv, dsum, B, h, w, C, heads, geom, dpad=dpad, **kw)
        torch.cuda.synchronize()
        assert torch.isfinite(dqkv.float()).all()
        if ref is None:
            ref = dqkv
        else:
            assert torch.equal(dqkv, ref), f"launch {i}: {int((dqkv != ref).sum())} elements differ"
