"""GPU (-m gpu): tcgen05 implicit-GEMM convolutions against the fp32-accumulate SIMT kernel on the same bf16 inputs and
against torch fp32 conv2d / conv_transpose2d."""
import pytest
import torch
import torch.nn.functional as F

from helpers import seeded

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("kind,B,Cin,Cout,H,W,gelu", [
    ("conv3", 2, 32, 32, 37, 21, True), ("conv3", 1, 32, 32, 130, 65, False),
    ("down", 2, 32, 64, 37, 21, True), ("down", 1, 64, 128, 66, 34, True), ("down", 1, 128, 256, 32, 17, True),
    ("up", 2, 64, 32, 18, 10, True), ("up", 1, 128, 64, 33, 16, True), ("up", 1, 256, 128, 16, 8, False),
])
def test_tc_conv_matches_reference(kind, B, Cin, Cout, H, W, gelu):
    from tfswa_unet_b200 import ops, _lib as L
    from tfswa_unet_b200.autograd import conv_layout
    x = seeded((B, Cin, H, W), 1).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    k = 3 if kind == "conv3" else 4
    w = (seeded((Cout, Cin, k, k), 2) / (Cin * k * k) ** 0.5).cuda().to(torch.bfloat16).float()      # Cout-first OIHW
    b = seeded((Cout,), 3, 0.1).cuda()
    if kind == "conv3":
        ref, out_hw = F.conv2d(x.float(), w, b, padding=1), (H, W)
    elif kind == "down":
        ref = F.conv2d(x.float(), w, b, stride=2, padding=1)
        out_hw = tuple(ref.shape[2:])
    else:
        ref = F.conv_transpose2d(x.float(), w.permute(1, 0, 2, 3), b, stride=2, padding=1)
        out_hw = tuple(ref.shape[2:])
    if gelu:
        ref = F.gelu(ref)
    wl = conv_layout(w, kind)
    kid = {"conv3": 0, "down": 1, "up": 2}[kind]
    y = ops.conv_tc(x, wl.to(torch.bfloat16).contiguous(), b, kid, out_hw, epilogue=L.EPI_GELU if gelu else 0)
    y2 = ops.conv(x, wl, b, kid, out_hw, epilogue=L.EPI_GELU if gelu else 0)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert y.shape == ref.shape
    assert float((y.float() - ref).abs().max()) <= 1e-2 * scale + 1e-3
    assert float((y.float() - y2.float()).abs().max()) <= 1e-2 * scale + 1e-3


@pytest.mark.parametrize("kind,B,Cin,Cout,H,W", [
    ("conv3", 2, 32, 32, 37, 21), ("down", 2, 32, 64, 37, 21), ("down", 1, 128, 256, 32, 17),
    ("up", 2, 64, 32, 18, 10), ("up", 1, 256, 128, 16, 8), ("up", 1, 128, 64, 33, 16),
])
def test_tc_conv_column_statistics(kind, B, Cin, Cout, H, W):
    """train-mode BatchNorm sums from the tcgen05 conv epilogue = channel sums / sums of squares of the stored output
    (ragged last tile, phase-grid cells without an output pixel and every Cout tile included)"""
    from tfswa_unet_b200 import ops
    from tfswa_unet_b200.autograd import conv_layout
    x = seeded((B, Cin, H, W), 11).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    k = 3 if kind == "conv3" else 4
    w = (seeded((Cout, Cin, k, k), 12) / (Cin * k * k) ** 0.5).cuda().to(torch.bfloat16).float()
    b = seeded((Cout,), 13, 0.3).cuda()
    out_hw = {"conv3": (H, W), "down": ((H - 2) // 2 + 1, (W - 2) // 2 + 1), "up": (2 * H, 2 * W)}[kind]
    kid = {"conv3": 0, "down": 1, "up": 2}[kind]
    stats = torch.zeros((2, Cout), dtype=torch.float32, device="cuda")
    y = ops.conv_tc(x, conv_layout(w, kind).to(torch.bfloat16).contiguous(), b, kid, out_hw, col_stats=stats)
    y_plain = ops.conv_tc(x, conv_layout(w, kind).to(torch.bfloat16).contiguous(), b, kid, out_hw)
    torch.cuda.synchronize()
    assert torch.equal(y, y_plain), "collecting statistics must not change the output"
    yd = y.double()
    s1, s2 = yd.sum((0, 2, 3)), (yd * yd).sum((0, 2, 3))
    assert float((stats[0].double() - s1).abs().max()) <= 1e-4 * float(yd.abs().sum((0, 2, 3)).max()) + 1e-3
    assert float((stats[1].double() - s2).abs().max()) <= 1e-4 * float(s2.max())


@pytest.mark.parametrize("kind,B,Cin,Cout,H,W", [
    ("conv3", 2, 32, 32, 37, 21), ("conv3", 1, 32, 32, 130, 65),
    ("down", 2, 32, 64, 37, 21), ("down", 1, 64, 128, 66, 34), ("down", 1, 128, 256, 32, 17),
    ("up", 2, 64, 32, 18, 10), ("up", 1, 128, 64, 33, 16), ("up", 1, 256, 128, 16, 8),
])
def test_conv_weight_gradient_mma_matches_simt_and_torch(kind, B, Cin, Cout, H, W):
    """tfswa_conv_wgrad in bf16: the warp-MMA implicit-GEMM kernel (default) against the CUDA-core kernel
    (TFSWA_CONV_WGRAD_SIMT=1) and against torch's fp32 autograd of the same convolution."""
    import os
    from tfswa_unet_b200 import ops
    from tfswa_unet_b200.autograd import conv_layout, up_phase_weights_inverse
    x = seeded((B, Cin, H, W), 21).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    k = 3 if kind == "conv3" else 4
    w = (seeded((Cout, Cin, k, k), 22) / (Cin * k * k) ** 0.5).cuda().requires_grad_(True)
    b = torch.zeros(Cout, device="cuda", requires_grad=True)
    if kind == "conv3":
        y = F.conv2d(x.float(), w, b, padding=1)
    elif kind == "down":
        y = F.conv2d(x.float(), w, b, stride=2, padding=1)
    else:
        y = F.conv_transpose2d(x.float(), w.permute(1, 0, 2, 3), b, stride=2, padding=1)
    g = seeded(tuple(y.shape), 23).cuda().to(torch.bfloat16).contiguous(memory_format=torch.channels_last)
    y.backward(g.float())
    kid = {"conv3": 0, "down": 1, "up": 2}[kind]
    wl_shape = tuple(conv_layout(w.detach(), kind).shape)
    res = {}
    for mode in ("mma", "simt"):
        if mode == "simt":
            os.environ["TFSWA_CONV_WGRAD_SIMT"] = "1"
        try:
            dwl, db = ops.conv_wgrad(x, g, kid, wl_shape)
            torch.cuda.synchronize()
        finally:
            os.environ.pop("TFSWA_CONV_WGRAD_SIMT", None)
        dw = dwl.permute(0, 3, 1, 2) if kind in ("conv3", "down") else up_phase_weights_inverse(dwl)
        res[mode] = (dw, db)
    for mode, (dw, db) in res.items():
        assert float((dw - w.grad).norm() / w.grad.norm()) <= 2e-3, mode      # same bf16 inputs, fp32 accumulation
        assert float((db - b.grad).abs().max()) <= 1e-3 * float(b.grad.abs().max()) + 1e-3, mode
    assert float((res["mma"][0] - res["simt"][0]).norm() / res["simt"][0].norm()) <= 1e-4
