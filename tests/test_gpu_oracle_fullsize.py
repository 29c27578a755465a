"""GPU (-m gpu): the product against the ORACLE ITSELF at BASELINE.json's full size (a 6 s / n_fft 2048 segment ->
(1, 2, 1025, 517)).  The oracle is plain torch, so it runs on the B200 in fp32 (TF32 off) where the CPU would take minutes:
the stage-1 TSA scores alone are 517 x 8 x 1025^2 fp32 = 17 GB.  This closes the gap the sampled-sequence and
bf16-vs-own-fp32 checks of test_gpu_fullsize.py leave (round-1 VERDICT, weak #1).

Tolerances: fp32 path max-abs <= 1e-3 * max|ref| (outputs), 3e-3 (gradients; long fp32 reductions in a different order);
bf16 path rel-L2 <= 2e-2 (outputs / logits), 6e-2 (gradients), masks max-abs <= 4e-2.
"""
import pytest
import torch

from oracle import tfswa_oracle as O
from helpers import seeded, assert_close, rel_l2

pytestmark = pytest.mark.gpu
H, W = 1025, 517
MODEL = dict(depths=[2, 2, 6, 2], dims=[32, 64, 128, 256], window_size=8, shift_size=4, num_heads=8)


@pytest.fixture(autouse=True)
def _no_tf32():
    a, b = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = a, b
    torch.cuda.empty_cache()


def test_fullsize_model_against_the_oracle_on_gpu():
    import tfswa_unet_b200 as T
    try:
        torch.manual_seed(0)
        model = T.TFSWAUNet(2, 2, **MODEL)
        sd = model.state_dict()
        O.randomize_state_(sd, 31, 0.7)                      # non-trivial LN / BN affine, biases and running statistics
        model.load_state_dict(sd)
        model = model.eval().cuda()
        x = seeded((1, 2, H, W), 940, 1.0).cuda()
        with torch.no_grad():
            ref_logits = O.unet_forward(x, {k: v.cuda() for k, v in sd.items()}, return_logits=True)
            ref_masks = torch.sigmoid(ref_logits)
        torch.cuda.synchronize()
        assert torch.isfinite(ref_logits).all() and float(ref_logits.abs().max()) < 1e4      # an informative test, not a saturated one
        assert 0.05 < float(ref_masks.mean()) < 0.95
        got = {}
        for prec in ("fp32", "bf16"):
            T.set_precision(prec)
            with torch.no_grad():
                got[prec] = model(x, return_logits=True)
        torch.cuda.synchronize()
        assert_close("fullsize.fp32.logits", got["fp32"][1], ref_logits, 1e-3)
        assert_close("fullsize.fp32.masks", got["fp32"][0], ref_masks, 1e-3)
        e = rel_l2(got["bf16"][1], ref_logits)
        assert e <= 2e-2, f"full-size bf16 logits vs oracle: rel-L2 {e:.3e}"
        assert float((got["bf16"][0] - ref_masks).abs().max()) <= 4e-2
    finally:
        T.set_precision("bf16")


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_fullsize_stage1_block_forward_backward_against_the_oracle_on_gpu(precision):
    """one stage-1 TFSWABlock (C=32, shift 4, 1025 x 517: padded, wrapped windows; ragged attention tiles), eval-mode
    BatchNorm, loss = sum(y * w): output, input gradient and every parameter gradient against oracle autograd in fp32."""
    import tfswa_unet_b200 as T
    try:
        T.set_precision(precision)
        blk = T.TFSWABlock(32, 32, 8, 4, 8)
        sd = blk.state_dict()
        O.randomize_state_(sd, 77, 1.0)
        blk.load_state_dict(sd)
        blk = blk.eval().cuda()
        x = seeded((1, 32, H, W), 78, 1.0).cuda()
        wgt = seeded((1, 32, H, W), 79, 1.0).cuda()
        # oracle, fp32 autograd on the GPU
        ps = {k: v.cuda().clone().requires_grad_(v.is_floating_point() and "running" not in k and "attn_mask" not in k)
              for k, v in sd.items()}
        xr = x.clone().requires_grad_(True)
        yr = O.tfswa_block(xr, ps, shift=4)
        (yr * wgt).sum().backward()
        ref = {"y": yr.detach(), "dx": xr.grad.detach(), "grads": {k: v.grad.detach() for k, v in ps.items() if v.grad is not None}}
        del yr, xr
        torch.cuda.empty_cache()
        xp = x.clone().requires_grad_(True)
        y = blk(xp)
        (y.float() * wgt).sum().backward()
        torch.cuda.synchronize()
        grads = {k: p.grad for k, p in blk.named_parameters()}
        assert set(grads) == set(ref["grads"])
        if precision == "fp32":
            assert_close("block.y", y, ref["y"], 1e-3)
            assert_close("block.dx", xp.grad, ref["dx"], 3e-3)
            for k, g in grads.items():
                assert_close(f"block.{k}", g, ref["grads"][k], 3e-3, atol=1e-3)
        else:
            assert rel_l2(y.float(), ref["y"]) <= 2e-2
            assert rel_l2(xp.grad.float(), ref["dx"]) <= 6e-2
            for k, g in grads.items():
                r = ref["grads"][k]
                if float(r.norm()) > 1e-3 * float(r.numel()) ** 0.5:       # skip analytically tiny gradients (relative error meaningless)
                    assert rel_l2(g.float(), r) <= 6e-2, k
    finally:
        T.set_precision("bf16")
