"""GPU (-m gpu): the spectrogram kernels around the model in overlap-add separation (csrc/spec.cu) against
the eager torch ops of the reference (stft_processor.py:186-204, 283-312; inference.py:132-145, 209-216), and the CUDA
``ShardedSeparator`` - which runs on them - against the output of the live reference's ``SourceSeparator``
(tests/golden/golden_ola_v1.pt)."""
import os

import pytest
import torch

from helpers import seeded

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _spec(B, F, T, seed):
    re, im = seeded((B, F, T), seed, 2.0), seeded((B, F, T), seed + 1, 0.5) + 0.3
    return torch.complex(re, im).cuda()


@pytest.mark.parametrize("B,F,T", [(2, 129, 41), (1, 1025, 517), (3, 17, 2), (1, 5, 33)])
@pytest.mark.parametrize("normalize", [True, False])
def test_pack_norm_matches_torch(B, F, T, normalize):
    from tfswa_unet_b200 import ops
    spec = _spec(B, F, T, 3)
    x, stats = ops.spec_pack_norm(spec, normalize)
    ref = torch.cat([spec.real[:, None], spec.imag[:, None]], dim=1)                  # to_model_input
    if normalize:
        mean = ref.mean(dim=-1, keepdim=True)
        std = ref.std(dim=-1, keepdim=True) + 1e-8                                    # SpectrogramNormalizer('instance')
        ref = (ref - mean) / std
        assert torch.allclose(stats[..., 0], mean[..., 0], rtol=1e-5, atol=1e-6)
        assert torch.allclose(stats[..., 1], std[..., 0], rtol=1e-5, atol=1e-6)
        assert float((x - ref).abs().max()) <= 1e-5 * float(ref.abs().max())          # fp32, different summation order
    else:
        assert stats is None and torch.equal(x, ref)


@pytest.mark.parametrize("B,S,F,T", [(2, 2, 129, 41), (1, 2, 1025, 517), (3, 1, 9, 7)])
@pytest.mark.parametrize("normalize", [True, False])
def test_mask_apply_matches_torch(B, S, F, T, normalize):
    from tfswa_unet_b200 import ops
    if normalize and S != 2:
        pytest.skip("the reference's mask denormalisation broadcasts (B, 2, F, 1) statistics: two stems")
    spec = _spec(B, F, T, 7)
    masks = torch.sigmoid(seeded((B, S, F, T), 9, 2.0)).cuda()
    stats = None
    want = masks
    if normalize:
        _, stats = ops.spec_pack_norm(spec, True)
        want = masks * stats[..., 1:2] + stats[..., 0:1]                               # SpectrogramNormalizer.denormalize
    out = ops.spec_mask_apply(masks, spec, stats)
    ref = spec[:, None] * want                                                         # inference.py:139-145
    assert out.shape == ref.shape and out.dtype == torch.complex64
    assert float((out - ref).abs().max()) <= 1e-6 * float(ref.abs().max())


@pytest.mark.parametrize("total,seg,L,hop,nseg,first", [(5000, 1000, 960, 750, 5, 0), (5000, 1000, 960, 750, 3, 2250),
                                                        (2100, 1000, 1000, 750, 2, 750), (700, 1000, 1024, 750, 1, 0)])
def test_ola_add_is_the_sequential_loop(total, seg, L, hop, nseg, first):
    from tfswa_unet_b200 import ops
    S = 2
    wav = seeded((nseg, S, L), 5, 1.0).cuda()
    win = torch.hann_window(seg, device="cuda")
    starts = [first + j * hop for j in range(nseg)]
    acc = seeded((S + 1, total), 6, 0.1).cuda()
    ref = acc.clone()
    for j, s in enumerate(starts):                                                     # inference.py:209-216
        n = min(seg, total - s, L)
        for i in range(S):
            ref[i, s:s + n] += wav[j, i, :n] * win[:n]
        ref[S, s:s + n] += win[:n]
    ops.ola_add(wav, starts, win, acc, seg)
    assert torch.equal(acc, ref)                                                       # same products, same order: bit-exact


def _stand_in(x):
    return torch.sigmoid(0.3 * x + 0.1 * x.flip(1))


@pytest.mark.parametrize("idx,batch", [(0, 3), (1, 8), (0, 1)])
def test_cuda_separator_matches_reference_separator(idx, batch):
    from tfswa_unet_b200 import ops
    from tfswa_unet_b200.separate import ShardedSeparator
    go = torch.load(os.path.join(ROOT, "tests", "golden", "golden_ola_v1.pt"), weights_only=False)
    case = go["cases"][idx]
    g = torch.Generator().manual_seed(go["audio_seed"])
    audio = (0.1 * torch.randn(2, go["samples"], generator=g)).cuda()
    sep = ShardedSeparator(_stand_in, n_fft=256, hop_length=64, sample_rate=8000, segment_length=0.5, overlap=0.25, batch=batch,
                           normalize=case["normalize"])
    n0 = ops.LAUNCHES
    out = sep.separate(audio)
    assert ops.LAUNCHES > n0, "the CUDA separator must run on the library's spectrogram kernels"
    for name in ("vocals", "other"):
        ref = case[name]
        assert out[name].shape == ref.shape
        err = float((out[name].cpu() - ref).abs().max())
        assert err <= 1e-4 * float(ref.abs().max()) + 1e-6, (name, err)              # cuFFT against the reference's CPU FFT


def _eager_mrstft(pred_audio, target_audio, fft_sizes=(2048, 1024, 512), hop_sizes=(512, 256, 128), win_lengths=(2048, 1024, 512),
                  magnitude_weight=1.0, log_magnitude_weight=1.0, eps=1e-5):
    """MultiResolutionSTFTLoss.forward (losses.py:143-189) as eager torch ops on whatever device the audio is on"""
    B, C, S = pred_audio.shape
    pa, ta = pred_audio.reshape(B * C, S), target_audio.reshape(B * C, S)
    total = 0.0
    for n_fft, hop, win in zip(fft_sizes, hop_sizes, win_lengths):
        w = torch.hann_window(win, device=pa.device)
        pm = torch.stft(pa, n_fft, hop, win, w, center=True, return_complex=True).abs()
        tm = torch.stft(ta, n_fft, hop, win, w, center=True, return_complex=True).abs()
        total = total + magnitude_weight * torch.nn.functional.l1_loss(pm, tm) \
            + log_magnitude_weight * torch.nn.functional.l1_loss(torch.log(pm + eps), torch.log(tm + eps))
    return total / len(fft_sizes)


def test_fused_mrstft_loss_matches_reference_goldens_and_eager():
    """mrstft_loss on CUDA (tfswa_mrstft_mag_loss: magnitudes, logs, both L1 terms and the gradient in one kernel per resolution)
    against the live reference's values / gradients (golden_losses_v1.pt, CPU FFT) and against the eager chain on the same GPU."""
    from tfswa_unet_b200 import ops
    from tfswa_unet_b200.losses import mrstft_loss
    gl = torch.load(os.path.join(ROOT, "tests", "golden", "golden_losses_v1.pt"), weights_only=False)
    for case in gl["mrstft"]:
        kw = case.get("kwargs", {})
        pred = seeded(case["shape"], case["seeds"][0], case["scale"]).cuda().requires_grad_(True)
        tgt = seeded(case["shape"], case["seeds"][1], case["scale"]).cuda()
        n0 = ops.LAUNCHES
        loss = mrstft_loss(pred, tgt, **kw)
        loss.backward()
        assert ops.LAUNCHES - n0 == len(kw.get("fft_sizes", (2048, 1024, 512))), "one fused kernel per resolution"
        assert abs(float(loss) - float(case["loss"])) <= 1e-4 * abs(float(case["loss"])), (float(loss), float(case["loss"]))
        g = pred.grad[:, :, ::case["grad_stride"]].cpu()
        # (the log term's gradient is sign(.) / (|P| + 1e-5): where |P| is tiny the CPU-FFT / cuFFT rounding difference is amplified;
        # the same-GPU comparison below is the tight one)
        assert float((g - case["grad"]).abs().max()) <= 5e-3 * float(case["grad"].abs().max()) + 1e-9
        assert abs(float(pred.grad.norm()) - float(case["grad_norm"])) <= 5e-3 * float(case["grad_norm"])
        pred2 = pred.detach().clone().requires_grad_(True)
        ref = _eager_mrstft(pred2, tgt, **kw)
        ref.backward()
        assert abs(float(loss) - float(ref)) <= 1e-5 * abs(float(ref))
        assert float((pred.grad - pred2.grad).norm()) <= 1e-3 * float(pred2.grad.norm())
