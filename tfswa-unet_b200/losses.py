"""Loss terms of the training step that sit next to the path (reference ``src/training/losses.py``).

``masked_magnitude_l1`` (the term the reference trainer actually uses) lives in ``train_step.py``.  This module adds the
multi-resolution STFT term BASELINE's training configuration names (``MultiResolutionSTFTLoss``, losses.py:67-189:
mean-absolute error of the magnitudes plus mean-absolute error of the log-magnitudes at FFT sizes 2048 / 1024 / 512,
averaged over the resolutions) and the weighted combination of ``SourceSeparationLoss`` (losses.py:235-283).  The
reference trainer switches the term off (scripts/train.py:247) and never produces the predicted audio it needs; a
caller that wants it passes waveforms (e.g. ``torch.istft`` of ``mixture_spec * mask``).  The transforms are
``torch.stft`` (cuFFT on the GPU).  On CUDA tensors the magnitude / log-magnitude arithmetic of a resolution and its gradient
with respect to the predicted spectrogram are ONE kernel (``tfswa_mrstft_mag_loss``, csrc/spec.cu): |P|, |T| and their
logarithms never reach HBM, and the target's transform stays out of the autograd graph (the eager form transformed
prediction and target in one batch and back-propagated through both halves).  On CPU tensors the same arithmetic is eager
torch code; both are pinned against the live reference by ``tests/golden/golden_losses_v1.pt``.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import torch

Tensor = torch.Tensor


class _MagLogL1(torch.autograd.Function):
    """w_mag * mean||P| - |T|| + w_log * mean|log(|P| + eps) - log(|T| + eps)| of two complex STFTs (CUDA)."""

    @staticmethod
    def forward(ctx, pred_spec: Tensor, target_spec: Tensor, w_mag: float, w_log: float, eps: float) -> Tensor:
        from . import ops
        loss, grad = ops.mrstft_mag_loss(pred_spec, target_spec, w_mag, w_log, eps, want_grad=pred_spec.requires_grad)
        ctx.save_for_backward(grad)
        return loss[0].float()

    @staticmethod
    def backward(ctx, gout: Tensor):
        (grad,) = ctx.saved_tensors
        return (grad * gout if grad is not None else None), None, None, None, None


def mrstft_loss(pred_audio: Tensor, target_audio: Tensor, fft_sizes: Sequence[int] = (2048, 1024, 512),
                hop_sizes: Sequence[int] = (512, 256, 128), win_lengths: Sequence[int] = (2048, 1024, 512),
                magnitude_weight: float = 1.0, log_magnitude_weight: float = 1.0, eps: float = 1e-5) -> Tensor:
    """pred_audio, target_audio: (B, channels, samples) -> scalar (losses.py:143-189)."""
    if not (len(fft_sizes) == len(hop_sizes) == len(win_lengths)):
        raise ValueError("fft_sizes, hop_sizes and win_lengths must have the same length")
    B, C, S = pred_audio.shape
    if pred_audio.is_cuda:
        pa, ta = pred_audio.reshape(B * C, S).float(), target_audio.reshape(B * C, S).float().detach()
        total = pa.new_zeros(())
        for n_fft, hop, win in zip(fft_sizes, hop_sizes, win_lengths):
            window = torch.hann_window(win, device=pa.device)
            ps = torch.stft(pa, n_fft=n_fft, hop_length=hop, win_length=win, window=window, center=True, return_complex=True)
            with torch.no_grad():
                ts = torch.stft(ta, n_fft=n_fft, hop_length=hop, win_length=win, window=window, center=True, return_complex=True)
            total = total + _MagLogL1.apply(ps, ts, float(max(magnitude_weight, 0.0)), float(max(log_magnitude_weight, 0.0)), eps)
        return total / len(fft_sizes)
    # one batched transform for prediction and target: (2*B*C, S)
    both = torch.cat([pred_audio.reshape(B * C, S), target_audio.reshape(B * C, S)]).float()
    total = both.new_zeros(())
    for n_fft, hop, win in zip(fft_sizes, hop_sizes, win_lengths):
        window = torch.hann_window(win, device=both.device)
        mag = torch.stft(both, n_fft=n_fft, hop_length=hop, win_length=win, window=window, center=True,
                         return_complex=True).abs()
        pm, tm = mag[:B * C], mag[B * C:]
        if magnitude_weight > 0:
            total = total + magnitude_weight * (pm - tm).abs().mean()
        if log_magnitude_weight > 0:
            total = total + log_magnitude_weight * (torch.log(pm + eps) - torch.log(tm + eps)).abs().mean()
    return total / len(fft_sizes)


def source_separation_loss(pred_specs: Dict[str, Tensor], target_specs: Dict[str, Tensor],
                           pred_audios: Optional[Dict[str, Tensor]] = None, target_audios: Optional[Dict[str, Tensor]] = None,
                           l1_weight: float = 1.0, mrstft_weight: float = 0.5, use_l1: bool = True, use_mrstft: bool = True
                           ) -> Dict[str, Tensor]:
    """``SourceSeparationLoss.forward`` (losses.py:235-283): per-stem L1 on (magnitude) spectrograms averaged over stems,
    plus the MR-STFT term averaged over stems when audio is given; returns the same dictionary keys."""
    out: Dict[str, Tensor] = {}
    total = 0.0
    if use_l1:
        acc = 0.0
        for name, pred in pred_specs.items():
            tgt = target_specs[name]
            pred = pred.abs() if torch.is_complex(pred) else pred
            tgt = tgt.abs() if torch.is_complex(tgt) else tgt
            stem = (pred.float() - tgt.float()).abs().mean()
            out[f"l1_{name}"] = stem
            acc = acc + stem
        out["l1_loss"] = acc / len(pred_specs)
        total = total + l1_weight * out["l1_loss"]
    if use_mrstft and pred_audios is not None and target_audios is not None:
        acc = 0.0
        for name, pred in pred_audios.items():
            stem = mrstft_loss(pred, target_audios[name])
            out[f"mrstft_{name}"] = stem
            acc = acc + stem
        out["mrstft_loss"] = acc / len(pred_audios)
        total = total + mrstft_weight * out["mrstft_loss"]
    out["total_loss"] = total
    return out
