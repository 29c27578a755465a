"""Batched, rank-sharded overlap-add separation (the caller of the hot path at reference
``src/evaluation/inference.py:60-237``).

Same arithmetic as the reference's ``SourceSeparator`` - mono down-mix (:84-85), per-segment STFT (:114), real/imag
packing (:121), optional instance normalisation over time (:124-125, stft_processor.py:240-312), model -> masks
(:128-129), mask * spectrogram per stem (:139-145), ISTFT (:148-150), Hann-weighted overlap-add and normalisation
(:209-223) - but segments are processed ``batch`` at a time instead of one by one, and the segment list is
sharded contiguously over the ranks of the process group; the only exchange is one all-reduce (SUM) of the
per-rank output and window-weight buffers.  STFT/ISTFT are torch.stft/istft (batched cuFFT).  On a CUDA device the
element-wise chains between them and the model are three kernels of libtfswa_b200 (csrc/spec.cu):
real/imag packing + instance normalisation, mask denormalisation + mask * spectrogram for all stems, and the windowed
overlap-add of a whole batch; the ISTFT of all stems of a batch is one call.  On CPU tensors (the gloo tests of the
sharding logic) the same steps are eager torch ops.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional

import torch
import torch.distributed as dist

from .parallel import shard_range

Tensor = torch.Tensor


def balanced_batches(lo: int, hi: int, max_batch: int):
    """[lo, hi) in the fewest batches of at most ``max_batch`` items, sizes as equal as possible: 17 segments -> 6 + 6 + 5,
    not 8 + 8 + 1 (a batch-of-1 forward costs 17 ms against 14.6 ms per segment at batch 8: pure tail latency)."""
    n = hi - lo
    if n <= 0:
        return []
    nb = -(-n // max_batch)
    base, extra = divmod(n, nb)
    out, s = [], lo
    for i in range(nb):
        e = s + base + (1 if i < extra else 0)
        out.append((s, e))
        s = e
    return out


class ShardedSeparator:
    def __init__(self, model: Callable[[Tensor], Tensor], n_fft: int = 2048, hop_length: int = 512, sample_rate: int = 44100,
                 segment_length: float = 6.0, overlap: float = 0.25, batch: int = 8, normalize: bool = True,
                 group: Optional[dist.ProcessGroup] = None):
        self.model, self.n_fft, self.hop, self.sr = model, n_fft, hop_length, sample_rate
        self.segment_samples = int(segment_length * sample_rate)                 # inference.py:57
        self.hop_samples = int(self.segment_samples * (1 - overlap))             # inference.py:58
        self.batch, self.normalize, self.group = batch, normalize, group

    # ---- segment plan (inference.py:187-201) --------------------------------------------------
    def plan(self, total: int) -> List[int]:
        if total <= self.segment_samples:
            return [0]                                           # (separate() takes the single-segment path for these)
        n = (total - self.segment_samples) // self.hop_samples + 1
        return [i * self.hop_samples for i in range(n)]          # samples past the last full hop stay uncovered (reference quirk)

    def _masks(self, seg: Tensor):
        """seg (b, S) mono -> complex spec (b, F, T), masked complex stems (b, stems, F, T)"""
        win = torch.hann_window(self.n_fft, device=seg.device)
        spec = torch.stft(seg, self.n_fft, self.hop, self.n_fft, win, center=True, pad_mode="reflect", normalized=False,
                          onesided=True, return_complex=True)
        if seg.is_cuda:
            from . import ops
            spec = spec.contiguous()
            x, stats = ops.spec_pack_norm(spec, self.normalize)                     # to_model_input + SpectrogramNormalizer in one pass
            masks = self.model(x)
            return spec, ops.spec_mask_apply(masks.float().contiguous(), spec, stats)   # (b, stems, F, T) complex stems
        x = torch.stack([spec.real, spec.imag], dim=1)                              # to_model_input, stft_processor.py:186-204
        if self.normalize:
            mean = x.mean(dim=-1, keepdim=True)
            std = x.std(dim=-1, keepdim=True) + 1e-8
            masks = self.model(((x - mean) / std).contiguous())
            # the reference "denormalises" the masks with the *input's* statistics (inference.py:132-133); the stats are
            # (b, 2, F, 1) and the masks (b, stems, F, T): same broadcasting as the reference when stems == 2
            masks = masks * std + mean
        else:
            masks = self.model(x.contiguous())
        return spec, spec[:, None] * masks                                          # inference.py:139-145

    @torch.no_grad()
    def separate(self, audio: Tensor, stem_names: Optional[List[str]] = None) -> Dict[str, Tensor]:
        stem_names = stem_names or ["vocals", "other"]
        if audio.dim() == 1:
            audio = audio[None]
        mono = audio.mean(dim=0) if audio.shape[0] > 1 else audio[0]               # inference.py:84-85
        total = mono.shape[0]
        if total <= self.segment_samples:                                          # inference.py:92-95: one segment as it is -
            spec, stems = self._masks(mono[None])                                  # no padding, no window, ISTFT's own length
            win = torch.hann_window(self.n_fft, device=mono.device)
            return {name: torch.istft(stems[:, i], self.n_fft, self.hop, self.n_fft, win, center=True, normalized=False,
                                      onesided=True) for i, name in enumerate(stem_names[:stems.shape[1]])}
        starts = self.plan(total)
        world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        lo, hi = shard_range(len(starts), world, rank)
        S = self.segment_samples
        win_seg = torch.hann_window(S, device=mono.device)                          # inference.py:227-237
        n_st = len(stem_names)
        acc = torch.zeros((n_st + 1, total), device=mono.device)                   # stems + window-weight row
        for b0, b1 in balanced_batches(lo, hi, self.batch):
            idx = starts[b0:b1]
            seg = torch.stack([torch.nn.functional.pad(mono[s:s + S], (0, max(0, S - (total - s)))) for s in idx])
            spec, stems = self._masks(seg)
            # The reference reconstructs without a target length (inference.py:148-150 -> stft_processor.py:136-184): a segment
            # comes back (frames-1)*hop samples long - 264 192 of 264 600 for 6 s at hop 512 - and only that many samples,
            # weighted by the FIRST part of the full-length Hann window, enter the overlap-add (inference.py:209-216).
            L = (spec.shape[-1] - 1) * self.hop
            k = min(n_st, stems.shape[1])
            fft_win = torch.hann_window(self.n_fft, device=seg.device)
            if seg.is_cuda and k == n_st:
                from . import ops
                wav = torch.istft(stems.reshape(-1, *stems.shape[2:]), self.n_fft, self.hop, self.n_fft, fft_win, center=True,
                                  normalized=False, onesided=True)                 # every stem of the batch in one cuFFT call
                ops.ola_add(wav.view(len(idx), k, -1).contiguous(), idx, win_seg, acc, S)
                continue
            for i in range(k):
                wav = torch.istft(stems[:, i], self.n_fft, self.hop, self.n_fft, fft_win, center=True, normalized=False, onesided=True)
                for j, s in enumerate(idx):
                    n = min(S, total - s, L)
                    acc[i, s:s + n] += wav[j, :n] * win_seg[:n]
            for s in idx:
                n = min(S, total - s, L)
                acc[n_st, s:s + n] += win_seg[:n]
        if world > 1:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.group)           # the single exchange step
        norm = acc[n_st].clamp_min(1e-8)
        return {name: (acc[i] / norm)[None] for i, name in enumerate(stem_names)}
