"""Sequential, batch-1 restatement of the reference's overlap-add inference loop
(``src/evaluation/inference.py:98-225`` with ``stft_processor.py:87-134,136-184,186-204,240-312``) for an
arbitrary ``model_fn``.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  torchaudio's
``Spectrogram(power=None)`` / ``InverseSpectrogram`` are restated with ``torch.stft`` / ``torch.istft`` (same
centre/reflect-pad/hann/onesided conventions).  PINNED: tests/test_ola_golden.py holds this file to the output of the live
reference ``SourceSeparator`` (tests/golden/make_golden_ola.py -> golden_ola_v1.pt, 1e-5), including the reference's
ISTFT-length behaviour (segments come back (frames-1)*hop samples long) and its single-segment path."""
import torch


def _single(mono, model_fn, n_fft, hop, win, normalize, n_stems):
    """inference.py:98-157 on the clip as it is (no padding, no window; output length = (frames-1)*hop)"""
    spec = torch.stft(mono, n_fft, hop, n_fft, win, center=True, pad_mode="reflect", return_complex=True)
    x = torch.stack([spec.real, spec.imag], dim=1)
    if normalize:
        mean, std = x.mean(-1, keepdim=True), x.std(-1, keepdim=True) + 1e-8
        masks = model_fn((x - mean) / std) * std + mean
    else:
        masks = model_fn(x)
    return torch.cat([torch.istft(spec * masks[:, s], n_fft, hop, n_fft, win, center=True) for s in range(n_stems)])


def separate_long(audio, model_fn, n_fft=2048, hop=512, sr=44100, segment_length=6.0, overlap=0.25, normalize=True, n_stems=2):
    if audio.dim() == 1:
        audio = audio[None]
    mono = audio.mean(dim=0, keepdim=True) if audio.shape[0] > 1 else audio          # inference.py:84-85
    total = mono.shape[1]
    S = int(segment_length * sr)
    H = int(S * (1 - overlap))
    win = torch.hann_window(n_fft)
    out = torch.zeros(n_stems, total)
    norm = torch.zeros(total)
    seg_win = torch.hann_window(S)
    if total <= S:                                                                     # inference.py:92-95: one segment, no overlap-add
        return _single(mono, model_fn, n_fft, hop, win, normalize, n_stems)
    n_seg = (total - S) // H + 1                                                        # inference.py:187
    for i in range(n_seg):
        start = i * H
        end = min(start + S, total)
        seg = mono[:, start:end]
        if seg.shape[1] < S:
            seg = torch.nn.functional.pad(seg, (0, S - seg.shape[1]))
        spec = torch.stft(seg, n_fft, hop, n_fft, win, center=True, pad_mode="reflect", return_complex=True)   # (1,F,T)
        x = torch.stack([spec.real, spec.imag], dim=1)                                  # (1,2,F,T)
        if normalize:
            mean, std = x.mean(-1, keepdim=True), x.std(-1, keepdim=True) + 1e-8
            masks = model_fn((x - mean) / std) * std + mean                             # inference.py:124-133
        else:
            masks = model_fn(x)
        n = end - start
        for s in range(n_stems):
            # stft_processor.istft is called without `length` (inference.py:148-150): the segment comes back
            # (frames-1)*hop samples long, i.e. SHORTER than the segment when hop does not divide it, and the
            # accumulation uses the first `actual_length` samples of the (untruncated) Hann window (inference.py:209-211)
            wav = torch.istft(spec * masks[:, s], n_fft, hop, n_fft, win, center=True)
            n = min(end - start, wav.shape[1])
            out[s, start:start + n] += wav[0, :n] * seg_win[:n]                        # inference.py:212-216
        norm[start:start + n] += seg_win[:n]
    return out / norm.clamp_min(1e-8)                                                   # inference.py:221-223
