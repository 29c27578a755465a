"""Golden output of the reference's overlap-add separator (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_ola.py

Runs the unmodified ``SourceSeparator`` of /root/reference/src/evaluation/inference.py (with its ``STFTProcessor`` and
instance ``SpectrogramNormalizer``) on CPU around a deterministic stand-in model, for a seeded stereo clip, with and
without the normaliser, and stores the separated stems in ``tests/golden/golden_ola_v1.pt``.  ``soundfile`` / ``musdb``
are stubbed (not installed here; only their import is needed).
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True
for name in ("soundfile", "musdb"):
    sys.modules.setdefault(name, types.ModuleType(name))

from src.data.stft_processor import SpectrogramNormalizer, STFTProcessor  # noqa: E402
from src.evaluation.inference import SourceSeparator  # noqa: E402


class StandIn(torch.nn.Module):
    """segment-local, deterministic (b,2,F,T) -> (b,2,F,T) "masks" (the same formula as tests/test_multirank_cpu.py)"""
    def forward(self, x):
        return torch.sigmoid(0.3 * x + 0.1 * x.flip(1))


def main():
    torch.set_num_threads(4)
    g = torch.Generator().manual_seed(77)
    audio = 0.1 * torch.randn(2, 8000 * 2 + 1777, generator=g)          # 2.22 s stereo at 8 kHz
    out = {"audio_seed": 77, "samples": audio.shape[1], "cases": []}
    for normalize in (True, False):
        proc = STFTProcessor(n_fft=256, hop_length=64, sample_rate=8000)
        sep = SourceSeparator(StandIn(), proc, SpectrogramNormalizer("instance") if normalize else None, device="cpu",
                              use_amp=False, segment_length=0.5, overlap=0.25)
        res = sep.separate(audio)
        out["cases"].append({"normalize": normalize, "vocals": res["vocals"].clone(), "other": res["other"].clone()})
    short = audio[:, :3000]                                               # shorter than one segment: single-segment path
    proc = STFTProcessor(n_fft=256, hop_length=64, sample_rate=8000)
    res = SourceSeparator(StandIn(), proc, SpectrogramNormalizer("instance"), device="cpu", use_amp=False, segment_length=0.5,
                          overlap=0.25).separate(short)
    out["short"] = {"samples": 3000, "vocals": res["vocals"].clone(), "other": res["other"].clone()}
    path = os.path.join(ROOT, "tests", "golden", "golden_ola_v1.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes", {k: tuple(v.shape) for k, v in res.items()})


if __name__ == "__main__":
    main()
