"""Golden values of the loss terms from the LIVE reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_losses.py

Imports ``src.training.losses`` from /root/reference (read-only, never copied; ``soundfile`` / ``musdb`` are stubbed
because ``src/data/__init__.py`` imports them and they are not installed here) and stores, for seeded inputs, the
MR-STFT loss with its gradient and the ``SourceSeparationLoss`` dictionary in ``tests/golden/golden_losses_v1.pt``.
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

from src.training.losses import MultiResolutionSTFTLoss, SourceSeparationLoss  # noqa: E402


def seeded(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return scale * torch.randn(shape, generator=g)


def main():
    torch.set_num_threads(4)
    out = {"mrstft": [], "combined": []}
    for i, (B, C, S) in enumerate([(2, 2, 6000), (1, 1, 4096), (3, 2, 5000)]):
        pred = seeded((B, C, S), 500 + i, 0.1).requires_grad_(True)
        tgt = seeded((B, C, S), 600 + i, 0.1)
        loss = MultiResolutionSTFTLoss()(pred, tgt)
        loss.backward()
        out["mrstft"].append({"shape": (B, C, S), "seeds": (500 + i, 600 + i), "scale": 0.1, "loss": loss.detach().clone(),
                              "grad": pred.grad.detach().clone()[:, :, ::7].contiguous(), "grad_stride": 7,
                              "grad_norm": pred.grad.norm().clone()})
    # non-default resolutions / weights
    pred = seeded((2, 1, 3000), 510, 0.2).requires_grad_(True)
    tgt = seeded((2, 1, 3000), 610, 0.2)
    lf = MultiResolutionSTFTLoss(fft_sizes=[512, 256], hop_sizes=[128, 64], win_lengths=[512, 256], magnitude_weight=0.5,
                                 log_magnitude_weight=2.0)
    loss = lf(pred, tgt)
    loss.backward()
    out["mrstft"].append({"shape": (2, 1, 3000), "seeds": (510, 610), "scale": 0.2, "loss": loss.detach().clone(),
                          "grad": pred.grad.detach().clone()[:, :, ::7].contiguous(), "grad_stride": 7, "grad_norm": pred.grad.norm().clone(),
                          "kwargs": dict(fft_sizes=(512, 256), hop_sizes=(128, 64), win_lengths=(512, 256), magnitude_weight=0.5,
                                         log_magnitude_weight=2.0)})
    # SourceSeparationLoss with both terms
    stems = ("vocals", "other")
    ps = {s: seeded((2, 33, 20), 700 + j).abs() for j, s in enumerate(stems)}
    ts_ = {s: seeded((2, 33, 20), 710 + j).abs() for j, s in enumerate(stems)}
    pa = {s: seeded((2, 2, 4500), 720 + j, 0.1) for j, s in enumerate(stems)}
    ta = {s: seeded((2, 2, 4500), 730 + j, 0.1) for j, s in enumerate(stems)}
    d = SourceSeparationLoss(l1_weight=1.0, mrstft_weight=0.5, use_l1=True, use_mrstft=True)(ps, ts_, pa, ta)
    out["combined"].append({k: (v.detach().clone() if torch.is_tensor(v) else torch.tensor(float(v))) for k, v in d.items()})
    d2 = SourceSeparationLoss(use_l1=True, use_mrstft=False)(ps, ts_)
    out["combined"].append({k: (v.detach().clone() if torch.is_tensor(v) else torch.tensor(float(v))) for k, v in d2.items()})
    path = os.path.join(ROOT, "tests", "golden", "golden_losses_v1.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
