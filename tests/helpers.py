"""Shared helpers for the parity tests (inputs/weights are regenerated from seeds)."""
import math

import torch

from oracle import tfswa_oracle as O


def seeded(shape, seed, scale=1.0, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return (scale * torch.randn(shape, generator=g)).to(dtype)


def unpack_grad(entry):
    """golden grads are either full tensors or {"sample","stride","norm"} (see make_golden.pack_grad)."""
    if torch.is_tensor(entry):
        return entry, None
    return entry["sample"], entry["stride"]


def assert_close(name, got, ref, rtol, atol=None):
    """max-abs error relative to max|ref| (the tolerance form SURVEY 8c states)."""
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    assert got.shape == ref.shape, f"{name}: shape {tuple(got.shape)} vs {tuple(ref.shape)}"
    assert torch.isfinite(got).all(), f"{name}: non-finite values"
    scale = float(ref.abs().max())
    err = float((got - ref).abs().max())
    bound = rtol * scale + (atol if atol is not None else 0.0)
    assert err <= bound, f"{name}: max-abs err {err:.3e} > {bound:.3e} (max|ref|={scale:.3e})"
    return err / max(scale, 1e-30)


def rel_l2(got, ref):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    return float((got - ref).norm() / ref.norm().clamp_min(1e-30))


def check_grad(name, got, entry, rtol, atol=None):
    """``atol`` covers gradients that are analytically zero (a conv bias feeding a train-mode BatchNorm)."""
    ref, stride = unpack_grad(entry)
    g = got.detach().flatten()[::stride] if stride else got.detach()
    return assert_close(name, g, ref, rtol, atol)
