"""Drop-in mechanisms for the reference code base.

* :func:`install_as_reference` registers this package's modules under the reference's import paths
  (``src.models.attention`` / ``.blocks`` / ``.tfswa_unet``) so the reference's unmodified
  ``scripts/train.py``, ``scripts/evaluate.py``, ``src/training/trainer.py`` and ``src/evaluation/inference.py``
  construct the B200 model when they ``from src.models.tfswa_unet import TFSWAUNet``.
* :func:`convert` turns an already constructed reference ``TFSWAUNet`` (or any module exposing the same
  attributes/state_dict) into the B200 model, sharing the *same* ``nn.Parameter`` objects.
"""
from __future__ import annotations

import sys
import types

import torch
from torch import nn


def install_as_reference(package: str = "src.models") -> None:
    """Route ``<package>.attention`` / ``.blocks`` / ``.tfswa_unet`` to this package's modules.

    The real packages along ``package`` (``src``, ``src.models``) are imported first when they exist, so their
    ``__path__`` stays intact and ``src.training.trainer`` / ``src.evaluation.inference`` / ``src.data...`` keep
    importing from the reference tree; only the three leaf modules are replaced.  Stub packages are created only
    where the real import fails (the reference is not on ``sys.path``).  If the reference's own leaf modules were
    already imported, names other modules bound with ``from src.models.x import Y`` before this call keep pointing
    at the reference classes - call this before importing the trainer / separator."""
    import importlib

    from . import attention, blocks, tfswa_unet
    parts = package.split(".")
    for i in range(1, len(parts) + 1):
        name = ".".join(parts[:i])
        if name in sys.modules:
            continue
        try:
            importlib.import_module(name)
        except ImportError:
            mod = types.ModuleType(name)
            mod.__path__ = []          # mark as (empty) package
            sys.modules[name] = mod
            if i > 1:
                setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], mod)
    for leaf, mod in (("attention", attention), ("blocks", blocks), ("tfswa_unet", tfswa_unet)):
        sys.modules[f"{package}.{leaf}"] = mod
        setattr(sys.modules[package], leaf, mod)
    # convenience re-exports some callers use (`from src.models import TFSWAUNet`)
    for name in ("TFSWAUNet",):
        if not hasattr(sys.modules[package], name):
            setattr(sys.modules[package], name, getattr(tfswa_unet, name))


def convert(model: nn.Module) -> nn.Module:
    """Build a tfswa_unet_b200.TFSWAUNet that shares parameters and buffers with ``model``."""
    from .tfswa_unet import TFSWAUNet
    blk = model.encoder_stages[0][0]
    shift = max((b.shift_size for b in model.encoder_stages[0]), default=0)
    dropout = float(getattr(blk.tsa.attn.dropout, "p", 0.0))
    hidden = blk.tsa.mlp[0].out_features
    new = TFSWAUNet(model.in_channels, model.out_channels, list(model.depths), list(model.dims), blk.window_size, shift,
                    blk.num_heads, dropout=dropout, mlp_ratio=hidden / blk.out_channels)
    src_params = dict(model.named_parameters())
    src_bufs = dict(model.named_buffers())

    def share(mod: nn.Module, prefix: str = "") -> None:
        for name in list(mod._parameters):
            full = prefix + name
            if mod._parameters[name] is not None:
                mod._parameters[name] = src_params[full]
        for name in list(mod._buffers):
            full = prefix + name
            if mod._buffers[name] is not None and full in src_bufs:
                mod._buffers[name] = src_bufs[full]
        for cname, child in mod._modules.items():
            if child is not None:
                share(child, prefix + cname + ".")

    share(new)
    new.train(model.training)
    return new
