"""Working precision of the activation stream ("bf16": throughput path, "fp32": full-precision parity path)."""
import os

import torch

_PRECISION = os.environ.get("TFSWA_B200_PRECISION", "bf16")


def set_precision(p: str) -> None:
    global _PRECISION
    if p not in ("bf16", "fp32"):
        raise ValueError("precision must be 'bf16' or 'fp32'")
    _PRECISION = p


def get_precision() -> str:
    return _PRECISION


def work_dtype() -> torch.dtype:
    return torch.bfloat16 if _PRECISION == "bf16" else torch.float32
