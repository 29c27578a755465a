"""Golden vectors for the window helpers, from the LIVE reference (build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_windows.py

window_partition / window_reverse outputs on seeded inputs and the attn_mask buffers the reference's
ShiftedWindowAttention registers (attention.py:241-277, :318-345).  Stored in tests/golden/golden_windows_v1.pt.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from src.models.attention import ShiftedWindowAttention, window_partition, window_reverse  # noqa: E402


def seeded(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g)


out = {"cases": []}
for (shape, ws, seed) in [((2, 3, 16, 24), 8, 1), ((1, 5, 8, 8), 8, 2), ((3, 2, 12, 20), 4, 3)]:
    x = seeded(shape, seed)
    win = window_partition(x, ws)
    back = window_reverse(win, ws, shape[2], shape[3])
    assert torch.equal(back, x)
    out["cases"].append({"shape": shape, "ws": ws, "seed": seed, "windows": win.clone()})
out["masks"] = {}
for (ws, shift) in [(8, 4), (8, 0), (4, 2)]:
    m = ShiftedWindowAttention(32, ws, 8, shift_size=shift)
    out["masks"][(ws, shift)] = None if m.attn_mask is None else m.attn_mask.to(torch.int8).clone()
torch.save(out, os.path.join(ROOT, "tests", "golden", "golden_windows_v1.pt"))
print("wrote golden_windows_v1.pt:", [c["windows"].shape for c in out["cases"]], {k: (None if v is None else tuple(v.shape)) for k, v in out["masks"].items()})
