"""Fused block head vs the unfused tensor-core sequence at the C3 stage shapes.  python tools/head_bench.py [B]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from tfswa_unet_b200 import _lib as L
from tfswa_unet_b200 import functional as Fn

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H, W = 1025, 517


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / n


for stage, C in ((1, 32), (2, 64)):
    h, w = H, W
    for _ in range(stage - 1):
        h, w = (h + 1) // 2, (w + 1) // 2
    M = B * h * w
    torch.manual_seed(0)
    x = torch.randn(M, 1, C, device="cuda").to(torch.bfloat16)
    mk = lambda n, k: Fn.LinW(torch.randn(1, n, k, device="cuda") / k ** 0.5, torch.randn(1, n, device="cuda") * 0.1)
    inp, qk = mk(C, C), mk(9 * C, C)

    def unfused():
        x1 = Fn.linear(x, inp)
        return Fn.linear(x1, qk, prologue=L.PRO_LNHAT, row_stats=Fn.row_stats(x1))

    with torch.no_grad():
        tu = timeit(unfused)
        tf = timeit(lambda: Fn.block_head(x, inp, qk))
    print(f"stage {stage} C={C} M={M}: unfused {tu:.3f} ms, fused {tf:.3f} ms  ({2 * M * C * 11 / tf / 1e6:.0f} GB/s algorithmic)")
