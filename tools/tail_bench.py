"""Fused branch tail vs the unfused tensor-core sequence at the C3 stage shapes.  python tools/tail_bench.py [B]"""
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


for stage, C in ((1, 32), (2, 64), (3, 128)):
    h, w = H, W
    for _ in range(stage - 1):
        h, w = (h + 1) // 2, (w + 1) // 2
    M = B * h * w
    torch.manual_seed(0)
    att = torch.randn(M, 3, C, device="cuda").to(torch.bfloat16)
    res = torch.randn(M, 1, C, device="cuda").to(torch.bfloat16)
    mk = lambda n, k: Fn.LinW(torch.randn(3, n, k, device="cuda") / k ** 0.5, torch.randn(3, n, device="cuda") * 0.1)
    proj, fc1, fc2 = mk(C, C), mk(4 * C, C), mk(C, 4 * C)

    def unfused():
        y = Fn.linear(att, proj, r1=res)
        st = Fn.row_stats(y)
        hd = Fn.linear(y, fc1, prologue=L.PRO_LNHAT, epilogue=L.EPI_GELU, row_stats=st)
        return Fn.linear(hd, fc2, r1=y)

    with torch.no_grad():
        tu = timeit(unfused)
        tf = timeit(lambda: Fn.branch_tail(att, res, proj, fc1, fc2))
    gel = M * 3 * 4 * C
    print(f"stage {stage} C={C} M={M}: unfused {tu:.3f} ms, fused {tf:.3f} ms  ({gel / tf / 1e6:.1f} G gelu/s, "
          f"{2 * M * C * 7 / tf / 1e6:.0f} GB/s algorithmic)")
