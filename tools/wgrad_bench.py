"""Micro-benchmark of the linear weight-gradient kernels at the C5 shapes (development helper).
TFSWA_WGRAD=mma selects the warp-level-MMA kernel, default = tcgen05 + TMA (wgrad_tc.cu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tfswa_unet_b200 import ops, _lib as L

B = 8
shapes = []   # (stage, M, K, N, nb, ln)
for st, (H, W, C) in enumerate([(1025, 517, 32), (512, 258, 64), (256, 129, 128), (128, 64, 256)], 1):
    M = B * H * W
    shapes += [(st, M, C, 9 * C, 1, True), (st, M, C, C, 3, False), (st, M, C, 4 * C, 3, True), (st, M, 4 * C, C, 3, False),
               (st, M, C, C, 1, False), (st, M, 3 * C, C, 1, False)]
tot = 0.0
for st, M, K, N, nb, ln in shapes:
    x = torch.randn(M, nb, K, device="cuda").to(torch.bfloat16)
    g = torch.randn(M, nb, N, device="cuda").to(torch.bfloat16)
    stats = torch.rand(nb, M, 2, device="cuda") if ln else None
    for _ in range(2):
        ops.linear_wgrad(x, g, prologue=L.PRO_LNHAT if ln else L.PRO_NONE, row_stats=stats)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.linear_wgrad(x, g, prologue=L.PRO_LNHAT if ln else L.PRO_NONE, row_stats=stats)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    gb = M * nb * (K + N) * 2 / 1e9
    tot += ms
    print(f"stage {st} K={K:4d} N={N:4d} nb={nb} ln={int(ln)}: {ms:7.3f} ms  {gb / ms:6.2f} TB/s ({gb:.2f} GB)")
    del x, g
print(f"sum {tot:.2f} ms")
