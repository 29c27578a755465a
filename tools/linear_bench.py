"""Micro-benchmark of the tcgen05 token GEMM at the C3 stage shapes (development helper)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tfswa_unet_b200 import ops, _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
cases = [("fc1 s1", 4239400, 3, 32, 128, True, True, False), ("qkv s1", 4239400, 1, 32, 288, True, False, False),
         ("fc2 s1", 4239400, 3, 128, 32, False, False, True), ("proj s1", 4239400, 3, 32, 32, False, False, True),
         ("fc1 s3", 264192, 3, 128, 512, True, True, False), ("fc2 s3", 264192, 3, 512, 128, False, False, True),
         ("qkv s3", 264192, 1, 128, 1152, True, False, False), ("in_proj s3", 264192, 1, 128, 128, False, False, False),
         ("fusion s1", 4239400, 1, 96, 32, False, True, False), ("fusion s2", 1056768, 1, 192, 64, False, True, False),
         ("fusion s3", 264192, 1, 384, 128, False, True, False), ("qkv s4", 65536, 1, 256, 2304, True, False, False)]
only = os.environ.get("LINEAR_BENCH_ONLY")
if only:
    cases = [c for c in cases if only in c[0]]
for name, M, nb, K, N, ln, gelu, res in cases:
    M = M * B // 8
    x = torch.randn(M, nb, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(nb, N, K, device="cuda") / K ** 0.5).to(torch.bfloat16).contiguous()
    wsum = w.float().sum(-1).contiguous()
    b = torch.randn(nb, N, device="cuda")
    st = ops.row_stats(x) if ln else None
    r1 = torch.randn(M, nb, N, device="cuda").to(torch.bfloat16) if res else None
    f = lambda: ops.linear_tc(x, w, wsum, b, prologue=L.PRO_LNHAT if ln else 0, epilogue=L.EPI_GELU if gelu else 0, row_stats=st, r1=r1)
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    byts = 2 * M * nb * (K + N) + (2 * M * nb * N if res else 0)
    print(f"{name}: M={M} nb={nb} K={K} N={N}: {ms:.3f} ms  {byts/ms/1e6:.0f} GB/s  {2*M*nb*N*K/ms/1e9:.1f} TFLOP/s")
