"""HBM write / read / copy bandwidth by torch kernels (development helper: what can a write-dominated GEMM hope for?)."""
import torch
def t(f, n=5):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for gb in (0.6, 2.4):
    n = int(gb * 2 ** 30) // 2
    x = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    y = torch.empty(n, dtype=torch.bfloat16, device="cuda")
    ms = t(lambda: x.fill_(1.0)); print(f"fill  {gb} GiB: {ms:.3f} ms  {n*2/ms/1e6:.0f} GB/s written")
    ms = t(lambda: x.sum()); print(f"sum   {gb} GiB: {ms:.3f} ms  {n*2/ms/1e6:.0f} GB/s read")
    ms = t(lambda: y.copy_(x)); print(f"copy  {gb} GiB: {ms:.3f} ms  {2*n*2/ms/1e6:.0f} GB/s read+written")
    ms = t(lambda: torch.cuda.memset if False else x.zero_()); print(f"zero  {gb} GiB: {ms:.3f} ms  {n*2/ms/1e6:.0f} GB/s written")
