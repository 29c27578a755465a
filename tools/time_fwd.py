"""Quick device-side timing of the full-model forward (development helper, not the bench contract)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tfswa_unet_b200 as T

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
H, W = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1025, 517)
T.set_precision(prec)
torch.manual_seed(0)
m = T.TFSWAUNet(2, 2, [2, 2, 6, 2], [32, 64, 128, 256], 8, 4, 8).eval().cuda()
x = torch.randn(B, 2, H, W, device="cuda")
with torch.no_grad():
    for _ in range(2):
        y = m(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 3
    e0.record()
    for _ in range(n):
        y = m(x)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"B={B} {prec} {H}x{W}: {ms:.2f} ms/fwd  -> {B*6.0/(ms/1e3):.1f} audio-s/s  peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB  finite={bool(torch.isfinite(y).all())}")
if "--profile" in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof, torch.no_grad():
        m(x); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=90))
