"""Quick device-side timing of one training step (fwd + L1-style loss + bwd), development helper."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import tfswa_unet_b200 as T

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
H, W = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1025, 517)
T.set_precision(prec)
torch.manual_seed(0)
m = T.TFSWAUNet(4, 4, [2, 2, 6, 2], [32, 64, 128, 256], 8, 4, 8).train().cuda()
x = torch.randn(B, 4, H, W, device="cuda")
tgt = torch.rand(B, 4, H, W, device="cuda")
def step():
    m.zero_grad(set_to_none=True)
    y = m(x)
    loss = (y - tgt).abs().mean()
    loss.backward()
    return loss
for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 3
e0.record()
for _ in range(n):
    l = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"train step B={B} {prec} {H}x{W}: {ms:.1f} ms  loss={float(l):.4f}  peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
if "--profile" in sys.argv:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=30, max_name_column_width=80))
