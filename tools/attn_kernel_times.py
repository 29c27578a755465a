"""Per-kernel device times (torch profiler) of one TSA and one FSA attention call at stage-1 size, B=1."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tfswa_unet_b200 import ops
from torch.profiler import profile, ProfilerActivity
B,H,W,C = 1,1025,517,32
M=B*H*W
torch.manual_seed(0)
qkv = torch.randn(M, 3*C, device="cuda").to(torch.bfloat16)
out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
for g in (0,1):
    for _ in range(2): ops.attention(qkv, out, B,H,W,C,8,g)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for g in (0,1):
        ops.attention(qkv, out, B,H,W,C,8,g)
    torch.cuda.synchronize()
for e in prof.events():
    if e.device_type == torch.autograd.DeviceType.CUDA: print(e.name[:60], round(e.device_time,1), "us")
