"""debug: per-(entry point, shape) CUDA-event times of one training step (B given), sorted"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import tfswa_unet_b200 as T
from tfswa_unet_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
T.set_precision("bf16")
torch.manual_seed(0)
m = T.TFSWAUNet(4, 4, [2, 2, 6, 2], [32, 64, 128, 256], 8, 4, 8).train().cuda()
x = torch.randn(B, 4, 1025, 517, device="cuda")
def step():
    m.zero_grad(set_to_none=True)
    m(x).square().mean().backward()
for _ in range(2): step()
torch.cuda.synchronize()
ops.enable_timing(True)
step()
torch.cuda.synchronize()
tm = ops.collect_timing()
ops.enable_timing(False)
tot = sum(v["ms"] for v in tm.values())
print(f"sum of tagged C-ABI calls: {tot:.1f} ms")
for k, v in sorted(tm.items(), key=lambda kv: -kv[1]["ms"])[:28]:
    gbs = v["bytes"] / v["ms"] / 1e6 if v["bytes"] else 0
    print(f"{k:44s} {v['ms']:8.2f} ms  n={v['launches']:3d}  {gbs:7.0f} GB/s")
