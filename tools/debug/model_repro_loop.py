"""debug: bitwise run-to-run reproducibility of the eval forward (full size), many repetitions"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import tfswa_unet_b200 as T
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
T.set_precision("bf16")
torch.manual_seed(0)
model = T.TFSWAUNet(2, 2, [2, 2, 6, 2], [32, 64, 128, 256], 8, 4, 8).eval().cuda()
x = torch.randn(B, 2, 1025, 517, device="cuda")
bad = 0
with torch.no_grad():
    ref = model(x, return_logits=True)[1]
    for i in range(reps):
        junk = torch.randn(48 << 20, device="cuda")
        got = model(x, return_logits=True)[1]
        del junk
        if not torch.equal(got, ref):
            bad += 1
            print(f"rep {i}: {int((got != ref).sum())} logits differ, max {float((got - ref).abs().max()):.3e}")
print(f"{bad}/{reps} forward passes differ from the first")
