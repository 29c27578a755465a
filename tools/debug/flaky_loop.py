"""debug: run-to-run determinism of the stage-1 axial attention launches (bitwise), many repetitions"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tfswa_unet_b200 import ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 30
B, H, W, C = 2, 1025, 517, 32
M = B * H * W
torch.manual_seed(0)
qkv = torch.randn(M, 3 * C, device="cuda").to(torch.bfloat16)
for geom in (0, 1):
    ref = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    ops.attention(qkv, ref, B, H, W, C, 8, geom)
    bad = 0
    worst = 0.0
    for i in range(reps):
        junk = torch.randn(64 << 20, device="cuda")            # disturb allocator / L2 / leftover shared memory between runs
        out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
        ops.attention(qkv, out, B, H, W, C, 8, geom)
        torch.cuda.synchronize()
        if not torch.equal(out, ref):
            bad += 1
            d = (out.float() - ref.float()).abs()
            worst = max(worst, float(d.max()))
            rows = int((d.max(1).values > 0).sum())
            print(f"  geom {geom} rep {i}: {rows} rows differ, max abs diff {float(d.max()):.4f}")
        del junk
    print(f"geom {geom}: {bad}/{reps} runs differ from the first, worst diff {worst}")
