"""debug: bitwise run-to-run reproducibility of the attention backward launches at the stage-1 shape, many repetitions"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tfswa_unet_b200 import ops
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
for (B, H, W, C) in ((2, 1025, 517, 32), (2, 512, 258, 64), (2, 256, 129, 128)):
    M = B * H * W
    torch.manual_seed(0)
    qkv = torch.randn(M, 3 * C, device="cuda").to(torch.bfloat16)
    dout = torch.randn(M, C, device="cuda").to(torch.bfloat16)
    pad_kv = torch.randn(2 * C, device="cuda")
    for geom in (0, 1, 2):
        kw = dict(ws=8, shift=4, pad_kv=pad_kv) if geom == 2 else {}
        out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
        lse = torch.empty(M, 8, device="cuda", dtype=torch.float32)
        ops.attention(qkv, out, B, H, W, C, 8, geom, lse=lse, **kw)
        ref, bad = None, 0
        for i in range(reps):
            junk = torch.randn(32 << 20, device="cuda")
            dqkv = torch.empty_like(qkv)
            dsum = torch.empty(M, 8, device="cuda", dtype=torch.float32)
            dpad = torch.zeros(2 * C, device="cuda") if geom == 2 else None
            ops.attention_bwd(qkv, out, lse, dout, dqkv, dsum, B, H, W, C, 8, geom, dpad=dpad, **kw)
            torch.cuda.synchronize()
            del junk
            if ref is None:
                ref = dqkv
            elif not torch.equal(dqkv, ref):
                bad += 1
                d = (dqkv.float() - ref.float()).abs()
                print(f"  C={C} geom {geom} rep {i}: {int((d > 0).sum())} elements differ, max {float(d.max()):.4f}")
        print(f"C={C} geom {geom}: {bad}/{reps - 1} backward launches differ from the first")
