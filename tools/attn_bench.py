"""Micro-benchmark of the attention kernels at the C3 stage shapes (development helper).
usage: attn_bench.py [B] [stage] [geoms, comma separated] [simt: 0/1]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tfswa_unet_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
stage = int(sys.argv[2]) if len(sys.argv) > 2 else 1
geoms = [int(g) for g in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1]
simt = bool(int(sys.argv[4])) if len(sys.argv) > 4 else False
H, W, C = {1: (1025, 517, 32), 2: (512, 258, 64), 3: (256, 129, 128), 4: (128, 64, 256)}[stage]
M = B * H * W
torch.manual_seed(0)
qkv = torch.randn(M, 3 * C, device="cuda").to(torch.bfloat16)
out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
pad_kv = torch.randn(2 * C, device="cuda")
kw = lambda g: dict(ws=8, shift=4, pad_kv=pad_kv) if g == 2 else {}
for tc in ((True, False) if simt else (True,)):
    ops.USE_TC_ATTENTION = tc
    for geom in geoms:
        for _ in range(3):
            ops.attention(qkv, out, B, H, W, C, 8, geom, **kw(geom))
        torch.cuda.synchronize()
        ts = []
        for _ in range(8):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.attention(qkv, out, B, H, W, C, 8, geom, **kw(geom))
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        ms = ts[len(ts) // 2]
        N = H if geom == 0 else (W if geom == 1 else 64)
        exps = M * N * 8
        print(f"stage {stage} B={B} geom={geom} tc={tc}: median {ms:.3f} ms (min {ts[0]:.3f})  {exps/(ms*1e-3)/1e12:.2f} Texp/s")
