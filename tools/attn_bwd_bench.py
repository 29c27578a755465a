"""Micro-benchmark of the attention backward (tfswa_attn_bwd) at the C3 stage shapes (development helper).
    python tools/attn_bwd_bench.py [B] [stage] [geoms e.g. 0,1,2]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tfswa_unet_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
stage = int(sys.argv[2]) if len(sys.argv) > 2 else 1
geoms = [int(g) for g in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1, 2]
H, W, C = {1: (1025, 517, 32), 2: (512, 258, 64), 3: (256, 129, 128), 4: (128, 64, 256)}[stage]
M = B * H * W
torch.manual_seed(0)
qkv = torch.randn(M, 3 * C, device="cuda").to(torch.bfloat16)
dout = torch.randn(M, C, device="cuda").to(torch.bfloat16)
out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(M, 8, device="cuda", dtype=torch.float32)
dqkv = torch.empty_like(qkv)
dsum = torch.empty(M, 8, device="cuda", dtype=torch.float32)
pad_kv = torch.randn(2 * C, device="cuda")
dpad = torch.zeros(2 * C, device="cuda")
for geom in geoms:
    kw = dict(ws=8, shift=4, pad_kv=pad_kv) if geom == 2 else {}
    ops.attention(qkv, out, B, H, W, C, 8, geom, lse=lse, **kw)
    bkw = dict(ws=8, shift=4, pad_kv=pad_kv, dpad=dpad) if geom == 2 else {}
    for _ in range(2):
        ops.attention_bwd(qkv, out, lse, dout, dqkv, dsum, B, H, W, C, 8, geom, **bkw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.attention_bwd(qkv, out, lse, dout, dqkv, dsum, B, H, W, C, 8, geom, **bkw)
    e1.record()
    torch.cuda.synchronize()
    print(f"stage {stage} B={B} geom={geom}: bwd {e0.elapsed_time(e1) / 3:.3f} ms")
