"""Per-tile timeline of one CTA of tc_attn_tma_kernel (development helper).
Build the trace variant first: tools/build_variant.sh trace tc_attn_tma.cu -DTFSWA_TMA_TRACE
Run: TFSWA_B200_LIB=tfswa-unet_b200/lib/libtfswa_b200_trace.so TFSWA_AXIAL_KERNEL=tma python tools/debug/tma_trace.py"""
import ctypes as C
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from tfswa_unet_b200 import ops, _lib

B, H, W, Cc = 8, 1025, 517, 32
M = B * H * W
torch.manual_seed(0)
qkv = torch.randn(M, 3 * Cc, device="cuda").to(torch.bfloat16)
out = torch.empty(M, Cc, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.attention(qkv, out, B, H, W, Cc, 8, 0)
torch.cuda.synchronize()
buf = (C.c_longlong * (8 * 128))()
rc = _lib.lib().tfswa_dbg_tma_trace(buf)
assert rc == 0, rc
ev = [[buf[e * 128 + t] for t in range(128)] for e in range(8)]
names = ["w0:S ready", "w0:ld done", "w0:math done", "w0:st done", "w0:arrived", "S issued", "PV wake"]
order = [0, 4, 5, 6, 1, 2, 3]
t0 = min(x for x in ev[2][:3] if x)
print("tile  " + "  ".join(f"{n:>12s}" for n in names))
for t in range(66):
    print(f"{t:4d}  " + "  ".join(f"{(ev[e][t] - t0) if ev[e][t] else -1:12d}" for e in order))
