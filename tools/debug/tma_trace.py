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
# item-level timeline of CTA 0, softmax warp 0 (persistent kernel): cycles relative to the first item's top
names = ["item top", "Q handed", "prev epi done", "S(0) in regs", "last P publ."]
order = [0, 1, 4, 5, 6]
t0 = ev[0][0]
print("item  " + "  ".join(f"{n:>13s}" for n in names) + "   item time")
for i in range(int(os.environ.get('TRACE_ITEMS', '24'))):
    if not ev[0][i]:
        break
    nxt = ev[0][i + 1] if ev[0][i + 1] else 0
    print(f"{i:4d}  " + "  ".join(f"{(ev[e][i] - t0) if ev[e][i] else -1:13d}" for e in order) + f"   {nxt - ev[0][i] if nxt else -1:9d}")
# drift over the kernel: item durations of the first 127 items of the traced CTA
tops = [ev[0][i] for i in range(128) if ev[0][i]]
d = [b - a for a, b in zip(tops, tops[1:])]
if d:
    print("item durations (cycles), every 8th:", d[::8])
    print(f"mean of first 16: {sum(d[:16]) / 16:.0f}   mean of last 16: {sum(d[-16:]) / 16:.0f}   mean of all {len(d)}: {sum(d) / len(d):.0f}")
