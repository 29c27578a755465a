"""Single-TFSWABlock microbenchmark (BASELINE.json configs[1]): fwd (eval) and fwd+bwd (train) at the stage-1/2/3
resolutions of the 6 s / n_fft 2048 workload, fp32 and bf16, batch 1.  Prints one JSON line per case:
model FLOPs (tfswa_unet_b200.flops.block_flops, backward counted as 2x forward), achieved TFLOP/s and the fraction of the measured
bf16 tensor peak (MEASURED_PEAKS.json) - the "tensor-pipe util" half of the headline metric.

    python tools/block_bench.py [B] > profiles/rXX_block_bench.jsonl
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import tfswa_unet_b200 as T
from tfswa_unet_b200.flops import block_flops as count_block_flops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
peak = 1408.6
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
if os.path.exists(p):
    peak = json.load(open(p)).get("bf16_tflops_sustained", peak)
STAGES = {1: (32, 1025, 517), 2: (64, 512, 258), 3: (128, 256, 129)}


def timed(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for prec in ("bf16", "fp32"):
    T.set_precision(prec)
    for stage, (C, H, W) in STAGES.items():
        torch.manual_seed(0)
        blk = T.TFSWABlock(C, C, num_heads=8, window_size=8, shift_size=4).cuda()
        x = torch.randn(B, C, H, W, device="cuda")
        flops = count_block_flops(B, C, H, W)
        blk.eval()
        with torch.no_grad():
            t_f = timed(lambda: blk(x), 5 if prec == "bf16" else 2)
        blk.train()
        xg = x.clone().requires_grad_(True)

        def step():
            blk.zero_grad(set_to_none=True)
            xg.grad = None
            blk(xg).float().square().mean().backward()

        t_fb = timed(step, 3 if prec == "bf16" else 1)
        for mode, ms, fl in (("fwd", t_f, flops), ("fwd+bwd", t_fb, 3 * flops)):
            tf = fl / (ms / 1e3) / 1e12
            print(json.dumps({"case": f"TFSWABlock stage {stage} C={C} {H}x{W} B={B}", "precision": prec, "mode": mode,
                              "ms": round(ms, 3), "model_gflop": round(fl / 1e9, 1), "tflops": round(tf, 1),
                              "frac_of_bf16_tensor_peak": round(tf / peak, 4)}), flush=True)
