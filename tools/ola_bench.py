#!/usr/bin/env python
"""Overlap-add separation benchmark (BASELINE.json configs[3], "C4"): a synthetic 10-minute stereo mix
(0.1*randn(2, 26 460 000) at 44.1 kHz) -> 133 segments of 6 s at 25 % overlap, sharded contiguously over the ranks,
batches of 8 through TFSWAUNet(2,2,...) eval bf16, Hann-weighted overlap-add, one all-reduce of the output buffers.

    python tools/ola_bench.py [--minutes 10] [--reps 2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/ola_bench.py

The caller being timed is ``tfswa_unet_b200.separate.ShardedSeparator`` (reference: src/evaluation/inference.py:60-237,
sequential batch-1).  STFT/ISTFT are torch.stft/istft (cuFFT), inside the timed region like in the reference's loop.
Timing: CUDA events, barrier + synchronize both sides, max over ranks; total work is fixed -> "scaling": "strong".
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import tfswa_unet_b200 as T
from tfswa_unet_b200.parallel import shard_range
from tfswa_unet_b200.separate import ShardedSeparator


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--minutes", type=float, default=10.0)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    T.set_precision("bf16")
    torch.manual_seed(0)
    model = T.TFSWAUNet(2, 2, [2, 2, 6, 2], [32, 64, 128, 256], 8, 4, 8).eval().cuda()
    g = torch.Generator(device="cuda").manual_seed(7)
    n = int(args.minutes * 60 * 44100)
    audio = 0.1 * torch.randn(2, n, device="cuda", generator=g)
    sep = ShardedSeparator(model, n_fft=2048, hop_length=512, sample_rate=44100, segment_length=6.0, overlap=0.25, batch=args.batch)
    n_seg = len(sep.plan(n))
    # warm-up on a short prefix (one full batch), then the timed mix
    sep.separate(audio[:, :sep.segment_samples + (args.batch - 1) * sep.hop_samples])
    torch.cuda.synchronize()
    times = []
    for _ in range(args.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = sep.separate(audio)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        times.append(float(t))
    ms = min(times)
    if rank == 0:
        v = out["vocals"]
        print(json.dumps({
            "metric": "separated audio-seconds/sec (overlap-add, whole mix)", "unit": "audio-s/s",
            "value": args.minutes * 60.0 / (ms / 1e3), "ms_per_mix": ms, "all_ms": times, "n_gpus": world, "scaling": "strong",
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"C4: {args.minutes:g}-minute stereo mix, {n_seg} segments of 6 s (25 % overlap), batch {args.batch}, "
                                   "segments sharded contiguously over ranks, one all-reduce of the overlap-add buffers",
                       "segments": n_seg, "segments_rank0": len(range(*shard_range(n_seg, world, 0)))},
            "finite": bool(torch.isfinite(v).all()), "out_rms": float(v.pow(2).mean().sqrt())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
