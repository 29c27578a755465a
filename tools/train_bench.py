#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[4], "C5"): fwd + trainer L1 loss + bwd + gradient all-reduce + clip +
AdamW of TFSWAUNet(4,4,...) on a per-GPU batch of synthetic 6-s stereo mixtures (n_fft 2048, hop 512), bf16 activations.

    python tools/train_bench.py [--batch 8] [--steps 5] [--warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        tools/train_bench.py --batch 8

One step = what src/training/trainer.py:129-219 does per batch: STFT of the mixture and of both target stems
(torch.stft / cuFFT: either side of the path), model forward, masked-magnitude L1, backward, clip_grad_norm_(1.0),
AdamW - through ``tfswa_unet_b200.train_step.TrainStep`` (flat arena, bucketed in-place NCCL all-reduce overlapped with
backward, fused norm+clip+AdamW kernels, no host sync inside the step).  The MR-STFT term of the BASELINE wording is
not part of the reference trainer's loss (scripts/train.py:247 use_mrstft=False); ``--mrstft`` adds it as
SourceSeparationLoss words it (losses.py:235-283: L1 + 0.5 * MR-STFT averaged over stems) on the predicted audio
ISTFT(mixture spectrogram x sigmoid(|mask|)) - the ISTFT the reference trainer never performs.
``--measure-exposed`` times the same steps a second time with the gradient exchange switched off: the difference is the
part of the all-reduce that backward did not hide.
Timing: CUDA events around K steps after W warm-up steps, barrier + synchronize on both sides, max over ranks.
Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import tfswa_unet_b200 as T
from tfswa_unet_b200 import ops
from tfswa_unet_b200.train_step import TrainStep

N_FFT, HOP, SAMPLES = 2048, 512, 264600


def stft(wave):                       # (B, C, S) -> (B, C, F, T) complex64; stft_processor.py:87-134 batched
    B, C, S = wave.shape
    win = torch.hann_window(N_FFT, device=wave.device)
    spec = torch.stft(wave.reshape(B * C, S), N_FFT, HOP, N_FFT, win, center=True, pad_mode="reflect", normalized=False,
                      onesided=True, return_complex=True)
    return spec.reshape(B, C, *spec.shape[-2:])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--bucket-mb", type=int, default=16)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--mrstft", action="store_true")
    ap.add_argument("--recompute", default="none", choices=["none", "blocks", "stage1"])
    ap.add_argument("--measure-exposed", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    T.set_precision(args.precision)
    torch.manual_seed(0)                                     # identical initial weights on every rank
    model = T.TFSWAUNet(4, 4, [2, 2, 6, 2], [32, 64, 128, 256], 8, 4, 8).train().cuda()
    step = TrainStep(model, lr=1e-3, weight_decay=1e-2, max_grad_norm=1.0, bucket_bytes=args.bucket_mb << 20, recompute=args.recompute)
    g = torch.Generator(device="cuda").manual_seed(1234 + rank)
    mixtures = 0.1 * torch.randn(args.batch, 2, SAMPLES, device="cuda", generator=g)
    targets = [0.1 * torch.randn(args.batch, 2, SAMPLES, device="cuda", generator=g) for _ in range(2)]

    from tfswa_unet_b200.losses import mrstft_loss
    window = torch.hann_window(N_FFT, device="cuda")

    def one_step():
        with torch.no_grad():
            spec = stft(mixtures)
            x = torch.cat([spec.real, spec.imag], dim=1)                      # to_model_input, stft_processor.py:186-204
            mix_mag = spec.mean(dim=1).abs()                                  # trainer.py:141-142
            tg = [stft(t).mean(dim=1).abs() for t in targets]                 # trainer.py:145-149
        extra = None
        if args.mrstft:
            def extra(out):                                                    # losses.py:235-283, mrstft_weight 0.5
                B = out.shape[0]
                acc = 0.0
                for i, t in enumerate(targets):
                    re, im = out[:, 2 * i].float(), out[:, 2 * i + 1].float()
                    mask = torch.sigmoid(torch.sqrt(re * re + im * im + 1e-8))            # trainer.py:182-183
                    pred = torch.istft((spec * mask[:, None]).reshape(B * 2, *spec.shape[-2:]), N_FFT, HOP, N_FFT, window,
                                       center=True, normalized=False, onesided=True, length=SAMPLES).reshape(B, 2, SAMPLES)
                    acc = acc + mrstft_loss(pred, t)
                return 0.5 * acc / len(targets)
        return step(x, mix_mag, tg, extra_loss=extra)

    def timed(n):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            out = one_step()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / n], device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dist.barrier()
        return float(t), out

    for _ in range(args.warmup):
        loss, norm = one_step()
    ops.reset_launch_count()
    ms, (loss, norm) = timed(args.steps)
    launches = ops.reset_launch_count()
    exposed = None
    if args.measure_exposed and world > 1:
        # the same steps without the collective (replicas drift apart afterwards: this is the last thing the script does)
        step.arena.skip_exchange = True
        one_step()
        ms_no, _ = timed(args.steps)
        step.arena.skip_exchange = False
        exposed = {"ms_per_step_without_allreduce": ms_no, "exposed_allreduce_ms": ms - ms_no,
                   "exposed_frac_of_step": (ms - ms_no) / ms, "allreduce_bytes": step.arena.numel * 4}
    if args.profile and rank == 0:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            one_step()
            torch.cuda.synchronize()
        print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=70), file=sys.stderr)
    if rank == 0:
        fwd_tflop = 13.41 * args.batch / 8.0                                  # C3/C5 forward, in=out=4, batch 8
        print(json.dumps({
            "metric": "training audio-seconds/sec (fwd+bwd+allreduce+clip+AdamW)", "unit": "audio-s/s",
            "value": world * args.batch * 6.0 / (ms / 1e3), "ms_per_step": ms, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "scaling": "weak", "dtype": args.precision, "data": "synthetic",
            "config": {"workload": f"C5: TFSWAUNet(4,4,[2,2,6,2],[32,64,128,256],8,4,8) train step, batch {args.batch}/GPU of 6 s "
                                   "stereo mixtures (STFT n_fft 2048 hop 512 -> (B,4,1025,517)), L1 masked-magnitude loss, "
                                   "clip 1.0, AdamW", "batch_per_gpu": args.batch, "bucket_mb": args.bucket_mb,
                       "parallelism": f"dp{world} (batch-sharded replicas, NCCL gradient all-reduce overlapped with backward)"},
            "model_tflop_per_step_per_gpu": 3 * fwd_tflop, "achieved_model_tflops_per_gpu": 3 * fwd_tflop / (ms / 1e3),
            "loss_terms": "L1 masked magnitude + 0.5 * MR-STFT(2048/1024/512) on ISTFT audio" if args.mrstft else "L1 masked magnitude (the reference trainer's loss)",
            "allreduce": exposed, "recompute": args.recompute,
            "loss": float(loss), "grad_norm": float(norm), "gpu_launches": launches,
            "peak_mem_gib": torch.cuda.max_memory_allocated() / 2 ** 30}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
