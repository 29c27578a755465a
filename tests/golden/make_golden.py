"""Generate golden vectors from the LIVE reference (run in the build container only).

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports the unmodified reference from /root/reference (read-only, never copied),
fills its parameters/buffers with ``oracle.tfswa_oracle.randomize_state_`` (so the
weights are reproducible from a seed and need not be stored), runs forward and
backward in fp32 on CPU and stores only inputs' seeds, outputs and gradients in
``tests/golden/golden_v1.pt``.  /root/reference does not exist on the GPU box, so
tests read this file, never the reference.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.dont_write_bytecode = True

from src.models.attention import (FrequencySequenceAttention, ShiftedWindowAttention,  # noqa: E402
                                  TemporalSequenceAttention)
from src.models.blocks import DownsampleBlock, TFSWABlock, UpsampleBlock  # noqa: E402
from src.models.tfswa_unet import TFSWAUNet  # noqa: E402

from oracle.tfswa_oracle import randomize_state_  # noqa: E402

torch.set_num_threads(8)
UNET_GAIN = 0.7  # keeps eval-mode logits O(10) through 22 residual blocks


def seeded(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return scale * torch.randn(shape, generator=g)


def pack_grad(g, full_limit, stride):
    """Small grads are stored whole; big ones as a strided sample + norm (keeps the file small)."""
    if g.numel() <= full_limit:
        return g.detach().clone()
    return {"sample": g.detach().flatten()[::stride].clone(), "stride": stride,
            "norm": float(g.double().norm())}


def fill(module, seed, gain=1.0):
    sd = module.state_dict()
    randomize_state_(sd, seed, gain)
    module.load_state_dict(sd)
    return module


def run_case(module, x, seed, train=False, extra=None):
    """forward + backward with loss = sum(y * w), w seeded."""
    module.train(train)
    x = x.clone().requires_grad_(True)
    args = (x,) if extra is None else (x, extra)
    y = module(*args)
    w = seeded(y.shape, seed + 7)
    (y * w).sum().backward()
    out = {"y": y.detach().clone(), "dx": x.grad.detach().clone(),
           "grads": {k: pack_grad(p.grad, 70000, 17) for k, p in module.named_parameters()}}
    if train:
        out["buffers"] = {k: b.detach().clone() for k, b in module.named_buffers() if "attn_mask" not in k}
    return out


def main():
    G = {"meta": {"torch": str(torch.__version__), "note": "reference fp32 CPU outputs; weights = randomize_state_(seed)"}}
    cases = {}

    # ---- a2/a3/a5: branch modules ------------------------------------------------
    for name, ctor, C, shape, seed in [
        ("tsa_c32", lambda C: TemporalSequenceAttention(C, 8), 32, (2, 32, 37, 21), 11),
        ("fsa_c32", lambda C: FrequencySequenceAttention(C, 8), 32, (2, 32, 37, 21), 12),
        ("swa_c32_s0", lambda C: ShiftedWindowAttention(C, 8, 8, 0), 32, (2, 32, 37, 21), 13),
        ("swa_c32_s4", lambda C: ShiftedWindowAttention(C, 8, 8, 4), 32, (2, 32, 37, 21), 14),
        ("tsa_c64", lambda C: TemporalSequenceAttention(C, 8), 64, (1, 64, 19, 24), 15),
        ("fsa_c128", lambda C: FrequencySequenceAttention(C, 8), 128, (1, 128, 9, 17), 16),
        ("swa_c256_s4", lambda C: ShiftedWindowAttention(C, 8, 8, 4), 256, (1, 256, 16, 8), 17),
    ]:
        m = fill(ctor(C), seed)
        x = seeded(shape, seed + 100)
        cases[name] = {"seed": seed, "shape": shape, **run_case(m, x, seed)}
        print(name, "ok", float(cases[name]["y"].abs().max()))

    # ---- a6: TFSWABlock ----------------------------------------------------------
    for name, C, shape, shift, with_skip, train, seed in [
        ("block_c32_s0_eval", 32, (2, 32, 20, 28), 0, False, False, 21),
        ("block_c32_s4_skip_eval", 32, (2, 32, 21, 13), 4, True, False, 22),
        ("block_c32_s4_train", 32, (2, 32, 20, 28), 4, False, True, 23),
        ("block_c64_s4_skip_train", 64, (2, 64, 12, 10), 4, True, True, 24),
    ]:
        m = fill(TFSWABlock(C, C, 8, shift, 8), seed)
        x = seeded(shape, seed + 100)
        skip = seeded(shape, seed + 200) if with_skip else None
        cases[name] = {"seed": seed, "shape": shape, "shift": shift, "with_skip": with_skip,
                       "train": train, **run_case(m, x, seed, train=train, extra=skip)}
        print(name, "ok", float(cases[name]["y"].abs().max()))

    # ---- a7/a8: down / up --------------------------------------------------------
    for name, ctor, shape, train, seed in [
        ("down_32_64_eval", lambda: DownsampleBlock(32, 64), (2, 32, 21, 13), False, 31),
        ("down_64_128_train", lambda: DownsampleBlock(64, 128), (2, 64, 12, 10), True, 32),
        ("up_64_32_eval", lambda: UpsampleBlock(64, 32), (2, 64, 10, 6), False, 33),
        ("up_128_64_train", lambda: UpsampleBlock(128, 64), (2, 128, 6, 5), True, 34),
    ]:
        m = fill(ctor(), seed)
        x = seeded(shape, seed + 100)
        cases[name] = {"seed": seed, "shape": shape, "train": train, **run_case(m, x, seed, train=train)}
        print(name, "ok", float(cases[name]["y"].abs().max()))

    # ---- a9/a10/a11: full model --------------------------------------------------
    for name, cin, cout, shape, train, seed in [
        ("unet_65x41_eval", 2, 2, (1, 2, 65, 41), False, 41),
        ("unet_64x96_train", 4, 4, (2, 4, 64, 96), True, 42),
    ]:
        m = fill(TFSWAUNet(cin, cout, [2, 2, 6, 2], [32, 64, 128, 256], 8, 4, 8), seed, gain=UNET_GAIN)
        m.train(train)
        x = seeded(shape, seed + 100).requires_grad_(True)
        # logits = everything but the final Sigmoid (output_head[4])
        taps = {}
        h = m.output_head[3].register_forward_hook(lambda mod, i, o: taps.__setitem__("logits", o.detach().clone()))
        y = m(x)
        h.remove()
        w = seeded(y.shape, seed + 7)
        (y * w).sum().backward()
        grads = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
        case = {"seed": seed, "shape": shape, "train": train, "cin": cin, "cout": cout,
                "y": y.detach().clone(), "logits": taps["logits"], "dx": x.grad.detach().clone(),
                "gain": UNET_GAIN, "grads": {k: pack_grad(v, 4096, 61) for k, v in grads.items()}}
        if train:
            case["buffers"] = {k: b.detach().clone() for k, b in m.named_buffers()
                               if "attn_mask" not in k and k.startswith(("stem.", "bottleneck.0.", "output_head."))}
        cases[name] = case
        print(name, "ok logits absmax", float(taps["logits"].abs().max()),
              "mask range", float(y.detach().min()), float(y.detach().max()))

    # ---- state_dict layout (a9): keys, shapes --------------------------------------
    m = TFSWAUNet(2, 2, [2, 2, 6, 2], [32, 64, 128, 256], 8, 4, 8)
    G["state_layout"] = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    G["num_parameters"] = m.get_num_parameters()
    G["attn_mask_ws8_s4"] = m.encoder_stages[0][1].swa.attn_mask.to(torch.int8).clone()
    G["cases"] = cases
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.pt")
    torch.save(G, out)
    print("wrote", out, os.path.getsize(out) / 1e6, "MB")


if __name__ == "__main__":
    main()
