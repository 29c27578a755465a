"""Training-step plumbing around the hot path (reference ``src/training/trainer.py:120-257``,
optimiser construction ``scripts/train.py:251-255``).

The reference step is ``zero_grad -> autocast forward -> L1 on masked magnitudes -> backward -> clip_grad_norm_(1.0)
-> AdamW.step`` with one ``.item()`` host sync per logged loss.  Here the same arithmetic is laid out for one process
per B200:

* ``FlatArena`` re-homes every parameter and its gradient as views into ONE contiguous fp32 buffer each (15.4 M
  parameters = 61.6 MB).  ``zero_grad`` is one memset; autograd accumulates into the views in place; the gradient
  exchange is an in-place NCCL all-reduce of contiguous slices of the gradient buffer (no pack / unpack copies),
  launched bucket by bucket from ``post_accumulate_grad`` hooks while the rest of backward is still running.
  Parameters stay ordinary ``nn.Parameter``s (``state_dict``, ``.to``, checkpointing keep working).
* ``FusedClipAdamW.step`` = two launches over the arena through the C ABI (``tfswa_grad_sumsq``,
  ``tfswa_adamw_clip_step``): global gradient norm, ``clip_grad_norm_`` coefficient and the AdamW update, with the
  norm staying on the device (no host synchronisation per step).  BatchNorm statistics stay rank-local
  (``parallel.broadcast_buffers`` at checkpoint time).
* ``masked_magnitude_l1`` is the trainer's loss (``trainer.py:176-199`` with ``L1SpectrogramLoss`` of
  ``losses.py:14-58``) on the model output.

The arena and the bucketed exchange are device-agnostic torch code (covered by the world-size-2 gloo tests on CPU);
the optimiser kernels exist only in the CUDA library - there is no CPU fallback.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import nn

from . import _lib as L

Tensor = torch.Tensor
_ALIGN = 128          # elements (512 B): every parameter view keeps the alignment the caching allocator would give it


class FlatArena:
    """Parameters and gradients of ``model`` as views into two flat fp32 buffers + bucketed in-place all-reduce."""

    def __init__(self, model: nn.Module, group: Optional[dist.ProcessGroup] = None, bucket_bytes: int = 16 << 20):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.skip_exchange = False               # measurement only: run the step without the gradient all-reduce
        self.params: List[nn.Parameter] = [p for p in model.parameters() if p.requires_grad]
        if not self.params:
            raise ValueError("FlatArena: the model has no trainable parameters")
        dev = self.params[0].device
        for p in self.params:
            if p.dtype != torch.float32 or p.device != dev:
                raise TypeError("FlatArena: parameters must be fp32 and live on one device (bf16 is an activation format here)")
        self.offsets: List[int] = []
        off = 0
        for p in self.params:
            self.offsets.append(off)
            off += -(-p.numel() // _ALIGN) * _ALIGN
        self.numel = off
        self.flat_p = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(off, dtype=torch.float32, device=dev)
        with torch.no_grad():
            for p, o in zip(self.params, self.offsets):
                self.flat_p[o:o + p.numel()].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[o:o + p.numel()].view(p.shape)
        self._point_grads()
        # buckets: contiguous [lo, hi) element ranges, filled from the END of the arena (backward reaches the last
        # registered parameters - head, decoder - first)
        self.buckets: List[Tuple[int, int, List[int]]] = []
        members: List[int] = []
        hi = off
        for i in reversed(range(len(self.params))):
            members.append(i)
            if (hi - self.offsets[i]) * 4 >= bucket_bytes or i == 0:
                self.buckets.append((self.offsets[i], hi, members))
                hi, members = self.offsets[i], []
        self._bucket_of = {i: b for b, (_, _, ms) in enumerate(self.buckets) for i in ms}
        self._ready = [0] * len(self.buckets)
        self._work: List[Optional[object]] = [None] * len(self.buckets)
        self._index = {id(p): i for i, p in enumerate(self.params)}
        self._hooks = [p.register_post_accumulate_grad_hook(self._on_grad) for p in self.params]

    # ---- gradient views -----------------------------------------------------------------------
    def _point_grads(self) -> None:
        for p, o in zip(self.params, self.offsets):
            p.grad = self.flat_g[o:o + p.numel()].view(p.shape)

    def check(self) -> None:
        """The arena is only as good as the views: ``model.to(...)``, ``.half()`` or assigning ``p.data`` re-allocates the
        parameters outside it, after which the fused optimiser would update memory the model no longer reads."""
        for i in (0, len(self.params) // 2, len(self.params) - 1):
            p, o = self.params[i], self.offsets[i]
            if p.data_ptr() != self.flat_p[o:].data_ptr() or p.dtype != torch.float32:
                raise RuntimeError("FlatArena: a parameter no longer lives in the arena (model.to()/.half() after TrainStep "
                                   "was built?) - rebuild the TrainStep / FlatArena after moving the model")

    def zero_grad(self) -> None:
        """One memset.  (``optimizer.zero_grad(set_to_none=True)`` would detach the views: ``adopt_grads`` repairs that.)"""
        self.flat_g.zero_()
        self._ready = [0] * len(self.buckets)
        self._work = [None] * len(self.buckets)

    def adopt_grads(self) -> None:
        """Copy gradients that autograd allocated outside the arena (after a ``set_to_none`` zero_grad) back into it."""
        for p, o in zip(self.params, self.offsets):
            view = self.flat_g[o:o + p.numel()].view(p.shape)
            if p.grad is None:
                view.zero_()
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
            p.grad = view

    # ---- exchange -----------------------------------------------------------------------------
    def _on_grad(self, p: nn.Parameter) -> None:
        b = self._bucket_of[self._index[id(p)]]
        self._ready[b] += 1
        if self._ready[b] == len(self.buckets[b][2]):
            self._launch(b)

    def _launch(self, b: int) -> None:
        if self.world == 1 or self._work[b] is not None:
            return
        if self.skip_exchange:                   # measurement knob (tools/train_bench.py --measure-exposed): no collective
            self._work[b] = _NoWork()
            return
        lo, hi, members = self.buckets[b]
        for i in members:                        # a hook fired on a gradient outside the arena: bring it in first
            p = self.params[i]
            if p.grad is not None and p.grad.data_ptr() != self.flat_g[self.offsets[i]:].data_ptr():
                self.flat_g[self.offsets[i]:self.offsets[i] + p.numel()].view(p.shape).copy_(p.grad)
        self._work[b] = dist.all_reduce(self.flat_g[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def finish(self) -> float:
        """After ``loss.backward()``: every bucket's all-reduce has been issued and the current stream waits for them.
        The arena then holds the SUM over ranks; returns the scale (1/world) the consumer must apply."""
        if self.world > 1:
            for b in range(len(self.buckets)):
                if self._work[b] is None:        # unused parameters: their hooks never fired, gradient = zeros
                    self._launch(b)
                self._work[b].wait()
        self._ready = [0] * len(self.buckets)
        self._work = [None] * len(self.buckets)
        return 1.0 / self.world

    def average_(self) -> None:
        """For consumers other than ``FusedClipAdamW`` (e.g. a stock torch optimiser): turn the sum into the mean."""
        scale = self.finish()
        if scale != 1.0:
            self.flat_g.mul_(scale)

    def remove(self) -> None:
        for h in self._hooks:
            h.remove()


class _NoWork:
    def wait(self) -> None:
        pass


class FusedClipAdamW:
    """``clip_grad_norm_(max_norm)`` + ``torch.optim.AdamW`` (one group, decoupled decay on every parameter, as the
    reference builds it) over a ``FlatArena``, as two kernels of libtfswa_b200."""

    def __init__(self, arena: FlatArena, lr: float = 1e-3, betas: Sequence[float] = (0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-2, max_grad_norm: float = 1.0):
        if not arena.flat_p.is_cuda:
            raise RuntimeError("FusedClipAdamW runs on CUDA (sm_100a) tensors only - there is no CPU fallback")
        self.arena, self.lr, self.betas, self.eps = arena, lr, tuple(betas), eps
        self.weight_decay, self.max_grad_norm = weight_decay, max_grad_norm
        self.exp_avg = torch.zeros_like(arena.flat_p)
        self.exp_avg_sq = torch.zeros_like(arena.flat_p)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=arena.flat_p.device)
        self.grad_norm = torch.zeros(1, dtype=torch.float32, device=arena.flat_p.device)   # unclipped total norm, on device
        self.step_count = 0          # host count of step() calls
        self._skipped = torch.zeros(1, dtype=torch.int64, device=arena.flat_p.device)   # device count of skipped (non-finite) updates

    def zero_grad(self) -> None:
        self.arena.zero_grad()

    def step(self, lr: Optional[float] = None) -> Tensor:
        """Waits (on the stream) for the gradient exchange, then norm + clip + AdamW.  Returns the device tensor holding
        the total gradient norm before clipping (what ``clip_grad_norm_`` returns); nothing synchronises with the host."""
        from . import ops
        a = self.arena
        a.check()
        scale = a.finish()
        self.step_count += 1
        st = torch.cuda.current_stream().cuda_stream
        ops._call("tfswa_grad_sumsq", a.flat_g.data_ptr(), a.numel, self._sumsq.data_ptr(), st)
        ops._call("tfswa_adamw_clip_step", a.flat_p.data_ptr(), a.flat_g.data_ptr(), self.exp_avg.data_ptr(),
                  self.exp_avg_sq.data_ptr(), a.numel, self._sumsq.data_ptr(), self.grad_norm.data_ptr(), scale,
                  float(self.max_grad_norm or 0.0), float(self.lr if lr is None else lr), self.betas[0], self.betas[1], self.eps,
                  self.weight_decay, self.step_count, self._skipped.data_ptr(), st)
        torch.autograd.graph.increment_version(a.params)     # prepared-weight caches key on the version counters
        return self.grad_norm

    # ---- checkpointing: the layout of torch.optim.AdamW.state_dict() (trainer.py:300-330 saves optimizer.state_dict()) ----
    def state_dict(self) -> Dict:
        a = self.arena
        state = {}
        eff_step = self.step_count - int(self._skipped.item())      # skipped (non-finite) updates do not count (GradScaler semantics)
        for i, (p, o) in enumerate(zip(a.params, a.offsets)):
            n = p.numel()
            state[i] = {"step": torch.tensor(float(eff_step)),
                        "exp_avg": self.exp_avg[o:o + n].view(p.shape).clone(),
                        "exp_avg_sq": self.exp_avg_sq[o:o + n].view(p.shape).clone()}
        group = {"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                 "params": list(range(len(a.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: Dict) -> None:
        a = self.arena
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps, self.weight_decay = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]
        for i, (p, o) in enumerate(zip(a.params, a.offsets)):
            s = sd["state"].get(i)
            if s is None:
                continue
            n = p.numel()
            self.exp_avg[o:o + n].copy_(s["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + n].copy_(s["exp_avg_sq"].reshape(-1))
            self.step_count = int(s["step"])
        self._skipped.zero_()


def cosine_lr(step: int, t_max: int, base_lr: float, eta_min: float = 1e-6) -> float:
    """Closed form of ``CosineAnnealingLR(T_max, eta_min)`` as built at scripts/train.py:258-262."""
    return eta_min + 0.5 * (base_lr - eta_min) * (1.0 + math.cos(math.pi * step / t_max))


def masked_magnitude_l1(model_output: Tensor, mixture_mag: Tensor, target_mags: Sequence[Tensor]) -> Tensor:
    """The trainer's loss: per stem ``sigmoid(|mask|) * |mixture|`` against the target magnitude, mean absolute error,
    averaged over stems (trainer.py:176-199 -> losses.py:235-283 with use_l1=True, use_mrstft=False).

    model_output (B, 2*stems, F, T): (real, imag) mask pair per stem; mixture_mag (B, F, T); target_mags: stems x (B, F, T)."""
    total = model_output.new_zeros(())
    for i, tgt in enumerate(target_mags):
        re, im = model_output[:, 2 * i].float(), model_output[:, 2 * i + 1].float()
        mask = torch.sigmoid(torch.sqrt(re * re + im * im + 1e-8))          # the second sigmoid is the trainer's (trainer.py:182-183)
        total = total + (mixture_mag * mask - tgt).abs().mean()
    return total / len(target_mags)


def enable_recompute(model: nn.Module, policy: str) -> int:
    """Wrap ``forward`` of the selected TFSWABlocks in ``torch.utils.checkpoint`` (non-reentrant).  -> number of blocks."""
    if policy in (None, "", "none"):
        return 0
    if policy not in ("blocks", "stage1"):
        raise ValueError(f"recompute policy {policy!r} not in ('none', 'blocks', 'stage1')")
    from torch.utils.checkpoint import checkpoint
    if policy == "stage1":
        mods = list(model.encoder_stages[0]) + list(model.decoder_stages[len(model.decoder_stages) - 1])
    else:
        mods = [m for m in model.modules() if "TFSWABlock" in type(m).__name__]
    n = 0
    for mod in mods:
        if getattr(mod, "_tfswa_recompute", False):
            continue
        orig = mod.forward

        def fwd(*a, _orig=orig, _mod=mod, **kw):
            if _mod.training and torch.is_grad_enabled():
                return checkpoint(_orig, *a, **kw, use_reentrant=False)
            return _orig(*a, **kw)
        mod.forward = fwd
        mod._tfswa_recompute = True
        n += 1
    return n


class TrainStep:
    """One data-parallel optimisation step of ``model`` (a ``TFSWAUNet`` of this package) on this rank's batch."""

    def __init__(self, model: nn.Module, lr: float = 1e-3, weight_decay: float = 1e-2, max_grad_norm: float = 1.0,
                 betas: Sequence[float] = (0.9, 0.999), eps: float = 1e-8,
                 group: Optional[dist.ProcessGroup] = None, bucket_bytes: int = 16 << 20, recompute: str = "none"):
        """``recompute``: activation-memory policy.  "none" keeps every TFSWABlock's intermediates for backward (115 GiB at
        batch 8 of 6 s segments); "blocks" keeps only each block's input and re-runs its forward kernels in backward
        (what the reference's ``enable_gradient_checkpointing`` does, gradient_checkpoint.py:44-69); "stage1" does that for
        the full-resolution stages only (encoder stage 0 / decoder stage 2: 4 of 22 blocks hold ~70 % of the activations)."""
        self.model = model
        enable_recompute(model, recompute)
        self.arena = FlatArena(model, group, bucket_bytes)
        self.optim = FusedClipAdamW(self.arena, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm)

    def __call__(self, model_input: Tensor, mixture_mag: Tensor, target_mags: Sequence[Tensor], lr: Optional[float] = None,
                 extra_loss=None):
        """-> (loss, grad_norm) as device tensors; no host synchronisation.  ``extra_loss(model_output) -> scalar`` (optional)
        is added to the trainer's L1 term, e.g. the weighted multi-resolution STFT term of ``losses.py``."""
        self.optim.zero_grad()
        out = self.model(model_input)
        loss = masked_magnitude_l1(out, mixture_mag, target_mags)
        if extra_loss is not None:
            loss = loss + extra_loss(out)
        loss.backward()
        norm = self.optim.step(lr)
        return loss.detach(), norm
