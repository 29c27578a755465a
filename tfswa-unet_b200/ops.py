"""Raw (non-autograd) Python wrappers over the C ABI.  Tensors are CUDA tensors owned by PyTorch;
the library only sees device pointers, sizes and the current stream.

Activation convention ("native" tensors): NHWC memory, viewed either as a logical-NCHW torch tensor
with channels_last strides, or as (M, C) / (M, nb, K) token matrices, M = B*H*W.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L

Tensor = torch.Tensor


_DEV = None      # device of the tensors of the op being launched (set by _cuda at the top of every wrapper)


def _stream() -> int:
    """current stream of the TENSORS' device (not of the process-wide current device)"""
    return torch.cuda.current_stream(_DEV).cuda_stream


USE_TC_ATTENTION = True   # tcgen05 attention for axial geometries at head_dim 4/8 (bf16); False forces the SIMT kernel


# ---- launch accounting: C-ABI compute calls (one kernel launch each, except the attention entry points, which launch their
# pre-pass / main / remainder kernels: 175 calls = ~305 kernels per C3 forward, profiles/r2h_launches_summary.md) ----
LAUNCHES = 0
_TIMING = None          # None, or {kernel tag: [(start_event, end_event, work_dict), ...]}
_TAG = None


def reset_launch_count() -> int:
    global LAUNCHES
    n, LAUNCHES = LAUNCHES, 0
    return n


def enable_timing(on: bool) -> None:
    """Per-launch CUDA-event timing on the launching stream (used by bench.py for the roofline figures)."""
    global _TIMING
    _TIMING = {} if on else None


def collect_timing():
    """-> {tag: {"launches": n, "ms": total, "flops": F, "bytes": B}}; call after torch.cuda.synchronize()."""
    out = {}
    for tag, recs in (_TIMING or {}).items():
        ms = sum(s.elapsed_time(e) for s, e, _ in recs)
        out[tag] = {"launches": len(recs), "ms": ms, "flops": sum(w.get("flops", 0) for _, _, w in recs),
                    "bytes": sum(w.get("bytes", 0) for _, _, w in recs), "exps": sum(w.get("exps", 0) for _, _, w in recs)}
    return out


def _call(name: str, *args, tag: str = None, work: dict = None) -> None:
    global LAUNCHES
    if _DEV is not None and _DEV.index is not None and _DEV.index != torch.cuda.current_device():
        with torch.cuda.device(_DEV):       # a model on cuda:1 in a process whose current device is cuda:0
            return _call(name, *args, tag=tag, work=work)
    fn = getattr(L.lib(), name)
    if _TIMING is not None:
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        rc = fn(*args)
        e.record()
        _TIMING.setdefault(tag or name, []).append((s, e, work or {}))
    else:
        rc = fn(*args)
    LAUNCHES += 1
    L.check(rc, name)


def _dt(t: Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return L.BF16
    if t.dtype == torch.float32:
        return L.F32
    raise TypeError(f"unsupported activation dtype {t.dtype}")


def _cuda(*ts) -> None:
    global _DEV
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("tfswa_unet_b200 runs on CUDA (sm_100a) tensors only - there is no CPU fallback")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"tfswa_unet_b200: tensors on different devices ({dev} vs {t.device})")
    _DEV = dev


def _p(t: Optional[Tensor]):
    return None if t is None else t.data_ptr()


def _f32c(t: Optional[Tensor]) -> Optional[Tensor]:
    if t is None:
        return None
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise TypeError("weights / statistics must be contiguous fp32 tensors")
    return t


def _tok3(t: Tensor, name: str) -> Tuple[int, int]:
    """(ld, batch_stride) of a (M, nb, K) token tensor whose last dim is dense."""
    if t.dim() != 3 or t.stride(2) != 1:
        raise ValueError(f"{name}: expected (M, nb, K) with a dense last dim, got {tuple(t.shape)} / {t.stride()}")
    return t.stride(0), (t.stride(1) if t.shape[1] > 1 else 0)


def linear(x: Tensor, w: Tensor, bias: Optional[Tensor] = None, *, prologue: int = 0, epilogue: int = 0,
           row_stats: Optional[Tensor] = None, in_scale: Optional[Tensor] = None, in_shift: Optional[Tensor] = None,
           r1: Optional[Tensor] = None, r2: Optional[Tensor] = None, save_pre: bool = False,
           col_stats: Optional[Tensor] = None, out: Optional[Tensor] = None):
    """y[:, b] = epi(pro(x[:, b]) @ w[b].T + bias[b]) + r1[:, b] + r2[:, b]   (see tfswa_linear_fwd)."""
    _cuda(x, w)
    M, nb, K = x.shape
    nbw, N, Kw = w.shape
    if nbw != nb or Kw != K:
        raise ValueError(f"linear: x {tuple(x.shape)} vs w {tuple(w.shape)}")
    _f32c(w), _f32c(bias), _f32c(row_stats), _f32c(in_scale), _f32c(in_shift), _f32c(col_stats)
    y = out if out is not None else torch.empty((M, nb, N), dtype=x.dtype, device=x.device)
    pre = torch.empty((M, nb, N), dtype=x.dtype, device=x.device) if save_pre else None
    a = L.LinearArgs()
    a.x = x.data_ptr(); a.ldx, a.x_bs = _tok3(x, "x")
    a.w = w.data_ptr(); a.w_bs = N * K
    a.bias = _p(bias); a.bias_bs = N
    a.row_stats = _p(row_stats); a.rs_bs = 2 * M
    a.in_scale = _p(in_scale); a.in_shift = _p(in_shift)
    for name, r in (("r1", r1), ("r2", r2)):
        if r is not None:
            if r.dtype != x.dtype or r.shape[0] != M or r.shape[2] != N or r.shape[1] not in (1, nb):
                raise ValueError(f"linear: residual {name} {tuple(r.shape)}/{r.dtype} incompatible")
            ld, bs = _tok3(r, name)
            setattr(a, name, r.data_ptr()); setattr(a, "ld" + name, ld); setattr(a, name + "_bs", bs)
    a.y = y.data_ptr(); a.ldy, a.y_bs = _tok3(y, "y")
    if pre is not None:
        a.pre = pre.data_ptr(); a.ldpre, a.pre_bs = _tok3(pre, "pre")
    a.col_stats = _p(col_stats)
    a.M, a.N, a.K = M, N, K
    a.prologue, a.epilogue, a.batch, a.dtype = prologue, epilogue, nb, _dt(x)
    _call("tfswa_linear_fwd", C.byref(a), _stream(), tag=f"linear[K={K},N={N},nb={nb}]",
          work={"flops": 2 * M * N * K * nb, "bytes": x.element_size() * M * nb * (K + N)})
    return (y, pre) if save_pre else y


def linear_tc(x: Tensor, w_bf16: Tensor, wsum: Optional[Tensor], bias: Optional[Tensor] = None, *, prologue: int = 0,
              epilogue: int = 0, row_stats: Optional[Tensor] = None, r1: Optional[Tensor] = None,
              r2: Optional[Tensor] = None, col_stats: Optional[Tensor] = None) -> Tensor:
    """tcgen05 version of :func:`linear` (bf16 activations, TMA-fed, TMEM accumulators; see tfswa_linear_tc_fwd).
    ``col_stats`` (2, N) fp32, zero-initialised: receives the column sums / sums of squares of the stored output."""
    _cuda(x, w_bf16)
    M, nb, K = x.shape
    nbw, N, Kw = w_bf16.shape
    if nbw != nb or Kw != K or w_bf16.dtype != torch.bfloat16 or not w_bf16.is_contiguous() or x.dtype != torch.bfloat16:
        raise ValueError(f"linear_tc: x {tuple(x.shape)}/{x.dtype} vs w {tuple(w_bf16.shape)}/{w_bf16.dtype}")
    _f32c(bias), _f32c(row_stats), _f32c(wsum)
    y = torch.empty((M, nb, N), dtype=x.dtype, device=x.device)
    a = L.LinearArgs()
    a.x = x.data_ptr(); a.ldx, a.x_bs = _tok3(x, "x")
    a.w = None; a.w_bs = N * K
    a.bias = _p(bias); a.bias_bs = N
    a.row_stats = _p(row_stats); a.rs_bs = 2 * M
    for name, r in (("r1", r1), ("r2", r2)):
        if r is not None:
            if r.dtype != x.dtype or r.shape[0] != M or r.shape[2] != N or r.shape[1] not in (1, nb):
                raise ValueError(f"linear_tc: residual {name} {tuple(r.shape)}/{r.dtype} incompatible")
            ld, bs = _tok3(r, name)
            setattr(a, name, r.data_ptr()); setattr(a, "ld" + name, ld); setattr(a, name + "_bs", bs)
    a.y = y.data_ptr(); a.ldy, a.y_bs = _tok3(y, "y")
    a.M, a.N, a.K = M, N, K
    a.prologue, a.epilogue, a.batch, a.dtype = prologue, epilogue, nb, L.BF16
    if col_stats is not None:
        if tuple(col_stats.shape) != (2, N) or nb != 1:
            raise ValueError(f"linear_tc: col_stats {tuple(col_stats.shape)} needs shape (2, {N}) and a single problem")
        _f32c(col_stats)
        a.col_stats = col_stats.data_ptr()
    _call("tfswa_linear_tc_fwd", C.byref(a), w_bf16.data_ptr(), _p(wsum), _stream(), tag=f"linear_tc[K={K},N={N},nb={nb}]",
          work={"flops": 2 * M * N * K * nb, "bytes": 2 * M * nb * (K + N) + (2 * M * nb * N if r1 is not None else 0)
                + (2 * M * nb * N if r2 is not None else 0)})
    return y


def branch_tail_tc(att: Tensor, res: Tensor, wp: Tensor, w1: Tensor, w2: Tensor, bp: Tensor, b1: Tensor, b2: Tensor,
                   eps: float = 1e-5) -> Tensor:
    """Fused proj + residual + LN + fc1 + GELU + fc2 + residual of `nb` branches (see tfswa_branch_tail_tc_fwd).
    att (M, nb, C) bf16, res (M, 1|nb, C) bf16, weights bf16 (nb,C,C) / (nb,4C,C) / (nb,C,4C), biases fp32."""
    _cuda(att, res, wp, w1, w2)
    M, nb, C_ = att.shape
    hid = w1.shape[1]
    if (att.dtype != torch.bfloat16 or res.dtype != torch.bfloat16 or res.shape[0] != M or res.shape[2] != C_
            or res.shape[1] not in (1, nb)):
        raise ValueError(f"branch_tail_tc: att {tuple(att.shape)}/{att.dtype} vs res {tuple(res.shape)}/{res.dtype}")
    for name, w, shape in (("wp", wp, (nb, C_, C_)), ("w1", w1, (nb, hid, C_)), ("w2", w2, (nb, C_, hid))):
        if tuple(w.shape) != shape or w.dtype != torch.bfloat16 or not w.is_contiguous():
            raise ValueError(f"branch_tail_tc: {name} {tuple(w.shape)}/{w.dtype}, expected contiguous bf16 {shape}")
    for name, b, n in (("bp", bp, C_), ("b1", b1, hid), ("b2", b2, C_)):
        if tuple(b.shape) != (nb, n):
            raise ValueError(f"branch_tail_tc: {name} {tuple(b.shape)}, expected {(nb, n)}")
        _f32c(b)
    out = torch.empty((M, nb, C_), dtype=att.dtype, device=att.device)
    a = L.TailArgs()
    a.att = att.data_ptr(); a.lda, a.att_bs = _tok3(att, "att")
    a.res = res.data_ptr(); a.ldr, a.res_bs = _tok3(res, "res")
    if res.shape[1] == 1:
        a.res_bs = 0
    a.wp, a.w1, a.w2 = wp.data_ptr(), w1.data_ptr(), w2.data_ptr()
    a.bp, a.b1, a.b2 = bp.data_ptr(), b1.data_ptr(), b2.data_ptr()
    a.out = out.data_ptr(); a.ldo, a.out_bs = _tok3(out, "out")
    a.M, a.C, a.hidden, a.batch, a.eps = M, C_, hid, nb, eps
    _call("tfswa_branch_tail_tc_fwd", C.byref(a), _stream(), tag=f"tail_tc[C={C_},nb={nb}]",
          work={"flops": 2 * M * nb * (C_ * C_ + 2 * C_ * hid), "bytes": 2 * M * C_ * (2 * nb + res.shape[1])})
    return out


def block_head_tc(x: Tensor, wi: Tensor, wq: Tensor, bi: Tensor, bq: Tensor, eps: float = 1e-5):
    """Fused input_proj + LayerNorm + q|k|v GEMM of a block (see tfswa_block_head_tc_fwd).  x (M,1,C) bf16;
    wi (1,C,C), wq (1,9C,C) bf16; bi (1,C), bq (1,9C) fp32 -> x1 (M,1,C), qkv (M,1,9C)."""
    _cuda(x, wi, wq)
    M, nb, C_ = x.shape
    if nb != 1 or x.dtype != torch.bfloat16:
        raise ValueError(f"block_head_tc: x {tuple(x.shape)}/{x.dtype}")
    for name, w, shape in (("wi", wi, (1, C_, C_)), ("wq", wq, (1, 9 * C_, C_))):
        if tuple(w.shape) != shape or w.dtype != torch.bfloat16 or not w.is_contiguous():
            raise ValueError(f"block_head_tc: {name} {tuple(w.shape)}/{w.dtype}, expected contiguous bf16 {shape}")
    for name, b, n in (("bi", bi, C_), ("bq", bq, 9 * C_)):
        if tuple(b.shape) != (1, n):
            raise ValueError(f"block_head_tc: {name} {tuple(b.shape)}, expected {(1, n)}")
        _f32c(b)
    x1 = torch.empty((M, 1, C_), dtype=x.dtype, device=x.device)
    qkv = torch.empty((M, 1, 9 * C_), dtype=x.dtype, device=x.device)
    a = L.HeadArgs()
    a.x, a.ldx = x.data_ptr(), _tok3(x, "x")[0]
    a.wi, a.wq, a.bi, a.bq = wi.data_ptr(), wq.data_ptr(), bi.data_ptr(), bq.data_ptr()
    a.x1, a.ld1 = x1.data_ptr(), C_
    a.qkv, a.ldq = qkv.data_ptr(), 9 * C_
    a.M, a.C, a.eps = M, C_, eps
    _call("tfswa_block_head_tc_fwd", C.byref(a), _stream(), tag=f"head_tc[C={C_}]",
          work={"flops": 2 * M * (C_ * C_ + 9 * C_ * C_), "bytes": 2 * M * C_ * 11})
    return x1, qkv


def row_stats(x: Tensor) -> Tensor:
    """(M, nb, K) -> (nb, M, 2) fp32 {mean, rstd} over K (LayerNorm statistics, eps 1e-5)."""
    _cuda(x)
    M, nb, K = x.shape
    ld, bs = _tok3(x, "x")
    st = torch.empty((nb, M, 2), dtype=torch.float32, device=x.device)
    _call("tfswa_row_stats", x.data_ptr(), ld, bs, st.data_ptr(), 2 * M, M, K, nb, _dt(x), _stream())
    return st


def attention(qkv: Tensor, out: Tensor, B: int, H: int, W: int, C_: int, heads: int, geom: int, *, ws: int = 8,
              shift: int = 0, pad_kv: Optional[Tensor] = None, rel_bias: Optional[Tensor] = None,
              use_shift_mask: bool = False, lse: Optional[Tensor] = None, force_exact: bool = False) -> Tensor:
    """qkv: (M, >=3C) view with q|k|v at columns 0|C|2C; out: (M, C) view (both row-strided, dense rows)."""
    _cuda(qkv, out)
    if qkv.dim() != 2 or out.dim() != 2 or qkv.stride(1) != 1 or out.stride(1) != 1:
        raise ValueError("attention: qkv/out must be 2-D row-strided views")
    if qkv.shape[0] != B * H * W or out.shape[0] != B * H * W or qkv.shape[1] != 3 * C_ or out.shape[1] != C_:
        raise ValueError(f"attention: bad shapes qkv {tuple(qkv.shape)} out {tuple(out.shape)}")
    _f32c(pad_kv), _f32c(rel_bias), _f32c(lse)
    a = L.AttnArgs()
    a.qkv, a.ldq, a.out, a.ldo = qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0)
    a.lse, a.pad_kv, a.rel_bias = _p(lse), _p(pad_kv), _p(rel_bias)
    a.B, a.H, a.W, a.C, a.heads = B, H, W, C_, heads
    a.geom, a.ws, a.shift, a.use_shift_mask, a.dtype = geom, ws, shift, int(use_shift_mask), _dt(qkv)
    a.flags = 1 if force_exact else 0          # TFSWA_ATTN_FORCE_EXACT (tests)
    d = C_ // heads
    tc = USE_TC_ATTENTION and qkv.dtype == torch.bfloat16 and geom != L.GEOM_SWA and d in (4, 8, 16) and heads * d == C_ and C_ <= 128
    if tc:
        # the tensor-core kernels work on 128-query tiles (keys in tiles of 4*d, zero-filled by TMA); short sequences that
        # fill the query tiles badly (e.g. 129 tokens -> 25 % useful work) stay on the SIMT kernel, which has no padding
        n = H if geom == L.GEOM_TSA else W
        kt = 4 * d
        rem = n % 128
        nq = n - rem if (n >= 128 and 0 < rem < 32) else n       # a short remainder goes to the key-split warp kernel
        fill = (nq / (-(-nq // 128) * 128)) * (n / (-(-n // kt) * kt))
        tc = fill >= 0.4
    win_tc = (USE_TC_ATTENTION and qkv.dtype == torch.bfloat16 and geom == L.GEOM_SWA and ws == 8 and rel_bias is None
              and not (use_shift_mask and shift > 0) and d in (4, 8, 16, 32) and heads * d == C_
              and (C_ == 32 or C_ % 64 == 0) and (d < 32 or C_ >= 64))
    tag = f"{'attn_tc' if (tc or win_tc) else 'attn'}[{('tsa', 'fsa', 'swa')[geom]},d={d}]"
    work = _attn_work(B, H, W, C_, geom, ws, qkv.element_size(), heads)
    if win_tc:
        _call("tfswa_attn_win_tc_fwd", C.byref(a), _stream(), tag=tag, work=work)
    elif tc:
        nbytes = L.lib().tfswa_attn_tc_scratch_bytes(C.byref(a))
        scratch = torch.empty((nbytes,), dtype=torch.uint8, device=qkv.device)
        _call("tfswa_attn_tc_fwd", C.byref(a), scratch.data_ptr(), nbytes, _stream(), tag=tag, work=work)
    else:
        _call("tfswa_attn_fwd", C.byref(a), _stream(), tag=tag, work=work)
    return out


def _attn_work(B, H, W, C_, geom, ws, esize, heads=8):
    """algorithmic work of one attention launch: 4*tokens*N*C FLOPs (QK^T + PV), one exp per score element per head"""
    if geom == L.GEOM_SWA:
        Hp, Wp = -(-H // ws) * ws, -(-W // ws) * ws
        toks, N = B * Hp * Wp, ws * ws
    else:
        toks, N = B * H * W, (H if geom == L.GEOM_TSA else W)
    return {"flops": 4 * toks * N * C_, "bytes": esize * B * H * W * 4 * C_, "exps": toks * N * heads}


def conv(x: Tensor, w: Tensor, bias: Tensor, kind: int, out_hw: Tuple[int, int], *, epilogue: int = 0,
         save_pre: bool = False, col_stats: Optional[Tensor] = None):
    """x: native (B, Cin, Hin, Win) channels_last; w already re-laid-out (see tfswa_conv_args)."""
    _cuda(x, w)
    B, Cin, Hin, Win = x.shape
    Hout, Wout = out_hw
    Cout = bias.shape[0]
    _f32c(w), _f32c(bias), _f32c(col_stats)
    y = torch.empty((B, Cout, Hout, Wout), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    pre = torch.empty_like(y) if save_pre else None
    a = L.ConvArgs()
    a.x, a.y, a.pre, a.w, a.bias, a.col_stats = x.data_ptr(), y.data_ptr(), _p(pre), w.data_ptr(), bias.data_ptr(), _p(col_stats)
    a.B, a.Hin, a.Win, a.Cin, a.Hout, a.Wout, a.Cout = B, Hin, Win, Cin, Hout, Wout, Cout
    a.kind, a.epilogue, a.dtype = kind, epilogue, _dt(x)
    _call("tfswa_conv_fwd", C.byref(a), _stream())
    return (y, pre) if save_pre else y


def conv_tc(x: Tensor, w_bf16: Tensor, bias: Tensor, kind: int, out_hw: Tuple[int, int], *, epilogue: int = 0,
            col_stats: Optional[Tensor] = None) -> Tensor:
    """tcgen05 version of :func:`conv` (bf16, no pre output); w_bf16 = kernel-layout weights cast to bf16.
    ``col_stats`` (2, Cout) fp32, zero-initialised: receives the channel sums / sums of squares of the stored output."""
    _cuda(x, w_bf16)
    B, Cin, Hin, Win = x.shape
    Hout, Wout = out_hw
    Cout = bias.shape[0]
    if x.dtype != torch.bfloat16 or w_bf16.dtype != torch.bfloat16 or not w_bf16.is_contiguous():
        raise TypeError("conv_tc: bf16 activations and contiguous bf16 weights required")
    _f32c(bias)
    y = torch.empty((B, Cout, Hout, Wout), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    a = L.ConvArgs()
    a.x, a.y, a.bias = x.data_ptr(), y.data_ptr(), bias.data_ptr()
    a.B, a.Hin, a.Win, a.Cin, a.Hout, a.Wout, a.Cout = B, Hin, Win, Cin, Hout, Wout, Cout
    a.kind, a.epilogue, a.dtype = kind, epilogue, L.BF16
    if col_stats is not None:
        if tuple(col_stats.shape) != (2, Cout):
            raise ValueError(f"conv_tc: col_stats {tuple(col_stats.shape)}, expected (2, {Cout})")
        _f32c(col_stats)
        a.col_stats = col_stats.data_ptr()
    taps = (9, 16, 16)[kind]
    _call("tfswa_conv_tc_fwd", C.byref(a), w_bf16.data_ptr(), _stream(), tag=f"conv_tc[kind={kind},Cin={Cin},Cout={Cout}]",
          work={"flops": 2 * B * Hout * Wout * Cout * Cin * (taps if kind != 2 else 4),
                "bytes": 2 * (x.numel() + y.numel())})
    return y


def stem(x_nchw: Tensor, w: Tensor, bias: Tensor, dtype: torch.dtype, *, epilogue: int = 0, save_pre: bool = False,
         col_stats: Optional[Tensor] = None):
    _cuda(x_nchw, w)
    if x_nchw.dtype != torch.float32 or not x_nchw.is_contiguous():
        raise TypeError("stem: model input must be a contiguous fp32 NCHW tensor")
    B, Cin, H, W = x_nchw.shape
    Cout = w.shape[0]
    _f32c(w), _f32c(bias), _f32c(col_stats)
    y = torch.empty((B, Cout, H, W), dtype=dtype, device=x_nchw.device, memory_format=torch.channels_last)
    pre = torch.empty_like(y) if save_pre else None
    _call("tfswa_stem_fwd", x_nchw.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), _p(pre), _p(col_stats),
                                   B, Cin, H, W, Cout, epilogue, _dt(y), _stream())
    return (y, pre) if save_pre else y


def head_tail(v: Tensor, w3: Tensor, b3: Tensor, scale: Optional[Tensor] = None, shift: Optional[Tensor] = None,
              want_logits: bool = False):
    _cuda(v, w3)
    B, C_, H, W = v.shape
    Cout = w3.shape[0]
    _f32c(w3), _f32c(b3), _f32c(scale), _f32c(shift)
    masks = torch.empty((B, Cout, H, W), dtype=torch.float32, device=v.device)
    logits = torch.empty_like(masks) if want_logits else None
    _call("tfswa_head_tail_fwd", v.data_ptr(), _p(scale), _p(shift), w3.data_ptr(), b3.data_ptr(), masks.data_ptr(),
                                        _p(logits), B, H, W, C_, Cout, _dt(v), _stream())
    return (masks, logits) if want_logits else masks


def bn_finalize(col_stats: Tensor, count: int, gamma: Tensor, beta: Tensor, running_mean: Optional[Tensor],
                running_var: Optional[Tensor], momentum: float, eps: float):
    """-> (scale, shift, save_mean_rstd); updates running stats in place (train-mode BatchNorm2d)."""
    _cuda(col_stats)
    Cn = gamma.shape[0]
    scale = torch.empty(Cn, dtype=torch.float32, device=col_stats.device)
    shift = torch.empty_like(scale)
    save = torch.empty((2, Cn), dtype=torch.float32, device=col_stats.device)
    _call("tfswa_bn_finalize", col_stats.data_ptr(), count, gamma.data_ptr(), beta.data_ptr(), _p(running_mean),
                                      _p(running_var), momentum, eps, scale.data_ptr(), shift.data_ptr(), save.data_ptr(),
                                      Cn, _stream())
    return scale, shift, save


def affine_act(v: Tensor, scale: Optional[Tensor], shift: Optional[Tensor], *, epilogue: int = 0,
               r1: Optional[Tensor] = None, r2: Optional[Tensor] = None) -> Tensor:
    """native (B,C,H,W) channels_last (or any dense (..., C) token tensor): y = act(v*scale+shift)+r1+r2."""
    _cuda(v)
    Cn = v.shape[1] if v.dim() == 4 else v.shape[-1]
    y = torch.empty_like(v)
    def dense(t):
        return t.is_contiguous(memory_format=torch.channels_last) if t.dim() == 4 else t.is_contiguous()
    if not dense(v):
        raise ValueError("affine_act: v must be a dense NHWC tensor")
    for r in (r1, r2):
        if r is not None and (r.shape != v.shape or not dense(r) or r.dtype != v.dtype):
            raise ValueError("affine_act: residual must match v's shape, layout and dtype")
    _call("tfswa_affine_act", v.data_ptr(), _p(scale), _p(shift), _p(r1), _p(r2), y.data_ptr(), v.numel() // Cn, Cn,
                                     epilogue, _dt(v), _stream())
    return y


def bilinear(x: Tensor, out_hw: Tuple[int, int]) -> Tensor:
    _cuda(x)
    B, Cn, Hin, Win = x.shape
    y = torch.empty((B, Cn, out_hw[0], out_hw[1]), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
    L.check(L.lib().tfswa_bilinear_fwd(x.data_ptr(), y.data_ptr(), B, Hin, Win, out_hw[0], out_hw[1], Cn, _dt(x), _stream()),
            "bilinear_fwd")
    return y


# ================================================================================================
# backward wrappers
# ================================================================================================
def _dense_tok(t: Tensor) -> Tensor:
    return t if t.is_contiguous() else t.contiguous()


def linear_wgrad(x: Tensor, g: Tensor, *, prologue: int = 0, row_stats: Optional[Tensor] = None, want_bias: bool = True):
    """dW (nb,N,K), dbias (nb,N) fp32 for y = pro(x) W^T + b given g = dL/d(pre-epilogue y)."""
    _cuda(x, g)
    M, nb, K = x.shape
    N = g.shape[2]
    dw = torch.zeros((nb, N, K), dtype=torch.float32, device=x.device)
    db = torch.zeros((nb, N), dtype=torch.float32, device=x.device) if want_bias else None
    a = L.LinearArgs()
    a.x = x.data_ptr(); a.ldx, a.x_bs = _tok3(x, "x")
    a.row_stats = _p(row_stats); a.rs_bs = 2 * M
    a.M, a.N, a.K = M, N, K
    a.prologue, a.batch, a.dtype = prologue, nb, _dt(x)
    ldg, g_bs = _tok3(g, "g")
    _call("tfswa_linear_wgrad", C.byref(a), g.data_ptr(), ldg, g_bs, dw.data_ptr(), _p(db), _stream(),
          tag=f"wgrad[K={K},N={N},nb={nb}]", work={"flops": 2 * M * N * K * nb})
    return dw, db


def conv_wgrad(x: Tensor, g: Tensor, kind: int, wl_shape, want_bias: bool = True):
    _cuda(x, g)
    B, Cin, Hin, Win = x.shape
    _, Cout, Hout, Wout = g.shape
    dw = torch.zeros(tuple(wl_shape), dtype=torch.float32, device=x.device)
    db = torch.zeros((Cout,), dtype=torch.float32, device=x.device) if want_bias else None
    a = L.ConvArgs()
    a.x = x.data_ptr()
    a.B, a.Hin, a.Win, a.Cin, a.Hout, a.Wout, a.Cout = B, Hin, Win, Cin, Hout, Wout, Cout
    a.kind, a.dtype = kind, _dt(x)
    _call("tfswa_conv_wgrad", C.byref(a), g.data_ptr(), dw.data_ptr(), _p(db), _stream())
    return dw, db


def act_bwd(g: Tensor, pre: Tensor, mode: int = 0, ds: Optional[Tensor] = None) -> Tensor:
    """mode 0: g*gelu'(pre); mode 1: g + ds[0][c] + 2*pre*ds[1][c]; dense tensors of identical shape/layout."""
    _cuda(g, pre)
    if g.shape != pre.shape or g.dtype != pre.dtype:
        raise ValueError("act_bwd: g/pre mismatch")
    Cn = g.shape[1] if g.dim() == 4 else g.shape[-1]
    out = torch.empty_like(pre)
    _call("tfswa_act_bwd", g.data_ptr(), pre.data_ptr(), _p(_f32c(ds)), out.data_ptr(), g.numel() // Cn, Cn, mode, _dt(g), _stream())
    return out


def affine_act_bwd(dy: Tensor, v: Tensor, scale: Optional[Tensor], shift: Optional[Tensor], epilogue: int, want_params: bool):
    _cuda(dy, v)
    Cn = v.shape[1] if v.dim() == 4 else v.shape[-1]
    dv = torch.empty_like(v)
    dsc = torch.zeros(Cn, dtype=torch.float32, device=v.device) if want_params else None
    dsh = torch.zeros(Cn, dtype=torch.float32, device=v.device) if want_params else None
    _call("tfswa_affine_act_bwd", dy.data_ptr(), v.data_ptr(), _p(scale), _p(shift), dv.data_ptr(), _p(dsc), _p(dsh),
          v.numel() // Cn, Cn, epilogue, _dt(v), _stream())
    return dv, dsc, dsh


def lnhat_bwd(da: Tensor, x: Tensor, stats: Tensor) -> Tensor:
    _cuda(da, x)
    M, nb, K = x.shape
    dx = torch.empty((M, nb, K), dtype=x.dtype, device=x.device)
    ldd, dbs = _tok3(da, "da"); ldx, xbs = _tok3(x, "x"); ldo, obs = _tok3(dx, "dx")
    _call("tfswa_lnhat_bwd", da.data_ptr(), ldd, dbs, x.data_ptr(), ldx, xbs, stats.data_ptr(), 2 * M, dx.data_ptr(), ldo, obs,
          M, K, nb, _dt(x), _stream())
    return dx


def sum_batch(g: Tensor) -> Tensor:
    """(M, nb, N) dense -> (M, 1, N)"""
    _cuda(g)
    g = _dense_tok(g)
    M, nb, N = g.shape
    out = torch.empty((M, 1, N), dtype=g.dtype, device=g.device)
    _call("tfswa_sum_batch", g.data_ptr(), out.data_ptr(), M, nb, N, _dt(g), _stream())
    return out


def bilinear_bwd(dy: Tensor, in_hw: Tuple[int, int]) -> Tensor:
    _cuda(dy)
    B, Cn, Hout, Wout = dy.shape
    dx = torch.empty((B, Cn, in_hw[0], in_hw[1]), dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
    _call("tfswa_bilinear_bwd", dy.data_ptr(), dx.data_ptr(), B, in_hw[0], in_hw[1], Hout, Wout, Cn, _dt(dy), _stream())
    return dx


def attention_bwd(qkv: Tensor, out: Tensor, lse: Tensor, dout: Tensor, dqkv: Tensor, dsum: Tensor, B: int, H: int, W: int,
                  C_: int, heads: int, geom: int, *, ws: int = 8, shift: int = 0, pad_kv: Optional[Tensor] = None,
                  dpad: Optional[Tensor] = None) -> None:
    _cuda(qkv, out, dout, dqkv)
    if out.stride(0) != dout.stride(0) or qkv.stride(0) != dqkv.stride(0):
        raise ValueError("attention_bwd: (out, dout) and (qkv, dqkv) must share their row strides")
    a = L.AttnArgs()
    a.qkv, a.ldq, a.out, a.ldo = qkv.data_ptr(), qkv.stride(0), out.data_ptr(), out.stride(0)
    a.lse, a.pad_kv = lse.data_ptr(), _p(pad_kv)
    a.B, a.H, a.W, a.C, a.heads = B, H, W, C_, heads
    a.geom, a.ws, a.shift, a.use_shift_mask, a.dtype = geom, ws, shift, 0, _dt(qkv)
    _call("tfswa_attn_bwd", C.byref(a), dout.data_ptr(), dqkv.data_ptr(), dsum.data_ptr(), _p(dpad), _stream(),
          tag=f"attn_bwd[{('tsa', 'fsa', 'swa')[geom]},d={C_ // heads}]")


def stem_bwd(x_nchw: Tensor, w: Tensor, g: Tensor, want_dx: bool):
    _cuda(x_nchw, g)
    B, Cin, H, W = x_nchw.shape
    Cout = w.shape[0]
    dw = torch.zeros_like(w)
    db = torch.zeros((Cout,), dtype=torch.float32, device=w.device)
    dx = torch.empty_like(x_nchw) if want_dx else None
    _call("tfswa_stem_bwd", x_nchw.data_ptr(), w.data_ptr(), g.data_ptr(), _p(dx), dw.data_ptr(), db.data_ptr(), B, Cin, H, W,
          Cout, _dt(g), _stream())
    return dx, dw, db


def head_tail_bwd(v: Tensor, scale: Optional[Tensor], shift: Optional[Tensor], w3: Tensor, b3: Tensor,
                  dmasks: Optional[Tensor], dlogits: Optional[Tensor]):
    _cuda(v, w3)
    B, C_, H, W = v.shape
    Cout = w3.shape[0]
    dv = torch.empty_like(v)
    dw3 = torch.zeros_like(w3)
    db3 = torch.zeros_like(b3)
    dsc = torch.zeros(C_, dtype=torch.float32, device=v.device) if scale is not None else None
    dsh = torch.zeros(C_, dtype=torch.float32, device=v.device) if scale is not None else None
    _call("tfswa_head_tail_bwd", v.data_ptr(), _p(scale), _p(shift), w3.data_ptr(), b3.data_ptr(), _p(dmasks), _p(dlogits),
          dv.data_ptr(), dw3.data_ptr(), db3.data_ptr(), _p(dsc), _p(dsh), B, H, W, C_, Cout, _dt(v), _stream())
    return dv, dw3, db3, dsc, dsh


# ---- spectrogram steps around the model in overlap-add separation -------------
def spec_pack_norm(spec: Tensor, normalize: bool = True, eps: float = 1e-8):
    """complex64 STFT (B, F, T) -> model input (B, 2, F, T) fp32 [real | imag] (stft_processor.py:186-204), instance-normalised
    over time when ``normalize`` (stft_processor.py:283-297); returns (x, stats (B, 2, F, 2) = {mean, std + eps} or None)."""
    _cuda(spec)
    if spec.dtype != torch.complex64 or spec.dim() != 3 or not spec.is_contiguous():
        raise ValueError("spec_pack_norm: contiguous complex64 (B, F, T) expected")
    B, F, T = spec.shape
    x = torch.empty((B, 2, F, T), dtype=torch.float32, device=spec.device)
    stats = torch.empty((B, 2, F, 2), dtype=torch.float32, device=spec.device) if normalize else None
    _call("tfswa_spec_pack_norm", spec.data_ptr(), x.data_ptr(), _p(stats), B, F, T, float(eps), int(normalize), _stream(),
          work={"bytes": 16 * B * F * T})
    return x, stats


def spec_mask_apply(masks: Tensor, spec: Tensor, stats: Optional[Tensor]) -> Tensor:
    """masks (B, S, F, T) fp32, spec (B, F, T) complex64 -> stems (B, S, F, T) complex64 = spec * (masks * std + mean)
    (inference.py:132-145); ``stats`` None = masks applied as they are."""
    _cuda(masks, spec)
    if masks.dtype != torch.float32 or not masks.is_contiguous() or spec.dtype != torch.complex64 or not spec.is_contiguous():
        raise ValueError("spec_mask_apply: contiguous fp32 masks and complex64 spec expected")
    B, S, F, T = masks.shape
    if tuple(spec.shape) != (B, F, T):
        raise ValueError(f"spec_mask_apply: spec {tuple(spec.shape)} does not match masks {tuple(masks.shape)}")
    out = torch.empty((B, S, F, T), dtype=torch.complex64, device=masks.device)
    _call("tfswa_spec_mask_apply", masks.data_ptr(), spec.data_ptr(), _p(stats), out.data_ptr(), B, S, F, T,
          int(stats is not None), _stream(), work={"bytes": (12 * S + 8) * B * F * T})
    return out


def ola_add(wav: Tensor, starts, win: Tensor, acc: Tensor, seg_len: int) -> None:
    """acc (S + 1, total) += Hann-weighted segments: wav (nseg, S, L) fp32, ``starts`` ascending python ints (inference.py:209-216)."""
    _cuda(wav, win, acc)
    if wav.dtype != torch.float32 or not wav.is_contiguous() or acc.dtype != torch.float32 or not acc.is_contiguous():
        raise ValueError("ola_add: contiguous fp32 buffers expected")
    nseg, S, Lw = wav.shape
    if len(starts) != nseg or acc.shape[0] != S + 1 or win.numel() < min(seg_len, Lw) or list(starts) != sorted(starts):
        raise ValueError("ola_add: inconsistent segment list / buffers")
    st = torch.tensor(list(starts), dtype=torch.int64).to(wav.device, non_blocking=True)
    _call("tfswa_ola_add", wav.data_ptr(), st.data_ptr(), int(starts[0]), int(starts[-1]), _f32c(win).data_ptr(), acc.data_ptr(),
          nseg, S, Lw, int(seg_len), acc.shape[1], _stream())


def mrstft_mag_loss(pred_spec: Tensor, target_spec: Tensor, w_mag: float, w_log: float, eps: float, want_grad: bool):
    """One resolution of the MR-STFT loss on two complex64 STFTs of identical dense layout (losses.py:125-141, 171-183):
    -> (loss fp64 (1,), d loss / d pred_spec or None)."""
    _cuda(pred_spec, target_spec)
    if pred_spec.dtype != torch.complex64 or target_spec.dtype != torch.complex64 or pred_spec.shape != target_spec.shape \
            or pred_spec.stride() != target_spec.stride():
        raise ValueError("mrstft_mag_loss: two complex64 tensors of the same shape and strides expected")
    n = pred_spec.numel()
    dense = torch.empty_like(pred_spec)          # preserve_format: the same dense layout as the input (e.g. stft's (B, T, F) storage)
    if dense.stride() != pred_spec.stride() or not (pred_spec.is_contiguous() or pred_spec.transpose(-1, -2).is_contiguous()):
        raise ValueError("mrstft_mag_loss: dense (non-overlapping, gap-free) spectrograms expected")
    grad = dense if want_grad else None
    loss = torch.zeros((1,), dtype=torch.float64, device=pred_spec.device)
    _call("tfswa_mrstft_mag_loss", pred_spec.data_ptr(), target_spec.data_ptr(), n, float(w_mag), float(w_log), float(eps),
          loss.data_ptr(), _p(grad), _stream(), work={"bytes": (16 + (8 if want_grad else 0)) * n})
    return loss, grad

