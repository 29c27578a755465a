"""``torch.autograd.Function`` wrappers: every forward and backward below is a sequence of C-ABI kernel calls
(``ops``).  PyTorch contributes graph bookkeeping, memory, and differentiable re-layouts of the (tiny) weight
tensors only.  Reference: plain autograd over the eager graph; here nothing N x N is saved - the
attention backward recomputes the weights from q, k and the saved log-sum-exp."""
from __future__ import annotations

import torch

from . import _lib as L
from . import ops

Tensor = torch.Tensor
_KIND = {"conv3": 0, "down": 1, "up": 2}
_UP_TAPS = ((1, 3), (0, 2))     # kernel rows used by output parity 0 / 1 (tap a=0 is the nearer input row)


def _contig(t: Tensor) -> Tensor:
    return t if t.is_contiguous() else t.contiguous()


def _cl(t: Tensor) -> Tensor:
    return t if t.is_contiguous(memory_format=torch.channels_last) else t.contiguous(memory_format=torch.channels_last)


def _tc_ok(x: Tensor, N: int, K: int) -> bool:
    from . import functional as Fn
    return Fn.USE_TC and x.dtype == torch.bfloat16 and K % 32 == 0 and N % 16 == 0


def _matmul(x: Tensor, w: Tensor, bias, **kw) -> Tensor:
    """forward-style token GEMM on the best available kernel (tcgen05 for bf16, SIMT otherwise)"""
    pro = kw.get("prologue", 0)
    if _tc_ok(x, w.shape[1], w.shape[2]) and pro in (L.PRO_NONE, L.PRO_LNHAT):
        wb = w.to(torch.bfloat16).contiguous()
        return ops.linear_tc(x, wb, wb.float().sum(-1).contiguous() if pro == L.PRO_LNHAT else None, bias, **kw)
    return ops.linear(x, w, bias, **kw)


class LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, bias, row_stats, r1, r2, prologue, epilogue, want_col_stats):
        w = _contig(w.detach())
        b = None if bias is None else _contig(bias.detach())
        ctx.cfg = (prologue, epilogue, want_col_stats, bias is not None,
                   None if r1 is None else r1.shape[1], None if r2 is None else r2.shape[1], x.shape[1])
        if want_col_stats:
            stats = torch.zeros((2, w.shape[1]), dtype=torch.float32, device=x.device)
            if _tc_ok(x, w.shape[1], w.shape[2]) and prologue == L.PRO_NONE and x.shape[1] == 1:
                pre = ops.linear_tc(x, w.to(torch.bfloat16).contiguous(), None, b, col_stats=stats)   # BatchNorm sums from the epilogue
            else:
                pre = ops.linear(x, w, b, prologue=prologue, row_stats=row_stats, col_stats=stats)
            ctx.save_for_backward(x, w, row_stats, pre)
            return pre, stats
        if epilogue == L.EPI_GELU:
            y, pre = ops.linear(x, w, b, prologue=prologue, epilogue=epilogue, row_stats=row_stats, r1=r1, r2=r2, save_pre=True)
        else:
            y, pre = _matmul(x, w, b, prologue=prologue, row_stats=row_stats, r1=r1, r2=r2), None
        ctx.save_for_backward(x, w, row_stats, pre)
        return y

    @staticmethod
    def backward(ctx, dy, dstats=None):
        x, w, row_stats, pre = ctx.saved_tensors
        prologue, epilogue, want_col_stats, has_bias, r1_nb, r2_nb, nb = ctx.cfg
        g = _contig(dy)
        if want_col_stats:
            gpre = ops.act_bwd(g, pre, 1, _contig(dstats.float())) if dstats is not None else g
        elif epilogue == L.EPI_GELU:
            gpre = ops.act_bwd(g, pre, 0)
        else:
            gpre = g
        dr1 = dr2 = None
        if r1_nb is not None and ctx.needs_input_grad[4]:
            dr1 = ops.sum_batch(g) if (r1_nb == 1 and nb > 1) else g
        if r2_nb is not None and ctx.needs_input_grad[5]:
            dr2 = ops.sum_batch(g) if (r2_nb == 1 and nb > 1) else g
        dw = db = dx = None
        if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
            dw, db = ops.linear_wgrad(x, gpre, prologue=prologue, row_stats=row_stats, want_bias=has_bias)
        if ctx.needs_input_grad[0]:
            wt = w.transpose(1, 2).contiguous()                      # (nb, K, N): dA = gpre @ W
            da = _matmul(gpre, wt, None)
            if prologue & L.PRO_LNHAT:
                dx = ops.lnhat_bwd(da, x, row_stats)
            elif prologue & L.PRO_GELU:
                dx = ops.act_bwd(da, _contig(x), 0)
            else:
                dx = da
        return dx, dw, db, None, dr1, dr2, None, None, None


class GeluFn(torch.autograd.Function):
    """h = GELU(u) as its own node (bf16 training path): keeping h lets fc2 run forward, data- and weight-gradient on
    the tensor-core kernels with a plain A operand instead of the CUDA-core GEMM with a GELU-on-load prologue."""
    @staticmethod
    def forward(ctx, u):
        ctx.save_for_backward(u)
        return ops.affine_act(u, None, None, epilogue=L.EPI_GELU)

    @staticmethod
    def backward(ctx, dh):
        (u,) = ctx.saved_tensors
        return ops.act_bwd(_contig(dh), u, 0)


class GeluLinearFn(torch.autograd.Function):
    """y = GELU(u) W^T + b (+ r1) as ONE node (bf16 training path of the MLP: attention.py:121-128 - fc2 over the activated hidden
    tensor).  Forward = the same two kernels as GeluFn + LinearFn; the point is the backward: dL/du = (g W) * gelu'(u) comes out of
    the data-gradient GEMM's epilogue (TFSWA_EPI_MUL_DGELU, the saved pre-activation u arrives as the GEMM's residual tile), so the
    4C-wide dL/dh never goes to HBM and back and the separate GELU-backward launch disappears."""
    @staticmethod
    def forward(ctx, u, w, bias, r1):
        w = _contig(w.detach())
        b = None if bias is None else _contig(bias.detach())
        h = ops.affine_act(u, None, None, epilogue=L.EPI_GELU)
        y = _matmul(h, w, b, r1=r1)
        ctx.cfg = (bias is not None, None if r1 is None else r1.shape[1], u.shape[1])
        ctx.save_for_backward(u, h, w)
        return y

    @staticmethod
    def backward(ctx, dy):
        u, h, w = ctx.saved_tensors
        has_bias, r1_nb, nb = ctx.cfg
        g = _contig(dy)
        dr1 = None
        if r1_nb is not None and ctx.needs_input_grad[3]:
            dr1 = ops.sum_batch(g) if (r1_nb == 1 and nb > 1) else g
        dw = db = du = None
        if ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]):
            dw, db = ops.linear_wgrad(h, g, want_bias=has_bias)
        if ctx.needs_input_grad[0]:
            wt = w.transpose(1, 2).contiguous()                      # (nb, K, N): dh = g @ W
            if _tc_ok(g, wt.shape[1], wt.shape[2]) and u.is_contiguous():
                du = ops.linear_tc(g, wt.to(torch.bfloat16).contiguous(), None, None, epilogue=L.EPI_MUL_DGELU, r1=u)
            else:
                du = ops.act_bwd(_matmul(g, wt, None), u, 0)
        return du, dw, db, dr1


class AttentionFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, qkv, pad_kv, rel_bias, B, H, W, C, heads, geom, ws, shift, use_shift_mask):
        if rel_bias is not None or use_shift_mask:
            raise NotImplementedError("the optional Swin mask / relative-position bias are forward-only (default-off features)")
        M = B * H * W
        out = torch.empty((M, C), dtype=qkv.dtype, device=qkv.device)
        lse = torch.empty((M, heads), dtype=torch.float32, device=qkv.device)
        pk = None if pad_kv is None else _contig(pad_kv.detach().float())
        ops.attention(qkv, out, B, H, W, C, heads, geom, ws=ws, shift=shift, pad_kv=pk, lse=lse)
        ctx.save_for_backward(qkv, out, lse, pk)
        ctx.cfg = (B, H, W, C, heads, geom, ws, shift, pad_kv is not None)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse, pk = ctx.saved_tensors
        B, H, W, C, heads, geom, ws, shift, has_pad = ctx.cfg
        M = B * H * W
        dout = _contig(dout)
        dqkv = torch.empty((M, 3 * C), dtype=qkv.dtype, device=qkv.device)
        dsum = torch.empty((M, heads), dtype=torch.float32, device=qkv.device)
        dpad = torch.zeros((2 * C,), dtype=torch.float32, device=qkv.device) if has_pad else None
        qv = qkv if qkv.stride(0) == 3 * C else _contig(qkv)
        ops.attention_bwd(qv, out, lse, dout, dqkv, dsum, B, H, W, C, heads, geom, ws=ws, shift=shift, pad_kv=pk, dpad=dpad)
        return (dqkv, dpad) + (None,) * 10


class Attention3Fn(torch.autograd.Function):
    """TSA | FSA | SW-MSA cores over one (M, 3, 3C) q|k|v buffer -> (M, 3, C)."""

    @staticmethod
    def forward(ctx, qkv3, pad_kv, rel_bias, B, H, W, C, heads, ws, shift, use_shift_mask):
        if rel_bias is not None or use_shift_mask:
            raise NotImplementedError("the optional Swin mask / relative-position bias are forward-only (default-off features)")
        M = B * H * W
        qkv3 = _contig(qkv3)
        att = torch.empty((M, 3, C), dtype=qkv3.dtype, device=qkv3.device)
        lse = torch.empty((3, M, heads), dtype=torch.float32, device=qkv3.device)
        pk = _contig(pad_kv.detach().float())
        for b, geom in enumerate((L.GEOM_TSA, L.GEOM_FSA, L.GEOM_SWA)):
            ops.attention(qkv3[:, b, :], att[:, b, :], B, H, W, C, heads, geom, ws=ws, shift=shift,
                          pad_kv=pk if geom == L.GEOM_SWA else None, lse=lse[b])
        ctx.save_for_backward(qkv3, att, lse, pk)
        ctx.cfg = (B, H, W, C, heads, ws, shift)
        return att

    @staticmethod
    def backward(ctx, datt):
        qkv3, att, lse, pk = ctx.saved_tensors
        B, H, W, C, heads, ws, shift = ctx.cfg
        M = B * H * W
        datt = _contig(datt)
        dqkv3 = torch.empty_like(qkv3)
        dsum = torch.empty((M, heads), dtype=torch.float32, device=qkv3.device)
        dpad = torch.zeros((2 * C,), dtype=torch.float32, device=qkv3.device)
        for b, geom in enumerate((L.GEOM_TSA, L.GEOM_FSA, L.GEOM_SWA)):
            ops.attention_bwd(qkv3[:, b, :], att[:, b, :], lse[b], datt[:, b, :], dqkv3[:, b, :], dsum, B, H, W, C, heads, geom,
                              ws=ws, shift=shift, pad_kv=pk if geom == L.GEOM_SWA else None,
                              dpad=dpad if geom == L.GEOM_SWA else None)
        return (dqkv3, dpad) + (None,) * 9


class AffineActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, scale, shift, r1, r2, epilogue):
        sc = None if scale is None else _contig(scale.detach().float())
        sh = None if shift is None else _contig(shift.detach().float())
        y = ops.affine_act(v, sc, sh, epilogue=epilogue, r1=r1, r2=r2)
        ctx.save_for_backward(v, sc, sh)
        ctx.cfg = (epilogue, r1 is not None, r2 is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        v, sc, sh = ctx.saved_tensors
        epilogue, has_r1, has_r2 = ctx.cfg
        dy = _cl(dy) if dy.dim() == 4 else _contig(dy)
        want_params = sc is not None and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2])
        dv, dsc, dsh = ops.affine_act_bwd(dy, v, sc, sh, epilogue, want_params)
        return dv, dsc, dsh, (dy if has_r1 else None), (dy if has_r2 else None), None


def up_phase_weights(w_oihw: Tensor) -> Tensor:
    """(Cout, Cin, 4, 4) -> (4 phases, Cout, 2, 2, Cin) as consumed by conv kind 2 (see igemm.cu a_offset<KIND_UP>)."""
    ph = []
    for py in (0, 1):
        for px in (0, 1):
            sel = w_oihw[:, :, list(_UP_TAPS[py]), :][:, :, :, list(_UP_TAPS[px])]
            ph.append(sel.permute(0, 2, 3, 1))
    return torch.stack(ph).contiguous()


def up_phase_weights_inverse(dwl: Tensor) -> Tensor:
    """(4, Cout, 2, 2, Cin) -> (Cout, Cin, 4, 4)"""
    _, Cout, _, _, Cin = dwl.shape
    dw = torch.zeros((Cout, Cin, 4, 4), dtype=dwl.dtype, device=dwl.device)
    for py in (0, 1):
        for px in (0, 1):
            blk = dwl[py * 2 + px].permute(0, 3, 1, 2)           # (Cout, Cin, 2, 2)
            for a in (0, 1):
                for b in (0, 1):
                    dw[:, :, _UP_TAPS[py][a], _UP_TAPS[px][b]] = blk[:, :, a, b]
    return dw


def conv_layout(w_oihw: Tensor, kind: str) -> Tensor:
    """kernel-side weight layout of a Cout-first (Cout, Cin, kh, kw) weight"""
    if kind == "stem":
        return w_oihw.contiguous()
    if kind in ("conv3", "down"):
        return w_oihw.permute(0, 2, 3, 1).contiguous()
    return up_phase_weights(w_oihw)


def _conv_any(x: Tensor, wl: Tensor, bias: Tensor, kind: int, out_hw) -> Tensor:
    """plain convolution (no epilogue / statistics) on the best available kernel: tcgen05 implicit GEMM for bf16
    activations with tileable channel counts, CUDA-core implicit GEMM otherwise (the data gradients of ConvFn)"""
    from . import functional as Fn
    if (Fn.USE_TC and x.dtype == torch.bfloat16 and x.shape[1] % 32 == 0 and bias.shape[0] % 16 == 0 and bias.shape[0] <= 256):
        return ops.conv_tc(x, wl.to(torch.bfloat16).contiguous(), bias, kind, out_hw)
    return ops.conv(x, wl, bias, kind, out_hw)


class ConvFn(torch.autograd.Function):
    """w is Cout-first OIHW (a ConvTranspose2d weight is passed already permuted to Cout-first)."""

    @staticmethod
    def forward(ctx, x, w, b, kind, out_hw, dtype, epilogue, want_col_stats):
        w = w.detach().float()
        b = _contig(b.detach().float())
        wl = conv_layout(w, kind)
        stats = torch.zeros((2, b.shape[0]), dtype=torch.float32, device=x.device) if want_col_stats else None
        save_pre = epilogue == L.EPI_GELU and not want_col_stats
        from . import functional as Fn
        if kind == "stem":
            res = ops.stem(x, wl, b, dtype, epilogue=epilogue, col_stats=stats, save_pre=save_pre)
        elif (want_col_stats and epilogue == L.EPI_NONE and Fn.USE_TC and x.dtype == torch.bfloat16 and x.shape[1] % 32 == 0
              and b.shape[0] % 16 == 0 and b.shape[0] <= 256):
            res = ops.conv_tc(x, wl.to(torch.bfloat16).contiguous(), b, _KIND[kind], out_hw, col_stats=stats)   # train-mode BN sums from the epilogue
        else:
            res = ops.conv(x, wl, b, _KIND[kind], out_hw, epilogue=epilogue, col_stats=stats, save_pre=save_pre)
        y, pre = res if save_pre else (res, None)
        ctx.save_for_backward(x, w, y if want_col_stats else pre)
        ctx.cfg = (kind, epilogue, want_col_stats, tuple(wl.shape))
        return (y, stats) if want_col_stats else y

    @staticmethod
    def backward(ctx, dy, dstats=None):
        x, w, pre = ctx.saved_tensors
        kind, epilogue, want_col_stats, wl_shape = ctx.cfg
        g = _cl(dy)
        if want_col_stats:
            gpre = ops.act_bwd(g, pre, 1, _contig(dstats.float())) if dstats is not None else g
        elif epilogue == L.EPI_GELU:
            gpre = ops.act_bwd(g, pre, 0)
        else:
            gpre = g
        if kind == "stem":
            dx, dw, db = ops.stem_bwd(x, _contig(w), gpre, ctx.needs_input_grad[0])
            return dx, dw, db, None, None, None, None, None
        dwl, db = ops.conv_wgrad(x, gpre, _KIND[kind], wl_shape)
        dw = dwl.permute(0, 3, 1, 2) if kind in ("conv3", "down") else up_phase_weights_inverse(dwl)
        dx = None
        if ctx.needs_input_grad[0]:
            zb = torch.zeros((x.shape[1],), dtype=torch.float32, device=x.device)
            hw = tuple(x.shape[2:])
            if kind == "down":      # data gradient of the stride-2 conv = transposed conv (kind 2) with (Cin_d, Cout_d) swapped
                dx = _conv_any(gpre, up_phase_weights(w.permute(1, 0, 2, 3)), zb, 2, hw)
            elif kind == "up":      # data gradient of the transposed conv = stride-2 conv (kind 1)
                dx = _conv_any(gpre, w.permute(1, 2, 3, 0).contiguous(), zb, 1, hw)
            else:                   # 3x3 s1 p1: flipped taps, swapped channels
                dx = _conv_any(gpre, w.flip(2, 3).permute(1, 2, 3, 0).contiguous(), zb, 0, hw)
        return dx, dw, db, None, None, None, None, None


class HeadTailFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v, w3, b3, scale, shift, want_logits):
        w3c, b3c = _contig(w3.detach().float()), _contig(b3.detach().float())
        sc = None if scale is None else _contig(scale.detach().float())
        sh = None if shift is None else _contig(shift.detach().float())
        res = ops.head_tail(v, w3c, b3c, sc, sh, want_logits=want_logits)
        ctx.save_for_backward(v, w3c, b3c, sc, sh)
        ctx.want_logits = want_logits
        return res

    @staticmethod
    def backward(ctx, dmasks, dlogits=None):
        v, w3, b3, sc, sh = ctx.saved_tensors
        dm = None if dmasks is None else _contig(dmasks.float())
        dl = None if dlogits is None else _contig(dlogits.float())
        if dm is None and dl is None:
            return (None,) * 6
        dv, dw3, db3, dsc, dsh = ops.head_tail_bwd(v, sc, sh, w3, b3, dm, dl)
        return dv, dw3, db3, dsc, dsh, None


class BilinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, out_hw):
        ctx.in_hw = tuple(x.shape[2:])
        return ops.bilinear(x, out_hw)

    @staticmethod
    def backward(ctx, dy):
        return ops.bilinear_bwd(_cl(dy), ctx.in_hw), None
