/*
 * tfswa_b200.h - C ABI of libtfswa_b200.so: the TFSWA-UNet hot path on NVIDIA B200 (sm_100a).
 *
 * The reference (chynggi/TFSWA-UNet) has no FFI layer: its hot path sits behind the Python
 * nn.Module surface (src/models/attention.py, blocks.py, tfswa_unet.py) and dispatches ATen ops.
 * Each entry point below replaces the ATen op sequence of one reference call site (cited per
 * function).  The Python mirror of the reference modules (the .py files of tfswa-unet_b200/) binds these
 * with ctypes; INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer owned by the caller (PyTorch's caching allocator);
 *    the library allocates nothing persistent on the device.
 *  - "act" tensors are NHWC ("tokens x channels", token m = (b*H + h)*W + w), element type
 *    given by `dtype` (TFSWA_F32 or TFSWA_BF16).  Weights, biases, statistics are fp32.
 *  - Every call takes the caller's cudaStream_t (as void*), launches asynchronously and never
 *    synchronises.  Re-entrant; no global device state.
 *  - Return 0 on success, negative on error; tfswa_last_error() returns a thread-local message.
 *    Nothing throws or exits across the ABI.  There is NO CPU fallback.
 */
#ifndef TFSWA_B200_H
#define TFSWA_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define TFSWA_F32 0
#define TFSWA_BF16 1

#define TFSWA_OK 0
#define TFSWA_EINVAL (-1)   /* bad argument / unsupported shape */
#define TFSWA_ECUDA (-2)    /* CUDA runtime error at launch */

/* prologue applied to the A operand of tfswa_linear_* while it is staged on chip */
#define TFSWA_PRO_NONE 0
#define TFSWA_PRO_LNHAT 1    /* (x - mean_row) * rstd_row from row_stats; LN affine is folded into W/bias by the host */
#define TFSWA_PRO_GELU 2     /* exact erf GELU */
#define TFSWA_PRO_AFFINE 4   /* x*in_scale[k] + in_shift[k] (train-mode BatchNorm apply); may be OR-ed with the others, applied first */
/* epilogue */
#define TFSWA_EPI_NONE 0
#define TFSWA_EPI_GELU 1
#define TFSWA_EPI_MUL_DGELU 2   /* tfswa_linear_tc_fwd only: y = (acc + bias) * gelu'(r1) - the backward of GELU fused into the data-gradient GEMM
                                 * of the layer above (r1 = the saved pre-activation, same shape as y; r2 must be NULL) */

/* attention geometries */
#define TFSWA_GEOM_TSA 0     /* sequences along H (reference dim 2), one per (b, w)   attention.py:143 */
#define TFSWA_GEOM_FSA 1     /* sequences along W (reference dim 3), one per (b, h)   attention.py:217 */
#define TFSWA_GEOM_SWA 2     /* ws x ws windows of the zero-padded, cyclically shifted map   attention.py:358-375 */

const char* tfswa_last_error(void);
const char* tfswa_version(void);
/* 1 if a kernel image for the current device exists in this library (sm_100a only). */
int tfswa_device_supported(void);

/* ---- token linear: Y = epi(pro(X) W^T + bias) (+R1) (+R2) ---------------------------------
 * Replaces nn.Linear / 1x1 nn.Conv2d call sites: attention.py:70 (qkv), :86 (proj), :121-128 (MLP),
 * blocks.py:53-56 (input_proj), :85-89 (fusion, K = 3C over the never-materialised concat).
 * `batch` independent problems (the three branches) are addressed by element strides *_bs. */
typedef struct {
  const void* x;  int64_t ldx;  int64_t x_bs;     /* (M,K) act, row stride ldx */
  const float* w;               int64_t w_bs;     /* (N,K) fp32 row-major (nn.Linear layout) */
  const float* bias;            int64_t bias_bs;  /* (N) or NULL */
  const float* row_stats;       int64_t rs_bs;    /* (M,2) {mean,rstd} for TFSWA_PRO_LNHAT */
  const float* in_scale; const float* in_shift;   /* (K) for TFSWA_PRO_AFFINE */
  const void* r1; int64_t ldr1; int64_t r1_bs;    /* residual added after the epilogue, or NULL */
  const void* r2; int64_t ldr2; int64_t r2_bs;
  void* y;        int64_t ldy;  int64_t y_bs;     /* (M,N) act */
  void* pre;      int64_t ldpre; int64_t pre_bs;  /* optional copy of the pre-epilogue value (saved for backward) */
  float* col_stats;                                /* optional (2,N): atomically accumulates sum / sum-of-squares of the pre-epilogue value (train-mode BN) */
  int64_t M; int32_t N; int32_t K;
  int32_t prologue; int32_t epilogue; int32_t batch; int32_t dtype;
} tfswa_linear_args;
int tfswa_linear_fwd(const tfswa_linear_args* a, void* stream);

/* Same contract on the tcgen05 tensor cores (bf16 activations only; TMA-fed, TMEM accumulators).  w_bf16: (batch,N,K)
 * bf16 copy of the weights.  TFSWA_PRO_LNHAT is applied algebraically in the epilogue,
 *   LN_hat(x) W^T = rstd*(x W^T - mean*wsum),  wsum[n] = sum_k w_bf16[n][k]  ((batch,N) fp32),
 * so the MMA reads the raw activations.  Supports prologue NONE/LNHAT, epilogue NONE/GELU, r1, r2; returns
 * TFSWA_EINVAL for pre / col_stats / AFFINE / GELU-prologue requests (use tfswa_linear_fwd for those). */
int tfswa_linear_tc_fwd(const tfswa_linear_args* a, const void* w_bf16, const float* wsum, void* stream);

/* Fused tail of a pre-LN transformer branch (attention.py:86 + :121-128 with the residuals of :146,:159 / :220,:233 /
 * :378,:387), bf16, tcgen05:    y = att Wp^T + bp + res;   out = y + W2 gelu(W1' LN_hat(y) + b1') + b2
 * with the LayerNorm affine folded into W1'/b1' by the host.  One launch replaces tfswa_linear_tc_fwd (proj) +
 * tfswa_row_stats + tfswa_linear_tc_fwd (fc1, GELU) + tfswa_linear_tc_fwd (fc2); y, its statistics and the 4C-wide
 * hidden activations never leave the SM.  C in {32, 64, 128} (at 128 the weights stream through a TMA ring instead of
 * staying resident), hidden == 4*C; other widths return TFSWA_EINVAL (callers use
 * the unfused sequence).  `batch` independent branches: att/out (M, batch, C) addressed by element strides; res is
 * (M, C), shared by all branches when res_bs == 0.  Weights bf16 row-major (nn.Linear layout), biases fp32. */
typedef struct {
  const void* att; int64_t lda; int64_t att_bs;
  const void* res; int64_t ldr; int64_t res_bs;
  const void* wp;  const void* w1; const void* w2;      /* (batch,C,C), (batch,hidden,C), (batch,C,hidden) bf16 */
  const float* bp; const float* b1; const float* b2;    /* (batch,C), (batch,hidden), (batch,C) */
  void* out; int64_t ldo; int64_t out_bs;
  int64_t M; int32_t C; int32_t hidden; int32_t batch; float eps;
} tfswa_tail_args;
int tfswa_branch_tail_tc_fwd(const tfswa_tail_args* a, void* stream);

/* Fused head of a TFSWABlock (blocks.py:53-56,115 + attention.py:70 of the three branches), bf16, tcgen05, inference:
 *   x1 = x Wi^T + bi  (input_proj with eval BatchNorm folded in);   qkv = LN_hat(x1) Wq^T + bq  (LN affine folded in)
 * One launch replaces tfswa_linear_tc_fwd (input_proj) + tfswa_row_stats + tfswa_linear_tc_fwd (qkv).  C in {32, 64};
 * wi (C,C), wq (9C,C) bf16 row-major; bi (C), bq (9C) fp32; x (M,C), x1 (M,C), qkv (M,9C) bf16 with row strides. */
typedef struct {
  const void* x;  int64_t ldx;
  const void* wi; const void* wq;
  const float* bi; const float* bq;
  void* x1;  int64_t ld1;
  void* qkv; int64_t ldq;
  int64_t M; int32_t C; float eps;
} tfswa_head_args;
int tfswa_block_head_tc_fwd(const tfswa_head_args* a, void* stream);

/* per-row LayerNorm statistics {mean, rstd} over K channels, eps 1e-5 (F.layer_norm, attention.py:146,159) */
int tfswa_row_stats(const void* x, int64_t ldx, int64_t x_bs, float* stats, int64_t st_bs,
                    int64_t M, int32_t K, int32_t batch, int32_t dtype, void* stream);

/* ---- attention core: O = softmax(Q K^T / sqrt(d)) V per head, never materialising the scores --
 * Replaces attention.py:70-85 together with the token regrouping of :143/:162 (TSA), :217/:236 (FSA)
 * and :358-375/:390-401 (SW-MSA pad + roll + window partition/reverse).
 * qkv: (M, ldq) act with q|k|v each C wide starting at column 0|C|2C of the given pointer.
 * out: (M, ldo) act, C wide.  lse (optional): (M, heads) fp32 log2-domain logsumexp for backward.
 * SWA: pad_kv = (2C) fp32 k|v of a zero-padded token (= folded qkv bias); rel_bias optional
 * (heads, ws*ws, ws*ws) fp32; use_shift_mask adds the Swin {0,-100} mask (both OFF = reference).
 * flags: bit 0 (TFSWA_ATTN_FORCE_EXACT) makes tfswa_attn_tc_fwd skip the row-maximum bound and run its exact two-pass
 * softmax (same result up to rounding; exists so tests can exercise the fallback path). */
#define TFSWA_ATTN_FORCE_EXACT 1
typedef struct {
  const void* qkv; int64_t ldq;
  void* out;       int64_t ldo;
  float* lse;
  const float* pad_kv;
  const float* rel_bias;
  int32_t B, H, W, C, heads;
  int32_t geom, ws, shift, use_shift_mask, dtype;
  int32_t flags;
} tfswa_attn_args;
int tfswa_attn_fwd(const tfswa_attn_args* a, void* stream);

/* Same contract for the axial geometries (TSA / FSA) at head_dim 4, 8 and 16, bf16, on the tcgen05 tensor cores:
 * K and V tiles arrive by TMA (the permutes of attention.py:143,217 are the tensor map's strides), QK^T and PV are
 * tcgen05.mma with TMEM accumulators, heads share the K slab through masked copies of Q, the softmax denominator is
 * accumulated by the tensor core through a ones tile appended to V. */
/* scratch: caller-owned device buffer of tfswa_attn_tc_scratch_bytes(a) bytes (per-sequence k extrema for the softmax
 * shift bound); contents are dead after the call. */
int64_t tfswa_attn_tc_scratch_bytes(const tfswa_attn_args* a);
int tfswa_attn_tc_fwd(const tfswa_attn_args* a, void* scratch, int64_t scratch_bytes, void* stream);

/* SW-MSA (8x8 windows, attention.py:347-403), bf16, default-off mask / relative-bias features excluded (they are served
 * by tfswa_attn_fwd).  head_dim 4 and 8: the windows whose 64 tokens exist and do not wrap around the rolled frame run
 * on tcgen05 - one TMA box (8 channels, 8, 8, 1) of the q|k|v token matrix is a window's K (K-major) and V (MN-major)
 * operand, two heads of a window fill the 128-row tile through a masked copy of Q, exact row maxima (one thread per
 * row), P in TMEM, the denominator through a ones tile; the bottom / right fringe (pad / wrap-around) and head_dim
 * 16 / 32 run on warp-level MMAs (m16n8k16) with scores and probabilities in registers.
 * tfswa_attn_win_tc_interior_launches: diagnostics - how many tcgen05 interior launches this process has made. */
int tfswa_attn_win_tc_fwd(const tfswa_attn_args* a, void* stream);
long long tfswa_attn_win_tc_interior_launches(void);

/* ---- convolutions (implicit GEMM over NHWC) ---------------------------------------------------
 * kind: 0 = 3x3 s1 p1 (output_head.0, tfswa_unet.py:140), 1 = 4x4 s2 p1 (DownsampleBlock, blocks.py:157),
 *       2 = transposed 4x4 s2 p1 (UpsampleBlock, blocks.py:172).
 * w: fp32, re-laid-out by the host to (Cout, kh, kw, Cin) for kinds 0/1 and to (4 phases, Cout, 2, 2, Cin)
 * for kind 2 (phase = (oy&1)*2 + (ox&1)).  y = epi(conv + bias); pre/col_stats as in tfswa_linear_args. */
typedef struct {
  const void* x; void* y; void* pre;
  const float* w; const float* bias; float* col_stats;
  int32_t B, Hin, Win, Cin, Hout, Wout, Cout;
  int32_t kind, epilogue, dtype;
} tfswa_conv_args;
int tfswa_conv_fwd(const tfswa_conv_args* a, void* stream);

/* Same contract on the tcgen05 tensor cores (bf16, eval-mode epilogues only: no pre / col_stats): the A operand is
 * gathered by producer warps into the UMMA swizzle pattern, the weight tiles come by TMA.  w_bf16: the weights of
 * tfswa_conv_args.w cast to bf16 (same layout). */
int tfswa_conv_tc_fwd(const tfswa_conv_args* a, const void* w_bf16, void* stream);

/* stem: Conv2d(Cin,Cout,7,p=3) on the NCHW fp32 network input -> NHWC act (tfswa_unet.py:58-62).
 * w: (Cout, Cin, 7, 7) fp32 in nn.Conv2d layout. */
int tfswa_stem_fwd(const float* x_nchw, const float* w, const float* bias, void* y, void* pre, float* col_stats,
                   int32_t B, int32_t Cin, int32_t H, int32_t W, int32_t Cout, int32_t epilogue, int32_t dtype, void* stream);

/* head tail: u = GELU(v*scale + shift); logits = W3 u + b3; masks = sigmoid(logits) -> NCHW fp32
 * (output_head.1-4, tfswa_unet.py:141-144).  scale/shift NULL = identity (eval: BN folded upstream). */
int tfswa_head_tail_fwd(const void* v, const float* scale, const float* shift, const float* w3, const float* b3,
                        float* masks_nchw, float* logits_nchw, int32_t B, int32_t H, int32_t W, int32_t C, int32_t Cout,
                        int32_t dtype, void* stream);

/* ---- BatchNorm helpers (train mode) ------------------------------------------------------------
 * finalize: from col_stats (2,C) over `count` values: scale = gamma*rsqrt(var+eps), shift = beta - mean*scale,
 * and the running-stat update (momentum, unbiased variance) in place.  save_mean_rstd (2,C) for backward. */
int tfswa_bn_finalize(const float* col_stats, int64_t count, const float* gamma, const float* beta,
                      float* running_mean, float* running_var, float momentum, float eps,
                      float* scale, float* shift, float* save_mean_rstd, int32_t C, void* stream);
/* y = act(v*scale[c] + shift[c]) (+r1) (+r2): BN apply + GELU + residuals over (M,C) act */
int tfswa_affine_act(const void* v, const float* scale, const float* shift, const void* r1, const void* r2, void* y,
                     int64_t M, int32_t C, int32_t epilogue, int32_t dtype, void* stream);

/* bilinear resize, align_corners=False, NHWC act (F.interpolate at tfswa_unet.py:210-216) */
int tfswa_bilinear_fwd(const void* x, void* y, int32_t B, int32_t Hin, int32_t Win, int32_t Hout, int32_t Wout,
                       int32_t C, int32_t dtype, void* stream);


/* =====================================================================================================
 * Backward (autograd of the call sites above).  Gradients w.r.t. weights/biases/statistics are
 * fp32 and ACCUMULATED with atomics into caller-zeroed buffers; activation gradients are written in the
 * activation dtype.  The data gradient of a linear / conv is itself a forward call with transposed weights
 * (tfswa_linear_fwd / tfswa_conv_fwd: the stride-2 conv's data gradient is kind 2 with Hout in {2Hin, 2Hin+1},
 * the transposed conv's is kind 1, the 3x3 conv's is kind 0 with flipped taps), so only the pieces below are new.
 * ===================================================================================================== */

/* dW[b][n][k] += sum_m G[m,b,n] * pro(X)[m,b,k];  dbias[b][n] += sum_m G[m,b,n].  `a` describes X exactly as in the
 * forward call (x, ldx, x_bs, prologue, row_stats, in_scale/in_shift, M, N, K, batch, dtype); G = gradient of the
 * pre-epilogue output, (M, ldg) with batch stride g_bs.  dbias may be NULL. */
int tfswa_linear_wgrad(const tfswa_linear_args* a, const void* g, int64_t ldg, int64_t g_bs, float* dw, float* dbias, void* stream);
/* same for the convolutions of tfswa_conv_fwd; dw has the layout of tfswa_conv_args.w; g = (B,Hout,Wout,Cout) */
int tfswa_conv_wgrad(const tfswa_conv_args* a, const void* g, float* dw, float* dbias, void* stream);

/* mode 0: out = g * gelu'(pre)   (gradient through a GELU epilogue)
 * mode 1: out = g + ds[c] + 2*pre*ds[C+c]   (adds what reaches `pre` through the train-mode BN column sums) */
int tfswa_act_bwd(const void* g, const void* pre, const float* ds, void* out, int64_t M, int32_t C, int32_t mode, int32_t dtype, void* stream);
/* backward of tfswa_affine_act: dv, and (optional) dscale/dshift (C) accumulated */
int tfswa_affine_act_bwd(const void* dy, const void* v, const float* scale, const float* shift, void* dv, float* dscale,
                         float* dshift, int64_t M, int32_t C, int32_t epilogue, int32_t dtype, void* stream);
/* LayerNorm input gradient for TFSWA_PRO_LNHAT: dx = rstd*(da - mean_k(da) - xhat*mean_k(da*xhat)) per row */
int tfswa_lnhat_bwd(const void* da, int64_t ldd, int64_t d_bs, const void* x, int64_t ldx, int64_t x_bs, const float* stats,
                    int64_t st_bs, void* dx, int64_t ldo, int64_t o_bs, int64_t M, int32_t K, int32_t batch, int32_t dtype, void* stream);
/* out[m][:] = sum_b g[m][b][:]  (gradient of a residual broadcast over the branch dimension), dense (M,nb,N) */
int tfswa_sum_batch(const void* g, void* out, int64_t M, int32_t nb, int32_t N, int32_t dtype, void* stream);
int tfswa_bilinear_bwd(const void* dy, void* dx, int32_t B, int32_t Hin, int32_t Win, int32_t Hout, int32_t Wout, int32_t C,
                       int32_t dtype, void* stream);

/* attention backward: `a` as in the forward call with a->out = O and a->lse = the saved log2-domain logsumexp.
 * dout: (M, ldo) gradient of O; dqkv: (M, ldq) receives dq|dk|dv; dsum: (M, heads) fp32 scratch (dO.O);
 * dpad: (2C) fp32, accumulates the gradient of pad_kv (may be NULL when nothing is padded). */
int tfswa_attn_bwd(const tfswa_attn_args* a, const void* dout, void* dqkv, float* dsum, float* dpad, void* stream);

/* stem backward: g = gradient of the conv output (NHWC act); dw (Cout,Cin,7,7) and dbias accumulated;
 * dx_nchw (optional) = gradient of the fp32 NCHW network input */
int tfswa_stem_bwd(const float* x_nchw, const float* w, const void* g, float* dx_nchw, float* dw, float* dbias, int32_t B,
                   int32_t Cin, int32_t H, int32_t W, int32_t Cout, int32_t dtype, void* stream);
/* head tail backward: from dmasks (and/or dlogits), NCHW fp32 -> dv (act), dw3 (Cout,C), db3 (Cout), dscale/dshift (C) accumulated */
int tfswa_head_tail_bwd(const void* v, const float* scale, const float* shift, const float* w3, const float* b3,
                        const float* dmasks_nchw, const float* dlogits_nchw, void* dv, float* dw3, float* db3, float* dscale,
                        float* dshift, int32_t B, int32_t H, int32_t W, int32_t C, int32_t Cout, int32_t dtype, void* stream);

/* =====================================================================================================
 * The spectrogram steps either side of the model in overlap-add separation (row f3 of SURVEY 8f).  STFT / ISTFT stay
 * batched cuFFT (torch.stft / torch.istft); these entry points replace the eager element-wise chains between them
 * and the model.  All buffers are dense fp32 / complex64 (interleaved re, im).
 * ===================================================================================================== */

/* spec (B, F, T) complex64 -> x (B, 2, F, T) fp32 = [real | imag] planes (stft_processor.py:186-204 to_model_input);
 * normalize != 0: instance normalisation over time per (b, channel, f) row, (v - mean) / (std + eps) with the unbiased
 * std (SpectrogramNormalizer.forward, stft_processor.py:283-297), and stats (B, 2, F, 2) receives {mean, std + eps}. */
int tfswa_spec_pack_norm(const void* spec_c64, float* x, float* stats, int32_t B, int32_t F, int32_t T, float eps,
                         int32_t normalize, void* stream);
/* out (B, S, F, T) complex64 = spec[b, f, t] * w,  w = masks[b, s, f, t]  (normalize == 0)  or
 * masks * std[b, s, f] + mean[b, s, f]: the reference "denormalises" the masks with the statistics of INPUT channel s
 * (inference.py:132-133, stft_processor.py:299-312), so S <= 2 there.  Replaces inference.py:139-145. */
int tfswa_spec_mask_apply(const float* masks, const void* spec_c64, const float* stats, void* out_c64, int32_t B, int32_t S,
                          int32_t F, int32_t T, int32_t normalize, void* stream);
/* Hann-weighted overlap-add of nseg reconstructed segments (inference.py:209-216): wav (nseg, S, L) fp32, starts (nseg)
 * int64 device array of output offsets in ascending order (first_start / last_start: its first and last entry, on the
 * host), win (>= min(seg_len, L)) fp32, acc (S + 1, total) fp32: rows 0..S-1 += wav * win, row S += win, over the first
 * min(seg_len, total - start, L) samples of each segment.  Per output sample the segments are added in ascending
 * order with separately rounded products - the arithmetic of the reference's sequential loop; no atomics. */
int tfswa_ola_add(const float* wav, const int64_t* starts, int64_t first_start, int64_t last_start, const float* win, float* acc,
                  int32_t nseg, int32_t S, int64_t L, int64_t seg_len, int64_t total, void* stream);

/* One resolution of MultiResolutionSTFTLoss (losses.py:125-141, 171-183; row f4 of SURVEY 8f) on the complex STFTs of the
 * predicted and the target audio (n complex64 elements each, any common dense layout):
 *   *loss += w_mag * mean| |P| - |T| | + w_log * mean| log(|P| + eps) - log(|T| + eps) |     (double, caller-zeroed)
 * grad_c64 (optional, n complex64): d(that value)/dP in PyTorch's convention for a real loss of a complex tensor
 * (dL/dRe + i dL/dIm) - the backward of abs / log / l1_loss of the eager chain, from the same pass. */
int tfswa_mrstft_mag_loss(const void* pred_c64, const void* target_c64, int64_t n, float w_mag, float w_log, float eps,
                          double* loss, void* grad_c64, void* stream);

/* =====================================================================================================
 * Optimiser step over a flat fp32 parameter arena (row f1 of SURVEY 8f).  Replaces
 * torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm) + optimizer.step() of src/training/trainer.py:214-219
 * for the AdamW of scripts/train.py:251-255 (one parameter group, decoupled weight decay on every parameter).
 * All buffers hold n fp32 elements (n % 4 == 0, 16-byte aligned); nothing synchronises with the host.
 * ===================================================================================================== */

/* *sumsq = sum_i g[i]^2 (zeroed by the call; double so that the order of the per-CTA atomics is immaterial) */
int tfswa_grad_sumsq(const float* g, int64_t n, double* sumsq, void* stream);
/* g_eff = g * grad_scale * min(1, max_norm / (sqrt(*sumsq)*grad_scale + 1e-6))   (max_norm <= 0: no clipping;
 *         grad_scale = 1/world when g holds the all-reduced SUM of the per-rank gradients)
 * p *= 1 - lr*weight_decay;  m = beta1 m + (1-beta1) g_eff;  v = beta2 v + (1-beta2) g_eff^2;
 * p -= lr/(1-beta1^step) * m / (sqrt(v)/sqrt(1-beta2^step) + eps)                 (torch.optim.AdamW, step >= 1)
 * norm_out (optional, device) receives the unclipped total norm; a non-finite norm skips the update.
 * skipped (optional, device int64, caller-zeroed once): number of skipped updates so far.  A skipped update increments
 * it and the bias corrections use step - *skipped, so a skipped step does not advance the optimiser state (the
 * semantics of GradScaler.step around torch.optim.AdamW, trainer.py:215-216). */
int tfswa_adamw_clip_step(float* p, const float* g, float* m, float* v, int64_t n, const double* sumsq, float* norm_out,
                          float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay,
                          int64_t step, long long* skipped, void* stream);

#ifdef __cplusplus
}
#endif
#endif
