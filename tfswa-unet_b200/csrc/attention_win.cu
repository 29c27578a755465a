// SW-MSA window attention (attention.py:347-403) with the scores, probabilities and outputs of a window living in
// registers: warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate) in the FlashAttention-2 register layout.
//
// Since round 2 this kernel serves what the tcgen05 + TMA kernel (tc_attn_win.cu) does not take: at head_dim 4 / 8 the
// FRINGE windows (those with zero-padded tokens or tokens that wrap around the rolled frame: the last two window rows /
// columns at shift 4, 4.6 % of the stage-1 windows) - p.win_edge enumerates only those - and every window at head_dim
// 16 / 32 (stages 3 / 4: 2 % of the C3 step; a 128-row tile of ONE 16-channel head would need the keys of two windows
// side by side in TMEM, i.e. one CTA per SM).  TFSWA_WIN_KERNEL=mma routes everything here (A/B, tests).
// A warp owns (head, 16 query rows): S = Q K^T is 8 n-tiles of accumulators (32 scores per thread), the row
// maximum is exact (the whole key range is in registers: no online rescaling, no bound), P is re-packed in place as
// the A operand of the PV MMA, and the row sum comes from one more MMA against a ones operand, so the bf16-rounded P
// feeds numerator and denominator alike.  No shared-memory round trip for S or P, no barriers after staging.
//
// One CTA = one window x one slab of <= 64 channels.  The window's q|k|v rows are gathered once through the token
// map (roll + pad are index arithmetic; zero-padded tokens are real keys whose k|v equal the folded qkv bias
// `pad_kv`) into padded shared-memory rows (pitch = slab + 8 elements: conflict-free fragment loads and ldmatrix).
#include "attn_common.cuh"
#include <stdlib.h>

namespace tfswa {

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack2_bf16(float lo, float hi) {
  uint32_t y;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo));
  return y;
}
// 2^x for x <= 0 on the FMA pipe (same polynomial as the axial tensor-core kernel: 7.5e-5 relative)
__device__ __forceinline__ float ex2_poly_w(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float r = x - (t - 12582912.0f);
  float p = fmaf(0.0551716685f, r, 0.2426111251f);
  p = fmaf(p, r, 0.6932609677f);
  p = fmaf(p, r, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

constexpr int WIN_TOK = 64;            // tokens per window (ws = 8)
constexpr int WIN_THREADS = 256;
#ifndef TFSWA_WIN_POLY_EVERY
#define TFSWA_WIN_POLY_EVERY 0
#endif
constexpr int WIN_POLY_EVERY = TFSWA_WIN_POLY_EVERY;   // every n-th exponential on the FMA pipe instead of MUFU (0 = none: this kernel is issue-bound)

// D = head_dim, CS = channels per CTA slab (min(C, 64))
template <int D, int CS>
__global__ void __launch_bounds__(WIN_THREADS) attn_win_mma_kernel(const AttnParams p) {
  constexpr int PITCH = CS + 8;                       // elements per shared-memory row
  constexpr int HS = CS / D;                          // heads per slab
  constexpr int ITEMS = HS * 4;                       // (head, 16-row tile) work items
  constexpr int IPW = ITEMS / 8;                      // items per warp
  constexpr int KS = (D + 15) / 16;                   // k-steps of the S MMA
  constexpr int DN = (D + 7) / 8;                     // 8-dim n-tiles of the PV MMA
  static_assert(ITEMS % 8 == 0 && IPW >= 1, "slab must give every warp at least one item");
  __shared__ __align__(16) bf16 Qs[WIN_TOK][PITCH];
  __shared__ __align__(16) bf16 Ks[WIN_TOK][PITCH];
  __shared__ __align__(16) bf16 Vs[WIN_TOK][PITCH];
  __shared__ int64_t s_tok[WIN_TOK];                  // token index, or -1 for a zero-padded window position

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  int win = blockIdx.x;
  const int ch0 = blockIdx.y * CS;
  if (p.win_edge) {
    // fringe windows only (the interior ran on tc_attn_win.cu): per image first the window rows below the interior, then
    // the columns right of it
    const int nb = (p.nWh - p.nWh_int) * p.nWw, nr = p.nWw - p.nWw_int;
    const int per_img = nb + p.nWh_int * nr;
    const int b = win / per_img, e = win - b * per_img;
    int wh, ww;
    if (e < nb) { wh = p.nWh_int + e / p.nWw; ww = e - (e / p.nWw) * p.nWw; }
    else { const int e2 = e - nb; wh = e2 / nr; ww = p.nWw_int + (e2 - wh * nr); }
    win = (b * p.nWh + wh) * p.nWw + ww;
  }
  const bf16* qkv = (const bf16*)p.qkv;

  if (tid < WIN_TOK) {
    bool valid;
    const int64_t tok = token_of<true>(p, win, tid, valid);
    s_tok[tid] = valid ? tok : -1;
  }
  __syncthreads();
  // ---- gather q|k|v rows of the window: 16-byte chunks, consecutive threads walk one token's slab ----
  constexpr int CPT = CS / 8;                         // chunks per token and part
  for (int v = tid; v < WIN_TOK * 3 * CPT; v += WIN_THREADS) {
    const int j = v / (3 * CPT), rem = v - j * 3 * CPT, part = rem / CPT, chunk = rem - part * CPT;
    const int64_t tok = s_tok[j];
    uint4 val;
    if (tok >= 0) {
      val = *reinterpret_cast<const uint4*>(qkv + tok * p.ldq + part * p.C + ch0 + chunk * 8);
    } else if (part == 0) {
      val = make_uint4(0, 0, 0, 0);                   // query of a pad position: its output row is dropped
    } else {
      const float* pk = p.pad_kv + (part - 1) * p.C + ch0 + chunk * 8;
      val.x = pack2_bf16(pk[0], pk[1]); val.y = pack2_bf16(pk[2], pk[3]);
      val.z = pack2_bf16(pk[4], pk[5]); val.w = pack2_bf16(pk[6], pk[7]);
    }
    bf16* dst = part == 0 ? &Qs[j][chunk * 8] : (part == 1 ? &Ks[j][chunk * 8] : &Vs[j][chunk * 8]);
    *reinterpret_cast<uint4*>(dst) = val;
  }
  __syncthreads();

  const float c = p.qscale;                           // head_dim^-0.5 * log2(e)
#pragma unroll 1
  for (int it = 0; it < IPW; ++it) {
    const int item = warp * IPW + it;
    const int hl = item >> 2, mt = item & 3;          // head within the slab, 16-row tile
    const int r0 = mt * 16 + g, r1 = r0 + 8;
    const int cbase = hl * D;

    // ---- S = Q K^T ----
    uint32_t qa[KS][4];
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int dk = D - ks * 16;                     // dims left in this k-step (>= 16 means a full step)
      const int col = cbase + ks * 16 + 2 * t;
      const bool lo = 2 * t < dk, hi = 2 * t + 8 < dk;
      qa[ks][0] = lo ? *reinterpret_cast<const uint32_t*>(&Qs[r0][col]) : 0u;
      qa[ks][1] = lo ? *reinterpret_cast<const uint32_t*>(&Qs[r1][col]) : 0u;
      qa[ks][2] = hi ? *reinterpret_cast<const uint32_t*>(&Qs[r0][col + 8]) : 0u;
      qa[ks][3] = hi ? *reinterpret_cast<const uint32_t*>(&Qs[r1][col + 8]) : 0u;
    }
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int dk = D - ks * 16;
        const int col = cbase + ks * 16 + 2 * t;
        const uint32_t b0 = 2 * t < dk ? *reinterpret_cast<const uint32_t*>(&Ks[nt * 8 + g][col]) : 0u;
        const uint32_t b1 = 2 * t + 8 < dk ? *reinterpret_cast<const uint32_t*>(&Ks[nt * 8 + g][col + 8]) : 0u;
        mma_bf16_16816(s[nt], qa[ks], b0, b1);
      }
    }
    // ---- exact row maxima (rows r0: elements 0,1; r1: elements 2,3), reduced over the 4 threads of a quad ----
    float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
      m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float mc0 = m0 * c, mc1 = m1 * c;

    // ---- P = exp2(S c - m c) packed as the A operand of the PV MMA; O += P V; l += P 1 ----
    float o[DN][4], l[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
    constexpr uint32_t ONES = 0x3F803F80u;            // bf16x2 {1, 1}
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {                  // 16 keys per step = score n-tiles 2kk, 2kk+1
      uint32_t pa[4];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int nt = 2 * kk + half;
        float e[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float x = fmaf(s[nt][i], c, i < 2 ? -mc0 : -mc1);
          e[i] = (WIN_POLY_EVERY > 0 && ((nt * 4 + i) % WIN_POLY_EVERY) == WIN_POLY_EVERY - 1) ? ex2_poly_w(x) : fast_exp2(x);
        }
        pa[half * 2 + 0] = pack2_bf16(e[0], e[1]);    // row r0, keys 16kk + 8half + 2t, +1
        pa[half * 2 + 1] = pack2_bf16(e[2], e[3]);    // row r1
      }
      // V^T fragments: ldmatrix.trans of two 8-key x 8-dim blocks (keys 16kk..+7 by lanes 0-7, 16kk+8..+15 by lanes 8-15)
#pragma unroll
      for (int dn = 0; dn < DN; ++dn) {
        const int vcol = D >= 8 ? cbase + dn * 8 : (cbase & ~7);       // d = 4: the aligned 8-dim block holding this head
        uint32_t b0, b1;
        ldmatrix_x2_trans(b0, b1, &Vs[kk * 16 + (lane & 15)][vcol]);
        mma_bf16_16816(o[dn], pa, b0, b1);
      }
      mma_bf16_16816(l, pa, ONES, ONES);
    }
    // ---- O / l -> out (rows of pad positions are dropped); every column of `l` holds the row sum ----
    const float inv0 = 1.0f / l[0], inv1 = 1.0f / l[2];
    const int64_t tok0 = s_tok[r0], tok1 = s_tok[r1];
    const int head = blockIdx.y * HS + hl;
    bf16* out = (bf16*)p.out;
    if (D >= 8) {
#pragma unroll
      for (int dn = 0; dn < DN; ++dn) {
        const int col = ch0 + cbase + dn * 8 + 2 * t;
        if (tok0 >= 0) *reinterpret_cast<uint32_t*>(out + tok0 * p.ldo + col) = pack2_bf16(o[dn][0] * inv0, o[dn][1] * inv0);
        if (tok1 >= 0) *reinterpret_cast<uint32_t*>(out + tok1 * p.ldo + col) = pack2_bf16(o[dn][2] * inv1, o[dn][3] * inv1);
      }
    } else {                                          // d = 4: my head's dims are columns (cbase & 7) .. +3 of the 8-dim block
      const int first = cbase & 7;                    // 0 or 4
      if (2 * t >= first && 2 * t < first + 4) {
        const int col = ch0 + (cbase & ~7) + 2 * t;
        if (tok0 >= 0) *reinterpret_cast<uint32_t*>(out + tok0 * p.ldo + col) = pack2_bf16(o[0][0] * inv0, o[0][1] * inv0);
        if (tok1 >= 0) *reinterpret_cast<uint32_t*>(out + tok1 * p.ldo + col) = pack2_bf16(o[0][2] * inv1, o[0][3] * inv1);
      }
    }
    if (p.lse && t == 0) {
      if (tok0 >= 0) p.lse[tok0 * p.heads + head] = mc0 + log2f(l[0]);
      if (tok1 >= 0) p.lse[tok1 * p.heads + head] = mc1 + log2f(l[2]);
    }
  }
}

}  // namespace tfswa

using namespace tfswa;

extern "C" int tfswa_attn_win_tc_fwd(const tfswa_attn_args* a, void* stream) {
  TFSWA_REQUIRE(a && a->qkv && a->out, "attn_win_tc: null pointer");
  TFSWA_REQUIRE(a->dtype == TFSWA_BF16, "attn_win_tc: bf16 activations only");
  TFSWA_REQUIRE(a->geom == TFSWA_GEOM_SWA && a->ws == 8, "attn_win_tc: 8x8 windows only (axial geometries use tfswa_attn_tc_fwd)");
  TFSWA_REQUIRE(!a->rel_bias && !(a->use_shift_mask && a->shift > 0), "attn_win_tc: mask / bias features run on tfswa_attn_fwd");
  TFSWA_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0 && a->heads > 0 && a->C % a->heads == 0, "attn_win_tc: bad shape");
  TFSWA_REQUIRE(a->shift >= 0 && a->shift < a->ws, "attn_win_tc: bad shift %d", a->shift);
  const int D = a->C / a->heads;
  TFSWA_REQUIRE(D == 4 || D == 8 || D == 16 || D == 32, "attn_win_tc: head_dim %d not in {4,8,16,32}", D);
  TFSWA_REQUIRE(a->C % 32 == 0 && (a->C < 64 || a->C % 64 == 0) && (D < 32 || a->C >= 64), "attn_win_tc: C=%d unsupported", a->C);
  TFSWA_REQUIRE(a->ldq % 8 == 0 && a->ldo % 2 == 0 && (((uintptr_t)a->qkv) & 15) == 0 && (((uintptr_t)a->out) & 3) == 0, "attn_win_tc: alignment");
  AttnParams p = {};
  p.qkv = a->qkv; p.ldq = a->ldq; p.out = a->out; p.ldo = a->ldo; p.lse = a->lse; p.pad_kv = a->pad_kv;
  p.B = a->B; p.H = a->H; p.W = a->W; p.C = a->C; p.heads = a->heads; p.geom = a->geom; p.ws = a->ws; p.shift = a->shift;
  p.qscale = (float)(1.4426950408889634 / sqrt((double)D));
  attn_fill_geometry(p);
  TFSWA_REQUIRE((p.Hp == a->H && p.Wp == a->W) || a->pad_kv, "attn_win_tc: padded windows need pad_kv");
  const int CS = p.C >= 64 ? 64 : 32;
  cudaStream_t st = (cudaStream_t)stream;
  int64_t n_win = (int64_t)p.B * p.nWh * p.nWw;
  // head_dim 4 / 8: the windows whose 64 tokens exist and do not wrap around the rolled frame run on tcgen05 with TMA-fed
  // operands (tc_attn_win.cu); this kernel keeps the bottom / right fringe.  TFSWA_WIN_KERNEL=mma: everything here (A/B, tests)
  static int force_mma = -1;
  if (force_mma < 0) { const char* e = getenv("TFSWA_WIN_KERNEL"); force_mma = (e && e[0] == 'm') ? 1 : 0; }
  if (!force_mma && (D == 4 || D == 8)) {
    const int nWh_int = (p.H - p.shift) / p.ws, nWw_int = (p.W - p.shift) / p.ws;
    const int rc = attn_win_tc_bf16(p, nWh_int, nWw_int, st);
    if (rc < 0 || rc > 1) return rc;
    if (rc == 0) {
      p.win_edge = 1; p.nWh_int = nWh_int; p.nWw_int = nWw_int;
      n_win = (int64_t)p.B * ((int64_t)p.nWh * p.nWw - (int64_t)nWh_int * nWw_int);
      if (n_win == 0) return TFSWA_OK;
    }
  }
  dim3 grid((unsigned)n_win, p.C / CS, 1);
#define TFSWA_WIN_CASE(d, cs) if (D == d && CS == cs) { attn_win_mma_kernel<d, cs><<<grid, WIN_THREADS, 0, st>>>(p); return check_launch("attn_win_tc"); }
  TFSWA_WIN_CASE(4, 32) TFSWA_WIN_CASE(4, 64) TFSWA_WIN_CASE(8, 32) TFSWA_WIN_CASE(8, 64)
  TFSWA_WIN_CASE(16, 32) TFSWA_WIN_CASE(16, 64) TFSWA_WIN_CASE(32, 64)
#undef TFSWA_WIN_CASE
  TFSWA_REQUIRE(false, "attn_win_tc: unsupported head_dim %d / slab %d", D, CS);
}
