// Backward of the axial (TSA / FSA) and 8x8-window (SW-MSA) attention on warp-level tensor-core MMAs, bf16 activations,
// head_dim 4 / 8 / 16 (autograd of attention.py:70-85 through the same index maps as the forward; for windows the
// pad / roll / partition chain of attention.py:358-375 is the token map, zero-padded tokens are real keys whose k|v is
// the folded qkv bias `pad_kv` and whose gradient reduces into `dpad`, and are absent as queries: their output is cropped).  Flash-style recomputation from q, k and
// the saved log2-sum-exp, everything in the FlashAttention-2 register layout (no N x N tensor, no atomics):
//     s = q.k * scale,  p = exp2(s*log2e - lse),  dp = dO.v,  ds = p (dp - D),  D_i = dO_i . O_i
//     dq_i = scale * sum_j ds_ij k_j        (dq kernel: a warp owns 16 query rows of one head, streams 64-key tiles)
//     dk_j = scale * sum_i ds_ij q_i,  dv_j = sum_i p_ij dO_i
//                                           (dkv kernel: a warp owns 16 keys of one head, streams 64-query tiles)
// In the dq kernel S = Q K^T and dP = dO V^T are two MMAs per 8 keys, dS is re-packed in place as the A operand of
// dQ += dS K.  The dkv kernel works on the transposed problem (S^T = K Q^T, dP^T = V dO^T), so P^T and dS^T come out of
// the accumulators already in A-operand layout for dV += P^T dO and dK += dS^T Q.  Operands with the reduction index
// along tokens are read from the token-major shared-memory tiles with ldmatrix.trans.
// One CTA = one sequence (or window: N = 64, one tile) x 64 rows x 8 heads (warp = head); tiles arrive by cp.async into a 3-deep ring.
// head_dim 4 uses 4 of the 8 K slots of its m16n8k8 MMAs; two of the spare slots carry -D_i (as a bf16 hi + lo pair,
// 16 mantissa bits) against ones in the other operand, so dP - D comes out of the dP MMA and dS costs one multiply.
// Replaces the CUDA-core attn_bwd_dq/dkv kernels for these shapes (55 % of a bf16 training step before).
#include "attn_common.cuh"
#include <stdlib.h>
#include <type_traits>

namespace tfswa {

namespace {

__device__ __forceinline__ void b_mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void b_mma16816_z(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
__device__ __forceinline__ void b_mma1688_z(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%7,%7,%7};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0), "f"(0.f));
}
__device__ __forceinline__ void b_ldsm_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ uint32_t b_pack(float lo, float hi) {
  uint32_t y;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo));
  return y;
}
// x as two bf16 values whose sum carries 16 mantissa bits: packed (hi, lo) for two adjacent K slots of an MMA operand
__device__ __forceinline__ uint32_t b_split2(float x) {
  const float hi = __bfloat162float(__float2bfloat16_rn(x));
  return b_pack(hi, x - hi);
}
constexpr uint32_t B_ONES2 = 0x3F803F80u;        // bf16 (1, 1)
__device__ __forceinline__ void b_cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void b_cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void b_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void b_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// resident CTAs per SM the window instantiations are compiled for: one window is little work behind a long load chain, so
// at head_dim 4 / 8 a third CTA (80 registers, some spills) beats two at 128 (stage 1, B=8: 4.03 -> 3.68 ms); at head_dim 16
// the spills cost more than the occupancy buys (0.60 vs 0.73 ms)
template <int D> __host__ __device__ constexpr int bw_win_ctas() { return D <= 8 ? 3 : 2; }
constexpr int BW_T = 64;             // rows per streamed tile
constexpr int BW_THREADS = 256;      // 8 warps = 8 heads
template <int D> __host__ __device__ constexpr int bw_mt() { return D == 16 ? 2 : 4; }   // 16-row tiles owned by a warp (register budget)

// S-type product of one 16-row A fragment set against 8 columns (one n-tile): D <= 8 uses one k8 step
template <int D>
__device__ __forceinline__ void s_mma(float (&d)[4], const uint32_t (&a)[(D + 15) / 16][4], const uint32_t (&b)[(D + 15) / 16][2]) {
  if (D <= 8) {
    b_mma1688_z(d, a[0][0], a[0][1], b[0][0]);
  } else {
    b_mma16816_z(d, a[0], b[0][0], b[0][1]);
#pragma unroll
    for (int ks = 1; ks < (D + 15) / 16; ++ks) b_mma16816(d, a[ks], b[ks][0], b[ks][1]);
  }
}

// A-operand fragments (16 rows x head dims) of rows (r0 + g, r0 + g + 8) of a token-major global tensor; absent rows = 0
template <int D>
__device__ __forceinline__ void load_a_frag(uint32_t (&a)[(D + 15) / 16][4], const bf16* base, int64_t ld, int64_t tok0, int64_t tok1,
                                            bool ok0, bool ok1, int t) {
#pragma unroll
  for (int ks = 0; ks < (D + 15) / 16; ++ks) {
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const int dcol = ks * 16 + hi * 8 + 2 * t;
      a[ks][hi * 2 + 0] = (dcol < D && ok0) ? *reinterpret_cast<const uint32_t*>(base + tok0 * ld + dcol) : 0u;
      a[ks][hi * 2 + 1] = (dcol < D && ok1) ? *reinterpret_cast<const uint32_t*>(base + tok1 * ld + dcol) : 0u;
    }
  }
}

// token of element n of sequence / window `row`.  Returns false when n lies outside the sequence; real = false for the
// zero-padded tokens of a window (present as keys with k|v = pad_kv, absent as queries).
// Windows: `wtok` is the CTA's shared table of the 64 window positions (token index, -1 for a zero-padded position),
// filled once per CTA - the pad / roll / partition map costs four integer divisions per lookup otherwise, which was
// more than half of the window kernels' instructions.
template <bool WIN>
__device__ __forceinline__ bool seq_token(const int64_t* wtok, int n, int N, int64_t tok_base, int64_t tok_stride,
                                          int64_t& tok, bool& real) {
  tok = 0; real = false;
  if (n >= N) return false;
  if (WIN) { const int64_t t = wtok[n]; real = t >= 0; tok = real ? t : 0; }
  else { tok = tok_base + (int64_t)n * tok_stride; real = true; }
  return true;
}

// rows (0: g, 1: g + 8) of an A fragment set filled from a per-channel fp32 vector (the pad token's k or v)
template <int D>
__device__ __forceinline__ void fill_a_frag_row(uint32_t (&a)[(D + 15) / 16][4], const float* vec, int which, int t) {
#pragma unroll
  for (int ks = 0; ks < (D + 15) / 16; ++ks) {
#pragma unroll
    for (int hi = 0; hi < 2; ++hi) {
      const int dcol = ks * 16 + hi * 8 + 2 * t;
      if (dcol < D) a[ks][hi * 2 + which] = b_pack(vec[dcol], vec[dcol + 1]);
    }
  }
}

}  // namespace

template <int D> __host__ __device__ constexpr int bw_smem_bytes() { return 3 * 2 * BW_T * (8 * D + 8) * 2; }

// ------------------------------------------------------------------------------------------------
// dq (and D_i = dO_i . O_i): grid (ceil(N/64), sequences, heads/8); windows: (windows, ceil(64/rows per CTA), heads/8)
// ------------------------------------------------------------------------------------------------
template <int D, bool WIN>
__global__ void __launch_bounds__(BW_THREADS, WIN ? bw_win_ctas<D>() : 2) attn_bwd_dq_mma_kernel(const AttnParams p) {
  constexpr int CS = 8 * D, PITCH = CS + 8, KS = (D + 15) / 16, DN = (D + 7) / 8, CPT = CS / 8, MT = bw_mt<D>();
  extern __shared__ __align__(16) uint8_t bw_smem[];
  typedef bf16 (*tile_t)[BW_T][PITCH];
  tile_t Ks = reinterpret_cast<tile_t>(bw_smem);
  tile_t Vs = reinterpret_cast<tile_t>(bw_smem + 3 * BW_T * PITCH * 2);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int row = WIN ? blockIdx.x : blockIdx.y, q0 = (WIN ? blockIdx.y : blockIdx.x) * (16 * MT), slab = blockIdx.z;
  const int N = WIN ? p.ws * p.ws : (p.geom == TFSWA_GEOM_TSA ? p.H : p.W);
  const int T = (N + BW_T - 1) / BW_T;
  int64_t tok_base = 0, tok_stride = 1;
  __shared__ int64_t s_wtok[WIN ? BW_T : 1];
  if (!WIN) {
    if (p.geom == TFSWA_GEOM_TSA) { const int b = row / p.W; tok_base = (int64_t)b * p.H * p.W + (row - b * p.W); tok_stride = p.W; }
    else { tok_base = (int64_t)row * p.W; }
  } else {
    if (tid < BW_T) { bool v; const int64_t tk = token_of<true>(p, row, tid, v); s_wtok[tid] = v ? tk : -1; }
    __syncthreads();
  }
  const bf16* qkv = (const bf16*)p.qkv;
  const int cbase = warp * D, head = slab * 8 + warp;
  const int mt_valid = min(MT, (N - q0 + 15) / 16);

  auto stage = [&](int tt, int b) {
    if (tt < T) {
      for (int v = tid; v < BW_T * 2 * CPT; v += BW_THREADS) {
        const int j = v / (2 * CPT), rem = v - j * 2 * CPT, part = rem / CPT, chunk = rem - part * CPT;
        bf16* dst = part ? &Vs[b][j][chunk * 8] : &Ks[b][j][chunk * 8];
        int64_t tok; bool real;
        const bool in = seq_token<WIN>(s_wtok, tt * BW_T + j, N, tok_base, tok_stride, tok, real);
        if (in && real) {
          b_cp_async16(dst, qkv + tok * p.ldq + (1 + part) * p.C + slab * CS + chunk * 8);
        } else if (WIN && in) {                                        // zero-padded window token: k|v = folded qkv bias
          const float* pk = p.pad_kv + part * p.C + slab * CS + chunk * 8;
          *reinterpret_cast<uint4*>(dst) = make_uint4(b_pack(pk[0], pk[1]), b_pack(pk[2], pk[3]), b_pack(pk[4], pk[5]), b_pack(pk[6], pk[7]));
        } else {
          *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        }
      }
    }
    b_cp_commit();
  };
  stage(0, 0);
  stage(1, 1);

  // ---- my rows: Q and dO fragments, lse, D = dO . O ----
  uint32_t qa[MT][KS][4], ga[MT][KS][4];
  float lse[MT][2], dsm[MT][2], dq[MT][DN][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    int64_t tok0, tok1; bool r0, r1;
    const bool ok0 = seq_token<WIN>(s_wtok, q0 + mt * 16 + g, N, tok_base, tok_stride, tok0, r0) && r0;
    const bool ok1 = seq_token<WIN>(s_wtok, q0 + mt * 16 + g + 8, N, tok_base, tok_stride, tok1, r1) && r1;
    load_a_frag<D>(qa[mt], qkv + slab * CS + cbase, p.ldq, tok0, tok1, ok0, ok1, t);
    load_a_frag<D>(ga[mt], (const bf16*)p.dout + slab * CS + cbase, p.ldo, tok0, tok1, ok0, ok1, t);
    uint32_t oa[KS][4];
    load_a_frag<D>(oa, (const bf16*)p.o + slab * CS + cbase, p.ldo, tok0, tok1, ok0, ok1, t);
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        const uint32_t g0 = ga[mt][ks][hi * 2], g1 = ga[mt][ks][hi * 2 + 1], o0 = oa[ks][hi * 2], o1 = oa[ks][hi * 2 + 1];
        d0 += __uint_as_float(g0 << 16) * __uint_as_float(o0 << 16) + __uint_as_float(g0 & 0xFFFF0000u) * __uint_as_float(o0 & 0xFFFF0000u);
        d1 += __uint_as_float(g1 << 16) * __uint_as_float(o1 << 16) + __uint_as_float(g1 & 0xFFFF0000u) * __uint_as_float(o1 & 0xFFFF0000u);
      }
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    dsm[mt][0] = d0; dsm[mt][1] = d1;
    if (D == 4 && t == 2) { ga[mt][0][0] = b_split2(-d0); ga[mt][0][1] = b_split2(-d1); }   // K slots 4, 5 of my two rows
    lse[mt][0] = ok0 ? p.lse[tok0 * p.heads + head] : CUDART_INF_F;
    lse[mt][1] = ok1 ? p.lse[tok1 * p.heads + head] : CUDART_INF_F;
    if (t == 0) {
      if (ok0) p.dsum[tok0 * p.heads + head] = d0;
      if (ok1) p.dsum[tok1 * p.heads + head] = d1;
    }
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) dq[mt][dn][0] = dq[mt][dn][1] = dq[mt][dn][2] = dq[mt][dn][3] = 0.f;
  }
  const float c = p.qscale;

  for (int tt = 0; tt < T; ++tt) {
    b_cp_wait<1>();
    __syncthreads();
    stage(tt + 2, (tt + 2) % 3);
    const int b = tt % 3;
    const int kcount = min(BW_T, N - tt * BW_T);
    // K and V as "n = key" operands (for S and dP), K as "k = key" operand (for dQ)
    uint32_t kb[8][KS][2], vb[8][KS][2], kt[4][DN][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int dcol = ks * 16 + 2 * t;
        kb[nt][ks][0] = dcol < D ? *reinterpret_cast<const uint32_t*>(&Ks[b][nt * 8 + g][cbase + dcol]) : 0u;
        kb[nt][ks][1] = dcol + 8 < D ? *reinterpret_cast<const uint32_t*>(&Ks[b][nt * 8 + g][cbase + dcol + 8]) : 0u;
        vb[nt][ks][0] = dcol < D ? *reinterpret_cast<const uint32_t*>(&Vs[b][nt * 8 + g][cbase + dcol]) : 0u;
        vb[nt][ks][1] = dcol + 8 < D ? *reinterpret_cast<const uint32_t*>(&Vs[b][nt * 8 + g][cbase + dcol + 8]) : 0u;
      }
      if (D == 4 && t == 2) vb[nt][0][0] = B_ONES2;
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dn = 0; dn < DN; ++dn) {
        const int col = D >= 8 ? cbase + dn * 8 : (cbase & ~7);
        b_ldsm_x2_trans(kt[kk][dn][0], kt[kk][dn][1], &Ks[b][kk * 16 + (lane & 15)][col]);
      }
    }
    // Only the last tile of a sequence holds absent keys; the per-element mask (a compare + select per exponential) is
    // compiled into that tile's copy of the loop only.
    auto tile_body = [&](auto masked) {
      constexpr bool MASK = decltype(masked)::value;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        if (mt < mt_valid) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            uint32_t pa[4];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
              const int nt = 2 * kk + half;
              float s[4], dp[4];
              s_mma<D>(s, qa[mt], kb[nt]);
              s_mma<D>(dp, ga[mt], vb[nt]);
              float ds[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int h2 = i >> 1;
                float pw = fast_exp2(fmaf(s[i], c, -lse[mt][h2]));
                if (MASK && nt * 8 + 2 * t + (i & 1) >= kcount) pw = 0.f;     // absent keys of the last tile
                ds[i] = D == 4 ? pw * dp[i] : pw * (dp[i] - dsm[mt][h2]);
              }
              pa[half * 2 + 0] = b_pack(ds[0], ds[1]);
              pa[half * 2 + 1] = b_pack(ds[2], ds[3]);
            }
#pragma unroll
            for (int dn = 0; dn < DN; ++dn) b_mma16816(dq[mt][dn], pa, kt[kk][dn][0], kt[kk][dn][1]);
          }
        }
      }
    };
    if (kcount < BW_T) tile_body(std::true_type{}); else tile_body(std::false_type{});
  }
  b_cp_wait<0>();
  // ---- dq = scale * acc ----
  bf16* dqkv = (bf16*)p.dqkv;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      int64_t tok; bool real;                              // recomputed rather than kept live across the key loop
      if (!(seq_token<WIN>(s_wtok, q0 + mt * 16 + g + h2 * 8, N, tok_base, tok_stride, tok, real) && real)) continue;
      if (D >= 8) {
#pragma unroll
        for (int dn = 0; dn < DN; ++dn)
          *reinterpret_cast<uint32_t*>(dqkv + tok * p.ldq + slab * CS + cbase + dn * 8 + 2 * t) =
              b_pack(dq[mt][dn][h2 * 2] * p.scale, dq[mt][dn][h2 * 2 + 1] * p.scale);
      } else {
        const int first = cbase & 7;
        if (2 * t >= first && 2 * t < first + 4)
          *reinterpret_cast<uint32_t*>(dqkv + tok * p.ldq + slab * CS + (cbase & ~7) + 2 * t) =
              b_pack(dq[mt][0][h2 * 2] * p.scale, dq[mt][0][h2 * 2 + 1] * p.scale);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dk, dv: same grids as the dq kernel; needs dsum from the dq kernel
// ------------------------------------------------------------------------------------------------
template <int D, bool WIN>
__global__ void __launch_bounds__(BW_THREADS, WIN ? bw_win_ctas<D>() : 2) attn_bwd_dkv_mma_kernel(const AttnParams p) {
  constexpr int CS = 8 * D, PITCH = CS + 8, KS = (D + 15) / 16, DN = (D + 7) / 8, CPT = CS / 8, MT = bw_mt<D>();
  extern __shared__ __align__(16) uint8_t bw_smem[];
  typedef bf16 (*tile_t)[BW_T][PITCH];
  tile_t Qs = reinterpret_cast<tile_t>(bw_smem);
  tile_t Gs = reinterpret_cast<tile_t>(bw_smem + 3 * BW_T * PITCH * 2);
  __shared__ __align__(8) float Ls[3][8][BW_T], Ds[3][8][BW_T];          // [ring slot][head][query]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int row = WIN ? blockIdx.x : blockIdx.y, k0 = (WIN ? blockIdx.y : blockIdx.x) * (16 * MT), slab = blockIdx.z;
  const int N = WIN ? p.ws * p.ws : (p.geom == TFSWA_GEOM_TSA ? p.H : p.W);
  const int T = (N + BW_T - 1) / BW_T;
  int64_t tok_base = 0, tok_stride = 1;
  __shared__ int64_t s_wtok[WIN ? BW_T : 1];
  if (!WIN) {
    if (p.geom == TFSWA_GEOM_TSA) { const int b = row / p.W; tok_base = (int64_t)b * p.H * p.W + (row - b * p.W); tok_stride = p.W; }
    else { tok_base = (int64_t)row * p.W; }
  } else {
    if (tid < BW_T) { bool v; const int64_t tk = token_of<true>(p, row, tid, v); s_wtok[tid] = v ? tk : -1; }
    __syncthreads();
  }
  const bf16* qkv = (const bf16*)p.qkv;
  const bf16* dout = (const bf16*)p.dout;
  const int cbase = warp * D;
  const int mt_valid = min(MT, (N - k0 + 15) / 16);

  auto stage = [&](int tt, int b) {
    if (tt < T) {
      for (int v = tid; v < BW_T * 2 * CPT; v += BW_THREADS) {
        const int j = v / (2 * CPT), rem = v - j * 2 * CPT, part = rem / CPT, chunk = rem - part * CPT;
        bf16* dst = part ? &Gs[b][j][chunk * 8] : &Qs[b][j][chunk * 8];
        int64_t tok; bool real;
        if (seq_token<WIN>(s_wtok, tt * BW_T + j, N, tok_base, tok_stride, tok, real) && real) {
          b_cp_async16(dst, part ? dout + tok * p.ldo + slab * CS + chunk * 8 : qkv + tok * p.ldq + slab * CS + chunk * 8);
        } else {                                                         // absent or zero-padded (cropped) query
          *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
        }
      }
      for (int v = tid; v < BW_T * 8; v += BW_THREADS) {                 // lse / D of the tile's queries, all 8 heads
        const int j = v >> 3, h = v & 7;
        int64_t tok; bool real;
        if (seq_token<WIN>(s_wtok, tt * BW_T + j, N, tok_base, tok_stride, tok, real) && real) {
          b_cp_async4(&Ls[b][h][j], p.lse + tok * p.heads + slab * 8 + h);      // asynchronous: no load-to-store stall per tile
          b_cp_async4(&Ds[b][h][j], p.dsum + tok * p.heads + slab * 8 + h);
        } else {
          Ls[b][h][j] = CUDART_INF_F; Ds[b][h][j] = 0.f;                 // +inf -> p = 0 for absent queries
        }
      }
    }
    b_cp_commit();
  };
  stage(0, 0);
  stage(1, 1);

  // ---- my keys: K and V fragments as A operands (rows = keys) ----
  uint32_t ka[MT][KS][4], va[MT][KS][4];
  float dk[MT][DN][4], dv[MT][DN][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    int64_t tok0, tok1; bool r0, r1;
    const bool in0 = seq_token<WIN>(s_wtok, k0 + mt * 16 + g, N, tok_base, tok_stride, tok0, r0);
    const bool in1 = seq_token<WIN>(s_wtok, k0 + mt * 16 + g + 8, N, tok_base, tok_stride, tok1, r1);
    load_a_frag<D>(ka[mt], qkv + p.C + slab * CS + cbase, p.ldq, tok0, tok1, in0 && r0, in1 && r1, t);
    load_a_frag<D>(va[mt], qkv + 2 * p.C + slab * CS + cbase, p.ldq, tok0, tok1, in0 && r0, in1 && r1, t);
    if (WIN) {
      if (in0 && !r0) { fill_a_frag_row<D>(ka[mt], p.pad_kv + slab * CS + cbase, 0, t); fill_a_frag_row<D>(va[mt], p.pad_kv + p.C + slab * CS + cbase, 0, t); }
      if (in1 && !r1) { fill_a_frag_row<D>(ka[mt], p.pad_kv + slab * CS + cbase, 1, t); fill_a_frag_row<D>(va[mt], p.pad_kv + p.C + slab * CS + cbase, 1, t); }
    }
    if (D == 4 && t == 2) { va[mt][0][0] = B_ONES2; va[mt][0][1] = B_ONES2; }
#pragma unroll
    for (int dn = 0; dn < DN; ++dn) {
      dk[mt][dn][0] = dk[mt][dn][1] = dk[mt][dn][2] = dk[mt][dn][3] = 0.f;
      dv[mt][dn][0] = dv[mt][dn][1] = dv[mt][dn][2] = dv[mt][dn][3] = 0.f;
    }
  }
  const float c = p.qscale;

  for (int tt = 0; tt < T; ++tt) {
    b_cp_wait<1>();
    __syncthreads();
    stage(tt + 2, (tt + 2) % 3);
    const int b = tt % 3;
    // Q and dO as "n = query" operands (S^T, dP^T) and as "k = query" operands (dK, dV)
    uint32_t qb[8][KS][2], gb[8][KS][2], qt[4][DN][2], gt[4][DN][2];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        const int dcol = ks * 16 + 2 * t;
        qb[nt][ks][0] = dcol < D ? *reinterpret_cast<const uint32_t*>(&Qs[b][nt * 8 + g][cbase + dcol]) : 0u;
        qb[nt][ks][1] = dcol + 8 < D ? *reinterpret_cast<const uint32_t*>(&Qs[b][nt * 8 + g][cbase + dcol + 8]) : 0u;
        gb[nt][ks][0] = dcol < D ? *reinterpret_cast<const uint32_t*>(&Gs[b][nt * 8 + g][cbase + dcol]) : 0u;
        gb[nt][ks][1] = dcol + 8 < D ? *reinterpret_cast<const uint32_t*>(&Gs[b][nt * 8 + g][cbase + dcol + 8]) : 0u;
      }
      if (D == 4 && t == 2) gb[nt][0][0] = b_split2(-Ds[b][warp][nt * 8 + g]);
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int dn = 0; dn < DN; ++dn) {
        const int col = D >= 8 ? cbase + dn * 8 : (cbase & ~7);
        b_ldsm_x2_trans(qt[kk][dn][0], qt[kk][dn][1], &Qs[b][kk * 16 + (lane & 15)][col]);
        b_ldsm_x2_trans(gt[kk][dn][0], gt[kk][dn][1], &Gs[b][kk * 16 + (lane & 15)][col]);
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      if (mt < mt_valid) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t pa[4], da[4];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int nt = 2 * kk + half;
            float s[4], dp[4];
            s_mma<D>(s, ka[mt], qb[nt]);                               // S^T: rows = my keys, columns = queries
            s_mma<D>(dp, va[mt], gb[nt]);                              // dP^T
            const float2 l2 = *reinterpret_cast<const float2*>(&Ls[b][warp][nt * 8 + 2 * t]);
            const float2 d2 = *reinterpret_cast<const float2*>(&Ds[b][warp][nt * 8 + 2 * t]);
            float pw[4], ds[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              pw[i] = fast_exp2(fmaf(s[i], c, -((i & 1) ? l2.y : l2.x)));
              ds[i] = D == 4 ? pw[i] * dp[i] : pw[i] * (dp[i] - ((i & 1) ? d2.y : d2.x));
            }
            pa[half * 2 + 0] = b_pack(pw[0], pw[1]); pa[half * 2 + 1] = b_pack(pw[2], pw[3]);
            da[half * 2 + 0] = b_pack(ds[0], ds[1]); da[half * 2 + 1] = b_pack(ds[2], ds[3]);
          }
#pragma unroll
          for (int dn = 0; dn < DN; ++dn) {
            b_mma16816(dv[mt][dn], pa, gt[kk][dn][0], gt[kk][dn][1]);
            b_mma16816(dk[mt][dn], da, qt[kk][dn][0], qt[kk][dn][1]);
          }
        }
      }
    }
  }
  b_cp_wait<0>();
  bf16* dqkv = (bf16*)p.dqkv;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      int64_t tok; bool real;                              // recomputed rather than kept live across the query loop
      if (!seq_token<WIN>(s_wtok, k0 + mt * 16 + g + h2 * 8, N, tok_base, tok_stride, tok, real)) continue;
      if (WIN && !real) {                                  // zero-padded key: its gradient belongs to the folded qkv bias
        if (p.dpad) {
          float* dp = p.dpad + slab * CS;
          if (D >= 8) {
#pragma unroll
            for (int dn = 0; dn < DN; ++dn) {
              const int col = cbase + dn * 8 + 2 * t;
              atomicAdd(dp + col, dk[mt][dn][h2 * 2] * p.scale); atomicAdd(dp + col + 1, dk[mt][dn][h2 * 2 + 1] * p.scale);
              atomicAdd(dp + p.C + col, dv[mt][dn][h2 * 2]); atomicAdd(dp + p.C + col + 1, dv[mt][dn][h2 * 2 + 1]);
            }
          } else {
            const int first = cbase & 7;
            if (2 * t >= first && 2 * t < first + 4) {
              const int col = (cbase & ~7) + 2 * t;
              atomicAdd(dp + col, dk[mt][0][h2 * 2] * p.scale); atomicAdd(dp + col + 1, dk[mt][0][h2 * 2 + 1] * p.scale);
              atomicAdd(dp + p.C + col, dv[mt][0][h2 * 2]); atomicAdd(dp + p.C + col + 1, dv[mt][0][h2 * 2 + 1]);
            }
          }
        }
        continue;
      }
      bf16* base = dqkv + tok * p.ldq + slab * CS;
      if (D >= 8) {
#pragma unroll
        for (int dn = 0; dn < DN; ++dn) {
          const int col = cbase + dn * 8 + 2 * t;
          *reinterpret_cast<uint32_t*>(base + p.C + col) = b_pack(dk[mt][dn][h2 * 2] * p.scale, dk[mt][dn][h2 * 2 + 1] * p.scale);
          *reinterpret_cast<uint32_t*>(base + 2 * p.C + col) = b_pack(dv[mt][dn][h2 * 2], dv[mt][dn][h2 * 2 + 1]);
        }
      } else {
        const int first = cbase & 7;
        if (2 * t >= first && 2 * t < first + 4) {
          const int col = (cbase & ~7) + 2 * t;
          *reinterpret_cast<uint32_t*>(base + p.C + col) = b_pack(dk[mt][0][h2 * 2] * p.scale, dk[mt][0][h2 * 2 + 1] * p.scale);
          *reinterpret_cast<uint32_t*>(base + 2 * p.C + col) = b_pack(dv[mt][0][h2 * 2], dv[mt][0][h2 * 2 + 1]);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// One-pass backward for head_dim 4 and 8 (TSA / FSA): the dq / dkv pair above computes every P_ij twice, and at head_dim 4 the
// exponentials and the instructions around them are the whole cost.  Here each (key block, query tile) pair is visited
// once: S^T, dP^T, P^T and dS^T are formed in the dkv orientation (rows = keys), dV += P^T dO and dK += dS^T Q accumulate in
// registers over the query loop, and dQ += dS K uses the SAME dS^T fragments, transposed in registers with movmatrix, into
// an fp32 accumulator of the whole sequence in shared memory (N x 32 channels, pitch 36 floats: conflict-free float2
// read-modify-write).  One CTA = one sequence x 8 heads, 16 warps: warp = (head, half of the query n-tiles), so the two
// warps of a head own disjoint rows of the dQ accumulator (no atomics, deterministic) and their dK / dV partials are summed
// through shared memory once per 32-key block.  D_i = dO_i . O_i comes from a small pre-kernel.
// ------------------------------------------------------------------------------------------------
constexpr int FB_THREADS = 512;
constexpr int FB_KB = 32;                 // keys per outer step: two 16-key m-tiles per warp
template <int D> __host__ __device__ constexpr int fb_qp() { return 8 * D + 4; }      // fp32 pitch of a dQ accumulator row
template <int D> __host__ __device__ constexpr int fb_pitch() { return 8 * D + 8; }   // bf16 pitch of the Q / dO tiles
template <int D> __host__ __device__ inline int fb_smem_bytes(int N) {
  const int T = (N + BW_T - 1) / BW_T;
  return T * BW_T * fb_qp<D>() * 4 + 3 * 2 * BW_T * fb_pitch<D>() * 2 + 2 * 3 * 8 * BW_T * 4 + 8 * 2 * 32 * 8 * 4;
}

__device__ __forceinline__ uint32_t b_movm_trans(uint32_t a) {
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(d) : "r"(a));
  return d;
}

// D_i = dO_i . O_i per (token, head), head_dim 4 or 8: one thread per 8 channels = two heads or one
template <int D>
__global__ void __launch_bounds__(256) attn_bwd_delta_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ o, int64_t ldo,
                                                              float* __restrict__ dsum, int64_t M, int C, int heads) {
  const int cpt = C / 8;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M * cpt; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t tok = i / cpt;
    const int c8 = (int)(i - tok * cpt);
    float a[8], b[8];
    load8(dout + tok * ldo + c8 * 8, a);
    load8(o + tok * ldo + c8 * 8, b);
    const float lo = a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3];
    const float hi = a[4] * b[4] + a[5] * b[5] + a[6] * b[6] + a[7] * b[7];
    if (D == 4) { dsum[tok * heads + 2 * c8] = lo; dsum[tok * heads + 2 * c8 + 1] = hi; }
    else dsum[tok * heads + c8] = lo + hi;
  }
}

template <int D>
__global__ void __launch_bounds__(FB_THREADS, 1) attn_bwd_fused_kernel(const AttnParams p) {
  constexpr int CS = 8 * D, CPT = D, FB_QP = fb_qp<D>(), FB_PITCH = fb_pitch<D>();
  constexpr int SROWS = FB_THREADS / (2 * CPT), SPASS = BW_T / SROWS;      // rows covered by one staging pass / passes per tile
  extern __shared__ __align__(16) uint8_t fb_smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, t = lane & 3;
  const int head_l = warp & 7, qh = warp >> 3;
  const int row = blockIdx.x, slab = blockIdx.y;
  const int N = p.geom == TFSWA_GEOM_TSA ? p.H : p.W;
  const int T = (N + BW_T - 1) / BW_T, KSTEPS = (N + FB_KB - 1) / FB_KB;
  int64_t tok_base, tok_stride;
  if (p.geom == TFSWA_GEOM_TSA) { const int b = row / p.W; tok_base = (int64_t)b * p.H * p.W + (row - b * p.W); tok_stride = p.W; }
  else { tok_base = (int64_t)row * p.W; tok_stride = 1; }
  float* dq_s = reinterpret_cast<float*>(fb_smem);
  typedef bf16 (*tile_t)[BW_T][FB_PITCH];
  uint8_t* ptr = fb_smem + (size_t)T * BW_T * FB_QP * 4;
  tile_t Qs = reinterpret_cast<tile_t>(ptr); ptr += 3 * BW_T * FB_PITCH * 2;
  tile_t Gs = reinterpret_cast<tile_t>(ptr); ptr += 3 * BW_T * FB_PITCH * 2;
  typedef float (*stat_t)[8][BW_T];
  stat_t Ls = reinterpret_cast<stat_t>(ptr); ptr += 3 * 8 * BW_T * 4;
  stat_t Ds = reinterpret_cast<stat_t>(ptr); ptr += 3 * 8 * BW_T * 4;
  float* red = reinterpret_cast<float*>(ptr);                       // [head][m-tile][lane][dk 4 | dv 4]
  const bf16* qkv = (const bf16*)p.qkv;
  const bf16* dout = (const bf16*)p.dout;
  const int cbase = head_l * D, head = slab * 8 + head_l;
  const int first = cbase & 7;
  const float c = p.qscale;

  for (int i = tid; i < T * BW_T * FB_QP; i += FB_THREADS) dq_s[i] = 0.f;

  // Staging state kept incrementally (one add per tile instead of 64-bit multiplies, a modulo by T and a modulo by 3 per
  // tile: the index arithmetic was ~45 % of this kernel's instructions in its first version).  Thread -> (row j, part, chunk)
  // of the Q | dO copy and (row j, head h) of the lse | D copy are fixed; per tile only the query index advances.
  const int total = KSTEPS * T;                                     // the query tiles are streamed once per key block
  const int sj = tid / (2 * CPT), srem = tid % (2 * CPT), spart = srem / CPT, schunk = srem % CPT;   // Q | dO copy: row sj (+ SROWS per pass)
  const int lj = tid >> 3, lh = tid & 7;                                                                 // lse | D copy: row lj, head lh
  const int64_t s_ld = spart ? p.ldo : p.ldq;
  const bf16* s_src0 = (spart ? dout + tok_base * p.ldo : qkv + tok_base * p.ldq) + slab * CS + schunk * 8
                       + (int64_t)sj * tok_stride * s_ld;
  const int64_t s_src_step = (int64_t)BW_T * tok_stride * s_ld;                                 // elements per 64-query tile
  const int64_t s_pass_step = (int64_t)SROWS * tok_stride * s_ld;                               // elements per staging pass
  const float* s_lse0 = p.lse + (tok_base + (int64_t)lj * tok_stride) * p.heads + slab * 8 + lh;
  const float* s_dsm0 = p.dsum + (tok_base + (int64_t)lj * tok_stride) * p.heads + slab * 8 + lh;
  const int64_t s_stat_step = (int64_t)BW_T * tok_stride * p.heads;
  const uint32_t s_dst0 = (uint32_t)__cvta_generic_to_shared(spart ? &Gs[0][sj][schunk * 8] : &Qs[0][sj][schunk * 8]);
  int s_gi = 0, s_tt = 0, s_slot = 0;                                // next tile to stage, its query tile and ring slot
  const bf16* s_src = s_src0; const float* s_lse = s_lse0; const float* s_dsm = s_dsm0;
  // Two halves: `stage_issue` starts the Q | dO copy (cp.async) and the lse | D loads into two registers, `stage_finish`
  // parks those registers in shared memory.  The consumer calls them before and after a tile's math, so the loads' latency
  // hides under the MMAs instead of stalling every warp right after the barrier (4-byte cp.async for them was slower: the
  // shared-memory instruction queue is the busiest unit of this kernel).
  float s_l = CUDART_INF_F, s_d = 0.f;
  float* s_lsp = &Ls[0][lh][lj]; float* s_dsp = &Ds[0][lh][lj];
  auto stage_issue = [&]() {
    s_l = CUDART_INF_F; s_d = 0.f;                                   // +inf -> p = 0 for absent queries
    s_lsp = &Ls[s_slot][lh][lj]; s_dsp = &Ds[s_slot][lh][lj];
    if (s_gi < total) {
      const uint32_t dst = s_dst0 + (uint32_t)s_slot * (BW_T * FB_PITCH * 2);
#pragma unroll
      for (int c = 0; c < SPASS; ++c) {
        if (s_tt * BW_T + sj + c * SROWS < N)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + c * (SROWS * FB_PITCH * 2)), "l"(s_src + c * s_pass_step) : "memory");
        else
          asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(dst + c * (SROWS * FB_PITCH * 2)), "r"(0u) : "memory");
      }
      if (s_tt * BW_T + lj < N) { s_l = *s_lse; s_d = *s_dsm; }
      ++s_gi; s_slot = s_slot == 2 ? 0 : s_slot + 1;
      if (++s_tt == T) { s_tt = 0; s_src = s_src0; s_lse = s_lse0; s_dsm = s_dsm0; }
      else { s_src += s_src_step; s_lse += s_stat_step; s_dsm += s_stat_step; }
    }
    b_cp_commit();
  };
  auto stage_finish = [&]() { *s_lsp = s_l; *s_dsp = s_d; };
  stage_issue(); stage_finish();
  stage_issue(); stage_finish();
  int cslot = 0;                                                      // ring slot of the tile being consumed

  bf16* dqkv = (bf16*)p.dqkv;
  float* const dq_mine = dq_s + (2 * qh * 16 + g) * FB_QP + cbase + 2 * t;   // my first accumulator row of a tile, my columns
  for (int ko = 0; ko < KSTEPS; ++ko) {
    const int k0 = ko * FB_KB;
    // ---- my keys: K, V as A operands (rows = keys), K once more as the B operand of dQ += dS K ----
    uint32_t ka[2][1][4], va[2][1][4], kq[2][2];
    float dk[2][4], dv[2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
      const int kn0 = k0 + mt * 16 + g, kn1 = kn0 + 8;
      const bool ok0 = kn0 < N, ok1 = kn1 < N;
      const int64_t tok0 = tok_base + (int64_t)(ok0 ? kn0 : 0) * tok_stride, tok1 = tok_base + (int64_t)(ok1 ? kn1 : 0) * tok_stride;
      load_a_frag<D>(ka[mt], qkv + p.C + slab * CS + cbase, p.ldq, tok0, tok1, ok0, ok1, t);
      load_a_frag<D>(va[mt], qkv + 2 * p.C + slab * CS + cbase, p.ldq, tok0, tok1, ok0, ok1, t);
      if (D == 4 && t == 2) { va[mt][0][0] = B_ONES2; va[mt][0][1] = B_ONES2; }   // the -D_i slots of dO meet ones here
#pragma unroll
      for (int r = 0; r < 2; ++r) {                                   // b_r = (keys 2t, 2t+1 of the r-th 8-key half; n = g = dim)
        const int ke = k0 + mt * 16 + r * 8 + 2 * t;
        uint32_t lo = 0u, hi = 0u;
        if (g < D) {
          if (ke < N) lo = *reinterpret_cast<const uint16_t*>(qkv + (tok_base + (int64_t)ke * tok_stride) * p.ldq + p.C + slab * CS + cbase + g);
          if (ke + 1 < N) hi = *reinterpret_cast<const uint16_t*>(qkv + (tok_base + (int64_t)(ke + 1) * tok_stride) * p.ldq + p.C + slab * CS + cbase + g);
        }
        kq[mt][r] = lo | (hi << 16);
      }
      dk[mt][0] = dk[mt][1] = dk[mt][2] = dk[mt][3] = 0.f;
      dv[mt][0] = dv[mt][1] = dv[mt][2] = dv[mt][3] = 0.f;
    }

    // rows of an m-tile that lie beyond the sequence (last key block only): their P must be zero - K = 0 gives s = 0, and
    // exp2(-lse) of a very negative lse would reach dQ as inf * 0
    const bool tail_block = k0 + FB_KB > N;
    auto tile_body = [&](auto masked, int tt) {
      constexpr bool MASK = decltype(masked)::value;
      b_cp_wait<1>();
      __syncthreads();
      stage_issue();
      const int b = cslot;
      cslot = cslot == 2 ? 0 : cslot + 1;
      // my four query n-tiles of this tile: Q and dO as "n = query" operands; the two 16-query blocks as "k = query" operands
      uint32_t qb[4], gb[4], qt[2][2], gt[2][2];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int nt = 4 * qh + j;
        qb[j] = 2 * t < D ? *reinterpret_cast<const uint32_t*>(&Qs[b][nt * 8 + g][cbase + 2 * t]) : 0u;
        gb[j] = 2 * t < D ? *reinterpret_cast<const uint32_t*>(&Gs[b][nt * 8 + g][cbase + 2 * t]) : 0u;
        if (D == 4 && t == 2) gb[j] = b_split2(-Ds[b][head_l][nt * 8 + g]);
      }
#pragma unroll
      for (int kl = 0; kl < 2; ++kl) {
        const int kk = 2 * qh + kl;
        b_ldsm_x2_trans(qt[kl][0], qt[kl][1], &Qs[b][kk * 16 + (lane & 15)][D >= 8 ? cbase : (cbase & ~7)]);
        b_ldsm_x2_trans(gt[kl][0], gt[kl][1], &Gs[b][kk * 16 + (lane & 15)][D >= 8 ? cbase : (cbase & ~7)]);
      }
      float dqa[2][4];
#pragma unroll
      for (int kl = 0; kl < 2; ++kl) dqa[kl][0] = dqa[kl][1] = dqa[kl][2] = dqa[kl][3] = 0.f;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
        for (int kl = 0; kl < 2; ++kl) {
          uint32_t pa[4], da[4];
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            const int j = 2 * kl + half, nt = 4 * qh + j;
            float s[4], dp[4];
            b_mma1688_z(s, ka[mt][0][0], ka[mt][0][1], qb[j]);            // S^T: rows = my keys, columns = queries
            b_mma1688_z(dp, va[mt][0][0], va[mt][0][1], gb[j]);           // dP^T - D
            const float2 l2 = *reinterpret_cast<const float2*>(&Ls[b][head_l][nt * 8 + 2 * t]);
            float2 d2 = make_float2(0.f, 0.f);
            if (D != 4) d2 = *reinterpret_cast<const float2*>(&Ds[b][head_l][nt * 8 + 2 * t]);   // head_dim 4 folds D_i into the MMA
            float pw[4], ds[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              pw[i] = fast_exp2(fmaf(s[i], c, -((i & 1) ? l2.y : l2.x)));
              if (MASK && k0 + mt * 16 + g + (i >> 1) * 8 >= N) pw[i] = 0.f;
              ds[i] = D == 4 ? pw[i] * dp[i] : pw[i] * (dp[i] - ((i & 1) ? d2.y : d2.x));
            }
            pa[half * 2 + 0] = b_pack(pw[0], pw[1]); pa[half * 2 + 1] = b_pack(pw[2], pw[3]);
            da[half * 2 + 0] = b_pack(ds[0], ds[1]); da[half * 2 + 1] = b_pack(ds[2], ds[3]);
          }
          b_mma16816(dv[mt], pa, gt[kl][0], gt[kl][1]);
          b_mma16816(dk[mt], da, qt[kl][0], qt[kl][1]);
          // dS (rows = queries) = the four 8x8 blocks of dS^T transposed: (q 0-7 | k 0-7), (q 8-15 | k 0-7), (q 0-7 | k 8-15), (q 8-15 | k 8-15)
          uint32_t ta[4];
          ta[0] = b_movm_trans(da[0]); ta[1] = b_movm_trans(da[2]); ta[2] = b_movm_trans(da[1]); ta[3] = b_movm_trans(da[3]);
          b_mma16816(dqa[kl], ta, kq[mt][0], kq[mt][1]);
        }
      }
      if (2 * t < D) {                                                // head_dim 4: my head's dims are columns 0..3 of the 8-wide tile
#pragma unroll
        for (int kl = 0; kl < 2; ++kl) {
          float2* r0 = reinterpret_cast<float2*>(dq_mine + (tt * BW_T + kl * 16) * FB_QP);
          float2* r1 = r0 + 8 * FB_QP / 2;
          float2 v0 = *r0, v1 = *r1;
          v0.x += dqa[kl][0]; v0.y += dqa[kl][1]; v1.x += dqa[kl][2]; v1.y += dqa[kl][3];
          *r0 = v0; *r1 = v1;
        }
      }
      stage_finish();
    };
    if (tail_block) { for (int tt = 0; tt < T; ++tt) tile_body(std::true_type{}, tt); }
    else { for (int tt = 0; tt < T; ++tt) tile_body(std::false_type{}, tt); }
    // ---- dK, dV of this key block: sum the two query halves, write ----
    float* myred = red + ((size_t)(head_l * 2) * 32 + lane) * 8;
    if (qh == 1) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        float* r = myred + (size_t)mt * 32 * 8;
        *reinterpret_cast<float4*>(r) = make_float4(dk[mt][0], dk[mt][1], dk[mt][2], dk[mt][3]);
        *reinterpret_cast<float4*>(r + 4) = make_float4(dv[mt][0], dv[mt][1], dv[mt][2], dv[mt][3]);
      }
    }
    __syncthreads();
    if (qh == 0 && (D >= 8 || (2 * t >= first && 2 * t < first + 4))) {
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const float* r = myred + (size_t)mt * 32 * 8;
        const float4 a = *reinterpret_cast<const float4*>(r), bq = *reinterpret_cast<const float4*>(r + 4);
        const float k4[4] = {dk[mt][0] + a.x, dk[mt][1] + a.y, dk[mt][2] + a.z, dk[mt][3] + a.w};
        const float v4[4] = {dv[mt][0] + bq.x, dv[mt][1] + bq.y, dv[mt][2] + bq.z, dv[mt][3] + bq.w};
#pragma unroll
        for (int h2 = 0; h2 < 2; ++h2) {
          const int kn = k0 + mt * 16 + g + h2 * 8;
          if (kn >= N) continue;
          bf16* base = dqkv + (tok_base + (int64_t)kn * tok_stride) * p.ldq + slab * CS + (D >= 8 ? cbase : (cbase & ~7)) + 2 * t;
          *reinterpret_cast<uint32_t*>(base + p.C) = b_pack(k4[h2 * 2] * p.scale, k4[h2 * 2 + 1] * p.scale);
          *reinterpret_cast<uint32_t*>(base + 2 * p.C) = b_pack(v4[h2 * 2], v4[h2 * 2 + 1]);
        }
      }
    }
  }
  b_cp_wait<0>();
  __syncthreads();
  // ---- dq = scale * accumulator ----
  for (int i = tid; i < N * CPT; i += FB_THREADS) {
    const int q = i / CPT, chunk = i % CPT;
    const float4 a = *reinterpret_cast<const float4*>(&dq_s[(size_t)q * FB_QP + chunk * 8]);
    const float4 bq = *reinterpret_cast<const float4*>(&dq_s[(size_t)q * FB_QP + chunk * 8 + 4]);
    uint4 o;
    o.x = b_pack(a.x * p.scale, a.y * p.scale); o.y = b_pack(a.z * p.scale, a.w * p.scale);
    o.z = b_pack(bq.x * p.scale, bq.y * p.scale); o.w = b_pack(bq.z * p.scale, bq.w * p.scale);
    *reinterpret_cast<uint4*>(dqkv + (tok_base + (int64_t)q * tok_stride) * p.ldq + slab * CS + chunk * 8) = o;
  }
}

template <int D>
static int launch_bwd_fused(const AttnParams& p, cudaStream_t st) {
  const int N = p.geom == TFSWA_GEOM_TSA ? p.H : p.W;
  const int rows = p.geom == TFSWA_GEOM_TSA ? p.B * p.W : p.B * p.H;
  const int smem = fb_smem_bytes<D>(N);
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    if (cudaFuncSetAttribute(attn_bwd_fused_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
      set_error("attn_bwd_fused: cudaFuncSetAttribute failed");
      return TFSWA_ECUDA;
    }
    attr_once.done();
  }
  const int64_t M = (int64_t)p.B * p.H * p.W;
  const int64_t work = M * (p.C / 8);
  const unsigned dgrid = (unsigned)((work + 255) / 256 < 148 * 16 ? (work + 255) / 256 : 148 * 16);
  attn_bwd_delta_kernel<D><<<dgrid, 256, 0, st>>>((const bf16*)p.dout, (const bf16*)p.o, p.ldo, p.dsum, M, p.C, p.heads);
  attn_bwd_fused_kernel<D><<<dim3(rows, p.heads / 8), FB_THREADS, smem, st>>>(p);
  return check_launch("attn_bwd_fused");
}

template <int D, bool WIN>
static int launch_bwd_mma(const AttnParams& p, cudaStream_t st) {
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    cudaError_t e1 = cudaFuncSetAttribute(attn_bwd_dq_mma_kernel<D, WIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, bw_smem_bytes<D>());
    cudaError_t e2 = cudaFuncSetAttribute(attn_bwd_dkv_mma_kernel<D, WIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, bw_smem_bytes<D>());
    if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("attn_bwd_mma: cudaFuncSetAttribute failed"); return TFSWA_ECUDA; }
    attr_once.done();
  }
  const int per_cta = 16 * bw_mt<D>();
  dim3 grid;
  if (WIN) {
    grid = dim3((unsigned)(p.B * p.nWh * p.nWw), (p.ws * p.ws + per_cta - 1) / per_cta, p.heads / 8);
  } else {
    const int N = p.geom == TFSWA_GEOM_TSA ? p.H : p.W;
    const int rows = p.geom == TFSWA_GEOM_TSA ? p.B * p.W : p.B * p.H;
    grid = dim3((N + per_cta - 1) / per_cta, rows, p.heads / 8);
  }
  attn_bwd_dq_mma_kernel<D, WIN><<<grid, BW_THREADS, bw_smem_bytes<D>(), st>>>(p);
  attn_bwd_dkv_mma_kernel<D, WIN><<<grid, BW_THREADS, bw_smem_bytes<D>(), st>>>(p);
  return check_launch("attn_bwd_mma");
}

// bf16 attention backward; returns 1 when the shape is not covered (caller falls back to the CUDA-core kernels)
int attn_bwd_mma_bf16(const AttnParams& p, cudaStream_t st) {
  const int D = p.C / p.heads;
  const bool win = p.geom == TFSWA_GEOM_SWA;
  const int rows = p.geom == TFSWA_GEOM_TSA ? p.B * p.W : p.B * p.H;
  if (p.heads % 8 != 0 || (D != 4 && D != 8 && D != 16)) return 1;
  if (win ? (p.ws != 8) : (rows > 65535)) return 1;
  if ((p.ldq % 8) || (p.ldo % 8) || (((uintptr_t)p.qkv | (uintptr_t)p.dqkv | (uintptr_t)p.dout) & 15) || (((uintptr_t)p.o) & 3)) return 1;
  // head_dim 4, axial: one-pass kernel (P computed once) when the sequence's fp32 dQ accumulator fits in shared memory
  const char* fe = getenv("TFSWA_ATTN_BWD_FUSED");               // "0": keep the dq + dkv pair (A/B, tests)
  const bool fused = !(fe && fe[0] == '0');
  const int Nseq = p.geom == TFSWA_GEOM_TSA ? p.H : p.W;
  if (!win && fused && (((uintptr_t)p.o) & 15) == 0) {
    if (D == 4 && fb_smem_bytes<4>(Nseq) <= 227 * 1024) return launch_bwd_fused<4>(p, st);
    if (D == 8 && fb_smem_bytes<8>(Nseq) <= 227 * 1024) return launch_bwd_fused<8>(p, st);
  }
  if (win) {
    if (D == 4) return launch_bwd_mma<4, true>(p, st);
    if (D == 8) return launch_bwd_mma<8, true>(p, st);
    return launch_bwd_mma<16, true>(p, st);
  }
  if (D == 4) return launch_bwd_mma<4, false>(p, st);
  if (D == 8) return launch_bwd_mma<8, false>(p, st);
  return launch_bwd_mma<16, false>(p, st);
}

}  // namespace tfswa
