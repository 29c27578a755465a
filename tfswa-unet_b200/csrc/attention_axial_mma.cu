// Axial (TSA / FSA) attention with register-resident scores: warp-level mma.sync.m16n8k16 in the FlashAttention-2
// register layout, the softmax shift fixed before the key loop by the same per-sequence bound as the tcgen05 kernel
// (tc_attention.cu), so there is no running maximum and no rescaling of O.
//
// At head_dim 4..16 this op is bound by exponentials (MUFU.EX2 at 16/clk/SM plus the polynomial share on the FMA pipe),
// not by MMA rate: either tensor path idles > 90 % of the time.  What differs is everything around the exponentials.
// Here S, P and O never leave the register file: S = Q K^T lands in accumulator registers, P is re-packed in place as
// the A operand of the PV MMA, the row sum comes from one more MMA against a ones operand - no TMEM load/store round
// trip, no mbarrier hand-offs between softmax and issuer warps, warps only meet at one __syncthreads per 64-key tile
// (K|V tiles arrive by cp.async into a 3-deep ring).
//
// One CTA = one sequence x QB = 16*MT queries x 8 heads (warp w = head w of the slab): K|V of a 64-key tile are
// loaded once per CTA for all 8 heads; each warp keeps its head's K / V^T fragments in registers for the MT row tiles.
#include "attn_common.cuh"
#include <type_traits>

namespace tfswa {

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// S tiles start from zero: separate outputs + a literal-zero C operand (RZ), so no accumulator has to be cleared and
// the A fragment is not copied into the destination registers
__device__ __forceinline__ void mma16816_z(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1), "f"(0.f));
}
// head_dim <= 8: the whole head fits one K = 8 step (A = 2 registers, B = 1)
__device__ __forceinline__ void mma1688_z(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%7,%7,%7,%7};"
               : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0), "f"(0.f));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ uint32_t pack_bf16_2(float lo, float hi) {
  uint32_t y;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo));
  return y;
}
__device__ __forceinline__ float ex2_poly_a(float x) {          // 2^x, x <= 0, FMA pipe (7.5e-5 relative)
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;
  const float r = x - (t - 12582912.0f);
  float p = fmaf(0.0551716685f, r, 0.2426111251f);
  p = fmaf(p, r, 0.6932609677f);
  p = fmaf(p, r, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

#ifndef TFSWA_AX_POLY_EVERY
#define TFSWA_AX_POLY_EVERY 8
#endif
constexpr int AX_POLY_EVERY = TFSWA_AX_POLY_EVERY;
constexpr int AX_KT = 64;                // keys per shared-memory tile
constexpr int AX_THREADS = 256;          // 8 warps = 8 heads

template <int D> __host__ __device__ constexpr int ax_smem_bytes() { return 3 * 2 * AX_KT * (8 * D + 8) * 2; }

template <int D, int MT, int MINB>
__global__ void __launch_bounds__(AX_THREADS, MINB) attn_axial_mma_kernel(const AttnParams p) {
  constexpr int CS = 8 * D;                           // channels of the 8-head slab
  constexpr int PITCH = CS + 8;                       // padded row: conflict-free fragment loads / ldmatrix
  constexpr int KS = (D + 15) / 16, DN = (D + 7) / 8;
  constexpr int QB = 16 * MT;
  constexpr int CPT = CS / 8;                         // 16-byte chunks per key and part
  extern __shared__ __align__(16) uint8_t ax_smem[];
  typedef bf16 (*tile_t)[AX_KT][PITCH];
  tile_t Ks = reinterpret_cast<tile_t>(ax_smem);                                   // [3][64][PITCH]
  tile_t Vs = reinterpret_cast<tile_t>(ax_smem + 3 * AX_KT * PITCH * 2);
  __shared__ int s_bad;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int row = blockIdx.y, q0 = p.q_begin + blockIdx.x * QB, slab = blockIdx.z;
  const int N = p.geom == TFSWA_GEOM_TSA ? p.H : p.W;
  const int q_end = p.q_end ? p.q_end : N;
  const int T = (N + AX_KT - 1) / AX_KT;
  int64_t tok_base, tok_stride;
  if (p.geom == TFSWA_GEOM_TSA) { const int b = row / p.W; tok_base = (int64_t)b * p.H * p.W + (row - b * p.W); tok_stride = p.W; }
  else { tok_base = (int64_t)row * p.W; tok_stride = 1; }
  const bf16* qkv = (const bf16*)p.qkv;
  const int cbase = warp * D;                         // my head's first channel inside the slab
  const int head = slab * 8 + warp;
  const float c = p.qscale;
  const int mt_valid = min(MT, (q_end - q0 + 15) / 16);

  auto stage = [&](int tt, int b) {                   // K|V rows of key tile tt -> ring slot b (absent keys: zeros)
    if (tt < T) {
      for (int v = tid; v < AX_KT * 2 * CPT; v += AX_THREADS) {
        const int j = v / (2 * CPT), rem = v - j * 2 * CPT, part = rem / CPT, chunk = rem - part * CPT;
        const int key = tt * AX_KT + j;
        bf16* dst = part ? &Vs[b][j][chunk * 8] : &Ks[b][j][chunk * 8];
        if (key < N) cp_async16(dst, qkv + (tok_base + (int64_t)key * tok_stride) * p.ldq + (1 + part) * p.C + slab * CS + chunk * 8);
        else *reinterpret_cast<uint4*>(dst) = make_uint4(0, 0, 0, 0);
      }
    }
    cp_async_commit();
  };
  stage(0, 0);
  stage(1, 1);

  // ---- Q fragments (registers, once) and the row-max bound m_i = sum_d max(q_d kmax_d, q_d kmin_d) ----
  uint32_t qa[MT][KS][4];
  float mrow[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    float bound[2] = {0.f, 0.f};
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      const int qn = q0 + mt * 16 + g + h2 * 8;
      const bool ok = qn < q_end;
      const bf16* qp = qkv + (tok_base + (int64_t)(ok ? qn : 0) * tok_stride) * p.ldq + slab * CS + cbase;
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
        for (int hi = 0; hi < 2; ++hi) {
          const int dcol = ks * 16 + hi * 8 + 2 * t;  // my two dims of this fragment register
          uint32_t v = 0u;
          if (dcol < D && ok) v = *reinterpret_cast<const uint32_t*>(qp + dcol);
          qa[mt][ks][hi * 2 + h2] = v;
          if (dcol < D) {
            const float2 kmin = *reinterpret_cast<const float2*>(p.kext + ((int64_t)row * 2 + 0) * p.C + slab * CS + cbase + dcol);
            const float2 kmax = *reinterpret_cast<const float2*>(p.kext + ((int64_t)row * 2 + 1) * p.C + slab * CS + cbase + dcol);
            const float qx = __uint_as_float(v << 16), qy = __uint_as_float(v & 0xFFFF0000u);
            bound[h2] += fmaxf(qx * kmax.x, qx * kmin.x) + fmaxf(qy * kmax.y, qy * kmin.y);
          }
        }
      }
    }
#pragma unroll
    for (int h2 = 0; h2 < 2; ++h2) {
      float b = bound[h2];
      b += __shfl_xor_sync(0xffffffffu, b, 1);
      b += __shfl_xor_sync(0xffffffffu, b, 2);
      mrow[mt][h2] = b;
    }
  }

  float o[MT][DN][4], l[MT][4];
  constexpr uint32_t ONES = 0x3F803F80u;

  bool first_sweep = true;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const bool exact = attempt == 1 || p.force_exact;
    // pass 0 of the exact path only computes the true row maxima; the last pass accumulates O and l
    for (int pass = exact ? 0 : 1; pass < 2; ++pass) {
      const bool max_pass = pass == 0;
      if (max_pass) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) mrow[mt][0] = mrow[mt][1] = -CUDART_INF_F;
      } else {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int dn = 0; dn < DN; ++dn) o[mt][dn][0] = o[mt][dn][1] = o[mt][dn][2] = o[mt][dn][3] = 0.f;
          l[mt][0] = l[mt][1] = l[mt][2] = l[mt][3] = 0.f;
        }
      }
      if (!first_sweep) {                             // another sweep over the keys: restart the ring
        __syncthreads();
        stage(0, 0);
        stage(1, 1);
      }
      first_sweep = false;
      for (int tt = 0; tt < T; ++tt) {
        cp_async_wait<1>();                           // tile tt landed (one newer group may be in flight)
        __syncthreads();                              // ... for every thread; and everyone is done with tile tt-1
        stage(tt + 2, (tt + 2) % 3);
        const int b = tt % 3;
        const int kcount = min(AX_KT, N - tt * AX_KT);
        auto tile = [&](auto tail_c) {
          constexpr bool tail = decltype(tail_c)::value;
          const int ntl = tail ? (kcount + 7) >> 3 : 8;          // key n-tiles holding at least one present key
          // my head's K fragments (8 key n-tiles) and V^T fragments (4 key steps) of this tile
          uint32_t kb[8][KS][2], vb[4][DN][2];
  #pragma unroll
          for (int nt = 0; nt < 8; ++nt) {
  #pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
              const int dcol = ks * 16 + 2 * t;
              kb[nt][ks][0] = dcol < D ? *reinterpret_cast<const uint32_t*>(&Ks[b][nt * 8 + g][cbase + dcol]) : 0u;
              kb[nt][ks][1] = dcol + 8 < D ? *reinterpret_cast<const uint32_t*>(&Ks[b][nt * 8 + g][cbase + dcol + 8]) : 0u;
            }
          }
          if (!max_pass) {
  #pragma unroll
            for (int kk = 0; kk < 4; ++kk) {
  #pragma unroll
              for (int dn = 0; dn < DN; ++dn) {
                const int vcol = D >= 8 ? cbase + dn * 8 : (cbase & ~7);
                ldsm_x2_trans(vb[kk][dn][0], vb[kk][dn][1], &Vs[b][kk * 16 + (lane & 15)][vcol]);
              }
            }
          }
  #pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            if (mt < mt_valid) {
              float s[8][4];
  #pragma unroll
              for (int nt = 0; nt < 8; ++nt) {
                if (tail && nt >= ntl) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f; continue; }
                if (D <= 8) {
                  mma1688_z(s[nt], qa[mt][0][0], qa[mt][0][1], kb[nt][0][0]);
                } else {
                  mma16816_z(s[nt], qa[mt][0], kb[nt][0][0], kb[nt][0][1]);
  #pragma unroll
                  for (int ks = 1; ks < KS; ++ks) mma16816(s[nt], qa[mt][ks], kb[nt][ks][0], kb[nt][ks][1]);
                }
              }
              if (max_pass) {
                float m0 = mrow[mt][0], m1 = mrow[mt][1];
  #pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
  #pragma unroll
                  for (int i = 0; i < 2; ++i) {
                    const bool in = !tail || nt * 8 + 2 * t + i < kcount;
                    if (in) { m0 = fmaxf(m0, s[nt][i]); m1 = fmaxf(m1, s[nt][2 + i]); }
                  }
                }
                mrow[mt][0] = m0; mrow[mt][1] = m1;
              } else {
                const float mc0 = mrow[mt][0] * c, mc1 = mrow[mt][1] * c;
  #pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                  if (tail && 2 * kk >= ntl) break;               // no present key in this 16-key step
                  uint32_t pa[4];
  #pragma unroll
                  for (int half = 0; half < 2; ++half) {
                    const int nt = 2 * kk + half;
                    float e[4];
  #pragma unroll
                    for (int i = 0; i < 4; ++i) {
                      const float x = fmaf(s[nt][i], c, i < 2 ? -mc0 : -mc1);
                      e[i] = (AX_POLY_EVERY > 0 && ((nt * 4 + i) % AX_POLY_EVERY) == AX_POLY_EVERY - 1) ? ex2_poly_a(x) : fast_exp2(x);
                      if (tail && nt * 8 + 2 * t + (i & 1) >= kcount) e[i] = 0.f;      // absent keys
                    }
                    pa[half * 2 + 0] = pack_bf16_2(e[0], e[1]);
                    pa[half * 2 + 1] = pack_bf16_2(e[2], e[3]);
                  }
  #pragma unroll
                  for (int dn = 0; dn < DN; ++dn) mma16816(o[mt][dn], pa, vb[kk][dn][0], vb[kk][dn][1]);
                  mma16816(l[mt], pa, ONES, ONES);
                }
              }
            }
          }
              };
        if (kcount < AX_KT) tile(std::true_type{}); else tile(std::false_type{});
      }
      cp_async_wait<0>();
      if (max_pass) {                                 // the 4 threads of a quad each saw a quarter of the keys
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            float m = mrow[mt][h2];
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
            m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
            mrow[mt][h2] = m;
          }
        }
      }
    }
    // the bound keeps every exponent <= 0; a bound so loose that a whole row underflowed (l ~ 0) sends the CTA through
    // the exact path once
    bool bad = false;
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      if (mt < mt_valid) {
        if (q0 + mt * 16 + g < q_end && !(l[mt][0] > 1e-30f)) bad = true;
        if (q0 + mt * 16 + g + 8 < q_end && !(l[mt][2] > 1e-30f)) bad = true;
      }
    }
    if (exact) break;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    if (bad) s_bad = 1;
    __syncthreads();
    if (!s_bad) break;
  }

  // ---- O / l -> out ----
  bf16* out = (bf16*)p.out;
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    if (mt < mt_valid) {
#pragma unroll
      for (int h2 = 0; h2 < 2; ++h2) {
        const int qn = q0 + mt * 16 + g + h2 * 8;
        if (qn >= q_end) continue;
        const int64_t tok = tok_base + (int64_t)qn * tok_stride;
        const float lsum = l[mt][h2 * 2];
        const float inv = 1.0f / lsum;
        if (D >= 8) {
#pragma unroll
          for (int dn = 0; dn < DN; ++dn)
            *reinterpret_cast<uint32_t*>(out + tok * p.ldo + slab * CS + cbase + dn * 8 + 2 * t) =
                pack_bf16_2(o[mt][dn][h2 * 2] * inv, o[mt][dn][h2 * 2 + 1] * inv);
        } else {
          const int first = cbase & 7;                // d = 4: my head's dims inside the aligned 8-dim V block
          if (2 * t >= first && 2 * t < first + 4)
            *reinterpret_cast<uint32_t*>(out + tok * p.ldo + slab * CS + (cbase & ~7) + 2 * t) =
                pack_bf16_2(o[mt][0][h2 * 2] * inv, o[mt][0][h2 * 2 + 1] * inv);
        }
        if (p.lse && t == 0) p.lse[tok * p.heads + head] = mrow[mt][h2] * c + log2f(lsum);
      }
    }
  }
}

template <int D, int MT, int MINB>
static int launch_axial(const AttnParams& p, int q_count, int rows, cudaStream_t st) {
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    cudaError_t e = cudaFuncSetAttribute(attn_axial_mma_kernel<D, MT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, ax_smem_bytes<D>());
    if (e != cudaSuccess) { set_error("attn_axial_mma: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return TFSWA_ECUDA; }
    attr_once.done();
  }
  dim3 grid((q_count + 16 * MT - 1) / (16 * MT), rows, p.heads / 8);
  attn_axial_mma_kernel<D, MT, MINB><<<grid, AX_THREADS, ax_smem_bytes<D>(), st>>>(p);
  return check_launch("attn_axial_mma");
}

// queries [p.q_begin, p.q_end) of every sequence (q_end = 0: to the end); p.kext must hold the per-sequence k extrema
int attn_axial_mma_bf16(const AttnParams& p, cudaStream_t st) {
  const int D = p.C / p.heads;
  const int N = p.geom == TFSWA_GEOM_TSA ? p.H : p.W;
  const int rows = p.geom == TFSWA_GEOM_TSA ? p.B * p.W : p.B * p.H;
  const int q_count = (p.q_end ? p.q_end : N) - p.q_begin;
  if (rows > 65535 || p.heads % 8 != 0) { set_error("attn_axial_mma: rows=%d heads=%d unsupported", rows, p.heads); return TFSWA_EINVAL; }
#ifdef TFSWA_AX_MT2
  if (D == 4) return launch_axial<4, 2, 3>(p, q_count, rows, st);
  if (D == 8) return launch_axial<8, 2, 3>(p, q_count, rows, st);
#endif
  if (D == 4) return launch_axial<4, 4, 2>(p, q_count, rows, st);
  if (D == 8) return launch_axial<8, 4, 2>(p, q_count, rows, st);
  if (D == 16) return launch_axial<16, 2, 2>(p, q_count, rows, st);
  set_error("attn_axial_mma: head_dim %d unsupported", D);
  return TFSWA_EINVAL;
}

}  // namespace tfswa
