// Axial (TSA / FSA) attention on the tcgen05 tensor cores for the small head dims of stages 1-2 (d = 4, 8).
//
// At d = 4 a bf16 MMA (K = 16) can hold 4 heads x 4 dims along K.  One CTA owns 128 queries x one "head quad"
// (16 channels) of one sequence:
//   S  : A = Q tile (128 x 16, the 16 channels as they lie in memory), B = "expanded" K: row (h', j) carries
//        k_j of head h' in K-slots [d*h', d*h'+d) and zeros elsewhere, so D[q, (h', j)] = q_h' . k_{j,h'} -
//        one MMA (N = 4*KT) yields the scores of all 4 heads; TMEM holds S (fp32).
//   P  : softmax threads (one query row each, two threads per row splitting the columns) read S with tcgen05.ld,
//        p = ex2(s*c - m*c) with ex2.approx.ftz.bf16x2 (two exps per MUFU op, P is bf16 for the PV MMA anyway) and
//        write P as the K-major A operand of the PV MMA.
//   O  : per head, D[q, 0..15] += P_h (128 x KT) * V'_h (KT x 16) where V'_h = [v dims | 1 | 0...]: the ones column
//        makes the tensor core accumulate the softmax denominator (from the same bf16-rounded P as the numerator).
// Exact maximum, no online rescaling: pass 1 streams S tiles (256 columns) and keeps the row max per head, pass 2
// recomputes S (128 columns per tile) and accumulates O in TMEM across ALL key tiles; O is read once at the end.
// Key tails are masked by zeroing V' (incl. its ones column) - the exponentials of absent keys multiply zeros.
// Operands are staged by the threads themselves (8-row x 16-byte core matrices, no swizzle) because the head
// expansion / zero padding is not expressible as a TMA box.  Two CTAs per SM (96 KB smem, 256 TMEM columns each)
// hide each other's MMA/barrier latencies.
//
// Replaces attention.py:70-85 (+ permutes :143,:162,:217,:236) for head_dim 4 and 8, bf16 activations.
#include "attn_common.cuh"
#include "sm100.cuh"

namespace tfswa {

using namespace sm100;

constexpr int TA_THREADS = 256;
constexpr int TA_QT = 128;                 // queries per CTA
constexpr uint32_t TA_TMEM_COLS = 256;
constexpr uint32_t TA_O_COL = 128;         // O accumulators live in columns [128, 128+64)

constexpr int TA_QS = 0;                   // Q operand, 128 rows x 32 B
constexpr int TA_KS = 4096;                // 2 x 8 KB expanded-K operand (pass 1: 256 rows, pass 2: 128 rows)
constexpr int TA_VS = TA_KS + 2 * 8192;    // 3 x 4 KB V' operands
constexpr int TA_PS = TA_VS + 3 * 4096;    // 2 x 32 KB P operands
constexpr int TA_SMEM = TA_PS + 2 * 32768; // 98304

__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) {
  uint32_t y;
  asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  uint32_t y;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo));
  return y;
}

// Stage the expanded-K operand of `kt` keys starting at key k0 (threads [0, kt)).
template <int D>
__device__ __forceinline__ void build_k(const AttnParams& p, uint8_t* ks, int row, int k0, int kt, int N, int quad, int tid) {
  constexpr int HPQ = 16 / D;
  if (tid >= kt) return;
  const int j = tid;
  uint4 a = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0);
  if (k0 + j < N) {
    bool valid; const int64_t tok = token_of<false>(p, row, k0 + j, valid);
    const uint4* src = reinterpret_cast<const uint4*>((const bf16*)p.qkv + tok * p.ldq + p.C + quad * 16);
    a = src[0]; b = src[1];
  }
  if (D == 4) {
    const uint2 parts[4] = {make_uint2(a.x, a.y), make_uint2(a.z, a.w), make_uint2(b.x, b.y), make_uint2(b.z, b.w)};
#pragma unroll
    for (int h = 0; h < HPQ; ++h) {
      const int n = h * kt + j;
      *reinterpret_cast<uint2*>(ks + (n >> 3) * 256 + (h >> 1) * 128 + (n & 7) * 16 + (h & 1) * 8) = parts[h];
    }
  } else {
#pragma unroll
    for (int h = 0; h < HPQ; ++h) {
      const int n = h * kt + j;
      *reinterpret_cast<uint4*>(ks + (n >> 3) * 256 + h * 128 + (n & 7) * 16) = h ? b : a;
    }
  }
}

// Stage V'_h (16 x kt, K-major) for all heads of the quad (threads [kt, 2kt)); absent keys become all-zero columns.
template <int D>
__device__ __forceinline__ void build_v(const AttnParams& p, uint8_t* vs, int row, int k0, int kt, int N, int quad, int tid) {
  constexpr int HPQ = 16 / D;
  if (tid < kt || tid >= 2 * kt) return;
  const int j = tid - kt;
  const bool present = k0 + j < N;
  uint4 raw[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
  if (present) {
    bool valid; const int64_t tok = token_of<false>(p, row, k0 + j, valid);
    const uint4* src = reinterpret_cast<const uint4*>((const bf16*)p.qkv + tok * p.ldq + 2 * p.C + quad * 16);
    raw[0] = src[0]; raw[1] = src[1];
  }
  const uint16_t* e = reinterpret_cast<const uint16_t*>(raw);
  const int head_bytes = 16 * kt * 2, sbo = (kt / 8) * 128;
  const int koff = (j >> 3) * 128 + (j & 7) * 2;
#pragma unroll
  for (int h = 0; h < HPQ; ++h) {
    uint8_t* base = vs + h * head_bytes;
#pragma unroll
    for (int d = 0; d < D; ++d) *reinterpret_cast<uint16_t*>(base + koff + d * 16) = e[h * D + d];
    const uint16_t one = present ? (uint16_t)0x3F80 : (uint16_t)0;          // bf16 1.0: the denominator column
    if (D == 4) *reinterpret_cast<uint16_t*>(base + koff + 4 * 16) = one;
    else *reinterpret_cast<uint16_t*>(base + sbo + koff) = one;
  }
}

template <int D>
__global__ void __launch_bounds__(TA_THREADS, 2) tc_attn_axial_kernel(const AttnParams p) {
  constexpr int HPQ = 16 / D;            // heads per CTA
  constexpr int KT2 = 128 / HPQ;         // keys per pass-2 tile (S tile = 128 columns)
  constexpr int KT1 = 256 / HPQ;         // keys per pass-1 tile (S tile = 256 columns)
  constexpr int HPT = HPQ / 2;           // heads per thread
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_s, bar_pv[2];
  __shared__ uint32_t s_tmem;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int row = blockIdx.x, q0 = blockIdx.y * TA_QT, quad = blockIdx.z;
  const int N = p.geom == TFSWA_GEOM_TSA ? p.H : p.W;
  const int quarter = warp & 3, half = warp >> 2;
  const int r = quarter * 32 + lane;                 // my query row == my TMEM lane
  const float c = p.qscale;                          // head_dim^-0.5 * log2(e)

  // ---- setup: zero the operand buffers (their zero patterns are permanent), barriers, TMEM ----
  for (int i = tid; i < (TA_PS - TA_KS) / 16; i += TA_THREADS) reinterpret_cast<uint4*>(smem + TA_KS)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) {
    if (lane == 0) { mbar_init(&bar_s, 1); mbar_init(&bar_pv[0], 1); mbar_init(&bar_pv[1], 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc(&s_tmem, TA_TMEM_COLS);
  }
  // Q operand: row r, 16 channels = 2 chunks of 16 B
  bool q_valid = false;
  int64_t q_tok = 0;
  if (tid < TA_QT) {
    uint4 a = make_uint4(0, 0, 0, 0), b = a;
    if (q0 + r < N) {
      q_tok = token_of<false>(p, row, q0 + r, q_valid);
      const uint4* src = reinterpret_cast<const uint4*>((const bf16*)p.qkv + q_tok * p.ldq + quad * 16);
      a = src[0]; b = src[1];
    }
    *reinterpret_cast<uint4*>(smem + TA_QS + (r >> 3) * 256 + (r & 7) * 16) = a;
    *reinterpret_cast<uint4*>(smem + TA_QS + (r >> 3) * 256 + 128 + (r & 7) * 16) = b;
  } else if (q0 + r < N) {
    q_tok = token_of<false>(p, row, q0 + r, q_valid);
  }
  __syncthreads();                                   // zero fill done before any build_* writes
  const uint32_t sbase = smem_u32(smem);
  const uint64_t qdesc = umma_smem_desc_ns(sbase + TA_QS, 128, 256);
  uint32_t ns = 0;                                   // number of S commits consumed so far (parity = ns & 1)

  // =========================== pass 1: exact row maxima ===========================
  const int T1 = (N + KT1 - 1) / KT1;
  build_k<D>(p, smem + TA_KS, row, 0, KT1, N, quad, tid);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t my_taddr = tmem + ((uint32_t)(quarter * 32) << 16);
  if (tid == 0) {
    umma_bf16_ss(tmem, qdesc, umma_smem_desc_ns(sbase + TA_KS, 128, 256), umma_idesc_bf16(128, 256), 0u);
    umma_commit(&bar_s);
  }
  float m[HPT];
#pragma unroll
  for (int i = 0; i < HPT; ++i) m[i] = -CUDART_INF_F;
  for (int t = 0; t < T1; ++t) {
    if (t + 1 < T1) build_k<D>(p, smem + TA_KS + ((t + 1) & 1) * 8192, row, (t + 1) * KT1, KT1, N, quad, tid);
    mbar_wait(&bar_s, ns & 1); ++ns;
    tc_fence_after();
    const int kcount = min(KT1, N - t * KT1);
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {                 // my 128 of the 256 columns, 32 at a time
      uint32_t s[32];
      __syncwarp();
      tmem_ld_x32(my_taddr + half * 128 + ch * 32, s);
      tmem_ld_wait();
      const int lc = ch * 32;                        // local column -> (head slot, key)
      const int hi = lc / KT1, jbase = lc % KT1;
      float mx = m[hi];
      if (jbase + 32 <= kcount) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(s[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) if (jbase + i < kcount) mx = fmaxf(mx, __uint_as_float(s[i]));
      }
      m[hi] = mx;
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0 && t + 1 < T1) {
      tc_fence_after();
      umma_bf16_ss(tmem, qdesc, umma_smem_desc_ns(sbase + TA_KS + ((t + 1) & 1) * 8192, 128, 256), umma_idesc_bf16(128, 256), 0u);
      umma_commit(&bar_s);
    }
  }
  float mc[HPT];
#pragma unroll
  for (int i = 0; i < HPT; ++i) mc[i] = m[i] * c;

  // =========================== pass 2: P = ex2(S*c - m*c), O += P V' ===========================
  const int T2 = (N + KT2 - 1) / KT2;
  constexpr int P_SBO = (KT2 / 8) * 128;             // 8-row group stride of the P / V' operands
  constexpr int P_HEAD = 128 * KT2 * 2;              // bytes of one head's P tile
  constexpr int V_HEAD = 16 * KT2 * 2;
  // the row -> head mapping of the expanded-K operand changes with the tile size: clear pass 1's pattern
  for (int i = tid; i < (2 * 8192) / 16; i += TA_THREADS) reinterpret_cast<uint4*>(smem + TA_KS)[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  build_k<D>(p, smem + TA_KS, row, 0, KT2, N, quad, tid);
  build_v<D>(p, smem + TA_VS, row, 0, KT2, N, quad, tid);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  if (tid == 0) {
    tc_fence_after();
    umma_bf16_ss(tmem, qdesc, umma_smem_desc_ns(sbase + TA_KS, 128, 256), umma_idesc_bf16(128, 128), 0u);
    umma_commit(&bar_s);
  }
  for (int t = 0; t < T2; ++t) {
    const int pb = t & 1;
    if (t >= 2) mbar_wait(&bar_pv[pb], ((t >> 1) - 1) & 1);        // PV(t-2) has finished reading Ps[pb], Vs[(t+1)%3]
    if (t + 1 < T2) {
      build_k<D>(p, smem + TA_KS + ((t + 1) & 1) * 8192, row, (t + 1) * KT2, KT2, N, quad, tid);
      build_v<D>(p, smem + TA_VS + ((t + 1) % 3) * 4096, row, (t + 1) * KT2, KT2, N, quad, tid);
    }
    mbar_wait(&bar_s, ns & 1); ++ns;
    tc_fence_after();
    const bool tail = (t + 1) * KT2 > N;             // only the last tile holds absent keys
    uint8_t* ps = smem + TA_PS + pb * 32768;
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {                 // my 64 of the 128 columns
      uint32_t s[32];
      __syncwarp();
      tmem_ld_x32(my_taddr + half * 64 + ch * 32, s);
      tmem_ld_wait();
      const int col = half * 64 + ch * 32;
      const int head = col / KT2, jbase = col % KT2; // head within the quad, first key of this chunk
      const float mcc = mc[(ch * 32) / KT2];
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float x0 = fmaf(__uint_as_float(s[2 * i]), c, -mcc);
        float x1 = fmaf(__uint_as_float(s[2 * i + 1]), c, -mcc);
        if (tail) { x0 = fminf(x0, 0.f); x1 = fminf(x1, 0.f); }   // absent keys score 0, which may exceed the max
        pk[i] = ex2_bf16x2(pack_bf16x2(x0, x1));
      }
      uint8_t* dst = ps + head * P_HEAD + (r >> 3) * P_SBO + (r & 7) * 16 + (jbase >> 3) * 128;
#pragma unroll
      for (int q4 = 0; q4 < 4; ++q4)
        *reinterpret_cast<uint4*>(dst + q4 * 128) = make_uint4(pk[4 * q4], pk[4 * q4 + 1], pk[4 * q4 + 2], pk[4 * q4 + 3]);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc_pv = umma_idesc_bf16(128, 16);
      const uint32_t pbase = sbase + TA_PS + pb * 32768, vbase = sbase + TA_VS + (t % 3) * 4096;
#pragma unroll
      for (int h = 0; h < HPQ; ++h) {
#pragma unroll
        for (int kk = 0; kk < KT2 / 16; ++kk) {
          umma_bf16_ss(tmem + TA_O_COL + 16 * h, umma_smem_desc_ns(pbase + h * P_HEAD + kk * 256, 128, P_SBO),
                       umma_smem_desc_ns(vbase + h * V_HEAD + kk * 256, 128, P_SBO), idesc_pv, (t | kk) ? 1u : 0u);
        }
      }
      umma_commit(&bar_pv[pb]);
      if (t + 1 < T2) {
        umma_bf16_ss(tmem, qdesc, umma_smem_desc_ns(sbase + TA_KS + ((t + 1) & 1) * 8192, 128, 256), umma_idesc_bf16(128, 128), 0u);
        umma_commit(&bar_s);
      }
    }
  }
  // ---- epilogue: O / l ----
  {
    const int last = T2 - 1;
    mbar_wait(&bar_pv[last & 1], (last >> 1) & 1);   // commits are ordered: the last PV implies all earlier ones
    tc_fence_after();
    uint32_t o[16 * HPT];
    __syncwarp();
    if (HPT == 2) {
      uint32_t t32[32];
      tmem_ld_x32(my_taddr + TA_O_COL + 32 * half, t32);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o[i % (16 * HPT)] = t32[i];
    } else {
      uint32_t t16[16];
      tmem_ld_x16(my_taddr + TA_O_COL + 16 * half, t16);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 16; ++i) o[i] = t16[i];
    }
    if (q_valid) {
#pragma unroll
      for (int hh = 0; hh < HPT; ++hh) {
        const int head = half * HPT + hh;            // head within the quad
        const float l = __uint_as_float(o[hh * 16 + D]);
        const float inv = 1.0f / l;
        bf16* op = (bf16*)p.out + q_tok * p.ldo + quad * 16 + head * D;
        if (D == 4) {
          float v[4];
#pragma unroll
          for (int d = 0; d < 4; ++d) v[d] = __uint_as_float(o[hh * 16 + d]) * inv;
          store4(op, v);
        } else {
          float v[8];
#pragma unroll
          for (int d = 0; d < 8; ++d) v[d] = __uint_as_float(o[hh * 16 + d]) * inv;
          store8(op, v);
        }
        if (p.lse) p.lse[q_tok * p.heads + quad * HPQ + head] = mc[hh] + log2f(l);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TA_TMEM_COLS);
}

}  // namespace tfswa

using namespace tfswa;

extern "C" int tfswa_attn_tc_fwd(const tfswa_attn_args* a, void* stream) {
  TFSWA_REQUIRE(a && a->qkv && a->out, "attn_tc: null pointer");
  TFSWA_REQUIRE(a->dtype == TFSWA_BF16, "attn_tc: bf16 activations only");
  TFSWA_REQUIRE(a->geom == TFSWA_GEOM_TSA || a->geom == TFSWA_GEOM_FSA, "attn_tc: axial geometries only (windows use tfswa_attn_fwd)");
  TFSWA_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0 && a->heads > 0 && a->C % a->heads == 0 && a->C % 16 == 0, "attn_tc: bad shape");
  const int D = a->C / a->heads;
  TFSWA_REQUIRE(D == 4 || D == 8, "attn_tc: head_dim %d not in {4,8} (use tfswa_attn_fwd)", D);
  TFSWA_REQUIRE(a->ldq % 8 == 0 && a->ldo % 4 == 0 && (((uintptr_t)a->qkv) & 15) == 0, "attn_tc: alignment");
  AttnParams p = {};
  p.qkv = a->qkv; p.ldq = a->ldq; p.out = a->out; p.ldo = a->ldo; p.lse = a->lse;
  p.B = a->B; p.H = a->H; p.W = a->W; p.C = a->C; p.heads = a->heads; p.geom = a->geom;
  p.qscale = (float)(1.4426950408889634 / sqrt((double)D));
  const int N = a->geom == TFSWA_GEOM_TSA ? a->H : a->W;
  const int rows = a->geom == TFSWA_GEOM_TSA ? a->B * a->W : a->B * a->H;
  dim3 grid(rows, (N + TA_QT - 1) / TA_QT, a->C / 16);
  TFSWA_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "attn_tc: sequence too long");
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e1 = cudaFuncSetAttribute(tc_attn_axial_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM);
    cudaError_t e2 = cudaFuncSetAttribute(tc_attn_axial_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, TA_SMEM);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("attn_tc: cudaFuncSetAttribute failed"); return TFSWA_ECUDA; }
    attr_set = true;
  }
  if (D == 4) tc_attn_axial_kernel<4><<<grid, TA_THREADS, TA_SMEM, st>>>(p);
  else tc_attn_axial_kernel<8><<<grid, TA_THREADS, TA_SMEM, st>>>(p);
  return check_launch("attn_tc");
}
