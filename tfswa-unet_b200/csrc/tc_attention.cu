// Axial (TSA / FSA) attention on the tcgen05 tensor cores for the small head dims of stages 1-2 (d = 4, 8).
//
// At d = 4 a bf16 MMA (K = 16) can hold 4 heads x 4 dims along K.  One CTA owns 128 queries x one "head quad"
// (16 channels) of one sequence:
//   S  : A = Q tile (128 x 16, the 16 channels as they lie in memory), B = "expanded" K: row (h', j) carries
//        k_j of head h' in K-slots [d*h', d*h'+d) and zeros elsewhere, so D[q, (h', j)] = q_h' . k_{j,h'} -
//        one MMA (N = 4*KT) yields the scores of all 4 heads; TMEM holds S (fp32).
//   P  : softmax threads (one query row each, two threads per row splitting the columns) read S with tcgen05.ld,
//        p = ex2(s*c - m*c) (MUFU.EX2; the packed bf16x2 form was measured to issue as two scalar MUFU ops on
//        sm_100a, so fp32 ex2 + one pack is both cheaper and more accurate) and write bf16 P as the K-major A
//        operand of the PV MMA.
//   O  : per head, D[q, 0..15] += P_h (128 x KT) * V'_h (KT x 16) where V'_h = [v dims | 1 | 0...]: the ones column
//        makes the tensor core accumulate the softmax denominator (from the same bf16-rounded P as the numerator).
// No online rescaling: the softmax shift m_i is fixed BEFORE the key loop, so O accumulates in TMEM across all key
// tiles and is read once.  m_i is an upper bound of the row maximum, sum_d max(q_d kmax_d, q_d kmin_d) with the
// per-channel extrema of k over the sequence (softmax is shift invariant; an upper bound only costs dynamic range,
// of which bf16/fp32 exponents have ~2^126).  If a bound is ever so loose that a row's denominator underflows, the
// CTA repeats the computation with the exact maxima (one extra streaming pass over S, double-buffered in TMEM) -
// `force_exact` selects that path for testing.  Each thread pulls its 64 scores into registers first, so the S MMA
// of tile t+1 and the PV MMA of tile t-1 run underneath the exponentials of tile t.
// Key tails are masked by zeroing V' (incl. its ones column) - the exponentials of absent keys multiply zeros.
// Operands are staged by the threads themselves (8-row x 16-byte core matrices, no swizzle) because the head
// expansion / zero padding is not expressible as a TMA box.  Two CTAs per SM (96 KB smem, 256 TMEM columns each)
// hide each other's MMA/barrier latencies.
//
// Replaces attention.py:70-85 (+ permutes :143,:162,:217,:236) for head_dim 4 and 8, bf16 activations.
#include "attn_common.cuh"
#include "sm100.cuh"
#include <stdlib.h>

namespace tfswa {

using namespace sm100;

constexpr int TA_THREADS = 256;                // softmax threads (warps 0-7)
constexpr int TA_NTHREADS = TA_THREADS + 64;   // + issuer warp (8) + producer warp (9)
constexpr int TA_QT = 128;                 // queries per CTA
constexpr uint32_t TA_TMEM_COLS = 256;
constexpr uint32_t TA_O_COL = 128;         // O accumulators live in columns [128, 128+64)
constexpr uint32_t TA_P_COL = 192;         // bf16 P (A operand of the PV MMA): 64 columns = 128 keys x 2 per 32-bit cell

constexpr int TA_QS = 0;                   // Q operand, 128 rows x 32 B
constexpr int TA_KS = 4096;                // 2 x 8 KB expanded-K operand (pass 1: 256 rows, pass 2: 128 rows)
constexpr int TA_VS = TA_KS + 2 * 8192;    // 3 V' operand buffers of vs_bytes<D>() each, then 1 KB exchange scratch
template <int D> __host__ __device__ constexpr int vs_bytes() { return D == 16 ? 8192 : 4096; }
template <int D> __host__ __device__ constexpr int ps_off() { return TA_VS + 3 * vs_bytes<D>(); }
template <int D> __host__ __device__ constexpr int smem_bytes() { return ps_off<D>() + 1024; }        // 33 KB (d=4,8) / 45 KB (d=16)

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

// ---- operand staging, split into a global->register half and a register->shared half so the global latency of
// tile t+1 overlaps the softmax of tile t.  Threads [0, KT) carry one key's k (32 B); threads [128, 256) carry one
// (key, head) slice of v (2*D bytes). ----
struct KvRegs { uint4 a, b; bool present; };

// token(n) = tok_base + n * tok_stride for the axial geometries (no division in the key loop)
template <int D, int KT>
__device__ __forceinline__ KvRegs load_kv(const AttnParams& p, int64_t tok_base, int64_t tok_stride, int k0, int N, int quad, int tid) {
  KvRegs r; r.a = make_uint4(0, 0, 0, 0); r.b = r.a; r.present = false;
  if (tid < KT) {
    if (k0 + tid < N) {
      const int64_t tok = tok_base + (int64_t)(k0 + tid) * tok_stride;
      const uint4* src = reinterpret_cast<const uint4*>((const bf16*)p.qkv + tok * p.ldq + p.C + quad * 16);
      r.a = src[0]; r.b = src[1]; r.present = true;
    }
  } else if (tid >= 128) {
    const int idx = tid - 128, j = idx % KT, h = idx / KT;
    if (k0 + j < N) {
      const int64_t tok = tok_base + (int64_t)(k0 + j) * tok_stride;
      const bf16* src = (const bf16*)p.qkv + tok * p.ldq + 2 * p.C + quad * 16 + h * D;
      if (D == 4) { const uint2 t = *reinterpret_cast<const uint2*>(src); r.a.x = t.x; r.a.y = t.y; }
      else if (D == 8) r.a = *reinterpret_cast<const uint4*>(src);
      else { r.a = reinterpret_cast<const uint4*>(src)[0]; r.b = reinterpret_cast<const uint4*>(src)[1]; }
      r.present = true;
    }
  }
  return r;
}

// expanded-K operand: row n = h*KT + j holds k_j of head h in K-slots [D*h, D*h+D), zeros elsewhere (zeros are permanent)
template <int D, int KT>
__device__ __forceinline__ void store_k(const KvRegs& kv, uint8_t* ks, int tid) {
  constexpr int HPQ = 16 / D;
  if (tid >= KT) return;
  const int j = tid;
  if (D == 4) {
    const uint2 parts[4] = {make_uint2(kv.a.x, kv.a.y), make_uint2(kv.a.z, kv.a.w), make_uint2(kv.b.x, kv.b.y), make_uint2(kv.b.z, kv.b.w)};
#pragma unroll
    for (int h = 0; h < HPQ; ++h) {
      const int n = h * KT + j;
      *reinterpret_cast<uint2*>(ks + (n >> 3) * 256 + (h >> 1) * 128 + (n & 7) * 16 + (h & 1) * 8) = parts[h];
    }
  } else if (D == 8) {
#pragma unroll
    for (int h = 0; h < HPQ; ++h) {
      const int n = h * KT + j;
      *reinterpret_cast<uint4*>(ks + (n >> 3) * 256 + h * 128 + (n & 7) * 16) = h ? kv.b : kv.a;
    }
  } else {                                           // d = 16: one head fills the 16 K-slots, no expansion
    *reinterpret_cast<uint4*>(ks + (j >> 3) * 256 + (j & 7) * 16) = kv.a;
    *reinterpret_cast<uint4*>(ks + (j >> 3) * 256 + 128 + (j & 7) * 16) = kv.b;
  }
}

// V'_h (16 x KT, K-major) per head: rows 0..D-1 = v dims, row D = 1 (denominator), rest 0; absent keys become
// all-zero columns, which is how key tails are masked.  One thread per (key, head).
template <int D, int KT>
__device__ __forceinline__ void store_v(const KvRegs& kv, uint8_t* vs, int tid) {
  if (tid < 128) return;
  constexpr int NV = D == 16 ? 32 : 16;
  const int idx = tid - 128, j = idx % KT, h = idx / KT;
  uint4 raw[2] = {kv.a, kv.b};
  const uint16_t* e = reinterpret_cast<const uint16_t*>(raw);
  constexpr int head_bytes = NV * KT * 2, sbo = (KT / 8) * 128;
  uint8_t* base = vs + h * head_bytes + (j >> 3) * 128 + (j & 7) * 2;
#pragma unroll
  for (int d = 0; d < D; ++d) *reinterpret_cast<uint16_t*>(base + (d >> 3) * sbo + (d & 7) * 16) = e[d];
  const uint16_t one = kv.present ? (uint16_t)0x3F80 : (uint16_t)0;       // bf16 1.0
  *reinterpret_cast<uint16_t*>(base + (D >> 3) * sbo + (D & 7) * 16) = one;
}

// Producer-warp staging of one whole key tile: lane -> keys lane, lane+32, ...; per key the quad's 16 k channels go to
// the expanded-K operand (store_k) and its 16 v channels to the per-head V' operands (v dims | 1 | 0...).
template <int D, int KT>
__device__ __forceinline__ void stage_tile(const AttnParams& p, int64_t tok_base, int64_t tok_stride, int k0, int N, int quad,
                                           uint8_t* ks, uint8_t* vs, int lane) {
  constexpr int HPQ = 16 / D;
  constexpr int NV = D == 16 ? 32 : 16;
  constexpr int head_bytes = NV * KT * 2, sbo = (KT / 8) * 128;
  KvRegs kr[KT / 32];
  uint4 vr[KT / 32][2];
#pragma unroll
  for (int i = 0; i < KT / 32; ++i) {                  // all loads first (independent, L2-resident after the prologue prefetch)
    const int j = lane + 32 * i;
    kr[i].a = make_uint4(0, 0, 0, 0); kr[i].b = kr[i].a; vr[i][0] = kr[i].a; vr[i][1] = kr[i].a;
    kr[i].present = k0 + j < N;
    if (kr[i].present) {
      const bf16* src = (const bf16*)p.qkv + (tok_base + (int64_t)(k0 + j) * tok_stride) * p.ldq + p.C + quad * 16;
      kr[i].a = reinterpret_cast<const uint4*>(src)[0]; kr[i].b = reinterpret_cast<const uint4*>(src)[1];
      vr[i][0] = reinterpret_cast<const uint4*>(src + p.C)[0]; vr[i][1] = reinterpret_cast<const uint4*>(src + p.C)[1];
    }
  }
#pragma unroll
  for (int i = 0; i < KT / 32; ++i) {
    const int j = lane + 32 * i;
    store_k<D, KT>(kr[i], ks, j);
    const uint16_t* e = reinterpret_cast<const uint16_t*>(vr[i]);
    const uint16_t one = kr[i].present ? (uint16_t)0x3F80 : (uint16_t)0;       // bf16 1.0; absent keys are all-zero columns
#pragma unroll
    for (int h = 0; h < HPQ; ++h) {
      uint8_t* base = vs + h * head_bytes + (j >> 3) * 128 + (j & 7) * 2;
#pragma unroll
      for (int d = 0; d < D; ++d) *reinterpret_cast<uint16_t*>(base + (d >> 3) * sbo + (d & 7) * 16) = e[h * D + d];
      *reinterpret_cast<uint16_t*>(base + (D >> 3) * sbo + (D & 7) * 16) = one;
    }
  }
}

__device__ __forceinline__ float ex2_f32(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for x <= 0 on the FMA/ALU pipes (no MUFU): round-to-nearest split x = n + r with the 1.5*2^23 trick, degree-3
// polynomial for 2^r on [-0.5, 0.5] (max rel. error 7.5e-5, far below the bf16 resolution of P), exponent patched in
// with an integer add.  Used for a fixed fraction of the elements so the MUFU and FMA pipes share the exponentials
// (the kernel is otherwise bound by MUFU.EX2 at 16 results/clk/SM).
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + 12582912.0f;                  // low mantissa bits of t = round(x) (two's complement)
  const float r = x - (t - 12582912.0f);
  float p = fmaf(0.0551716685f, r, 0.2426111251f);  // minimax (relative) fit of 2^r on [-0.5, 0.5]: 7.5e-5
  p = fmaf(p, r, 0.6932609677f);
  p = fmaf(p, r, 0.9999280572f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
#ifndef TFSWA_TA_POLY_EVERY
#define TFSWA_TA_POLY_EVERY 4
#endif
constexpr int TA_POLY_EVERY = TFSWA_TA_POLY_EVERY;  // every TA_POLY_EVERY-th exponential goes to ex2_poly (0 = never); -D for A/B builds

// Per-channel extrema of k over every sequence: kext[row][0][c] = min_j k[j][c], kext[row][1][c] = max_j k[j][c].
// One CTA per sequence; thread t owns 8 channels (one 16-byte load per key) of key lane t / (C/8).
__global__ void __launch_bounds__(256) attn_kext_kernel(const AttnParams p) {
  __shared__ float red[2][256][9];
  const int row = blockIdx.x, tid = threadIdx.x;
  const int groups = p.C / 8, g = tid % groups, kl = tid / groups, KL = 256 / groups;
  const int N = p.geom == TFSWA_GEOM_TSA ? p.H : p.W;
  int64_t tok_base, tok_stride;
  if (p.geom == TFSWA_GEOM_TSA) { const int b = row / p.W; tok_base = (int64_t)b * p.H * p.W + (row - b * p.W); tok_stride = p.W; }
  else { tok_base = (int64_t)row * p.W; tok_stride = 1; }
  float mn[8], mx[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { mn[i] = CUDART_INF_F; mx[i] = -CUDART_INF_F; }
  if (kl < KL) {
    for (int j = kl; j < N; j += KL) {
      float v[8];
      load8((const bf16*)p.qkv + (tok_base + (int64_t)j * tok_stride) * p.ldq + p.C + g * 8, v);
#pragma unroll
      for (int i = 0; i < 8; ++i) { mn[i] = fminf(mn[i], v[i]); mx[i] = fmaxf(mx[i], v[i]); }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[0][tid][i] = mn[i]; red[1][tid][i] = mx[i]; }
  __syncthreads();
  if (tid < 2 * p.C) {                              // thread -> (min|max, channel): reduce over the key lanes
    const int which = tid / p.C, c = tid % p.C, gg = c / 8, i = c % 8;
    float v = which ? -CUDART_INF_F : CUDART_INF_F;
    for (int k = 0; k < KL; ++k) {
      const float o = red[which][k * groups + gg][i];
      v = which ? fmaxf(v, o) : fminf(v, o);
    }
    p.kext[((int64_t)row * 2 + which) * p.C + c] = v;
  }
}

// Roles: warps 0-7 ("softmax", 256 threads) own one query row per thread pair and turn S into P - in the key loop
// they touch neither global nor shared memory (TMEM in, TMEM out); warp 9 ("producer") stages the K / V' operands of
// tile t+2 in shared memory while tile t is being consumed (the generic->async proxy fence, a MEMBAR, stays off the
// softmax warps); warp 8 ("issuer") does nothing but wait on mbarriers and issue tcgen05.mma.  There is no CTA-wide
// barrier in the key loop:
//   bar_s   (1)  S(t) complete in TMEM                          issuer commit  -> softmax
//   bar_a   (9)  S(t-1) pulled into registers (8 softmax warps) + operands(t) staged (producer) -> issuer (may issue S(t)).
//                Every source arrives exactly ONCE per phase: a softmax warp's arrival for S(t-1) needs S(t-1), i.e. phase
//                t-1 over, and the producer - which stages up to two tiles ahead - holds its arrival for tile t until
//                bar_s says S(t-1) has completed.  (Without that hold a producer that was a tile ahead stood in for a softmax
//                warp that had not pulled its rows of S(t-1) yet, S(t) overwrote them, and about one full-size launch in ten
//                returned 32 wrong rows in one CTA.  Two separate barriers fix it too, at +1.8 %; the hold is free.)
//   bar_free(1)  S(t) and PV(t-1) complete: Ks[t&1], Vs[(t-1)%3] may be overwritten    softmax warp 0 -> producer
//   bar_b   (8)  P(t) written to TMEM                                 softmax  -> issuer (may issue PV(t))
//   bar_pv  (1)  PV(t) complete: P columns and V' buffer t%3 free     issuer commit -> softmax
// P never touches shared memory: the softmax threads write it to TMEM with tcgen05.st (their own lane = their query
// row) and the PV MMA takes its A operand from TMEM, so the only generic->async proxy hand-off per tile is K / V'.
template <int D>
__global__ void __launch_bounds__(TA_NTHREADS, 2) tc_attn_axial_kernel(const AttnParams p) {
  constexpr int HPQ = 16 / D;            // heads per CTA (4, 2, 1)
  constexpr int KT = 128 / HPQ;          // keys per tile: one S tile = 128 TMEM columns = HPQ heads x KT keys
  constexpr int HPT = HPQ >= 2 ? HPQ / 2 : 1;   // head slots per thread (d = 16: both threads of a row share the head)
  constexpr int NV = D == 16 ? 32 : 16;  // PV MMA N: v dims + ones column (+ zero padding)
  constexpr int P_SBO = (KT / 8) * 128;  // 8-row group stride of the P / V' operands
  constexpr int V_HEAD = NV * KT * 2;
  constexpr int TA_PS = ps_off<D>();
  constexpr int VSB = vs_bytes<D>();
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar_s, bar_a, bar_b, bar_pv, bar_free, bar_sx[2], bar_ax[2];
  __shared__ uint32_t s_tmem;
  __shared__ float s_kext[2][16];        // per channel of the quad: min / max of k over the whole sequence

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool softmax = warp < 8, issuer = warp == 8, producer = warp == 9;
  // grid.x enumerates the (query tile, quad) pairs of one sequence, grid.y the sequences: CTAs that share a
  // sequence's k|v rows are scheduled together, so those rows come from HBM once and from L2 afterwards
  const int nquads = p.C / 16;
  const int row = blockIdx.y, q0 = (blockIdx.x / nquads) * TA_QT, quad = blockIdx.x % nquads;
  const int N = p.geom == TFSWA_GEOM_TSA ? p.H : p.W;
  const int quarter = warp & 3, half = (warp >> 2) & 1;
  const int r = quarter * 32 + lane;                 // my query row == my TMEM lane
  const float c = p.qscale;                          // head_dim^-0.5 * log2(e)
  const int T = (N + KT - 1) / KT;
  int64_t tok_base, tok_stride;
  if (p.geom == TFSWA_GEOM_TSA) { const int b = row / p.W; tok_base = (int64_t)b * p.H * p.W + (row - b * p.W); tok_stride = p.W; }
  else { tok_base = (int64_t)row * p.W; tok_stride = 1; }

  // ---- setup: zero the operand buffers (their zero patterns are permanent), barriers, TMEM ----
  for (int i = tid; i < (TA_PS - TA_KS) / 16; i += TA_NTHREADS) reinterpret_cast<uint4*>(smem + TA_KS)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) {
    if (lane == 0) {
      mbar_init(&bar_s, 1); mbar_init(&bar_a, 9); mbar_init(&bar_b, 8); mbar_init(&bar_pv, 1); mbar_init(&bar_free, 1);
#pragma unroll
      for (int i = 0; i < 2; ++i) { mbar_init(&bar_sx[i], 1); mbar_init(&bar_ax[i], 8); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&s_tmem, TA_TMEM_COLS);
  }
  bool q_valid = false;
  int64_t q_tok = 0;
  uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;        // my row's 16 q channels
  if (softmax) {
    if (q0 + r < (p.q_end ? p.q_end : N)) { q_tok = tok_base + (int64_t)(q0 + r) * tok_stride; q_valid = true; }
    if (q_valid) {
      const uint4* src = reinterpret_cast<const uint4*>((const bf16*)p.qkv + q_tok * p.ldq + quad * 16);
      qa = src[0]; qb = src[1];
    }
    if (tid < TA_QT) {                               // Q operand: row r, 16 channels = 2 chunks of 16 B
      *reinterpret_cast<uint4*>(smem + TA_QS + (r >> 3) * 256 + (r & 7) * 16) = qa;
      *reinterpret_cast<uint4*>(smem + TA_QS + (r >> 3) * 256 + 128 + (r & 7) * 16) = qb;
    }
  }
  // Pull this sequence's k|v rows (32 B each per quad) towards L2 now: the key loop only prefetches one tile ahead
  // into registers, which hides an L2 hit but not an HBM miss.
  for (int j = tid; j < 2 * N; j += TA_NTHREADS) {
    const int key = j >> 1;
    const bf16* ptr = (const bf16*)p.qkv + (tok_base + (int64_t)key * tok_stride) * p.ldq + ((j & 1) ? 2 * p.C : p.C) + quad * 16;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
  }
  if (tid < 32) {   // per-channel extrema of k over the sequence, computed once per (sequence, channel) by attn_kext_kernel
    const float* ke = p.kext + ((int64_t)row * 2 + (tid >> 4)) * p.C + quad * 16 + (tid & 15);
    s_kext[tid >> 4][tid & 15] = *ke;
  }
  // The zero patterns of the K / V' buffers and the Q operand were written through the generic proxy by every thread and
  // are read by tcgen05.mma through the async proxy: each WRITER has to fence before the barrier.  (Without this fence a
  // full-size launch occasionally - about one run in three - returned results that differed from run to run.)
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t sbase = smem_u32(smem);
  const uint64_t qdesc = umma_smem_desc_ns(sbase + TA_QS, 128, 256);
  const uint32_t idesc_s = umma_idesc_bf16(128, 128);
  const uint32_t tmem = s_tmem;
  const uint32_t my_taddr = tmem + ((uint32_t)(quarter * 32) << 16);
  // barrier completions consumed so far by this thread (wait parity = count & 1)
  uint32_t n_s = 0, n_a = 0, n_b = 0, n_pv = 0, n_free = 0, n_sx[2] = {0u, 0u}, n_ax[2] = {0u, 0u};

  // row-max upper bound per head: s_ij = sum_d q_d k_jd <= sum_d max(q_d kmax_d, q_d kmin_d)   (raw score units)
  float m[HPT];
#pragma unroll
  for (int hh = 0; hh < HPT; ++hh) m[hh] = 0.f;
  if (softmax) {
    uint4 raw[2] = {qa, qb};
    const __nv_bfloat16* qe = reinterpret_cast<const __nv_bfloat16*>(raw);
#pragma unroll
    for (int hh = 0; hh < HPT; ++hh) {
      const int head = HPQ >= 2 ? half * HPT + hh : 0;
      float b = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) {
        const float qv = __bfloat162float(qe[head * D + d]);
        b += fmaxf(qv * s_kext[1][head * D + d], qv * s_kext[0][head * D + d]);
      }
      m[hh] = b;
    }
  }

  for (int attempt = 0; attempt < 2; ++attempt) {
    if (attempt == 1 || p.force_exact) {
      // ================= exact row maxima (fallback): stream all S tiles once, double-buffered in TMEM =================
      //   bar_ax[b] (8): K(t) staged in Ks[b] and TMEM buffer b drained  -> issuer issues S(t);   bar_sx[b] (1): S(t) complete
      if (issuer) {
        for (int t = 0; t < T; ++t) {
          const int b = t & 1;
          mbar_wait(&bar_ax[b], n_ax[b] & 1); ++n_ax[b];
          if (lane == 0) {
            tc_fence_after();
            umma_bf16_ss(tmem + b * 128, qdesc, umma_smem_desc_ns(sbase + TA_KS + b * 8192, 128, 256), idesc_s, 0u);
            umma_commit(&bar_sx[b]);
          }
          __syncwarp();
        }
      } else if (softmax) {
#pragma unroll
        for (int i = 0; i < HPT; ++i) m[i] = -CUDART_INF_F;
        // stage K(0), K(1)
        for (int t0 = 0; t0 < 2 && t0 < T; ++t0) {
          KvRegs kv = load_kv<D, KT>(p, tok_base, tok_stride, t0 * KT, N, quad, tid);
          store_k<D, KT>(kv, smem + TA_KS + t0 * 8192, tid);
          fence_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bar_ax[t0]);
        }
        for (int t = 0; t < T; ++t) {
          const int b = t & 1;
          KvRegs kv2 = load_kv<D, KT>(p, tok_base, tok_stride, (t + 2) * KT, N, quad, tid);
          mbar_wait(&bar_sx[b], n_sx[b] & 1); ++n_sx[b];
          tc_fence_after();
          const int kcount = min(KT, N - t * KT);
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {             // my 64 of the 128 columns
            uint32_t s[32];
            __syncwarp();
            tmem_ld_x32(my_taddr + b * 128 + half * 64 + ch * 32, s);
            tmem_ld_wait();
            const int jbase = (half * 64 + ch * 32) % KT;
            float mx = m[(ch * 32) / KT];
            if (jbase + 32 <= kcount) {
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(s[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i) if (jbase + i < kcount) mx = fmaxf(mx, __uint_as_float(s[i]));
            }
            m[(ch * 32) / KT] = mx;
          }
          if (t + 2 < T) {
            store_k<D, KT>(kv2, smem + TA_KS + b * 8192, tid);      // S(t) complete -> Ks[b] free
            fence_async_smem();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_ax[b]);
          }
        }
      }
      if (HPQ == 1 && softmax) reinterpret_cast<float*>(smem + TA_PS)[half * 128 + r] = m[0];
      tc_fence_before();
      __syncthreads();                                              // every S tile consumed before the buffers are reused
      tc_fence_after();
      if (HPQ == 1) {                                               // the two threads of a row each saw half of the keys
        if (softmax) { const float* ex = reinterpret_cast<const float*>(smem + TA_PS); m[0] = fmaxf(ex[r], ex[128 + r]); }
        __syncthreads();
      }
    }
    float mc[HPT];
#pragma unroll
    for (int i = 0; i < HPT; ++i) mc[i] = m[i] * c;

    // =========================== P = ex2(S*c - m*c), O += P V' ===========================
    if (issuer) {
      const uint32_t idesc_pv = umma_idesc_bf16(128, NV);
      for (int t = 0; t <= T; ++t) {
        mbar_wait(&bar_a, n_a & 1); ++n_a;                            // S(t-1) consumed, operands(t) staged
        if (t < T && lane == 0) {
          tc_fence_after();
          umma_bf16_ss(tmem, qdesc, umma_smem_desc_ns(sbase + TA_KS + (t & 1) * 8192, 128, 256), idesc_s, 0u);
          umma_commit(&bar_s);
        }
        if (t >= 1) {
          const int u = t - 1;
          mbar_wait(&bar_b, n_b & 1); ++n_b;                          // P(u) in TMEM
          if (lane == 0) {
            tc_fence_after();
            const uint32_t vbase = sbase + TA_VS + (u % 3) * VSB;
#pragma unroll
            for (int h = 0; h < HPQ; ++h) {
#pragma unroll
              for (int kk = 0; kk < KT / 16; ++kk) {       // head h's keys occupy KT/2 columns; one K=16 step = 8 columns
                umma_bf16_ts(tmem + TA_O_COL + NV * h, tmem + TA_P_COL + h * (KT / 2) + kk * 8,
                             umma_smem_desc_ns(vbase + h * V_HEAD + kk * 256, 128, P_SBO), idesc_pv, (u | kk) ? 1u : 0u);
              }
            }
            umma_commit(&bar_pv);
          }
        }
        __syncwarp();
      }
    } else if (producer) {
      // operands(k) for k = 0..T-1 (k >= 2 waits until tile k-2 released its buffers); one arrival per bar_a phase, held until
      // S(k-1) has completed so that it can never land in phase k-1 (see the barrier table above)
      for (int k = 0; k <= T; ++k) {
        if (k < T) {
          if (k >= 2) { mbar_wait(&bar_free, n_free & 1); ++n_free; }
          stage_tile<D, KT>(p, tok_base, tok_stride, k * KT, N, quad, smem + TA_KS + (k & 1) * 8192, smem + TA_VS + (k % 3) * VSB, lane);
          fence_async_smem();
        }
        if (k >= 1) { mbar_wait(&bar_s, n_s & 1); ++n_s; }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_a);
      }
    } else {
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_a);                             // phase 0: nothing to consume yet
      for (int t = 0; t < T; ++t) {
        mbar_wait(&bar_s, n_s & 1); ++n_s;
        tc_fence_after();
        const bool tail = (t + 1) * KT > N;          // only the last tile holds absent keys: their score 0 may exceed the bound
        // my 64 scores in two halves of 32 (keeps only 32 score registers live): exp of the first half runs while
        // nothing else of S(t) is needed; S is released to the issuer after the second TMEM load
#pragma unroll
        for (int ch = 0; ch < 2; ++ch) {
          uint32_t sc[32];
          __syncwarp();
          tmem_ld_x32(my_taddr + half * 64 + ch * 32, sc);
          tmem_ld_wait();
          if (ch == 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_a);                       // S(t) consumed
          }
          const int col = half * 64 + ch * 32;
          const float mcc = mc[(ch * 32) / KT];
          if (tail) {
#pragma unroll
            for (int i = 0; i < 32; ++i) sc[i] = __float_as_uint(fminf(__uint_as_float(sc[i]), m[(ch * 32) / KT]));
          }
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float x0 = fmaf(__uint_as_float(sc[2 * i]), c, -mcc);
            const float x1 = fmaf(__uint_as_float(sc[2 * i + 1]), c, -mcc);
            const float e0 = (TA_POLY_EVERY > 0 && ((2 * i) % TA_POLY_EVERY) == TA_POLY_EVERY - 1) ? ex2_poly(x0) : ex2_f32(x0);
            const float e1 = (TA_POLY_EVERY > 0 && ((2 * i + 1) % TA_POLY_EVERY) == TA_POLY_EVERY - 1) ? ex2_poly(x1) : ex2_f32(x1);
            pk[i] = pack_bf16x2(e0, e1);
          }
          if (ch == 0) {
            if (t >= 1) {                            // PV(t-1) complete: the P columns (and Vs[(t-1)%3]) are free
              mbar_wait(&bar_pv, n_pv & 1); ++n_pv;
              tc_fence_after();
            }
            if (tid == 0 && t + 2 < T) mbar_arrive(&bar_free);       // S(t), PV(t-1) complete -> producer may stage tile t+2
          }
          __syncwarp();
          tmem_st_x16(my_taddr + TA_P_COL + (uint32_t)(col >> 1), pk);   // column pair (2c, 2c+1) -> 32-bit cell c
        }
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_b);                           // P(t) in TMEM
      }
      // (measured: consuming S in four 16-column chunks with the tcgen05.ld of chunk k+1 / k+2 in flight under the
      // exponentials of chunk k is 4-6 % slower - the load latency is already covered by the other warps of the SMSP)
      // (measured: splitting P / PV per 32-column chunk with two barrier pairs, to give each chunk a whole tile of
      // slack, is 11 % slower - the extra tcgen05.wait::st, arrives and issuer wake-ups cost more than the sleep they remove)
      // ---- drain: PV(0..T-2) were consumed inside the loop, PV(T-1) is the one outstanding completion ----
      mbar_wait(&bar_pv, n_pv & 1); ++n_pv;
      tc_fence_after();
    }
    // ---- epilogue: O / l ----
    uint32_t o[32];
    bool bad = false;
    if (softmax) {
      __syncwarp();
      if (D == 8) {
        uint32_t t16[16];
        tmem_ld_x16(my_taddr + TA_O_COL + 16 * half, t16);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = t16[i];
      } else {
        tmem_ld_x32(my_taddr + TA_O_COL + (D == 4 ? 32 * half : 0), o);
        tmem_ld_wait();
      }
      // the bound keeps every exponent <= 0; if it was so loose that a whole row underflowed (denominator ~ 0) the CTA
      // repeats the computation with the exact maximum
      if (D == 16) bad = q_valid && !(__uint_as_float(o[16]) > 1e-30f);
      else {
#pragma unroll
        for (int hh = 0; hh < HPT; ++hh) bad = bad || (q_valid && !(__uint_as_float(o[hh * 16 + D]) > 1e-30f));
      }
    }
    tc_fence_before();
    const bool redo = attempt == 0 && !p.force_exact && __syncthreads_or(bad);
    if (redo) continue;
    if (softmax && q_valid) {
      if (D == 16) {                                 // both threads of the row hold the same 32 columns: split the 16 dims
        const float l = __uint_as_float(o[16]);
        const float inv = 1.0f / l;
        float v[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) v[d] = __uint_as_float(half ? o[8 + d] : o[d]) * inv;
        store8((bf16*)p.out + q_tok * p.ldo + quad * 16 + half * 8, v);
        if (p.lse && half == 0) p.lse[q_tok * p.heads + quad] = mc[0] + log2f(l);
      } else {
#pragma unroll
        for (int hh = 0; hh < HPT; ++hh) {
          const int head = half * HPT + hh;          // head within the quad
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
    break;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TA_TMEM_COLS);
}

}  // namespace tfswa

using namespace tfswa;

static int64_t kext_bytes(const tfswa_attn_args* a) {
  const int64_t rows = a->geom == TFSWA_GEOM_TSA ? (int64_t)a->B * a->W : (int64_t)a->B * a->H;
  return (rows * 2 * a->C * (int64_t)sizeof(float) + 255) / 256 * 256;
}

// [per-sequence k extrema] [work space of the TMA kernel: redo count, flags, list]
extern "C" int64_t tfswa_attn_tc_scratch_bytes(const tfswa_attn_args* a) {
  if (!a || a->C <= 0) return 0;
  AttnParams p = {};
  p.B = a->B; p.H = a->H; p.W = a->W; p.C = a->C; p.geom = a->geom;
  return kext_bytes(a) + attn_axial_tma_work_bytes(p);
}

extern "C" int tfswa_attn_tc_fwd(const tfswa_attn_args* a, void* scratch, int64_t scratch_bytes, void* stream) {
  TFSWA_REQUIRE(a && a->qkv && a->out, "attn_tc: null pointer");
  TFSWA_REQUIRE(a->dtype == TFSWA_BF16, "attn_tc: bf16 activations only");
  TFSWA_REQUIRE(a->geom == TFSWA_GEOM_TSA || a->geom == TFSWA_GEOM_FSA, "attn_tc: axial geometries only (windows use tfswa_attn_fwd)");
  TFSWA_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0 && a->heads > 0 && a->C % a->heads == 0 && a->C % 16 == 0, "attn_tc: bad shape");
  const int D = a->C / a->heads;
  TFSWA_REQUIRE(D == 4 || D == 8 || D == 16, "attn_tc: head_dim %d not in {4,8,16} (use tfswa_attn_fwd)", D);
  TFSWA_REQUIRE(a->ldq % 8 == 0 && a->ldo % 4 == 0 && (((uintptr_t)a->qkv) & 15) == 0, "attn_tc: alignment");
  AttnParams p = {};
  p.qkv = a->qkv; p.ldq = a->ldq; p.out = a->out; p.ldo = a->ldo; p.lse = a->lse;
  p.B = a->B; p.H = a->H; p.W = a->W; p.C = a->C; p.heads = a->heads; p.geom = a->geom;
  p.qscale = (float)(1.4426950408889634 / sqrt((double)D));
  TFSWA_REQUIRE(scratch && scratch_bytes >= tfswa_attn_tc_scratch_bytes(a), "attn_tc: scratch buffer too small");
  TFSWA_REQUIRE(a->C <= 128, "attn_tc: C=%d > 128 unsupported", a->C);
  p.kext = (float*)scratch;
  p.force_exact = (a->flags & TFSWA_ATTN_FORCE_EXACT) ? 1 : 0;      // test hook: exact two-pass path
  const int N = a->geom == TFSWA_GEOM_TSA ? a->H : a->W;
  const int rows = a->geom == TFSWA_GEOM_TSA ? a->B * a->W : a->B * a->H;
  // Ragged remainder: the MMA needs 128-query tiles; when the last tile would hold only a few queries (1025 = 8*128 + 1,
  // 517 = 4*128 + 5) those queries go to the SIMT flash kernel instead of a 128-row tile that is >75 % padding.
  const int rem = N % TA_QT;
  const int q_tc = (N >= TA_QT && rem > 0 && rem < 32) ? N - rem : N;
  p.q_begin = 0; p.q_end = q_tc;
  dim3 grid(((q_tc + TA_QT - 1) / TA_QT) * (a->C / 16), rows, 1);
  TFSWA_REQUIRE(rows <= 65535, "attn_tc: more than 65535 sequences in one launch (split the batch)");
  cudaStream_t st = (cudaStream_t)stream;
  attn_kext_kernel<<<rows, 256, 0, st>>>(p);
  // TFSWA_AXIAL_KERNEL=tma|umma|mma selects the main kernel (A/B): "tma" = tcgen05 + TMEM with TMA-fed operands
  // (tc_attn_tma.cu), "umma" = round 1's thread-staged tcgen05 kernel (this file), "mma" = register-resident warp-level
  // MMAs (attention_axial_mma.cu).
  // Measured on B200 (tools/attn_bench.py, C3 stage shapes, B=8, TSA / FSA ms per launch): head_dim 4: tma 9.49 / 5.20, umma
  // (round 1) 10.58 / 6.26; head_dim 8: tma 1.34 / 0.99, mma 1.46 / 1.02; head_dim 16: tma 0.30 / 0.31, mma 0.37 / 0.33
  // -> the TMA-fed tcgen05 kernel is the default at every head_dim.
  static int force = -1;                              // 0 = default choice, 1 = umma, 2 = mma, 3 = tma
  if (force < 0) { const char* e = getenv("TFSWA_AXIAL_KERNEL"); force = !e ? 0 : (e[0] == 'm' ? 2 : (e[0] == 't' ? 3 : 1)); }
  const bool use_mma = force == 2;
  if (use_mma && a->heads % 8 == 0) {                // 16-row granularity: no separate remainder pass
    AttnParams pm = p; pm.q_begin = 0; pm.q_end = 0;
    return attn_axial_mma_bf16(pm, st);
  }
  if (force == 3 || force == 0) {
    int rc = attn_axial_tma_bf16(p, (char*)scratch + kext_bytes(a), st);
    if (rc) return rc;
    if (q_tc < N) {                                  // ragged remainder (< 32 queries per sequence), see below
      AttnParams ps = p;
      ps.q_begin = q_tc; ps.q_end = 0;
      if (a->heads % 8 == 0 && N - q_tc > 2) return attn_axial_mma_bf16(ps, st);
      return attn_simt_axial_bf16(ps, st);
    }
    return TFSWA_OK;
  }
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    cudaError_t e1 = cudaFuncSetAttribute(tc_attn_axial_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<4>());
    cudaError_t e2 = cudaFuncSetAttribute(tc_attn_axial_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<8>());
    if (e2 == cudaSuccess) e2 = cudaFuncSetAttribute(tc_attn_axial_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<16>());
    cudaFuncSetAttribute(tc_attn_axial_kernel<16>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("attn_tc: cudaFuncSetAttribute failed"); return TFSWA_ECUDA; }
    attr_once.done();
  }
  if (D == 4) tc_attn_axial_kernel<4><<<grid, TA_NTHREADS, smem_bytes<4>(), st>>>(p);
  else if (D == 8) tc_attn_axial_kernel<8><<<grid, TA_NTHREADS, smem_bytes<8>(), st>>>(p);
  else tc_attn_axial_kernel<16><<<grid, TA_NTHREADS, smem_bytes<16>(), st>>>(p);
  if (q_tc < N) {                                    // ragged remainder (< 32 queries per sequence)
    int rc = check_launch("attn_tc");
    if (rc) return rc;
    AttnParams ps = p;
    ps.q_begin = q_tc; ps.q_end = 0;
    // (measured: forking this launch onto a side stream so that it runs under the main kernel changes nothing -
    // 10.594 vs 10.585 ms at B=8 - the main kernel owns every SM until its last wave)
    // 3+ left-over queries: one 16-row tile of the register-resident kernel (FSA, 5 queries: 90 -> 42 us at B=1);
    // 1-2 queries: the key-split warp kernel is still ahead (36 vs 40 us)
    if (a->heads % 8 == 0 && force != 1 && N - q_tc > 2) return attn_axial_mma_bf16(ps, st);
    return attn_simt_axial_bf16(ps, st);
  }
  return check_launch("attn_tc");
}
