// fp32-accumulate SIMT implicit GEMM over NHWC tokens:  Y = epi(pro(A) W^T + bias) (+R1) (+R2)
//
// This is the full-precision ("fp32 parity") engine and the generic fallback for shapes the
// tcgen05 kernels do not cover.  A rows are tokens; the A gather is a template policy so the same
// kernel serves nn.Linear / 1x1 convs (attention.py:70,86,121-128; blocks.py:53,85), the 3x3 head
// conv (tfswa_unet.py:140), the 4x4/s2 down conv (blocks.py:157) and the 4-phase transposed conv
// (blocks.py:172).  Tiles: TM x TN outputs per CTA (256 threads, 8x4 per thread), K in chunks of 16
// through double-buffered shared memory; 128-bit global accesses throughout.
#include "common.cuh"
#include <stdlib.h>

namespace tfswa {

enum { KIND_LINEAR = 0, KIND_CONV3 = 1, KIND_DOWN = 2, KIND_UP = 3 };

// bf16 linear weight gradient on the tensor cores (wgrad_mma.cu); returns 1 when the shape is not covered
int wgrad_mma_bf16(const tfswa_linear_args* a, const void* g, int64_t ldg, int64_t g_bs, float* dw, float* dbias, cudaStream_t st);
int wgrad_tc_bf16(const tfswa_linear_args* a, const void* g, int64_t ldg, int64_t g_bs, float* dw, float* dbias, cudaStream_t st);
int conv_wgrad_mma_bf16(const tfswa_conv_args* a, const void* g, float* dw, float* dbias, cudaStream_t st);

struct IgemmParams {
  const void* x; int64_t ldx;
  const float* w; const float* bias;
  const float* row_stats; const float* in_scale; const float* in_shift;
  const void* r1; int64_t ldr1; const void* r2; int64_t ldr2;
  void* y; int64_t ldy; void* pre; int64_t ldpre;
  float* col_stats;
  int64_t M; int N; int K;
  int prologue; int epilogue;
  // batch strides (elements)
  int64_t x_bs, w_bs, bias_bs, rs_bs, r1_bs, r2_bs, y_bs, pre_bs;
  // conv geometry
  int B, Hin, Win, Cin, Hout, Wout;
  int Hq, Wq;      // KIND_UP: phase grid = ceil(Hout/2) x ceil(Wout/2)
  // weight-gradient mode
  const void* g; int64_t ldg; int64_t g_bs; float* dw; float* dbias; int64_t rows_per_cta; int msplit;
};

constexpr int TK = 16;
constexpr int NTHREADS = 256;

// Element offset into x of the 8-element run of A starting at logical (row m, k), or -1 for zero padding.
template <int KIND>
__device__ __forceinline__ int64_t a_offset(const IgemmParams& p, int64_t m, int k, int phase) {
  if (KIND == KIND_LINEAR) return m * p.ldx + k;
  const int tap = k / p.Cin, ci = k - tap * p.Cin;
  const int hw = (KIND == KIND_UP) ? p.Hq * p.Wq : p.Hout * p.Wout;
  const int wdim = (KIND == KIND_UP) ? p.Wq : p.Wout;
  const int b = (int)(m / hw);
  const int rem = (int)(m - (int64_t)b * hw);
  const int oy = rem / wdim, ox = rem - oy * wdim;
  int iy, ix;
  if (KIND == KIND_CONV3) { iy = oy + tap / 3 - 1; ix = ox + tap % 3 - 1; }
  else if (KIND == KIND_DOWN) { iy = 2 * oy + (tap >> 2) - 1; ix = 2 * ox + (tap & 3) - 1; }
  else {  // KIND_UP: (oy, ox) index the phase grid; tap = a*2 + b
    const int py = phase >> 1, px = phase & 1, a = tap >> 1, bb = tap & 1;
    iy = py ? (a ? oy : oy + 1) : (a ? oy - 1 : oy);
    ix = px ? (bb ? ox : ox + 1) : (bb ? ox - 1 : ox);
    if (2 * oy + py >= p.Hout || 2 * ox + px >= p.Wout) return -1;   // phase-grid cell outside the output
  }
  if (iy < 0 || iy >= p.Hin || ix < 0 || ix >= p.Win) return -1;
  return (((int64_t)b * p.Hin + iy) * p.Win + ix) * p.Cin + ci;
}

// element offset of output row m (-1: the phase-grid cell lies outside the output, nothing to store)
template <int KIND>
__device__ __forceinline__ int64_t out_row_offset(const IgemmParams& p, int64_t m, int phase, int64_t ld) {
  if (KIND != KIND_UP) return m * ld;
  const int hw = p.Hq * p.Wq;
  const int b = (int)(m / hw);
  const int rem = (int)(m - (int64_t)b * hw);
  const int oy2 = rem / p.Wq, ox2 = rem - oy2 * p.Wq;
  const int oy = 2 * oy2 + (phase >> 1), ox = 2 * ox2 + (phase & 1);
  if (oy >= p.Hout || ox >= p.Wout) return -1;
  return (((int64_t)b * p.Hout + oy) * p.Wout + ox) * ld;
}

template <typename T, int KIND, int TN>
__global__ void __launch_bounds__(NTHREADS) igemm_kernel(const IgemmParams p) {
  constexpr int TXN = TN / 4;            // threads across N
  constexpr int TYN = NTHREADS / TXN;    // threads across M
  constexpr int TM = TYN * 8;
  constexpr int A_VEC = TM * TK / 8 / NTHREADS;   // load8 vectors per thread per chunk
  constexpr int LDA = TM + 4, LDB = TN + 4;
  __shared__ __align__(16) float As[2][TK][LDA];
  __shared__ __align__(16) float Bs[2][TK][LDB];
  __shared__ float s_stats[2][TN];

  const int tid = threadIdx.x;
  const int z = blockIdx.z;               // batch index (linear) or phase (up-conv)
  const int phase = (KIND == KIND_UP) ? z : 0;
  const int64_t m0 = (int64_t)blockIdx.x * TM;
  const int n0 = blockIdx.y * TN;
  const T* x = (const T*)p.x + (KIND == KIND_LINEAR ? z * p.x_bs : 0);
  const float* w = p.w + (int64_t)z * p.w_bs;
  const float* rs = p.row_stats ? p.row_stats + (int64_t)z * p.rs_bs : nullptr;

  const int ty = tid / TXN, tx = tid % TXN;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // per-thread A-load coordinates
  int a_row[A_VEC]; const int a_kseg = (tid & 1) * 8;
  float a_mean[A_VEC], a_rstd[A_VEC];
#pragma unroll
  for (int i = 0; i < A_VEC; ++i) {
    a_row[i] = (tid + i * NTHREADS) >> 1;
    a_mean[i] = 0.f; a_rstd[i] = 1.f;
    const int64_t m = m0 + a_row[i];
    if ((p.prologue & TFSWA_PRO_LNHAT) && m < p.M) { a_mean[i] = rs[m * 2]; a_rstd[i] = rs[m * 2 + 1]; }
  }
  const int b_n = tid >> 2, b_kseg = (tid & 3) * 4;
  const bool b_active = (tid < TN * 4);

  float a_reg[A_VEC][8];
  float4 b_reg = make_float4(0.f, 0.f, 0.f, 0.f);

  auto load_chunk = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_VEC; ++i) {
      const int64_t m = m0 + a_row[i];
      const int k = k0 + a_kseg;
      int64_t off = -1;
      if (m < p.M && k < p.K) off = a_offset<KIND>(p, m, k, phase);
      if (off >= 0) {
        load8(x + off, a_reg[i]);
        if (KIND == KIND_LINEAR && p.prologue != TFSWA_PRO_NONE) {
          if (p.prologue & TFSWA_PRO_AFFINE) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a_reg[i][j] = a_reg[i][j] * p.in_scale[k + j] + p.in_shift[k + j];
          }
          if (p.prologue & TFSWA_PRO_LNHAT) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a_reg[i][j] = (a_reg[i][j] - a_mean[i]) * a_rstd[i];
          }
          if (p.prologue & TFSWA_PRO_GELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a_reg[i][j] = gelu_erf(a_reg[i][j]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) a_reg[i][j] = 0.f;
      }
    }
    if (b_active) {
      const int n = n0 + b_n, k = k0 + b_kseg;
      if (n < p.N && k < p.K) b_reg = *reinterpret_cast<const float4*>(w + (int64_t)n * p.K + k);
      else b_reg = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto store_chunk = [&](int buf) {
#pragma unroll
    for (int i = 0; i < A_VEC; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) As[buf][a_kseg + j][a_row[i]] = a_reg[i][j];
    if (b_active) {
      Bs[buf][b_kseg + 0][b_n] = b_reg.x; Bs[buf][b_kseg + 1][b_n] = b_reg.y;
      Bs[buf][b_kseg + 2][b_n] = b_reg.z; Bs[buf][b_kseg + 3][b_n] = b_reg.w;
    }
  };

  const int nchunks = (p.K + TK - 1) / TK;
  load_chunk(0);
  store_chunk(0);
  __syncthreads();
  for (int kc = 0; kc < nchunks; ++kc) {
    const int buf = kc & 1;
    if (kc + 1 < nchunks) load_chunk((kc + 1) * TK);
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 8 + 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    if (kc + 1 < nchunks) store_chunk(buf ^ 1);
    __syncthreads();
  }

  // ---------------- epilogue ----------------
  const int n = n0 + tx * 4;
  const bool n_ok = n < p.N;   // N % 4 == 0 is required by the host wrapper
  float bias[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.bias && n_ok) {
    const float4 t = *reinterpret_cast<const float4*>(p.bias + (int64_t)(KIND == KIND_LINEAR ? z : 0) * p.bias_bs + n);
    bias[0] = t.x; bias[1] = t.y; bias[2] = t.z; bias[3] = t.w;
  }
  if (p.col_stats) {
    for (int i = tid; i < 2 * TN; i += NTHREADS) (&s_stats[0][0])[i] = 0.f;
    __syncthreads();
  }
  float cs[4] = {0.f, 0.f, 0.f, 0.f}, cq[4] = {0.f, 0.f, 0.f, 0.f};
  T* y = (T*)p.y + (KIND == KIND_LINEAR ? z * p.y_bs : 0);
  T* pre = p.pre ? (T*)p.pre + (KIND == KIND_LINEAR ? z * p.pre_bs : 0) : nullptr;
  const T* r1 = p.r1 ? (const T*)p.r1 + (KIND == KIND_LINEAR ? z * p.r1_bs : 0) : nullptr;
  const T* r2 = p.r2 ? (const T*)p.r2 + (KIND == KIND_LINEAR ? z * p.r2_bs : 0) : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ty * 8 + i;
    if (m >= p.M || !n_ok) continue;
    const int64_t yoff = out_row_offset<KIND>(p, m, phase, p.ldy);
    if (yoff < 0) continue;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = acc[i][j] + bias[j];
      cs[j] += v[j]; cq[j] += v[j] * v[j];
    }
    if (pre) store4(pre + out_row_offset<KIND>(p, m, phase, p.ldpre) + n, v);
    if (p.epilogue == TFSWA_EPI_GELU) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = gelu_erf(v[j]);
    }
    if (r1) { float t[4]; load4(r1 + m * p.ldr1 + n, t);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += t[j]; }
    if (r2) { float t[4]; load4(r2 + m * p.ldr2 + n, t);
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] += t[j]; }
    store4(y + yoff + n, v);
  }
  if (p.col_stats) {
    if (n_ok) {
#pragma unroll
      for (int j = 0; j < 4; ++j) { atomicAdd(&s_stats[0][tx * 4 + j], cs[j]); atomicAdd(&s_stats[1][tx * 4 + j], cq[j]); }
    }
    __syncthreads();
    for (int i = tid; i < 2 * TN; i += NTHREADS) {
      const int which = i / TN, c = i % TN;
      if (n0 + c < p.N) atomicAdd(p.col_stats + (int64_t)which * p.N + n0 + c, s_stats[which][c]);
    }
  }
}

template <typename T, int KIND>
static int launch_igemm(const IgemmParams& p, int zdim, cudaStream_t st) {
  if (p.N <= 32) {
    constexpr int TN = 32, TM = (NTHREADS / (TN / 4)) * 8;
    dim3 grid((unsigned)ceil_div64(p.M, TM), (p.N + TN - 1) / TN, zdim);
    igemm_kernel<T, KIND, TN><<<grid, NTHREADS, 0, st>>>(p);
  } else {
    constexpr int TN = 64, TM = (NTHREADS / (TN / 4)) * 8;
    dim3 grid((unsigned)ceil_div64(p.M, TM), (p.N + TN - 1) / TN, zdim);
    igemm_kernel<T, KIND, TN><<<grid, NTHREADS, 0, st>>>(p);
  }
  return check_launch("igemm");
}


// ------------------------------------------------------------------------------------------------
// weight gradient:  dW[n][k] += sum_m G[m][n] * pro(A)[m][k]   (and dbias[n] += sum_m G[m][n])
// The reduction runs over millions of tokens while the result is at most 256 x 1024: split M across CTAs
// (msplit), accumulate a 64x64 tile per CTA in registers and finish with fp32 atomics.  The A gather and the
// prologues are the forward's, so nn.Linear, 1x1 and k x k convolutions share this kernel.
// ------------------------------------------------------------------------------------------------
constexpr int WG_T = 64;      // tile edge (n and k)
constexpr int WG_MC = 16;     // rows per shared-memory chunk

template <typename T, int KIND>
__global__ void __launch_bounds__(NTHREADS) wgrad_kernel(const IgemmParams p) {
  __shared__ __align__(16) float Gs[WG_MC][WG_T + 4];
  __shared__ __align__(16) float As[WG_MC][WG_T + 4];
  const int tid = threadIdx.x;
  const int k0 = blockIdx.x * WG_T, n0 = blockIdx.y * WG_T;
  const int zb = blockIdx.z / p.msplit, split = blockIdx.z % p.msplit;   // zb: batch (linear) or phase (up-conv)
  const int phase = (KIND == KIND_UP) ? zb : 0;
  const T* x = (const T*)p.x + (KIND == KIND_LINEAR ? zb * p.x_bs : 0);
  const T* g = (const T*)p.g + (KIND == KIND_LINEAR ? zb * p.g_bs : 0);
  const float* rs = p.row_stats ? p.row_stats + (int64_t)zb * p.rs_bs : nullptr;
  float* dw = p.dw + (int64_t)zb * p.w_bs;
  const int64_t m_begin = (int64_t)split * p.rows_per_cta;
  const int64_t m_end = m_begin + p.rows_per_cta < p.M ? m_begin + p.rows_per_cta : p.M;

  const int tn = tid / 16, tk = tid % 16;          // 4 n x 4 k outputs per thread
  const int lr = tid / 16, lc = (tid % 16) * 4;    // loader: row lr, columns lc..lc+3
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};

  for (int64_t mc = m_begin; mc < m_end; mc += WG_MC) {
    const int64_t m = mc + lr;
    float gv[4] = {0.f, 0.f, 0.f, 0.f}, av[4] = {0.f, 0.f, 0.f, 0.f};
    if (m < m_end) {
      const int64_t goff = out_row_offset<KIND>(p, m, phase, p.ldg);
      if (goff >= 0) {
        if (n0 + lc < p.N) load4(g + goff + n0 + lc, gv);
        const int k = k0 + lc;
        if (k < p.K) {
          const int64_t off = a_offset<KIND>(p, m, k, phase);
          if (off >= 0) {
            load4(x + off, av);
            if (KIND == KIND_LINEAR && p.prologue != TFSWA_PRO_NONE) {
              if (p.prologue & TFSWA_PRO_AFFINE) {
#pragma unroll
                for (int j = 0; j < 4; ++j) av[j] = av[j] * p.in_scale[k + j] + p.in_shift[k + j];
              }
              if (p.prologue & TFSWA_PRO_LNHAT) {
                const float mean = rs[m * 2], rstd = rs[m * 2 + 1];
#pragma unroll
                for (int j = 0; j < 4; ++j) av[j] = (av[j] - mean) * rstd;
              }
              if (p.prologue & TFSWA_PRO_GELU) {
#pragma unroll
                for (int j = 0; j < 4; ++j) av[j] = gelu_erf(av[j]);
              }
            }
          }
        }
      }
    }
    __syncthreads();
    *reinterpret_cast<float4*>(&Gs[lr][lc]) = make_float4(gv[0], gv[1], gv[2], gv[3]);
    *reinterpret_cast<float4*>(&As[lr][lc]) = make_float4(av[0], av[1], av[2], av[3]);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < WG_MC; ++r) {
      const float4 g4 = *reinterpret_cast<const float4*>(&Gs[r][tn * 4]);
      const float4 a4 = *reinterpret_cast<const float4*>(&As[r][tk * 4]);
      const float gg[4] = {g4.x, g4.y, g4.z, g4.w};
      const float aa[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(gg[i], aa[j], acc[i][j]);
        if (tk == 0) bsum[i] += gg[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + tn * 4 + i;
    if (n >= p.N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tk * 4 + j;
      if (k < p.K) atomicAdd(dw + (int64_t)n * p.K + k, acc[i][j]);
    }
    if (p.dbias && tk == 0 && blockIdx.x == 0 && (KIND != KIND_UP || true))
      atomicAdd(p.dbias + (int64_t)(KIND == KIND_LINEAR ? zb : 0) * p.bias_bs + n, bsum[i]);
  }
}

template <typename T, int KIND>
static int launch_wgrad(IgemmParams& p, int zdim, cudaStream_t st) {
  const int kt = (p.K + WG_T - 1) / WG_T, nt = (p.N + WG_T - 1) / WG_T;
  int64_t want = 148 * 16 / ((int64_t)kt * nt * zdim);
  if (want < 1) want = 1;
  int64_t max_split = (p.M + 255) / 256;
  if (want > max_split) want = max_split;
  if (want > 4096) want = 4096;
  p.msplit = (int)want;
  p.rows_per_cta = ((p.M + p.msplit - 1) / p.msplit + WG_MC - 1) / WG_MC * WG_MC;
  p.msplit = (int)((p.M + p.rows_per_cta - 1) / p.rows_per_cta);
  dim3 grid(kt, nt, zdim * p.msplit);
  wgrad_kernel<T, KIND><<<grid, NTHREADS, 0, st>>>(p);
  return check_launch("wgrad");
}

}  // namespace tfswa

using namespace tfswa;

extern "C" {

int tfswa_linear_fwd(const tfswa_linear_args* a, void* stream) {
  TFSWA_REQUIRE(a && a->x && a->w && a->y, "linear: null pointer");
  TFSWA_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0 && a->batch > 0, "linear: empty problem");
  TFSWA_REQUIRE(a->K % 8 == 0 && a->N % 4 == 0, "linear: need K%%8==0 and N%%4==0 (K=%d N=%d)", a->K, a->N);
  TFSWA_REQUIRE(a->ldx % 8 == 0 && a->ldy % 4 == 0 && a->x_bs % 8 == 0 && a->y_bs % 4 == 0,
                "linear: leading dims / batch strides must keep 16-byte alignment");
  TFSWA_REQUIRE(!(a->prologue & TFSWA_PRO_LNHAT) || a->row_stats, "linear: PRO_LNHAT needs row_stats");
  TFSWA_REQUIRE(!(a->prologue & TFSWA_PRO_AFFINE) || (a->in_scale && a->in_shift), "linear: PRO_AFFINE needs in_scale/in_shift");
  TFSWA_REQUIRE(!a->r1 || a->ldr1 % 4 == 0, "linear: ldr1 alignment");
  TFSWA_REQUIRE(!a->r2 || a->ldr2 % 4 == 0, "linear: ldr2 alignment");
  TFSWA_REQUIRE(!a->pre || a->ldpre % 4 == 0, "linear: ldpre alignment");
  TFSWA_REQUIRE(!a->col_stats || a->batch == 1, "linear: col_stats only with batch==1");
  TFSWA_REQUIRE(a->batch <= 65535, "linear: batch too large");
  IgemmParams p = {};
  p.x = a->x; p.ldx = a->ldx; p.w = a->w; p.bias = a->bias; p.row_stats = a->row_stats;
  p.in_scale = a->in_scale; p.in_shift = a->in_shift;
  p.r1 = a->r1; p.ldr1 = a->ldr1; p.r2 = a->r2; p.ldr2 = a->ldr2;
  p.y = a->y; p.ldy = a->ldy; p.pre = a->pre; p.ldpre = a->ldpre; p.col_stats = a->col_stats;
  p.M = a->M; p.N = a->N; p.K = a->K; p.prologue = a->prologue; p.epilogue = a->epilogue;
  p.x_bs = a->x_bs; p.w_bs = a->w_bs; p.bias_bs = a->bias_bs; p.rs_bs = a->rs_bs; p.r1_bs = a->r1_bs; p.r2_bs = a->r2_bs;
  p.y_bs = a->y_bs; p.pre_bs = a->pre_bs;
  if (a->dtype == TFSWA_F32) return launch_igemm<float, KIND_LINEAR>(p, a->batch, (cudaStream_t)stream);
  if (a->dtype == TFSWA_BF16) return launch_igemm<bf16, KIND_LINEAR>(p, a->batch, (cudaStream_t)stream);
  TFSWA_REQUIRE(false, "linear: bad dtype %d", a->dtype);
}

int tfswa_conv_fwd(const tfswa_conv_args* a, void* stream) {
  TFSWA_REQUIRE(a && a->x && a->w && a->y, "conv: null pointer");
  TFSWA_REQUIRE(a->B > 0 && a->Hin > 0 && a->Win > 0, "conv: empty problem");
  TFSWA_REQUIRE(a->Cin % 16 == 0 && a->Cout % 4 == 0, "conv: need Cin%%16==0, Cout%%4==0 (Cin=%d Cout=%d)", a->Cin, a->Cout);
  IgemmParams p = {};
  p.x = a->x; p.w = a->w; p.bias = a->bias; p.y = a->y; p.pre = a->pre; p.col_stats = a->col_stats;
  p.ldy = a->Cout; p.ldpre = a->Cout; p.N = a->Cout; p.epilogue = a->epilogue;
  p.B = a->B; p.Hin = a->Hin; p.Win = a->Win; p.Cin = a->Cin; p.Hout = a->Hout; p.Wout = a->Wout;
  int kind, zdim = 1;
  if (a->kind == 0) {
    TFSWA_REQUIRE(a->Hout == a->Hin && a->Wout == a->Win, "conv3x3: output size must equal input size");
    kind = KIND_CONV3; p.K = 9 * a->Cin; p.M = (int64_t)a->B * a->Hout * a->Wout;
  } else if (a->kind == 1) {
    TFSWA_REQUIRE(a->Hout == (a->Hin - 2) / 2 + 1 && a->Wout == (a->Win - 2) / 2 + 1, "down conv: bad output size");
    kind = KIND_DOWN; p.K = 16 * a->Cin; p.M = (int64_t)a->B * a->Hout * a->Wout;
  } else if (a->kind == 2) {
    // forward ConvTranspose2d gives exactly 2x; as the data-gradient of the stride-2 conv the output is the conv's
    // (possibly odd) input size, 2*Hin or 2*Hin+1
    TFSWA_REQUIRE(a->Hout >= 2 * a->Hin && a->Hout <= 2 * a->Hin + 1 && a->Wout >= 2 * a->Win && a->Wout <= 2 * a->Win + 1,
                  "up conv: output must be 2x (or 2x+1) the input");
    p.Hq = (a->Hout + 1) / 2; p.Wq = (a->Wout + 1) / 2;
    kind = KIND_UP; p.K = 4 * a->Cin; p.M = (int64_t)a->B * p.Hq * p.Wq; zdim = 4;
    p.w_bs = (int64_t)a->Cout * 4 * a->Cin;
  } else TFSWA_REQUIRE(false, "conv: bad kind %d", a->kind);
  cudaStream_t st = (cudaStream_t)stream;
#define TFSWA_DISPATCH(T)                                                  \
  switch (kind) {                                                          \
    case KIND_CONV3: return launch_igemm<T, KIND_CONV3>(p, zdim, st);      \
    case KIND_DOWN: return launch_igemm<T, KIND_DOWN>(p, zdim, st);        \
    default: return launch_igemm<T, KIND_UP>(p, zdim, st);                 \
  }
  if (a->dtype == TFSWA_F32) { TFSWA_DISPATCH(float) }
  if (a->dtype == TFSWA_BF16) { TFSWA_DISPATCH(bf16) }
#undef TFSWA_DISPATCH
  TFSWA_REQUIRE(false, "conv: bad dtype %d", a->dtype);
}


int tfswa_linear_wgrad(const tfswa_linear_args* a, const void* g, int64_t ldg, int64_t g_bs, float* dw, float* dbias, void* stream) {
  TFSWA_REQUIRE(a && a->x && g && dw, "linear_wgrad: null pointer");
  TFSWA_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0 && a->batch > 0 && a->batch <= 16, "linear_wgrad: bad problem size");
  TFSWA_REQUIRE(a->K % 4 == 0 && a->N % 4 == 0 && a->ldx % 4 == 0 && ldg % 4 == 0 && a->x_bs % 4 == 0 && g_bs % 4 == 0,
                "linear_wgrad: alignment (multiples of 4 elements)");
  TFSWA_REQUIRE(!(a->prologue & TFSWA_PRO_LNHAT) || a->row_stats, "linear_wgrad: PRO_LNHAT needs row_stats");
  if (a->dtype == TFSWA_BF16) {                      // tensor-core paths; 1 = shape not covered, fall through
    static int force_mma = -1;                       // TFSWA_WGRAD=mma: skip the tcgen05 kernel (A/B)
    if (force_mma < 0) { const char* e = getenv("TFSWA_WGRAD"); force_mma = (e && e[0] == 'm') ? 1 : 0; }
    int rc = force_mma ? 1 : wgrad_tc_bf16(a, g, ldg, g_bs, dw, dbias, (cudaStream_t)stream);   // tcgen05 + TMA (wgrad_tc.cu)
    if (rc != 1) return rc;
    rc = wgrad_mma_bf16(a, g, ldg, g_bs, dw, dbias, (cudaStream_t)stream);                        // warp-level MMA (wgrad_mma.cu)
    if (rc != 1) return rc;
  }
  IgemmParams p = {};
  p.x = a->x; p.ldx = a->ldx; p.x_bs = a->x_bs; p.row_stats = a->row_stats; p.rs_bs = a->rs_bs;
  p.in_scale = a->in_scale; p.in_shift = a->in_shift; p.prologue = a->prologue;
  p.g = g; p.ldg = ldg; p.g_bs = g_bs; p.dw = dw; p.dbias = dbias; p.bias_bs = a->N; p.w_bs = (int64_t)a->N * a->K;
  p.M = a->M; p.N = a->N; p.K = a->K;
  if (a->dtype == TFSWA_F32) return launch_wgrad<float, KIND_LINEAR>(p, a->batch, (cudaStream_t)stream);
  if (a->dtype == TFSWA_BF16) return launch_wgrad<bf16, KIND_LINEAR>(p, a->batch, (cudaStream_t)stream);
  TFSWA_REQUIRE(false, "linear_wgrad: bad dtype %d", a->dtype);
}

int tfswa_conv_wgrad(const tfswa_conv_args* a, const void* g, float* dw, float* dbias, void* stream) {
  TFSWA_REQUIRE(a && a->x && g && dw, "conv_wgrad: null pointer");
  TFSWA_REQUIRE(a->Cin % 16 == 0 && a->Cout % 4 == 0, "conv_wgrad: need Cin%%16==0, Cout%%4==0");
  IgemmParams p = {};
  p.x = a->x; p.g = g; p.ldg = a->Cout; p.dw = dw; p.dbias = dbias; p.N = a->Cout;
  p.B = a->B; p.Hin = a->Hin; p.Win = a->Win; p.Cin = a->Cin; p.Hout = a->Hout; p.Wout = a->Wout;
  int kind, zdim = 1;
  if (a->kind == 0) { kind = KIND_CONV3; p.K = 9 * a->Cin; p.M = (int64_t)a->B * a->Hout * a->Wout; }
  else if (a->kind == 1) { kind = KIND_DOWN; p.K = 16 * a->Cin; p.M = (int64_t)a->B * a->Hout * a->Wout; }
  else if (a->kind == 2) {
    p.Hq = (a->Hout + 1) / 2; p.Wq = (a->Wout + 1) / 2;
    kind = KIND_UP; p.K = 4 * a->Cin; p.M = (int64_t)a->B * p.Hq * p.Wq; zdim = 4;
    p.w_bs = (int64_t)a->Cout * 4 * a->Cin;
  } else TFSWA_REQUIRE(false, "conv_wgrad: bad kind %d", a->kind);
  cudaStream_t st = (cudaStream_t)stream;
  if (a->dtype == TFSWA_BF16 && !getenv("TFSWA_CONV_WGRAD_SIMT")) {   // tensor-core path (wgrad_mma.cu); 1 = shape not covered
    const int rc = conv_wgrad_mma_bf16(a, g, dw, dbias, st);
    if (rc != 1) return rc;
  }
#define TFSWA_DISPATCH(T)                                                  \
  switch (kind) {                                                          \
    case KIND_CONV3: return launch_wgrad<T, KIND_CONV3>(p, zdim, st);      \
    case KIND_DOWN: return launch_wgrad<T, KIND_DOWN>(p, zdim, st);        \
    default: return launch_wgrad<T, KIND_UP>(p, zdim, st);                 \
  }
  if (a->dtype == TFSWA_F32) { TFSWA_DISPATCH(float) }
  if (a->dtype == TFSWA_BF16) { TFSWA_DISPATCH(bf16) }
#undef TFSWA_DISPATCH
  TFSWA_REQUIRE(false, "conv_wgrad: bad dtype %d", a->dtype);
}

}  // extern "C"
