// Weight gradient of a (batched) token linear on the tensor cores, bf16 activations:
//     dW[b][n][k] += sum_m G[m,b,n] * pro(X)[m,b,k],      dbias[b][n] += sum_m G[m,b,n]
// (the backward of nn.Linear / 1x1 conv in attention.py:70,86,121-128 and blocks.py:53-56,85-89).
//
// The reduction runs over millions of tokens while the result is at most 1152 x 1024, so the GEMM is "tall": one CTA
// owns a 64 (n) x 64 (k) output tile and a slice of the tokens (split-M), keeps the tile in MMA accumulators and
// finishes with fp32 atomics.  Both operands are token-major in memory (G: tokens x N, X: tokens x K) while the MMA
// wants the tokens along its K dimension, so the fragments are read from the shared-memory token tiles with
// ldmatrix.trans - no transposed copy is ever made.  The LayerNorm / GELU / BatchNorm-affine prologues of the forward
// are applied while X is staged (fp32 math, then rounded to bf16 for the MMA).  The bias gradient rides along as one
// more MMA against a ones operand.  Replaces the CUDA-core wgrad_kernel (igemm.cu) for bf16, where the weight
// gradients were 28 % of a training step.
#include "common.cuh"

namespace tfswa {

struct WgradMmaParams {
  const bf16* x; int64_t ldx, x_bs;
  const bf16* g; int64_t ldg, g_bs;
  const float* row_stats; int64_t rs_bs;
  const float* in_scale; const float* in_shift;
  float* dw; int64_t w_bs;
  float* dbias; int64_t bias_bs;
  int64_t M, rows_per_cta;
  int N, K, prologue, msplit;
  // convolution mode (implicit GEMM over output pixels; X rows are gathered per filter tap, see igemm.cu a_offset)
  int kind;                  // 0: 3x3 s1 p1, 1: 4x4 s2 p1, 2: one phase of the 4x4 s2 transposed conv (blockIdx.z / msplit = phase)
  int B, Hin, Win, Cin, Hout, Wout, Hq, Wq, cblocks;
};

constexpr int WM_T = 64;            // output tile edge (n and k)
constexpr int WM_MC = 64;           // tokens per shared-memory chunk
constexpr int WM_PITCH = WM_T + 8;  // padded row (conflict-free ldmatrix)
constexpr int WM_THREADS = 256;

__device__ __forceinline__ void wm_mma(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void wm_ldsm_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void wm_ldsm_x2_trans(uint32_t& r0, uint32_t& r1, const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

constexpr int WM_STAGES = 3;        // cp.async ring depth: two chunks (2 x 18 KB per CTA) in flight under the MMAs of the third
constexpr int WM_SMEM = WM_STAGES * (2 * WM_MC * WM_PITCH * 2 + WM_MC * 8);

// 16-byte asynchronous copy global -> shared; bytes = 0 zero-fills the destination (rows / columns outside the problem)
__device__ __forceinline__ void wm_cp_async16(void* smem_dst, const void* gsrc, int bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void wm_cp_async8(void* smem_dst, const void* gsrc, int bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc), "r"(bytes) : "memory");
}

// The kernel is a stream over the tokens at an arithmetic intensity of ~50 FLOP/B (far below the ridge), i.e. HBM-bound:
// what matters is bytes in flight.  G and X chunks go global -> shared with cp.async through a WM_STAGES-deep ring (the
// first version staged one chunk through registers and was latency-bound: 1.4 TB/s, long-scoreboard stall 10 per issue,
// profiles/r1zz_train_ncu_summary.md); the prologue, when there is one, is applied in place in shared memory.
template <bool CONV>
__global__ void __launch_bounds__(WM_THREADS) wgrad_mma_kernel(const WgradMmaParams p) {
  extern __shared__ __align__(16) uint8_t wm_smem[];
  typedef bf16 (*tile_t)[WM_MC][WM_PITCH];
  tile_t Gs = reinterpret_cast<tile_t>(wm_smem);
  tile_t As = reinterpret_cast<tile_t>(wm_smem + WM_STAGES * WM_MC * WM_PITCH * 2);
  float2 (*Rs)[WM_MC] = reinterpret_cast<float2 (*)[WM_MC]>(wm_smem + 2 * WM_STAGES * WM_MC * WM_PITCH * 2);   // (mean, rstd) per token
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // linear: k-tile = 64 columns of X.  conv: k-tile = (filter tap, 64-channel block); zb = phase of the transposed conv
  const int tap = CONV ? blockIdx.x / p.cblocks : 0;
  const int k0 = CONV ? (blockIdx.x - tap * p.cblocks) * WM_T : blockIdx.x * WM_T, n0 = blockIdx.y * WM_T;
  const int kdim = CONV ? p.Cin : p.K;                               // valid extent of this k-tile's column index
  const int zb = blockIdx.z / p.msplit, split = blockIdx.z % p.msplit;
  const bf16* x = p.x + (int64_t)zb * p.x_bs;
  const bf16* g = p.g + (int64_t)zb * p.g_bs;
  const float* rs = p.row_stats ? p.row_stats + (int64_t)zb * p.rs_bs : nullptr;
  const int64_t m_begin = (int64_t)split * p.rows_per_cta;
  const int64_t m_end = m_begin + p.rows_per_cta < p.M ? m_begin + p.rows_per_cta : p.M;
  const int nchunks = (int)((m_end - m_begin + WM_MC - 1) / WM_MC);

  // staging: thread -> (token row lr + 32 i, 8 columns at lc), i = 0, 1, for G and for X
  const int lr = tid >> 3, lc = (tid & 7) * 8;
  const bool g_ok = n0 + lc < p.N, a_ok = k0 + lc < kdim;         // N, K, Cin are multiples of 8
  // compute: warp -> 16 n rows (nt) x 32 k columns (kh)
  const int nt = warp & 3, kh = warp >> 2;
  float acc[4][4], bacc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

  auto issue = [&](int c) {
    if (c < nchunks) {
      const int slot = c % WM_STAGES;
      const int64_t mc = m_begin + (int64_t)c * WM_MC;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int r = lr + 32 * i;
        const int64_t m = mc + r;
        const bool in = m < m_end;
        const int64_t ms = in ? m : m_begin;                       // any valid address when nothing is copied
        if (!CONV) {
          wm_cp_async16(&Gs[slot][r][lc], g + ms * p.ldg + (g_ok ? n0 + lc : 0), (in && g_ok) ? 16 : 0);
          wm_cp_async16(&As[slot][r][lc], x + ms * p.ldx + (a_ok ? k0 + lc : 0), (in && a_ok) ? 16 : 0);
          if (rs && lc == 0) wm_cp_async8(&Rs[slot][r], rs + ms * 2, in ? 8 : 0);
        } else {
          // output pixel (or phase-grid cell) of row m, the input pixel this tap reads, zero-fill outside the image
          const int wdim = p.kind == 2 ? p.Wq : p.Wout, hw = (p.kind == 2 ? p.Hq : p.Hout) * wdim;
          const int b = (int)(ms / hw), rem = (int)(ms - (int64_t)b * hw);
          const int oy = rem / wdim, ox = rem - oy * wdim;
          int iy, ix;
          int64_t goff = ms * p.ldg;
          bool g_in = in;
          if (p.kind == 0) { iy = oy + tap / 3 - 1; ix = ox + tap % 3 - 1; }
          else if (p.kind == 1) { iy = 2 * oy + (tap >> 2) - 1; ix = 2 * ox + (tap & 3) - 1; }
          else {
            const int py = zb >> 1, px = zb & 1, ta = tap >> 1, tb = tap & 1;
            iy = py ? (ta ? oy : oy + 1) : (ta ? oy - 1 : oy);
            ix = px ? (tb ? ox : ox + 1) : (tb ? ox - 1 : ox);
            const int yy = 2 * oy + py, xx = 2 * ox + px;
            g_in = in && yy < p.Hout && xx < p.Wout;                // phase-grid cell outside the output
            goff = g_in ? (((int64_t)b * p.Hout + yy) * p.Wout + xx) * p.ldg : 0;
          }
          const bool x_in = g_in && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
          const int64_t xoff = x_in ? (((int64_t)b * p.Hin + iy) * p.Win + ix) * p.Cin : 0;
          wm_cp_async16(&Gs[slot][r][lc], g + goff + (g_ok ? n0 + lc : 0), (g_in && g_ok) ? 16 : 0);
          wm_cp_async16(&As[slot][r][lc], x + xoff + (a_ok ? k0 + lc : 0), (x_in && a_ok) ? 16 : 0);
        }
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  constexpr uint32_t ONES = 0x3F803F80u;
#pragma unroll
  for (int c = 0; c < WM_STAGES - 1; ++c) issue(c);
  for (int c = 0; c < nchunks; ++c) {
    const int slot = c % WM_STAGES;
    asm volatile("cp.async.wait_group %0;" ::"n"(WM_STAGES - 2) : "memory");
    __syncthreads();                                               // chunk c visible; chunk c-1's fragments have been read
    issue(c + WM_STAGES - 1);                                      // refills the slot chunk c-1 used
    if (p.prologue != TFSWA_PRO_NONE) {                            // pro(X) in place (fp32 math, rounded to bf16 for the MMA)
      if (a_ok) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int r = lr + 32 * i;
          float v[8];
          load8(&As[slot][r][lc], v);
          if (p.prologue & TFSWA_PRO_AFFINE) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = v[j] * p.in_scale[k0 + lc + j] + p.in_shift[k0 + lc + j];
          }
          if (p.prologue & TFSWA_PRO_LNHAT) {
            const float2 st = Rs[slot][r];                         // zero-filled (mean 0, rstd 0) for rows past m_end
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = (v[j] - st.x) * st.y;
          }
          if (p.prologue & TFSWA_PRO_GELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = gelu_erf(v[j]);
          }
          if (m_begin + (int64_t)c * WM_MC + r >= m_end) {         // affine / GELU of a zero-filled row must stay zero
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = 0.f;
          }
          store8(&As[slot][r][lc], v);
        }
      }
      __syncthreads();
    }
#pragma unroll
    for (int ks = 0; ks < WM_MC / 16; ++ks) {
      // A = G^T: (n rows, token cols).  x4.trans blocks: (tokens 0-7, n 0-7), (tokens 0-7, n 8-15), (tokens 8-15, n 0-7), (tokens 8-15, n 8-15)
      uint32_t a[4];
      wm_ldsm_x4_trans(a, &Gs[slot][ks * 16 + (lane & 7) + ((lane >> 4) << 3)][nt * 16 + ((lane >> 3) & 1) * 8]);
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        uint32_t b0, b1;                                           // B = X: tokens 0-7 / 8-15 x 8 k columns
        wm_ldsm_x2_trans(b0, b1, &As[slot][ks * 16 + (lane & 15)][kh * 32 + kt * 8]);
        wm_mma(acc[kt], a, b0, b1);
      }
      if (kh == 0) wm_mma(bacc, a, ONES, ONES);                    // every column of bacc = sum over tokens of G
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // ---- split-M partial -> global with fp32 atomics ----
  const int gq = lane >> 2, tq = lane & 3;
  float* dw = p.dw + (int64_t)zb * p.w_bs;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int n = n0 + nt * 16 + gq + h * 8;
    if (n >= p.N) continue;
#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      const int k = k0 + kh * 32 + kt * 8 + 2 * tq;
      if (k < kdim) {
        float* dst = dw + (int64_t)n * p.K + (CONV ? tap * p.Cin : 0) + k;
        atomicAdd(dst, acc[kt][h * 2]);
        atomicAdd(dst + 1, acc[kt][h * 2 + 1]);
      }
    }
    if (p.dbias && kh == 0 && tq == 0 && blockIdx.x == 0) atomicAdd(p.dbias + (int64_t)zb * p.bias_bs + n, bacc[h * 2]);
  }
}

// bf16 linear weight gradient; returns TFSWA_EINVAL-free "not handled" (1) when the shape is outside this kernel
int wgrad_mma_bf16(const tfswa_linear_args* a, const void* g, int64_t ldg, int64_t g_bs, float* dw, float* dbias, cudaStream_t st) {
  if (a->K % 8 || a->N % 8 || a->ldx % 8 || ldg % 8 || a->x_bs % 8 || g_bs % 8 || (((uintptr_t)a->x | (uintptr_t)g) & 15)) return 1;
  WgradMmaParams p = {};
  p.x = (const bf16*)a->x; p.ldx = a->ldx; p.x_bs = a->x_bs; p.g = (const bf16*)g; p.ldg = ldg; p.g_bs = g_bs;
  p.row_stats = a->row_stats; p.rs_bs = a->rs_bs; p.in_scale = a->in_scale; p.in_shift = a->in_shift;
  p.dw = dw; p.w_bs = (int64_t)a->N * a->K; p.dbias = dbias; p.bias_bs = a->N;
  p.M = a->M; p.N = a->N; p.K = a->K; p.prologue = a->prologue;
  const int kt = (p.K + WM_T - 1) / WM_T, nt = (p.N + WM_T - 1) / WM_T;
  int64_t want = 148 * 8 / ((int64_t)kt * nt * a->batch);
  if (want < 1) want = 1;
  const int64_t max_split = (p.M + 511) / 512;
  if (want > max_split) want = max_split;
  if (want > 4096) want = 4096;
  p.rows_per_cta = ((p.M + want - 1) / want + WM_MC - 1) / WM_MC * WM_MC;
  p.msplit = (int)((p.M + p.rows_per_cta - 1) / p.rows_per_cta);
  dim3 grid(kt, nt, a->batch * p.msplit);
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    if (cudaFuncSetAttribute(wgrad_mma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, WM_SMEM) != cudaSuccess) {
      set_error("wgrad_mma: cudaFuncSetAttribute failed");
      return TFSWA_ECUDA;
    }
    attr_once.done();
  }
  wgrad_mma_kernel<false><<<grid, WM_THREADS, WM_SMEM, st>>>(p);
  return check_launch("linear_wgrad");
}

// bf16 convolution weight gradient (3x3 / 4x4-stride-2 / 4-phase transposed conv; dw in the layout of tfswa_conv_args.w);
// returns 1 when the shape is outside this kernel
int conv_wgrad_mma_bf16(const tfswa_conv_args* a, const void* g, float* dw, float* dbias, cudaStream_t st) {
  if (a->Cin % 8 || a->Cout % 8 || (((uintptr_t)a->x | (uintptr_t)g) & 15) || a->kind < 0 || a->kind > 2) return 1;
  WgradMmaParams p = {};
  p.x = (const bf16*)a->x; p.g = (const bf16*)g; p.ldg = a->Cout; p.dw = dw; p.dbias = dbias; p.bias_bs = 0;
  p.kind = a->kind; p.B = a->B; p.Hin = a->Hin; p.Win = a->Win; p.Cin = a->Cin; p.Hout = a->Hout; p.Wout = a->Wout;
  p.N = a->Cout; p.prologue = TFSWA_PRO_NONE;
  int taps, zdim = 1;
  if (a->kind == 0) { taps = 9; p.M = (int64_t)a->B * a->Hout * a->Wout; }
  else if (a->kind == 1) { taps = 16; p.M = (int64_t)a->B * a->Hout * a->Wout; }
  else { taps = 4; p.Hq = (a->Hout + 1) / 2; p.Wq = (a->Wout + 1) / 2; p.M = (int64_t)a->B * p.Hq * p.Wq; zdim = 4; }
  p.K = taps * a->Cin;
  p.w_bs = (int64_t)a->Cout * p.K;                     // phase stride of the transposed conv's (4, Cout, 2, 2, Cin) layout
  p.cblocks = (a->Cin + WM_T - 1) / WM_T;
  const int kt = taps * p.cblocks, nt = (p.N + WM_T - 1) / WM_T;
  int64_t want = 148 * 8 / ((int64_t)kt * nt * zdim);
  if (want < 1) want = 1;
  const int64_t max_split = (p.M + 511) / 512;
  if (want > max_split) want = max_split;
  p.rows_per_cta = ((p.M + want - 1) / want + WM_MC - 1) / WM_MC * WM_MC;
  p.msplit = (int)((p.M + p.rows_per_cta - 1) / p.rows_per_cta);
  dim3 grid(kt, nt, zdim * p.msplit);
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    if (cudaFuncSetAttribute(wgrad_mma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, WM_SMEM) != cudaSuccess) {
      set_error("conv_wgrad_mma: cudaFuncSetAttribute failed");
      return TFSWA_ECUDA;
    }
    attr_once.done();
  }
  wgrad_mma_kernel<true><<<grid, WM_THREADS, WM_SMEM, st>>>(p);
  return check_launch("conv_wgrad");
}

}  // namespace tfswa
