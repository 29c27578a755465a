// Weight gradient of a (batched) token linear on the tcgen05 tensor cores (bf16 activations):
//     dW[b][n][k] += sum_m G[m,b,n] * pro(X)[m,b,k],      dbias[b][n] += sum_m G[m,b,n]
// (the backward of nn.Linear / 1x1 conv in attention.py:70,86,121-128 and blocks.py:53-56,85-89; autograd row a13).
//
// The reduction runs over millions of tokens while the result is at most 1152 x 1024: the kernel is a stream over the
// tokens, HBM-bound (~50 FLOP/B).  Both operands are TOKEN-major in memory (G: tokens x N, X: tokens x K) while the MMA
// reduces over tokens, i.e. both are MN-MAJOR UMMA operands (the layout the attention kernel's V operand uses): a TMA box
// (32 or 64 columns x 64 / 128 tokens, 64B / 128B swizzle) lands in shared memory exactly as `tcgen05.mma` wants it, with the
// token axis as the MMA's K.  No thread touches the operands unless the forward had a LayerNorm prologue: then four warps
// rewrite the X boxes in place ((x - mean) * rstd per token, swizzle-agnostic because the op is per row) and hand the
// stage over with a proxy fence.
//   D (TMEM, 128 n x <= 256 k fp32) += G_tile^T (M = 128 n, K = 16 tokens) x X_tile (N = k columns, K = 16 tokens)
// The bias gradient rides along as one more MMA against a ones operand.  One CTA = one 128 x <= 256 output tile and a
// slice of the tokens (split-M); a 4-stage TMA ring; one elected lane issues the MMAs; the epilogue adds the tile to the
// caller-zeroed fp32 result with atomics (the outputs are small: 9 k .. 200 k elements x split).
// Replaces `wgrad_mma_kernel` (warp-level mma.sync + cp.async, wgrad_mma.cu) for the linear shapes it covers.
#include "common.cuh"
#include "sm100.cuh"

namespace tfswa {

using namespace sm100;

namespace wgtc {

constexpr int MAX_TOK = 256;            // tokens per stage: 64 / 128 / 256, chosen per shape so that a stage is 32-48 KB
constexpr int NT = 128;                 // output rows (n) per CTA = UMMA M
constexpr int STAGES = 4;
#ifndef TFSWA_WGTC_TW
#define TFSWA_WGTC_TW 8
#endif
constexpr int TW = TFSWA_WGTC_TW;       // transform / epilogue warps (a multiple of 4)
constexpr int THREADS = 64 + TW * 32;   // warp 0: TMA, warp 1: MMA issue + TMEM alloc, warps 2-9: prologue transform + epilogue

struct Params {
  const float* row_stats;               // (M, 2) per batch: (mean, rstd) per token (PRO_LNHAT) or nullptr
  int64_t rs_bs;
  float* dw; int64_t w_bs;              // (N, K) fp32 per batch
  float* dbias; int64_t bias_bs;
  int64_t M, rows_per_cta;
  int N, K, msplit;
  int cbA, cbX;                         // columns per TMA box of G / X (32 -> 64B swizzle, 64 -> 128B swizzle)
  int BKW;                              // k columns per CTA (UMMA N): multiple of 16, <= 256
  int TOK;                              // tokens per stage (TMA box rows)
  uint32_t tmem_cols;
};

// MN-major operand in a swizzled TMA layout: boxes of `cb` columns ([token][cb * 2 bytes], swizzle span = row), tokens = K.
//   LBO = distance between column groups (one box), SBO = 8 token rows, layout type 2 (128B) / 4 (64B)
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, int cb, uint32_t box_bytes) {
  const uint32_t row = (uint32_t)cb * 2u;
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(box_bytes >> 4) << 16;
  d |= (uint64_t)((8u * row) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(row == 128 ? 2 : 4) << 61;
  return d;
}

__global__ void __launch_bounds__(THREADS, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_g, const __grid_constant__ CUtensorMap tm_x, const Params p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[STAGES], bar_ready[STAGES], bar_empty[STAGES], bar_acc;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) uint16_t s_ones[256];           // 16 x 16 bf16 ones (K-major core matrices: any layout of ones is ones)
  __shared__ __align__(8) float2 s_rs[STAGES][MAX_TOK];     // LayerNorm prologue: (mean, rstd) of a stage's tokens, cp.async ring

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k0 = blockIdx.x * p.BKW, n0 = blockIdx.y * NT;
  const int zb = blockIdx.z / p.msplit, split = blockIdx.z % p.msplit;
  const int64_t m_begin = (int64_t)split * p.rows_per_cta;
  const int64_t m_end = m_begin + p.rows_per_cta < p.M ? m_begin + p.rows_per_cta : p.M;
  const int TOK = p.TOK;
  const uint32_t A_BYTES = (uint32_t)NT * TOK * 2;           // G tile (boxes of cbA columns, [token][cbA * 2 B] each)
  const int nchunks = m_end > m_begin ? (int)((m_end - m_begin + TOK - 1) / TOK) : 0;
  const int bkw = min(p.BKW, p.K - k0);                     // valid k columns of this tile (multiple of 16)
  const uint32_t x_bytes = (uint32_t)TOK * p.BKW * 2;
  const uint32_t stage_bytes = A_BYTES + x_bytes;
  const bool lnhat = p.row_stats != nullptr;
  const bool want_bias = p.dbias != nullptr && blockIdx.x == 0;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_g); prefetch_tmap(&tm_x);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_ready[s], TW); mbar_init(&bar_empty[s], 1); }
    mbar_init(&bar_acc, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&s_tmem, p.tmem_cols);
  for (int i = tid; i < 256; i += THREADS) s_ones[i] = 0x3F80;
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer ----------------
      const int nboxA = NT / p.cbA, nboxX = p.BKW / p.cbX;
      const uint32_t boxA = (uint32_t)TOK * p.cbA * 2, boxX = (uint32_t)TOK * p.cbX * 2;
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % STAGES;
        if (c >= STAGES) mbar_wait(&bar_empty[s], ((c / STAGES) - 1) & 1);
        mbar_arrive_expect_tx(&bar_full[s], stage_bytes);
        uint8_t* dst = sm + (size_t)s * stage_bytes;
        const int m = (int)(m_begin + (int64_t)c * TOK);
        for (int b = 0; b < nboxA; ++b) tma_load_3d(dst + b * boxA, &tm_g, &bar_full[s], n0 + b * p.cbA, m, zb);
        for (int b = 0; b < nboxX; ++b) tma_load_3d(dst + A_BYTES + b * boxX, &tm_x, &bar_full[s], k0 + b * p.cbX, m, zb);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------- MMA issuer ----------------
      // kind::f16, D = f32, A = B = bf16, both MN-major (bits 15, 16); the bias MMA takes a K-major ones operand
      const uint32_t idesc = umma_idesc_bf16(NT, p.BKW) | (1u << 15) | (1u << 16);
      const uint32_t idesc_b = umma_idesc_bf16(NT, 16) | (1u << 15);
      const uint32_t boxA = (uint32_t)TOK * p.cbA * 2, boxX = (uint32_t)TOK * p.cbX * 2;
      const uint32_t stepA = 16u * p.cbA * 2, stepX = 16u * p.cbX * 2;       // 16 tokens along the MMA's K
      const uint64_t ones_desc = umma_smem_desc_ns(smem_u32(s_ones), 256, 128);
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % STAGES;
        mbar_wait(lnhat ? &bar_ready[s] : &bar_full[s], (c / STAGES) & 1);
        tc_fence_after();
        const uint32_t sa = base + s * stage_bytes;
        for (int k = 0; k < TOK / 16; ++k) {
          const uint64_t ad = desc_mn(sa + k * stepA, p.cbA, boxA);
          umma_bf16_ss(tmem, ad, desc_mn(sa + A_BYTES + k * stepX, p.cbX, boxX), idesc, (c | k) ? 1u : 0u);
          if (want_bias) umma_bf16_ss(tmem + 256, ad, ones_desc, idesc_b, (c | k) ? 1u : 0u);
        }
        umma_commit(&bar_empty[s]);          // frees this stage when the MMAs above have read it
      }
      umma_commit(&bar_acc);                 // accumulators complete
    }
  } else {
    constexpr int ET = TW * 32;
    const int et = tid - 64;                 // 0..ET-1: transform / epilogue threads; warp % 4 selects the TMEM lane quarter
    if (lnhat) {
      // ---------------- LayerNorm prologue on the X boxes, in place: x <- (x - mean_m) * rstd_m (0 beyond M) ----------------
      // (mean, rstd) of the tokens of chunk c + 2 are fetched with 4-byte cp.async while chunk c is transformed: a plain
      // load after the TMA barrier put a global-memory latency on every stage's critical path (1.5 TB/s instead of 3+)
      const float* rs = p.row_stats + (int64_t)zb * p.rs_bs;
      const uint32_t rowb = (uint32_t)p.cbX * 2;                     // bytes per token row inside a box
      const uint32_t cpr_shift = rowb == 128 ? 3u : 2u, chunks = x_bytes / 16;     // 16-byte chunks per token row: 8 or 4; TOK is 64 or 128
      const uint32_t tok_mask = (uint32_t)TOK - 1u;
      auto prefetch = [&](int c) {
        if (c < nchunks) {
          for (int row = et; row < TOK; row += ET) {
            int64_t mm = m_begin + (int64_t)c * TOK + row;
            if (mm >= p.M) mm = p.M - 1;                             // (value unused: the row is zeroed)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(&s_rs[c % STAGES][row])), "l"(rs + 2 * mm) : "memory");
          }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
      };
      prefetch(0); prefetch(1);
      for (int c = 0; c < nchunks; ++c) {
        const int s = c % STAGES;
        prefetch(c + 2);
        asm volatile("cp.async.wait_group 2;" ::: "memory");         // chunk c's statistics have landed (this thread's copies) ...
        asm volatile("bar.sync 1, %0;" ::"n"(ET) : "memory");         // ... and everybody else's
        mbar_wait(&bar_full[s], (c / STAGES) & 1);
        uint8_t* xs = sm + (size_t)s * stage_bytes + A_BYTES;
        const int64_t m = m_begin + (int64_t)c * TOK;
        for (uint32_t i = et; i < chunks; i += ET) {
          const uint32_t row = (i >> cpr_shift) & tok_mask;          // token inside the box (swizzle permutes chunks within a row only; no divisions here:
                                                                     // with `/ chunks_per_row % TOK` this loop was the kernel's bottleneck, 3100 cycles per stage)
          uint4* ptr = reinterpret_cast<uint4*>(xs + (size_t)i * 16);
          uint4 v = *ptr;
          if (m + row < p.M) {
            const float2 st = s_rs[s][row];
            const float mean = st.x, rstd = st.y;
            __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&v);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 f = __bfloat1622float2(h[j]);
              h[j] = __floats2bfloat162_rn((f.x - mean) * rstd, (f.y - mean) * rstd);
            }
          } else {
            v = make_uint4(0, 0, 0, 0);
          }
          *ptr = v;
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_ready[s]);
      }
    }
    // ---------------- epilogue: TMEM -> fp32 atomics ----------------
    mbar_wait(&bar_acc, 0);
    tc_fence_after();
    const int quad = warp & 3;
    const int n = n0 + quad * 32 + lane;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
    if (nchunks > 0) {
      float* dwrow = p.dw + (int64_t)zb * p.w_bs + (int64_t)n * p.K + k0;
      const int whalf = (warp - 2) >> 2;       // the TW / 4 warps of a lane quarter take alternate 16-column chunks
      for (int c0 = whalf * 16; c0 < bkw; c0 += (TW / 4) * 16) {
        uint32_t raw[16];
        __syncwarp();
        tmem_ld_x16(lane_addr + c0, raw);
        tmem_ld_wait();
        if (n < p.N) {
#pragma unroll
          for (int j = 0; j < 16; ++j) atomicAdd(dwrow + c0 + j, __uint_as_float(raw[j]));
        }
      }
      if (want_bias && whalf == 0) {
        uint32_t raw[16];
        __syncwarp();
        tmem_ld_x16(lane_addr + 256, raw);
        tmem_ld_wait();
        if (n < p.N) atomicAdd(p.dbias + (int64_t)zb * p.bias_bs + n, __uint_as_float(raw[0]));
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, p.tmem_cols);
}

}  // namespace wgtc

// token-major bf16 matrix (cols, M rows, batch) with boxes of `cb` columns x 64 tokens; swizzle span = cb * 2 bytes
static int make_tmap_tokens(CUtensorMap* out, const void* base, int cols, int64_t M, int batch, int64_t ld, int64_t bs, int cb, int tok) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return TFSWA_ECUDA; }
  cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)M, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? bs : ld * M) * 2};
  cuuint32_t box[3] = {(cuuint32_t)cb, (cuuint32_t)tok, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, cb == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("wgrad_tc: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return TFSWA_ECUDA; }
  return TFSWA_OK;
}

// bf16 linear weight gradient on tcgen05; returns 1 ("not handled") when the shape / prologue is outside this kernel
int wgrad_tc_bf16(const tfswa_linear_args* a, const void* g, int64_t ldg, int64_t g_bs, float* dw, float* dbias, cudaStream_t st) {
  using namespace wgtc;
  if (a->prologue != TFSWA_PRO_NONE && a->prologue != TFSWA_PRO_LNHAT) return 1;
  if (a->K % 32 || a->N % 32 || a->ldx % 8 || ldg % 8 || a->x_bs % 8 || g_bs % 8 || (((uintptr_t)a->x | (uintptr_t)g) & 15)) return 1;
  if (a->M >= (1ll << 31) || a->M < MAX_TOK) return 1;
  Params p = {};
  p.row_stats = a->prologue == TFSWA_PRO_LNHAT ? a->row_stats : nullptr; p.rs_bs = a->rs_bs;
  p.dw = dw; p.w_bs = (int64_t)a->N * a->K; p.dbias = dbias; p.bias_bs = a->N;
  p.M = a->M; p.N = a->N; p.K = a->K;
  p.cbA = a->N % 64 == 0 ? 64 : 32;
  p.cbX = a->K % 64 == 0 ? 64 : 32;
  // k tile: the largest divisor-friendly width <= 256 (multiple of the box width)
  p.BKW = a->K <= 256 ? a->K : (a->K % 256 == 0 ? 256 : (a->K % 192 == 0 ? 192 : 128));
  if (p.BKW % p.cbX || a->K % p.BKW) return 1;
  p.TOK = p.BKW <= 64 ? 128 : 64;                     // stage = (128 + BKW) * TOK * 2 bytes: 40 / 48 / 32..48 KB
  p.tmem_cols = 512;                                  // accumulator at [0, BKW), bias accumulator at [256, 272)
  const int kt = a->K / p.BKW, nt = (a->N + NT - 1) / NT;
  static int sms = 0;
  static DeviceOnce attr_once;
  const size_t smem = (size_t)STAGES * ((size_t)NT * 64 * 2 + (size_t)64 * 256 * 2) + 1024;      // the largest stage: 48 KB
  if (attr_once.needed()) {
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, attr_once.dev);
    if (cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess || sms <= 0) {
      set_error("wgrad_tc: cudaFuncSetAttribute failed");
      return TFSWA_ECUDA;
    }
    attr_once.done();
  }
  int64_t want = (2 * (int64_t)sms) / ((int64_t)kt * nt * a->batch);   // ~two waves of one CTA per SM
  if (want < 1) want = 1;
  const int64_t max_split = (p.M + 2047) / 2048;
  if (want > max_split) want = max_split;
  p.rows_per_cta = ((p.M + want - 1) / want + p.TOK - 1) / p.TOK * p.TOK;
  p.msplit = (int)((p.M + p.rows_per_cta - 1) / p.rows_per_cta);
  CUtensorMap tm_g, tm_x;
  int rc = make_tmap_tokens(&tm_g, g, a->N, a->M, a->batch, ldg, g_bs, p.cbA, p.TOK);
  if (rc) return rc;
  rc = make_tmap_tokens(&tm_x, a->x, a->K, a->M, a->batch, a->ldx, a->x_bs, p.cbX, p.TOK);
  if (rc) return rc;
  dim3 grid(kt, nt, a->batch * p.msplit);
  const size_t smem_launch = (size_t)STAGES * ((size_t)NT * p.TOK * 2 + (size_t)p.TOK * p.BKW * 2) + 1024;
  wgrad_tc_kernel<<<grid, THREADS, smem_launch, st>>>(tm_g, tm_x, p);
  return check_launch("linear_wgrad(tc)");
}

}  // namespace tfswa
