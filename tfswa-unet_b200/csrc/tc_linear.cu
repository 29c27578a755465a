// Token GEMM on the 5th-generation tensor cores:  Y = epi(LN?(X) W^T + bias) (+R1) (+R2), bf16 in, fp32 accumulate.
//
//   * X (tokens x K) and W (N x K) tiles are brought to shared memory by TMA (cp.async.bulk.tensor, 128B/64B swizzle)
//     through a multi-stage mbarrier ring; one elected thread issues tcgen05.mma (M=128, N=BN<=256, K=16) with the
//     accumulator in TMEM; four epilogue warps read it back with tcgen05.ld (one token row per thread).
//   * LayerNorm never touches the A operand: with the LN affine folded into W by the host,
//        LN_hat(x) W^T = rstd * (x W^T - mean * rowsum(W)),
//     so the MMA consumes the raw bf16 activations and the epilogue applies (mean, rstd, wsum) per row/column.
//     (This is also more accurate than rounding the normalised activations to bf16 before the MMA.)
//   * bias, exact-erf GELU and up to two residual streams are fused in the epilogue.
//
// Replaces the ATen mm/addmm + layer_norm + gelu + add sequences at attention.py:70,86,121-128,146,159 and the
// 1x1 convs of blocks.py:53-56,85-89.  One output tile (128 x BN) per CTA, 2-3 CTAs resident per SM so one CTA's
// epilogue overlaps another's loads/MMAs; these GEMMs are HBM-bound (K <= 1024), so the aim is bytes/s, not MMA issue.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>
#include "sm100.cuh"
#include <mutex>

namespace tfswa {

using namespace sm100;

// ------------------------------------------------------------------------------------------------
// host: driver entry point + tensor-map helper
// ------------------------------------------------------------------------------------------------
PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = (PFN_encodeTiled)p;
  });
  return fn;
}

int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t cols, uint64_t rows, uint64_t batch, uint64_t row_stride_elems,
                      uint64_t batch_stride_elems, uint32_t box_cols, uint32_t box_rows) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return TFSWA_ECUDA; }
  cuuint64_t dims[3] = {cols, rows, batch};
  uint64_t bstride = batch > 1 ? batch_stride_elems : row_stride_elems * rows;
  cuuint64_t strides[2] = {row_stride_elems * 2, bstride * 2};
  cuuint32_t box[3] = {box_cols, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  const uint32_t span = box_cols * 2;
  CUtensorMapSwizzle sw = span == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : (span == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  if (((uintptr_t)base & 15) || (strides[0] & 15) || (strides[1] & 15)) {
    set_error("tensor map: base/strides must be 16-byte aligned"); return TFSWA_EINVAL;
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return TFSWA_ECUDA; }
  return TFSWA_OK;
}

// ------------------------------------------------------------------------------------------------
struct TcLinearParams {
  const float* bias; int64_t bias_bs;
  const float* row_stats; int64_t rs_bs;
  const float* wsum; int64_t wsum_bs;
  const bf16* r1; int64_t ldr1, r1_bs;
  const bf16* r2; int64_t ldr2, r2_bs;
  bf16* y; int64_t ldy, y_bs;
  int64_t M; int N, K;
  int BN, BK, KB, stages;
  int epilogue, ln;
  float* col_stats;          // optional (2, N) fp32: sum and sum of squares of the (bf16-rounded) outputs over the rows (train-mode BatchNorm)
  uint32_t tmem_cols;
  int OB;                    // columns per output TMA box (64 / 32 / 16): BN/OB boxes of 128 rows x OB*2 bytes, swizzled
};

constexpr int TC_BM = 128;
constexpr int TC_THREADS = 320;     // warp 0: TMA + TMEM alloc, warp 1: MMA issue, warps 2-9: epilogue (2 per TMEM lane quadrant)
constexpr int TC_MAX_STAGES = 4;

__global__ void __launch_bounds__(TC_THREADS) tc_linear_kernel(const __grid_constant__ CUtensorMap tmx,
                                                               const __grid_constant__ CUtensorMap tmw,
                                                               const __grid_constant__ CUtensorMap tmy,
                                                               const TcLinearParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[TC_MAX_STAGES], bar_empty[TC_MAX_STAGES], bar_acc;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_bias[256], s_wsum[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
  const int n0 = blockIdx.y * p.BN;
  const int z = blockIdx.z;
  const uint32_t a_bytes = TC_BM * p.BK * 2, b_bytes = p.BN * p.BK * 2;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;     // shared-space address, 1024-aligned
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tmx);
      prefetch_tmap(&tmw);
      prefetch_tmap(&tmy);
      for (int s = 0; s < p.stages; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
      mbar_init(&bar_acc, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&s_tmem, p.tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer ----------------
      for (int kb = 0; kb < p.KB; ++kb) {
        const int s = kb % p.stages;
        if (kb >= p.stages) mbar_wait(&bar_empty[s], ((kb / p.stages) - 1) & 1);
        mbar_arrive_expect_tx(&bar_full[s], a_bytes + b_bytes);
        uint8_t* sa = tiles + (size_t)s * (a_bytes + b_bytes);
        tma_load_3d(sa, &tmx, &bar_full[s], kb * p.BK, (int)m0, z);
        tma_load_3d(sa + a_bytes, &tmw, &bar_full[s], kb * p.BK, n0, z);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------- MMA issuer ----------------
      const uint32_t idesc = umma_idesc_bf16(TC_BM, p.BN);
      const uint32_t row_bytes = p.BK * 2;
      for (int kb = 0; kb < p.KB; ++kb) {
        const int s = kb % p.stages;
        mbar_wait(&bar_full[s], (kb / p.stages) & 1);
        tc_fence_after();
        const uint32_t sa = base + s * (a_bytes + b_bytes);
        const uint64_t adesc = umma_smem_desc(sa, row_bytes);
        const uint64_t bdesc = umma_smem_desc(sa + a_bytes, row_bytes);
        for (int k = 0; k < p.BK / 16; ++k) {
          // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the (addr >> 4) field
          umma_bf16_ss(tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(&bar_empty[s]);          // frees this smem stage when the MMAs above have read it
      }
      umma_commit(&bar_acc);                 // accumulator complete
    }
  } else {
    // ---------------- epilogue: TMEM -> registers -> bias / LN algebra / GELU / residuals -> global ----------------
    const int et = threadIdx.x - 64;         // 0..255
    for (int i = et; i < p.BN; i += 256) {
      s_bias[i] = p.bias ? p.bias[(int64_t)z * p.bias_bs + n0 + i] : 0.f;
      s_wsum[i] = p.ln ? p.wsum[(int64_t)z * p.wsum_bs + n0 + i] : 0.f;
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int quad = warp & 3;               // TMEM lane quadrant this warp may access
    const int nch = p.BN / 16;               // 16-column chunks; the two warps of a quadrant split them
    const int c_begin = ((warp - 2) >> 2) ? (nch / 2) * 16 : 0;
    const int c_end = ((warp - 2) >> 2) ? p.BN : (nch / 2) * 16;
    const int64_t m = m0 + quad * 32 + lane;
    const bool row_ok = m < p.M;
    float rstd = 1.f, nrm = 0.f;                  // nrm = -rstd*mean
    if (p.ln && row_ok) {
      const float2 st = *reinterpret_cast<const float2*>(p.row_stats + (int64_t)z * p.rs_bs + m * 2);
      rstd = st.y; nrm = -st.x * st.y;
    }
    // Output staging: once the accumulator is complete every smem pipeline stage is dead, so the tile is staged in
    // the same memory as BN/OB boxes of [128 rows x OB columns] in the TMA swizzle pattern (bank-conflict-free 16-byte
    // row-per-thread writes) and leaves the SM as coalesced bulk-tensor stores; rows beyond M are clipped by the TMA.
    const int row_in_tile = quad * 32 + lane;
    const uint32_t ob_bytes = p.OB * 2, box_bytes = 128u * ob_bytes, swz_mask = (ob_bytes >> 4) - 1u;
    const uint32_t ob_shift = p.OB == 64 ? 6u : (p.OB == 32 ? 5u : 4u);
    const bf16* r1row = p.r1 ? p.r1 + (int64_t)z * p.r1_bs + m * p.ldr1 + n0 : nullptr;
    const bf16* r2row = p.r2 ? p.r2 + (int64_t)z * p.r2_bs + m * p.ldr2 + n0 : nullptr;

    mbar_wait(&bar_acc, 0);
    tc_fence_after();
    for (int c = c_begin; c < c_end; c += 16) {
      uint32_t raw[16];
      __syncwarp();                                      // tcgen05.ld is .sync.aligned: keep the warp converged
      tmem_ld_x16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, raw);
      tmem_ld_wait();
      float v[16];
      {
        float bs[16];
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[c + 4 * q4]);
          bs[4 * q4] = b4.x; bs[4 * q4 + 1] = b4.y; bs[4 * q4 + 2] = b4.z; bs[4 * q4 + 3] = b4.w;
        }
        if (p.ln) {                                  // y = rstd*acc - (rstd*mean)*wsum + bias
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 w4 = *reinterpret_cast<const float4*>(&s_wsum[c + 4 * q4]);
            bs[4 * q4] = fmaf(nrm, w4.x, bs[4 * q4]); bs[4 * q4 + 1] = fmaf(nrm, w4.y, bs[4 * q4 + 1]);
            bs[4 * q4 + 2] = fmaf(nrm, w4.z, bs[4 * q4 + 2]); bs[4 * q4 + 3] = fmaf(nrm, w4.w, bs[4 * q4 + 3]);
          }
        }
        if (p.epilogue == TFSWA_EPI_GELU) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = gelu_erf_fast(fmaf(rstd, __uint_as_float(raw[j]), bs[j]));
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = fmaf(rstd, __uint_as_float(raw[j]), bs[j]);
        }
      }
      if (row_ok) {
      if (r1row) {
        float t[8];
        if (p.epilogue == TFSWA_EPI_MUL_DGELU) {       // r1 = saved pre-activation: dL/dpre = dL/dh * gelu'(pre)
          load8(r1row + c, t);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] *= gelu_erf_grad_fast(t[j]);
          load8(r1row + c + 8, t);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[8 + j] *= gelu_erf_grad_fast(t[j]);
        } else {
          load8(r1row + c, t);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += t[j];
          load8(r1row + c + 8, t);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[8 + j] += t[j];
        }
      }
      if (r2row) {
        float t[8];
        load8(r2row + c, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += t[j];
        load8(r2row + c + 8, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[8 + j] += t[j];
      }
      }
      {
        float lo[8], hi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { lo[j] = v[j]; hi[j] = v[8 + j]; }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int cc = c + hf * 8;                                   // first column of this 16-byte chunk
          const uint32_t blk = (uint32_t)cc >> ob_shift, chunk = ((uint32_t)cc & (p.OB - 1)) >> 3;
          uint32_t off = row_in_tile * ob_bytes + chunk * 16u;
          off ^= ((off >> 7) & swz_mask) << 4;                         // Swizzle<log2(span/16),4,3>, as the TMA expects
          store8(reinterpret_cast<bf16*>(tiles + blk * box_bytes + off), hf ? hi : lo);
        }
      }
    }
    fence_async_smem();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (threadIdx.x == 64) {
      for (int b = 0; b < p.BN / p.OB; ++b) tma_store_3d(&tmy, tiles + b * box_bytes, n0 + b * p.OB, (int)m0, z);
    }
    if (p.col_stats) {
      // Train-mode BatchNorm statistics of the tile, from the staged (bf16-rounded, i.e. exactly what is stored) values:
      // two threads per column walk the 128 rows of the swizzled tile, one fp32 atomic pair per column half and CTA.
      const int col = et >> 1, rhalf = et & 1;
      if (col < p.BN) {
        const int64_t left = p.M - m0;
        const int rows_valid = left < TC_BM ? (int)left : TC_BM;     // rows beyond M hold bias, not data
        const uint32_t blk = (uint32_t)col >> ob_shift, chunk = ((uint32_t)col & (p.OB - 1)) >> 3;
        float s1 = 0.f, s2 = 0.f;
        for (int r = rhalf * 64; r < rhalf * 64 + 64 && r < rows_valid; ++r) {
          uint32_t off = (uint32_t)r * ob_bytes + chunk * 16u;
          off ^= ((off >> 7) & swz_mask) << 4;
          const float v = __bfloat162float(*reinterpret_cast<const bf16*>(tiles + blk * box_bytes + off + (col & 7) * 2));
          s1 += v; s2 = fmaf(v, v, s2);
        }
        if (rhalf * 64 < rows_valid) {
          atomicAdd(p.col_stats + n0 + col, s1);
          atomicAdd(p.col_stats + p.N + n0 + col, s2);
        }
      }
    }
    if (threadIdx.x == 64) tma_store_commit_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, p.tmem_cols);
}

// ------------------------------------------------------------------------------------------------
// Persistent form (round 2): one CTA per SM loops over the output tiles (tile = blockIdx.x + i * gridDim.x, n-tile
// fastest so the CTAs that run together share an activation tile in L2), set-up once per CTA, and the four roles only meet
// on mbarriers:
//   warp 0       TMA: X and W k-blocks of tile i, i+1, ... through an S-stage ring that runs across tile boundaries; with a
//                residual, also the r1 tile, loaded straight INTO the output staging buffer (same box geometry / swizzle as
//                the store), where the epilogue adds it in place
//   warp 1       tcgen05.mma into one of two TMEM accumulators (tile i -> buffer i & 1)
//   warps 2-9    epilogue group 0: tiles 0, 2, 4, ...   (TMEM -> bias / LN algebra / GELU / residuals -> bf16 -> staging[0] -> TMA store)
//   warps 10-17  epilogue group 1: tiles 1, 3, 5, ...   (staging[1])
// so the loads and MMAs of the next tiles and the other group's epilogue run under a group's epilogue.  The one-tile-per-CTA
// kernel above reaches 2.1-4.3 TB/s on the block's GEMMs (every CTA pays TMEM allocation, barrier set-up, descriptor fetch
// and a cold pipeline for 40-64 KB of traffic); it remains the path for problems with fewer than two tiles per SM.
constexpr int PL_MAXG = 3;                 // epilogue groups (8 warps each); tile i -> group i % NG.  Two instantiations: NG = 2 (96 registers:
                                           // tiles of 96-128 columns) and NG = 3 (72 registers: narrow tiles, where a third group hides more latency)
constexpr int PL_MAX_STAGES = 8;

struct PlParams {
  TcLinearParams g;
  int tiles_n, nstages, has_r1;
  int ng;                              // epilogue groups in use (2 or 3) = TMEM accumulators = tiles in flight behind the MMA warp
  uint32_t tmem_total;                 // power of two >= ng * tmem_cols
  int nsb;                             // staging buffers per epilogue group (1 or 2): stores of that many tiles may be in flight per group
  int tiles_m, total, rows, row_major;   // rows = tiles_m * batch; row_major: a CTA walks all n-tiles of a row block before the next one
  uint32_t ring_bytes, stg_bytes;      // bytes of the load ring / of ONE staging buffer (128 x BN bf16)
};

template <int NG>
__global__ void __launch_bounds__(64 + NG * 256, 1) tc_linear_persist_kernel(const __grid_constant__ CUtensorMap tmx,
                                                                          const __grid_constant__ CUtensorMap tmw,
                                                                          const __grid_constant__ CUtensorMap tmy,
                                                                          const __grid_constant__ CUtensorMap tmr,
                                                                          const PlParams pp) {
  const TcLinearParams& p = pp.g;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[PL_MAX_STAGES], bar_empty[PL_MAX_STAGES], acc_full[PL_MAXG], acc_empty[PL_MAXG], stg_free[PL_MAXG], r1_full[PL_MAXG], stg_ok[PL_MAXG];
  __shared__ uint32_t s_tmem;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = TC_BM * p.BK * 2, b_bytes = p.BN * p.BK * 2, st_bytes = a_bytes + b_bytes;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* ring = smem_raw + (base - smem_u32(smem_raw));
  uint8_t* staging = ring + pp.ring_bytes;
  const int S = pp.nstages;
  const uint32_t ob_bytes = p.OB * 2, box_bytes = 128u * ob_bytes;
  const int nbox = p.BN / p.OB;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tmx); prefetch_tmap(&tmw); prefetch_tmap(&tmy);
      if (pp.has_r1) prefetch_tmap(&tmr);
      for (int s = 0; s < S; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
      for (int b = 0; b < NG; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], 8); mbar_init(&stg_free[b], 1); mbar_init(&r1_full[b], 1); mbar_init(&stg_ok[b], 1); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&s_tmem, pp.tmem_total);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;

  // i-th tile of this CTA -> (batch z, row block m0, column block n0).  row_major: the CTA walks all n-tiles of row block
  // blockIdx.x + q * gridDim.x before the next one, so that a block of 128 token rows is written in full within a few
  // microseconds by one SM; otherwise tile = blockIdx.x + i * gridDim.x with n fastest (better balance when there are few row blocks)
  auto tile_of = [&](int i, int64_t& m0, int& n0, int& z) {
    int rt, nt;
    if (pp.row_major) {
      const int q = i / pp.tiles_n;
      nt = i - q * pp.tiles_n;
      rt = (int)blockIdx.x + q * (int)gridDim.x;
    } else {
      const int t = (int)blockIdx.x + i * (int)gridDim.x;
      rt = t / pp.tiles_n;
      nt = t - rt * pp.tiles_n;
    }
    if (rt >= pp.rows) return false;
    z = rt / pp.tiles_m;
    m0 = (int64_t)(rt - z * pp.tiles_m) * TC_BM;
    n0 = nt * p.BN;
    return true;
  };

  if (warp == 0) {
    if (elect_one()) {
      // ---------------- TMA producer ----------------
      uint32_t g = 0;
      int64_t m0; int n0, z;
      for (int i = 0; tile_of(i, m0, n0, z); ++i) {
        const int b = i % NG;
        if (pp.has_r1) {
          const int k = i / NG;                      // k-th tile of group b: staging buffer k % nsb of that group
          if (k >= 1) mbar_wait(&stg_free[b], (uint32_t)(k - 1) & 1u);   // the group's store of tile k - nsb has read that buffer
          uint8_t* dst = staging + (size_t)(b * pp.nsb + k % pp.nsb) * pp.stg_bytes;
          mbar_arrive_expect_tx(&r1_full[b], 128u * p.BN * 2u);
          for (int bx = 0; bx < nbox; ++bx) tma_load_3d(dst + bx * box_bytes, &tmr, &r1_full[b], n0 + bx * p.OB, (int)m0, z);
        }
        for (int kb = 0; kb < p.KB; ++kb, ++g) {
          const uint32_t s = g % S;
          if (g >= (uint32_t)S) mbar_wait(&bar_empty[s], ((g / S) - 1) & 1u);
          mbar_arrive_expect_tx(&bar_full[s], st_bytes);
          uint8_t* sa = ring + (size_t)s * st_bytes;
          tma_load_3d(sa, &tmx, &bar_full[s], kb * p.BK, (int)m0, z);
          tma_load_3d(sa + a_bytes, &tmw, &bar_full[s], kb * p.BK, n0, z);
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------- MMA issuer ----------------
      const uint32_t idesc = umma_idesc_bf16(TC_BM, p.BN);
      const uint32_t row_bytes = p.BK * 2;
      uint32_t g = 0;
      int64_t m0; int n0, z;
      for (int i = 0; tile_of(i, m0, n0, z); ++i) {
        const int b = i % NG;
        if (i >= NG) { mbar_wait(&acc_empty[b], (uint32_t)(i / NG - 1) & 1u); tc_fence_after(); }   // epilogue of tile i - NG has drained buffer b
        for (int kb = 0; kb < p.KB; ++kb, ++g) {
          const uint32_t s = g % S;
          mbar_wait(&bar_full[s], (g / S) & 1u);
          tc_fence_after();
          const uint32_t sa = base + s * st_bytes;
          const uint64_t adesc = umma_smem_desc(sa, row_bytes);
          const uint64_t bdesc = umma_smem_desc(sa + a_bytes, row_bytes);
          for (int k = 0; k < p.BK / 16; ++k)
            umma_bf16_ss(tmem + b * p.tmem_cols, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
          umma_commit(&bar_empty[s]);
        }
        umma_commit(&acc_full[b]);
      }
    }
  } else {
    // ---------------- epilogue groups ----------------
    const int gi = (warp - 2) >> 3;                   // group = TMEM buffer = staging buffer
    const int et = threadIdx.x - 64 - gi * 256;       // 0..255 inside the group
    const int w8 = (warp - 2) & 7;
    const int quad = warp & 3;                        // TMEM lane quadrant this warp may access
    const int nch = p.BN / 16;
    const int c_begin = (w8 >> 2) ? (nch / 2) * 16 : 0;
    const int c_end = (w8 >> 2) ? p.BN : (nch / 2) * 16;
    const int row_in_tile = quad * 32 + lane;
    const uint32_t swz_mask = (ob_bytes >> 4) - 1u;
    const uint32_t ob_shift = p.OB == 64 ? 6u : (p.OB == 32 ? 5u : 4u);
    // per-group copies of the bias and rowsum(W) rows of the current batch entry (all N columns): reloaded only when z changes
    float* sb_all = reinterpret_cast<float*>(staging + (size_t)NG * pp.nsb * pp.stg_bytes) + (size_t)gi * 2 * p.N;
    float* sw_all = sb_all + p.N;
    int cur_z = -1;
    const uint32_t bar_id = 1 + gi;
    uint32_t n_use = 0;                               // tiles this group has processed (phase of acc_full / r1_full / stg_ok)
    int64_t m0; int n0, zt;
    for (int i = gi; tile_of(i, m0, n0, zt); i += NG, ++n_use) {
      uint8_t* stg = staging + (size_t)(gi * pp.nsb + n_use % pp.nsb) * pp.stg_bytes;
      if (zt != cur_z) {                              // (group-uniform) new batch entry: its bias / rowsum(W) rows
        asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");   // nobody still reads the old rows
        for (int ii = et; ii < p.N; ii += 256) {
          sb_all[ii] = p.bias ? p.bias[(int64_t)zt * p.bias_bs + ii] : 0.f;
          sw_all[ii] = p.ln ? p.wsum[(int64_t)zt * p.wsum_bs + ii] : 0.f;
        }
        asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
        cur_z = zt;
      }
      const float* sb = sb_all + n0;
      const float* sw = sw_all + n0;
      const int64_t m = m0 + row_in_tile;
      const bool row_ok = m < p.M;
      float rstd = 1.f, nrm = 0.f;
      if (p.ln && row_ok) {
        const float2 st = *reinterpret_cast<const float2*>(p.row_stats + (int64_t)zt * p.rs_bs + m * 2);
        rstd = st.y; nrm = -st.x * st.y;
      }
      const bf16* r2row = p.r2 ? p.r2 + (int64_t)zt * p.r2_bs + m * p.ldr2 + n0 : nullptr;
      // staging[gi] is free again once the group's previous store has read it: without a residual the storing thread says so
      // on stg_ok (no group-wide rendezvous); with one, the landed residual tile (r1_full) implies it
      if (!pp.has_r1) {
        if (et == 0) {
          if (pp.nsb == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store of tile n_use - 2 (same buffer)
          else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive(&stg_ok[gi]);
        }
      }
      mbar_wait(&acc_full[gi], n_use & 1u);
      tc_fence_after();
      if (pp.has_r1) mbar_wait(&r1_full[gi], n_use & 1u);
      else mbar_wait(&stg_ok[gi], n_use & 1u);
      // 16 accumulator columns starting at column c -> epilogue arithmetic -> bf16 -> staging (adding the residual found there)
      auto cols16 = [&](const uint32_t* raw, int c) {
        float v[16];
        {
          float bs[16];
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            const float4 b4 = *reinterpret_cast<const float4*>(&sb[c + 4 * q4]);
            bs[4 * q4] = b4.x; bs[4 * q4 + 1] = b4.y; bs[4 * q4 + 2] = b4.z; bs[4 * q4 + 3] = b4.w;
          }
          if (p.ln) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
              const float4 w4 = *reinterpret_cast<const float4*>(&sw[c + 4 * q4]);
              bs[4 * q4] = fmaf(nrm, w4.x, bs[4 * q4]); bs[4 * q4 + 1] = fmaf(nrm, w4.y, bs[4 * q4 + 1]);
              bs[4 * q4 + 2] = fmaf(nrm, w4.z, bs[4 * q4 + 2]); bs[4 * q4 + 3] = fmaf(nrm, w4.w, bs[4 * q4 + 3]);
            }
          }
          if (p.epilogue == TFSWA_EPI_GELU) {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = gelu_erf_fast(fmaf(rstd, __uint_as_float(raw[j]), bs[j]));
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = fmaf(rstd, __uint_as_float(raw[j]), bs[j]);
          }
        }
        if (r2row && row_ok) {
          float tt[8];
          load8(r2row + c, tt);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += tt[j];
          load8(r2row + c + 8, tt);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[8 + j] += tt[j];
        }
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int cc = c + hf * 8;
          const uint32_t blk = (uint32_t)cc >> ob_shift, chunk = ((uint32_t)cc & (p.OB - 1)) >> 3;
          uint32_t off = row_in_tile * ob_bytes + chunk * 16u;
          off ^= ((off >> 7) & swz_mask) << 4;
          bf16* dst = reinterpret_cast<bf16*>(stg + blk * box_bytes + off);
          float o8[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o8[j] = v[hf * 8 + j];
          if (pp.has_r1) {                             // the residual tile was loaded into this very position by TMA
            float tt[8];
            load8(dst, tt);
            if (p.epilogue == TFSWA_EPI_MUL_DGELU) {   // ... or the saved pre-activation: dL/dpre = dL/dh * gelu'(pre)
#pragma unroll
              for (int j = 0; j < 8; ++j) o8[j] *= gelu_erf_grad_fast(tt[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) o8[j] += tt[j];
            }
          }
          store8(dst, o8);
        }
      };
      const uint32_t acc_addr = tmem + gi * p.tmem_cols + ((uint32_t)(quad * 32) << 16);
      for (int c = c_begin; c < c_end; c += 32) {
        const bool two = c + 32 <= c_end;              // warp-uniform
        uint32_t raw[32];
        __syncwarp();
        if (two) tmem_ld_x32(acc_addr + (uint32_t)c, raw);
        else {
          uint32_t r16[16];
          tmem_ld_x16(acc_addr + (uint32_t)c, r16);
#pragma unroll
          for (int j = 0; j < 16; ++j) raw[j] = r16[j];
        }
        tmem_ld_wait();
        if (c + 32 >= c_end) {                         // last read of the accumulator: hand the buffer back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&acc_empty[gi]);
        }
        cols16(raw, c);
        if (two) cols16(raw + 16, c + 16);
      }
      fence_async_smem();
      asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");
      if (et == 0) {
        for (int bx = 0; bx < nbox; ++bx) tma_store_3d(&tmy, stg + bx * box_bytes, n0 + bx * p.OB, (int)m0, zt);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      if (p.col_stats) {
        const int col = et >> 1, rhalf = et & 1;
        if (col < p.BN) {
          const int64_t left = p.M - m0;
          const int rows_valid = left < TC_BM ? (int)left : TC_BM;
          const uint32_t blk = (uint32_t)col >> ob_shift, chunk = ((uint32_t)col & (p.OB - 1)) >> 3;
          float s1 = 0.f, s2 = 0.f;
          for (int r = rhalf * 64; r < rhalf * 64 + 64 && r < rows_valid; ++r) {
            uint32_t off = (uint32_t)r * ob_bytes + chunk * 16u;
            off ^= ((off >> 7) & swz_mask) << 4;
            const float v = __bfloat162float(*reinterpret_cast<const bf16*>(stg + blk * box_bytes + off + (col & 7) * 2));
            s1 += v; s2 = fmaf(v, v, s2);
          }
          if (rhalf * 64 < rows_valid) {
            atomicAdd(p.col_stats + n0 + col, s1);
            atomicAdd(p.col_stats + p.N + n0 + col, s2);
          }
        }
      }
      if (p.col_stats) asm volatile("bar.sync %0, 256;" ::"r"(bar_id) : "memory");   // the column sums above read staging[gi]
      if (et == 0 && pp.has_r1) {
        // with a residual the TMA warp wants staging[gi] back as early as possible (it loads the group's next residual tile
        // into it, two tiles ahead): wait for the store's reads here rather than at the top of the next tile
        if (pp.nsb == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the buffer of the group's NEXT tile is free
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        mbar_arrive(&stg_free[gi]);
      }
    }
    if (et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, pp.tmem_total);
}

static int pick_bn(int N) {
  static int forced = -1;               // TFSWA_LINEAR_BN=<n>: tile-width experiment (development A/B)
  if (forced < 0) { const char* e = getenv("TFSWA_LINEAR_BN"); forced = e ? atoi(e) : 0; }
  if (forced >= 16 && forced <= 256 && forced % 16 == 0 && N % forced == 0) return forced;
  // <= 128 accumulator columns keeps 4 CTAs resident per SM (TMEM) so the epilogue of one tile overlaps the loads and
  // MMAs of others; prefer tiles whose output box is 64 (then 32) columns wide
  for (int step = 64; step >= 16; step >>= 1)
    for (int bn = 128; bn >= step; bn -= step)
      if (bn % step == 0 && N % bn == 0) return bn;
  for (int bn = 256; bn >= 16; bn -= 16)
    if (N % bn == 0) return bn;
  return 0;
}

}  // namespace tfswa

using namespace tfswa;

extern "C" int tfswa_linear_tc_fwd(const tfswa_linear_args* a, const void* w_bf16, const float* wsum, void* stream) {
  TFSWA_REQUIRE(a && a->x && w_bf16 && a->y, "linear_tc: null pointer");
  TFSWA_REQUIRE(a->dtype == TFSWA_BF16, "linear_tc: bf16 activations only");
  TFSWA_REQUIRE(a->M > 0 && a->batch > 0 && a->batch <= 65535, "linear_tc: empty problem");
  TFSWA_REQUIRE(a->K % 32 == 0 && a->K >= 32 && a->N % 16 == 0, "linear_tc: need K%%32==0 and N%%16==0 (K=%d N=%d)", a->K, a->N);
  TFSWA_REQUIRE(a->prologue == TFSWA_PRO_NONE || a->prologue == TFSWA_PRO_LNHAT, "linear_tc: prologue %d unsupported", a->prologue);
  TFSWA_REQUIRE(a->prologue != TFSWA_PRO_LNHAT || (a->row_stats && wsum), "linear_tc: LN needs row_stats and wsum");
  TFSWA_REQUIRE(!a->pre, "linear_tc: the pre-activation output is not produced by this kernel");
  TFSWA_REQUIRE(a->epilogue == TFSWA_EPI_NONE || a->epilogue == TFSWA_EPI_GELU || a->epilogue == TFSWA_EPI_MUL_DGELU, "linear_tc: epilogue %d unsupported", a->epilogue);
  TFSWA_REQUIRE(a->epilogue != TFSWA_EPI_MUL_DGELU || (a->r1 && !a->r2 && !a->col_stats), "linear_tc: TFSWA_EPI_MUL_DGELU needs r1 (the pre-activation) and no r2 / col_stats");
  TFSWA_REQUIRE(!a->col_stats || (a->batch == 1 && a->epilogue == TFSWA_EPI_NONE && !a->r1 && !a->r2),
                "linear_tc: col_stats needs batch 1 and a plain epilogue (the statistics are those of the stored tensor)");
  TFSWA_REQUIRE(a->ldy % 8 == 0 && a->y_bs % 8 == 0 && a->ldx % 8 == 0 && a->x_bs % 8 == 0, "linear_tc: 16-byte alignment of ld/strides");
  TFSWA_REQUIRE((!a->r1 || (a->ldr1 % 8 == 0 && a->r1_bs % 8 == 0)) && (!a->r2 || (a->ldr2 % 8 == 0 && a->r2_bs % 8 == 0)),
                "linear_tc: residual alignment");
  TcLinearParams p = {};
  p.BN = pick_bn(a->N);
  TFSWA_REQUIRE(p.BN >= 16, "linear_tc: no N tile for N=%d", a->N);
  p.BK = (a->K % 64 == 0) ? 64 : 32;
  p.KB = a->K / p.BK;
  const int stage_bytes = (TC_BM + p.BN) * p.BK * 2;
  p.stages = p.KB < TC_MAX_STAGES ? p.KB : TC_MAX_STAGES;
  while (p.stages > 1 && p.stages * stage_bytes > 96 * 1024) --p.stages;
  p.tmem_cols = 32;
  while ((int)p.tmem_cols < p.BN) p.tmem_cols <<= 1;
  p.bias = a->bias; p.bias_bs = a->bias_bs; p.row_stats = a->row_stats; p.rs_bs = a->rs_bs; p.wsum = wsum; p.wsum_bs = a->N;
  p.r1 = (const bf16*)a->r1; p.ldr1 = a->ldr1; p.r1_bs = a->r1_bs; p.r2 = (const bf16*)a->r2; p.ldr2 = a->ldr2; p.r2_bs = a->r2_bs;
  p.y = (bf16*)a->y; p.ldy = a->ldy; p.y_bs = a->y_bs;
  p.M = a->M; p.N = a->N; p.K = a->K; p.epilogue = a->epilogue; p.ln = a->prologue == TFSWA_PRO_LNHAT;
  p.col_stats = a->col_stats;
  p.OB = (p.BN % 64 == 0) ? 64 : ((p.BN % 32 == 0) ? 32 : 16);
  CUtensorMap tmx, tmw, tmy;
  int rc = make_tmap_bf16_3d(&tmx, a->x, a->K, a->M, a->batch, a->ldx, a->x_bs, p.BK, TC_BM);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tmw, w_bf16, a->K, a->N, a->batch, a->K, (uint64_t)a->N * a->K, p.BK, p.BN);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tmy, a->y, a->N, a->M, a->batch, a->ldy, a->y_bs, p.OB, TC_BM);
  if (rc) return rc;
  size_t smem = (size_t)p.stages * stage_bytes;
  if (smem < (size_t)TC_BM * p.BN * 2) smem = (size_t)TC_BM * p.BN * 2;     // the output tile is staged in the same memory
  smem += 1024;
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    cudaError_t e = cudaFuncSetAttribute(tc_linear_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { set_error("linear_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return TFSWA_ECUDA; }
    attr_once.done();
  }
  // persistent form when every SM gets at least two tiles (TFSWA_LINEAR_KERNEL=v1 forces the one-tile-per-CTA kernel)
  static int sms = 0, force_v1 = -1;
  static DeviceOnce pl_once;
  if (force_v1 < 0) { const char* e = getenv("TFSWA_LINEAR_KERNEL"); force_v1 = (e && !strcmp(e, "v1")) ? 1 : 0; }
  if (pl_once.needed()) {
    cudaError_t e = cudaFuncSetAttribute(tc_linear_persist_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tc_linear_persist_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl_once.dev);
    if (e != cudaSuccess || sms <= 0) { set_error("linear_tc: cudaFuncSetAttribute(persistent): %s", cudaGetErrorString(e)); return TFSWA_ECUDA; }
    pl_once.done();
  }
  PlParams pp = {};
  const int64_t tiles_m64 = ceil_div64(a->M, TC_BM);
  const int64_t total64 = tiles_m64 * (a->N / p.BN) * a->batch;
  pp.tiles_m = (int)tiles_m64;
  pp.tiles_n = a->N / p.BN;
  pp.rows = (int)(tiles_m64 * a->batch);
  pp.total = (int)total64;
  pp.row_major = (pp.tiles_n > 1 && pp.rows >= 8 * sms) ? 1 : 0;
  pp.stg_bytes = (uint32_t)TC_BM * p.BN * 2;
  pp.ng = p.BN <= 64 ? 3 : 2;
  {
    static int force_ng = -1;           // TFSWA_LINEAR_GROUPS=2|3 (development A/B)
    if (force_ng < 0) { const char* e = getenv("TFSWA_LINEAR_GROUPS"); force_ng = e ? atoi(e) : 0; }
    if (force_ng == 2 || (force_ng == 3 && 3 * p.tmem_cols <= 512)) pp.ng = force_ng;
  }
  pp.tmem_total = pp.ng * p.tmem_cols <= 256 ? (pp.ng * p.tmem_cols <= 128 ? (pp.ng * p.tmem_cols <= 64 ? 64 : 128) : 256) : 512;
  const int64_t rows_bytes = 2 * (int64_t)pp.ng * a->N * (int64_t)sizeof(float);   // bias + rowsum(W) rows, one copy per epilogue group
  pp.nsb = 1;        // (2 = two stores in flight per group: measured no gain on the write-dominated shapes and it costs ring depth)
  const int64_t ring_room = 220 * 1024 - 1024 - (int64_t)pp.ng * pp.nsb * pp.stg_bytes - rows_bytes;
  pp.nstages = (int)(ring_room / stage_bytes);
  if (pp.nstages > PL_MAX_STAGES) pp.nstages = PL_MAX_STAGES;
  if (!force_v1 && total64 < (1ll << 30) && total64 >= 2 * (int64_t)sms && pp.nstages >= 2 && p.tmem_cols <= 256 && p.BN >= 32 &&
      !(a->epilogue == TFSWA_EPI_GELU && a->N > a->K)) {   // (wide erf-GELU epilogues - fc1: N = 4K - are issue-bound: measured equal or
                                                                // faster on the one-tile-per-CTA kernel with its 3 CTAs / SM; the fusion convs, N = K/3, are not)
    pp.g = p;
    pp.has_r1 = a->r1 ? 1 : 0;
    pp.ring_bytes = (uint32_t)pp.nstages * stage_bytes;
    CUtensorMap tmr = tmy;
    if (a->r1) {
      rc = make_tmap_bf16_3d(&tmr, a->r1, a->N, a->M, a->batch, a->ldr1, a->r1_bs, p.OB, TC_BM);
      if (rc) return rc;
    }
    const size_t psmem = 1024 + pp.ring_bytes + (size_t)pp.ng * pp.nsb * pp.stg_bytes + (size_t)rows_bytes;
    if (pp.ng == 3) tc_linear_persist_kernel<3><<<sms, 64 + 3 * 256, psmem, (cudaStream_t)stream>>>(tmx, tmw, tmy, tmr, pp);
    else tc_linear_persist_kernel<2><<<sms, 64 + 2 * 256, psmem, (cudaStream_t)stream>>>(tmx, tmw, tmy, tmr, pp);
    return check_launch("linear_tc(persistent)");
  }
  dim3 grid((unsigned)ceil_div64(a->M, TC_BM), a->N / p.BN, a->batch);
  tc_linear_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(tmx, tmw, tmy, p);
  return check_launch("linear_tc");
}
