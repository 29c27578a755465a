// Implicit-GEMM convolutions on the tcgen05 tensor cores (bf16, eval-mode epilogues): 3x3 s1 p1 (output_head.0,
// tfswa_unet.py:140), 4x4 s2 p1 (DownsampleBlock, blocks.py:157) and the 4-phase transposed 4x4 s2 p1
// (UpsampleBlock, blocks.py:172).
//
// Same skeleton as tc_linear.cu (mbarrier ring, one elected thread issuing tcgen05.mma into TMEM, eight epilogue
// warps), but the A operand cannot come from a TMA box: a 128-token tile of an NHWC feature map under a k x k
// window is a gather with zero padding.  Eight producer warps (two threads per output pixel) copy the Cin-contiguous
// 64-channel run of one filter tap into shared memory in the 128B-swizzle pattern the UMMA descriptor expects
// (64-byte runs / 64B swizzle when Cin = 32); the (N x K) weight tile still arrives by TMA.  K is ordered
// (tap, channel), i.e. the host re-lays the weights out as (Cout, kh, kw, Cin) - the same layout the SIMT igemm uses.
// After the last k-block the producer warps turn into the epilogue warps (bias, erf-GELU, bf16, TMA store; the
// transposed conv scatters its rows directly because a phase's 128 output pixels are not a TMA box).
#include "common.cuh"
#include "sm100.cuh"

namespace tfswa {

using namespace sm100;

enum { CK_CONV3 = 0, CK_DOWN = 1, CK_UP = 2 };

struct TcConvParams {
  const bf16* x; bf16* y; const float* bias;
  int B, Hin, Win, Cin, Hout, Wout, Cout;
  int Hq, Wq;                 // CK_UP: phase grid = ceil(Hout/2) x ceil(Wout/2)
  int kind, epilogue;
  int64_t M;                  // GEMM rows per launch z (output pixels, or phase-grid cells)
  int K, BN, BK, KB, stages, OB;
  uint32_t tmem_cols;
  float* col_stats;          // optional (2, Cout) fp32: sum / sum of squares of the stored outputs (train-mode BatchNorm)
};

constexpr int CV_BM = 128;
constexpr int CV_THREADS = 320;     // warp 0: weight TMA + TMEM alloc, warp 1: MMA issue, warps 2-9: A gather, then epilogue
constexpr int CV_MAX_STAGES = 4;

__global__ void __launch_bounds__(CV_THREADS) tc_conv_kernel(const __grid_constant__ CUtensorMap tmw,
                                                             const __grid_constant__ CUtensorMap tmy, const TcConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_full[CV_MAX_STAGES], bar_empty[CV_MAX_STAGES], bar_acc;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_bias[256];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t m0 = (int64_t)blockIdx.x * CV_BM;
  const int n0 = blockIdx.y * p.BN;
  const int phase = blockIdx.z;                                     // CK_UP only
  const uint32_t a_bytes = CV_BM * p.BK * 2, b_bytes = p.BN * p.BK * 2;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* tiles = smem_raw + (base - smem_u32(smem_raw));

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tmw);
      prefetch_tmap(&tmy);
      for (int s = 0; s < p.stages; ++s) { mbar_init(&bar_full[s], 1 + 8); mbar_init(&bar_empty[s], 1); }
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
      // ---------------- weight tiles by TMA ----------------
      for (int kb = 0; kb < p.KB; ++kb) {
        const int s = kb % p.stages;
        if (kb >= p.stages) mbar_wait(&bar_empty[s], ((kb / p.stages) - 1) & 1);
        mbar_arrive_expect_tx(&bar_full[s], b_bytes);
        tma_load_3d(tiles + (size_t)s * (a_bytes + b_bytes) + a_bytes, &tmw, &bar_full[s], kb * p.BK, n0, phase);
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ---------------- MMA issuer ----------------
      const uint32_t idesc = umma_idesc_bf16(CV_BM, p.BN);
      const uint32_t row_bytes = p.BK * 2;
      for (int kb = 0; kb < p.KB; ++kb) {
        const int s = kb % p.stages;
        mbar_wait(&bar_full[s], (kb / p.stages) & 1);
        tc_fence_after();
        const uint32_t sa = base + s * (a_bytes + b_bytes);
        const uint64_t adesc = umma_smem_desc(sa, row_bytes);
        const uint64_t bdesc = umma_smem_desc(sa + a_bytes, row_bytes);
        for (int k = 0; k < p.BK / 16; ++k)
          umma_bf16_ss(tmem, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) ? 1u : 0u);
        umma_commit(&bar_empty[s]);
      }
      umma_commit(&bar_acc);
    }
  } else {
    // ---------------- A gather: thread (row, half) copies half of the BK-channel run of each tap ----------------
    const int quad = warp & 3, hsel = (warp - 2) >> 2;              // TMEM lane quadrant / which half of the run
    const int row_in_tile = quad * 32 + lane;
    const int64_t m = m0 + row_in_tile;
    // decode the output pixel (or phase-grid cell) once
    const int wdim = p.kind == CK_UP ? p.Wq : p.Wout;
    const int hw = (p.kind == CK_UP ? p.Hq : p.Hout) * wdim;
    const bool row_ok = m < p.M;
    const int b = row_ok ? (int)(m / hw) : 0;
    const int rem = row_ok ? (int)(m - (int64_t)b * hw) : 0;
    const int oy = rem / wdim, ox = rem - oy * wdim;
    const uint32_t row_bytes = p.BK * 2, half_bytes = row_bytes >> 1, swz_mask = (row_bytes >> 4) - 1u;
    const int per_tap = p.Cin / p.BK;                               // k-blocks per filter tap
    for (int kb = 0; kb < p.KB; ++kb) {
      const int s = kb % p.stages;
      const int tap = kb / per_tap, ci0 = (kb - tap * per_tap) * p.BK;
      int iy, ix;
      if (p.kind == CK_CONV3) { iy = oy + tap / 3 - 1; ix = ox + tap % 3 - 1; }
      else if (p.kind == CK_DOWN) { iy = 2 * oy + (tap >> 2) - 1; ix = 2 * ox + (tap & 3) - 1; }
      else {
        const int py = phase >> 1, px = phase & 1, a = tap >> 1, bb = tap & 1;
        iy = py ? (a ? oy : oy + 1) : (a ? oy - 1 : oy);
        ix = px ? (bb ? ox : ox + 1) : (bb ? ox - 1 : ox);
      }
      const bool ok = row_ok && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win &&
                      (p.kind != CK_UP || (2 * oy + (phase >> 1) < p.Hout && 2 * ox + (phase & 1) < p.Wout));
      uint4 v[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = make_uint4(0, 0, 0, 0);
      if (ok) {
        const uint4* src = reinterpret_cast<const uint4*>(p.x + (((int64_t)b * p.Hin + iy) * p.Win + ix) * p.Cin + ci0) + hsel * (half_bytes >> 4);
#pragma unroll
        for (int c = 0; c < 4; ++c) if (c < (int)(half_bytes >> 4)) v[c] = src[c];
      }
      if (kb >= p.stages) mbar_wait(&bar_empty[s], ((kb / p.stages) - 1) & 1);
      uint8_t* sa = tiles + (size_t)s * (a_bytes + b_bytes);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        if (c < (int)(half_bytes >> 4)) {
          uint32_t off = row_in_tile * row_bytes + (hsel * (half_bytes >> 4) + c) * 16u;
          off ^= ((off >> 7) & swz_mask) << 4;
          *reinterpret_cast<uint4*>(sa + off) = v[c];
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_full[s]);
    }

    // ---------------- epilogue ----------------
    const int et = threadIdx.x - 64;
    for (int i = et; i < p.BN; i += 256) s_bias[i] = p.bias ? p.bias[n0 + i] : 0.f;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const int nch = p.BN / 16;
    const int c_begin = hsel ? (nch / 2) * 16 : 0;
    const int c_end = hsel ? p.BN : (nch / 2) * 16;
    const uint32_t ob_bytes = p.OB * 2, box_bytes = 128u * ob_bytes, oswz = (ob_bytes >> 4) - 1u;
    const uint32_t ob_shift = p.OB == 64 ? 6u : (p.OB == 32 ? 5u : 4u);
    int64_t up_off = -1;                                            // CK_UP: element offset of my output row
    if (p.kind == CK_UP && row_ok) {
      const int yy = 2 * oy + (phase >> 1), xx = 2 * ox + (phase & 1);
      if (yy < p.Hout && xx < p.Wout) up_off = (((int64_t)b * p.Hout + yy) * p.Wout + xx) * p.Cout + n0;
    }
    mbar_wait(&bar_acc, 0);
    tc_fence_after();
    for (int c = c_begin; c < c_end; c += 16) {
      uint32_t raw[16];
      __syncwarp();
      tmem_ld_x16(tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)c, raw);
      tmem_ld_wait();
      float lo[8], hi[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float a0 = __uint_as_float(raw[j]) + s_bias[c + j], a1 = __uint_as_float(raw[8 + j]) + s_bias[c + 8 + j];
        if (p.epilogue == TFSWA_EPI_GELU) { a0 = gelu_erf_fast(a0); a1 = gelu_erf_fast(a1); }
        lo[j] = a0; hi[j] = a1;
      }
      if (p.kind == CK_UP) {
        if (up_off >= 0) { store8(p.y + up_off + c, lo); store8(p.y + up_off + c + 8, hi); }
      }
      if (p.col_stats && !(p.kind == CK_UP ? up_off >= 0 : row_ok)) {   // rows that produce no output must not count
#pragma unroll
        for (int j = 0; j < 8; ++j) { lo[j] = 0.f; hi[j] = 0.f; }
      }
      if (p.kind != CK_UP || p.col_stats) {
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
          const int cc = c + hf * 8;
          const uint32_t blk = (uint32_t)cc >> ob_shift, chunk = ((uint32_t)cc & (p.OB - 1)) >> 3;
          uint32_t off = row_in_tile * ob_bytes + chunk * 16u;
          off ^= ((off >> 7) & oswz) << 4;
          store8(reinterpret_cast<bf16*>(tiles + blk * box_bytes + off), hf ? hi : lo);
        }
      }
    }
    if (p.kind != CK_UP || p.col_stats) {
      fence_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (p.kind != CK_UP && threadIdx.x == 64) {
        for (int bx = 0; bx < p.BN / p.OB; ++bx) tma_store_3d(&tmy, tiles + bx * box_bytes, n0 + bx * p.OB, (int)m0, 0);
      }
      if (p.col_stats) {
        // BatchNorm statistics of the tile from the staged bf16 values (rows without an output were staged as zeros):
        // two threads per column walk the 128 rows of the swizzled tile, one fp32 atomic pair per column half and CTA
        const int col = et >> 1, rhalf = et & 1;
        if (col < p.BN) {
          const uint32_t blk = (uint32_t)col >> ob_shift, chunk = ((uint32_t)col & (p.OB - 1)) >> 3;
          float s1 = 0.f, s2 = 0.f;
          for (int r = rhalf * 64; r < rhalf * 64 + 64; ++r) {
            uint32_t off = (uint32_t)r * ob_bytes + chunk * 16u;
            off ^= ((off >> 7) & oswz) << 4;
            const float v = __bfloat162float(*reinterpret_cast<const bf16*>(tiles + blk * box_bytes + off + (col & 7) * 2));
            s1 += v; s2 = fmaf(v, v, s2);
          }
          atomicAdd(p.col_stats + n0 + col, s1);
          atomicAdd(p.col_stats + p.Cout + n0 + col, s2);
        }
      }
      if (p.kind != CK_UP && threadIdx.x == 64) tma_store_commit_wait();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, p.tmem_cols);
}

}  // namespace tfswa

using namespace tfswa;

extern "C" int tfswa_conv_tc_fwd(const tfswa_conv_args* a, const void* w_bf16, void* stream) {
  TFSWA_REQUIRE(a && a->x && w_bf16 && a->y && a->bias, "conv_tc: null pointer");
  TFSWA_REQUIRE(a->dtype == TFSWA_BF16, "conv_tc: bf16 activations only");
  TFSWA_REQUIRE(!a->pre, "conv_tc: the pre-activation output is not produced by this kernel");
  TFSWA_REQUIRE(!a->col_stats || a->epilogue == TFSWA_EPI_NONE, "conv_tc: col_stats are those of the stored tensor (no epilogue)");
  TFSWA_REQUIRE(a->Cin % 32 == 0 && a->Cout % 16 == 0 && a->Cout <= 256, "conv_tc: need Cin%%32==0, Cout%%16==0, Cout<=256");
  TcConvParams p = {};
  p.x = (const bf16*)a->x; p.y = (bf16*)a->y; p.bias = a->bias;
  p.B = a->B; p.Hin = a->Hin; p.Win = a->Win; p.Cin = a->Cin; p.Hout = a->Hout; p.Wout = a->Wout; p.Cout = a->Cout;
  p.kind = a->kind; p.epilogue = a->epilogue; p.col_stats = a->col_stats;
  int taps, zdim = 1;
  if (a->kind == CK_CONV3) {
    TFSWA_REQUIRE(a->Hout == a->Hin && a->Wout == a->Win, "conv_tc 3x3: output size must equal input size");
    taps = 9; p.M = (int64_t)a->B * a->Hout * a->Wout;
  } else if (a->kind == CK_DOWN) {
    TFSWA_REQUIRE(a->Hout == (a->Hin - 2) / 2 + 1 && a->Wout == (a->Win - 2) / 2 + 1, "conv_tc down: bad output size");
    taps = 16; p.M = (int64_t)a->B * a->Hout * a->Wout;
  } else if (a->kind == CK_UP) {
    TFSWA_REQUIRE(a->Hout >= 2 * a->Hin && a->Hout <= 2 * a->Hin + 1 && a->Wout >= 2 * a->Win && a->Wout <= 2 * a->Win + 1,
                  "conv_tc up: output must be 2x (or 2x+1) the input");
    taps = 4; p.Hq = (a->Hout + 1) / 2; p.Wq = (a->Wout + 1) / 2; p.M = (int64_t)a->B * p.Hq * p.Wq; zdim = 4;
  } else TFSWA_REQUIRE(false, "conv_tc: bad kind %d", a->kind);
  TFSWA_REQUIRE(p.M < (1ll << 31), "conv_tc: too many output pixels");
  p.K = taps * a->Cin;
  p.BK = (a->Cin % 64 == 0) ? 64 : 32;
  p.KB = p.K / p.BK;
  p.BN = a->Cout <= 128 ? a->Cout : 128;
  TFSWA_REQUIRE(a->Cout % p.BN == 0, "conv_tc: Cout=%d not tileable", a->Cout);
  p.OB = (p.BN % 64 == 0) ? 64 : ((p.BN % 32 == 0) ? 32 : 16);
  const int stage_bytes = (CV_BM + p.BN) * p.BK * 2;
  p.stages = p.KB < CV_MAX_STAGES ? p.KB : CV_MAX_STAGES;
  while (p.stages > 2 && p.stages * stage_bytes > 64 * 1024) --p.stages;      // keep >= 2 CTAs resident per SM
  p.tmem_cols = 32;
  while ((int)p.tmem_cols < p.BN) p.tmem_cols <<= 1;
  CUtensorMap tmw, tmy;
  int rc = make_tmap_bf16_3d(&tmw, w_bf16, p.K, a->Cout, zdim, p.K, (uint64_t)a->Cout * p.K, p.BK, p.BN);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tmy, a->y, a->Cout, p.kind == CK_UP ? (uint64_t)a->B * a->Hout * a->Wout : (uint64_t)p.M, 1, a->Cout,
                         0, p.OB, CV_BM);
  if (rc) return rc;
  size_t smem = (size_t)p.stages * stage_bytes;
  if (smem < (size_t)CV_BM * p.BN * 2) smem = (size_t)CV_BM * p.BN * 2;
  smem += 1024;
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    cudaError_t e = cudaFuncSetAttribute(tc_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) { set_error("conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return TFSWA_ECUDA; }
    attr_once.done();
  }
  dim3 grid((unsigned)ceil_div64(p.M, CV_BM), a->Cout / p.BN, zdim);
  tc_conv_kernel<<<grid, CV_THREADS, smem, (cudaStream_t)stream>>>(tmw, tmy, p);
  return check_launch("conv_tc");
}
