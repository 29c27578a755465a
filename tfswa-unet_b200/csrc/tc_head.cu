// Fused head of a TFSWABlock on the tcgen05 tensor cores, for the narrow stages (C = 32, 64), inference:
//
//     x1  = x Wi^T + bi                       input_proj: 1x1 conv with eval BatchNorm folded in   (blocks.py:53-56,115)
//     qkv = LN_hat(x1) Wqkv'^T + bqkv'        q|k|v of all three branches, LN affine folded        (attention.py:70,146,220,378)
//
// Unfused this is a GEMM, a statistics pass and a second GEMM that read x1 three times; fused, a 128-token tile of x is
// read once, x1 is written once (the block tail needs it as the residual) and the 9C-wide qkv rows stream out through
// TMA stores while the next chunk's MMA runs: 11*C*2 bytes per token instead of 26*C*2.
//
// One persistent CTA per SM slot, both weight matrices resident in shared memory.  Per tile:
//     TMA(x; prefetched one tile ahead) -> MMA1 (N=C, K=C) -> epilogue 1: + bias -> x1 (global, full 32-byte sectors per
//     thread), row mean / rstd, LN_hat(x1) -> smem -> MMA2 in three chunks of 3C columns (one branch each, TMEM double
//     buffered) -> epilogue 2: + bias -> swizzled staging tile -> TMA store.
// Thread layout as in tc_tail.cu: one token row per TMEM lane, NT = C/16 threads per row.
#include "common.cuh"
#include "sm100.cuh"

namespace tfswa {

using namespace sm100;

struct HeadParams {
  const float* bi; const float* bq;          // (C), (9C)
  bf16* x1; int64_t ld1;
  int64_t M;
  int tiles;
  float eps;
};

template <int C> struct HeadCfg {
  static constexpr int NT = C / 16;
  static constexpr int THREADS = 128 * NT;
  static constexpr int ROWB = C * 2;
  static constexpr uint32_t SWZ = ROWB / 16 - 1;
  static constexpr int TILE = 128 * ROWB;                // one 128 x C bf16 tile (8 / 16 KB)
  static constexpr int NQ = 9 * C, CH = 3 * C;           // qkv width, columns per chunk
  static constexpr int NSTG = C == 32 ? 2 : 1;           // staging buffers for the qkv chunks
  static constexpr int WI = 0;
  static constexpr int WQ = WI + C * ROWB;
  static constexpr int X0 = WQ + NQ * ROWB;              // 2 x tiles (prefetch double buffer)
  static constexpr int A1 = X0 + 2 * TILE;               // LN_hat(x1) operand
  static constexpr int STG = A1 + TILE;                  // NSTG x (3 boxes of 128 rows x C columns)
  static constexpr int BYTES = STG + NSTG * 3 * TILE;    // 92 KB (C = 32) / 176 KB (C = 64)
  static constexpr uint32_t TMEM_COLS = C == 32 ? 256 : 512;
  static constexpr uint32_t Q_COL = 64;                  // x1 accumulator at column 0, qkv chunk buffers at 64 + b*CH
  static constexpr uint32_t W_BYTES = C * ROWB + NQ * ROWB;
};

__device__ __forceinline__ void tma_store_commit_h() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read_h() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_store_wait_all_h() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int C>
__global__ void __launch_bounds__(HeadCfg<C>::THREADS, C == 32 ? 2 : 1)
tc_head_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_wi,
               const __grid_constant__ CUtensorMap tm_wq, const __grid_constant__ CUtensorMap tm_q, const HeadParams p) {
  using Cfg = HeadCfg<C>;
  constexpr int NT = Cfg::NT, ROWB = Cfg::ROWB, CH = Cfg::CH, PT = CH / NT;   // PT = 48 qkv columns per thread and chunk
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_w, bar_ld[2], bar_m1, bar_q[2];
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_bi[C], s_bq[9 * C];
  __shared__ float s_sum[NT][128], s_sq[NT][128];

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, t = warp >> 2;
  const int r = quad * 32 + lane;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tm_x); prefetch_tmap(&tm_q);
      mbar_init(&bar_w, 1); mbar_init(&bar_ld[0], 1); mbar_init(&bar_ld[1], 1); mbar_init(&bar_m1, 1);
      mbar_init(&bar_q[0], 1); mbar_init(&bar_q[1], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&s_tmem, Cfg::TMEM_COLS);
  }
  for (int i = tid; i < C; i += Cfg::THREADS) s_bi[i] = p.bi[i];
  for (int i = tid; i < 9 * C; i += Cfg::THREADS) s_bq[i] = p.bq[i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);
  const int cta = blockIdx.x, ncta = gridDim.x;

  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(&bar_w, Cfg::W_BYTES);
    tma_load_3d(sm + Cfg::WI, &tm_wi, &bar_w, 0, 0, 0);
#pragma unroll
    for (int j = 0; j < 3; ++j) tma_load_3d(sm + Cfg::WQ + j * CH * ROWB, &tm_wq, &bar_w, 0, j * CH, 0);
    if (cta < p.tiles) {
      mbar_arrive_expect_tx(&bar_ld[0], Cfg::TILE);
      tma_load_3d(sm + Cfg::X0, &tm_x, &bar_ld[0], 0, cta * 128, 0);
    }
  }

  const uint32_t idesc_c = umma_idesc_bf16(128, C), idesc_q = umma_idesc_bf16(128, CH);
  uint32_t n_m1 = 0, n_q0 = 0, n_q1 = 0;
  int it = 0, nstore = 0;                                  // nstore: chunk stores issued so far by thread 0 (staging slot = nstore % NSTG)
  for (int tile = cta; tile < p.tiles; tile += ncta, ++it) {
    const int buf = it & 1;
    const int64_t m = (int64_t)tile * 128 + r;
    if (warp == 0 && elect_one()) {
      mbar_wait(&bar_ld[buf], (it >> 1) & 1);
      const int nxt = tile + ncta;
      if (nxt < p.tiles) {
        mbar_arrive_expect_tx(&bar_ld[buf ^ 1], Cfg::TILE);
        tma_load_3d(sm + Cfg::X0 + (buf ^ 1) * Cfg::TILE, &tm_x, &bar_ld[buf ^ 1], 0, nxt * 128, 0);
      }
      if (it == 0) mbar_wait(&bar_w, 0);
      tc_fence_after();
      const uint64_t ad = umma_smem_desc(base + Cfg::X0 + buf * Cfg::TILE, ROWB), bd = umma_smem_desc(base + Cfg::WI, ROWB);
#pragma unroll
      for (int k = 0; k < C / 16; ++k) umma_bf16_ss(tmem, ad + 2 * k, bd + 2 * k, idesc_c, k ? 1u : 0u);
      umma_commit(&bar_m1);
    }
    __syncwarp();

    // ---------------- epilogue 1: x1 = acc + bi -> global; statistics; LN_hat(x1) -> A1 ----------------
    mbar_wait(&bar_m1, n_m1 & 1); ++n_m1;
    tc_fence_after();
    float y[16];
    {
      uint32_t raw[16];
      __syncwarp();
      tmem_ld_x16(lane_addr + t * 16, raw);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) y[j] = __uint_as_float(raw[j]) + s_bi[t * 16 + j];
      if (m < p.M) {
        bf16* dst = p.x1 + m * p.ld1 + t * 16;
        store8(dst, *reinterpret_cast<float(*)[8]>(&y[0]));
        store8(dst + 8, *reinterpret_cast<float(*)[8]>(&y[8]));
      }
      // the statistics and the normalised operand use the bf16-rounded x1 - the tensor every later consumer sees
#pragma unroll
      for (int j = 0; j < 16; ++j) y[j] = __bfloat162float(__float2bfloat16(y[j]));
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) s += y[j];
    s_sum[t][r] = s;
    __syncthreads();
    float mean = 0.f;
#pragma unroll
    for (int tt = 0; tt < NT; ++tt) mean += s_sum[tt][r];
    mean *= 1.0f / C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) { const float d = y[j] - mean; q = fmaf(d, d, q); }
    s_sq[t][r] = q;
    __syncthreads();
    float var = 0.f;
#pragma unroll
    for (int tt = 0; tt < NT; ++tt) var += s_sq[tt][r];
    const float rstd = rsqrtf(var * (1.0f / C) + p.eps);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = (y[h * 8 + j] - mean) * rstd;
      uint32_t off = r * ROWB + (t * 2 + h) * 16;
      off ^= ((off >> 7) & Cfg::SWZ) << 4;
      store8(reinterpret_cast<bf16*>(sm + Cfg::A1 + off), v);
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();

    // ---------------- MMA2: chunks 0 and 1 now (two TMEM buffers), chunk 2 when buffer 0 has been drained ----------------
    auto issue_chunk = [&](int j) {
      const uint64_t ad = umma_smem_desc(base + Cfg::A1, ROWB), bd = umma_smem_desc(base + Cfg::WQ + j * CH * ROWB, ROWB);
#pragma unroll
      for (int k = 0; k < C / 16; ++k) umma_bf16_ss(tmem + Cfg::Q_COL + (j & 1) * CH, ad + 2 * k, bd + 2 * k, idesc_q, k ? 1u : 0u);
      umma_commit(&bar_q[j & 1]);
    };
    if (warp == 0 && elect_one()) { tc_fence_after(); issue_chunk(0); issue_chunk(1); }
    __syncwarp();

#pragma unroll
    for (int j = 0; j < 3; ++j) {
      if ((j & 1) == 0) { mbar_wait(&bar_q[0], n_q0 & 1); ++n_q0; } else { mbar_wait(&bar_q[1], n_q1 & 1); ++n_q1; }
      tc_fence_after();
      const int slot = Cfg::NSTG == 2 ? ((it * 3 + j) & 1) : 0;
      if (warp == 0 && elect_one()) {                                   // the staging slot's previous store must have been read out
        if (Cfg::NSTG == 2) { if (nstore >= 2) tma_store_wait_read_h<1>(); }
        else { if (nstore >= 1) tma_store_wait_read_h<0>(); }
      }
      __syncthreads();
      uint8_t* stg = sm + Cfg::STG + slot * 3 * Cfg::TILE;
#pragma unroll
      for (int c3 = 0; c3 < PT / 16; ++c3) {
        uint32_t raw[16];
        __syncwarp();
        tmem_ld_x16(lane_addr + Cfg::Q_COL + (j & 1) * CH + t * PT + c3 * 16, raw);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int col = t * PT + c3 * 16 + h * 8;     // column inside the chunk
          float v[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) v[e] = __uint_as_float(raw[h * 8 + e]) + s_bq[j * CH + col + e];
          const int box = col / C, cin = col - box * C;
          uint32_t off = r * ROWB + (cin >> 3) * 16;
          off ^= ((off >> 7) & Cfg::SWZ) << 4;
          store8(reinterpret_cast<bf16*>(stg + box * Cfg::TILE + off), v);
        }
      }
      fence_async_smem();
      tc_fence_before();
      __syncthreads();
      if (warp == 0 && elect_one()) {
#pragma unroll
        for (int b = 0; b < 3; ++b) tma_store_3d(&tm_q, stg + b * Cfg::TILE, j * CH + b * C, tile * 128, 0);
        tma_store_commit_h();
        ++nstore;
        if (j == 0) { tc_fence_after(); issue_chunk(2); }   // TMEM buffer 0 drained by every thread (barrier above)
      }
      __syncwarp();
    }
  }
  if (warp == 0 && elect_one()) tma_store_wait_all_h();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, Cfg::TMEM_COLS);
}

template <int C>
static int launch_head(const tfswa_head_args* a, cudaStream_t st) {
  using Cfg = HeadCfg<C>;
  CUtensorMap tm_x, tm_wi, tm_wq, tm_q;
  int rc = make_tmap_bf16_3d(&tm_x, a->x, C, a->M, 1, a->ldx, 0, C, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tm_wi, a->wi, C, C, 1, C, 0, C, C);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tm_wq, a->wq, C, 9 * C, 1, C, 0, C, 3 * C);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tm_q, a->qkv, 9 * C, a->M, 1, a->ldq, 0, C, 128);
  if (rc) return rc;
  static int sms = 0;
  static DeviceOnce attr_once;
  const size_t smem = Cfg::BYTES + 1024;
  if (attr_once.needed()) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(tc_head_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess || sms <= 0) { set_error("block_head_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return TFSWA_ECUDA; }
    attr_once.done();
  }
  HeadParams p = {};
  p.bi = a->bi; p.bq = a->bq; p.x1 = (bf16*)a->x1; p.ld1 = a->ld1; p.M = a->M; p.eps = a->eps;
  p.tiles = (int)ceil_div64(a->M, 128);
  int grid = sms * (C == 32 ? 2 : 1);
  if (grid > p.tiles) grid = p.tiles;
  tc_head_kernel<C><<<grid, Cfg::THREADS, smem, st>>>(tm_x, tm_wi, tm_wq, tm_q, p);
  return check_launch("block_head_tc");
}

}  // namespace tfswa

using namespace tfswa;

extern "C" int tfswa_block_head_tc_fwd(const tfswa_head_args* a, void* stream) {
  TFSWA_REQUIRE(a && a->x && a->wi && a->wq && a->bi && a->bq && a->x1 && a->qkv, "block_head_tc: null pointer");
  TFSWA_REQUIRE(a->M > 0 && a->M / 128 < (1 << 24), "block_head_tc: bad M");
  TFSWA_REQUIRE(a->C == 32 || a->C == 64, "block_head_tc: C=%d not in {32, 64} (use the unfused tfswa_linear_tc_fwd sequence)", a->C);
  TFSWA_REQUIRE(a->ldx % 8 == 0 && a->ld1 % 8 == 0 && a->ldq % 8 == 0, "block_head_tc: 16-byte alignment of row strides");
  TFSWA_REQUIRE((((uintptr_t)a->x1) & 15) == 0, "block_head_tc: x1 alignment");
  TFSWA_REQUIRE(a->eps > 0.f, "block_head_tc: eps must be positive");
  return a->C == 32 ? launch_head<32>(a, (cudaStream_t)stream) : launch_head<64>(a, (cudaStream_t)stream);
}
