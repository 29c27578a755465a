// Fused tail of a pre-LN transformer branch on the tcgen05 tensor cores, for the narrow stages (C = 32, 64):
//
//     y = att Wp^T + bp + x1                       attention.py:86 (+ residual :146/:220/:378)
//     z = y + W2 gelu(W1' LN_hat(y) + b1') + b2    attention.py:121-128, :159/:233/:387   (W1' = W1 diag(gamma), b1' = W1 beta + b1)
//
// Unfused this is three HBM-bound GEMM launches plus a statistics pass that move 86*C bytes per token and branch
// (y written + read three times, the 4C-wide hidden tile written + read); fused, a 128-token tile of one branch stays on
// the SM - y in registers, LN_hat(y) and the hidden tile in shared memory as the next MMA's A operand, all three
// accumulators in TMEM - and 7*C bytes per token cross HBM (att in, x1 in, z out).  The kernel is then bound by the
// epilogue arithmetic (12*C erf-GELUs per token), not by memory.
//
// One CTA = one branch, persistent over 128-token tiles; all three weight matrices of the branch are TMA-loaded once
// and stay in shared memory (18 KB at C = 32, 72 KB at C = 64).  Per tile:
//     TMA(att, x1; prefetched one tile ahead) -> MMA1 (N=C,  K=C)  -> epilogue 1: + bias + residual, row mean / rstd,
//                                                                     LN_hat(y) -> smem (K-major, swizzled)
//                                             -> MMA2 (N=4C, K=C)  -> epilogue 2: + bias, erf-GELU -> smem
//                                             -> MMA3 (N=C,  K=4C) -> epilogue 3: + bias + y -> smem -> TMA store
// No warp specialisation: every phase depends on the previous one, so thread 0 issues the (tiny) MMAs and every warp
// is an epilogue warp - one token row per TMEM lane, NT = C/16 threads per row splitting the columns, so each thread
// owns 16 columns of y/z and 64 hidden columns (= one 128-byte swizzle row of the hidden operand).  Latency is hidden
// by the co-resident CTA (C = 32: 2 per SM) or the 16 warps of the CTA (C = 64).
#include "common.cuh"
#include "sm100.cuh"

namespace tfswa {

using namespace sm100;

struct TailParams {
  const float* bp; const float* b1; const float* b2;     // (nb, C), (nb, 4C), (nb, C)
  int64_t M;
  int nb, res_batched, tiles;
  float eps;
};

template <int C> struct TailCfg {
  static constexpr int HID = 4 * C;
  static constexpr int NT = C / 16;                      // threads per token row
  static constexpr int THREADS = 128 * NT;               // 256 / 512
  static constexpr int ROWB = C * 2;                     // bytes per row of the C-wide operand tiles == their swizzle span
  static constexpr uint32_t SWZ = ROWB / 16 - 1;
  static constexpr int KB2 = HID / 64;                   // 128-byte K blocks of the hidden operand (== NT)
  static constexpr int TILE = 128 * ROWB;
  static constexpr int WP = 0;                           // C x C
  static constexpr int W1 = WP + C * ROWB;               // 4C x C
  static constexpr int W2 = W1 + HID * ROWB;             // KB2 blocks of (C rows x 128 B)
  static constexpr int A0 = W2 + KB2 * C * 128;          // 2 attention tiles (prefetch double buffer)
  static constexpr int RS = A0 + 2 * TILE;               // 2 residual tiles
  static constexpr int A1 = RS + 2 * TILE;               // LN_hat(y) operand, later the output staging tile
  static constexpr int A2 = A1 + TILE;                   // hidden operand: KB2 blocks of (128 rows x 128 B)
  static constexpr int BYTES = A2 + KB2 * 16384;         // 90 KB (C = 32) / 216 KB (C = 64)
  static constexpr uint32_t TMEM_COLS = C == 32 ? 256 : 512;
  static constexpr uint32_t H_COL = 64;                  // y / z accumulator at column 0, hidden accumulator at 64
  static constexpr uint32_t W_BYTES = C * ROWB + HID * ROWB + KB2 * C * 128;
};

__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int C>
__global__ void __launch_bounds__(TailCfg<C>::THREADS, C == 32 ? 2 : 1)
tc_tail_kernel(const __grid_constant__ CUtensorMap tm_att, const __grid_constant__ CUtensorMap tm_res,
               const __grid_constant__ CUtensorMap tm_wp, const __grid_constant__ CUtensorMap tm_w1,
               const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_out, const TailParams p) {
  using Cfg = TailCfg<C>;
  constexpr int NT = Cfg::NT, ROWB = Cfg::ROWB, HID = Cfg::HID;
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t bar_w, bar_ld[2], bar_mma;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_bp[C], s_b1[HID], s_b2[C];
  __shared__ float s_sum[NT][128], s_sq[NT][128];

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int quad = warp & 3, t = warp >> 2;              // TMEM lane quadrant of this warp, column split of this thread
  const int r = quad * 32 + lane;                        // token row inside the tile == TMEM lane
  const int z = blockIdx.x % p.nb;                       // adjacent CTAs = the branches of one tile sequence (x1 shared via L2)
  const int cta = blockIdx.x / p.nb, ncta = gridDim.x / p.nb;
  const int rz = p.res_batched ? z : 0;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tm_att); prefetch_tmap(&tm_res); prefetch_tmap(&tm_out);
      mbar_init(&bar_w, 1); mbar_init(&bar_ld[0], 1); mbar_init(&bar_ld[1], 1); mbar_init(&bar_mma, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&s_tmem, Cfg::TMEM_COLS);
  }
  for (int i = tid; i < C; i += Cfg::THREADS) { s_bp[i] = p.bp[z * C + i]; s_b2[i] = p.b2[z * C + i]; }
  for (int i = tid; i < HID; i += Cfg::THREADS) s_b1[i] = p.b1[z * HID + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);

  if (warp == 0 && elect_one()) {
    mbar_arrive_expect_tx(&bar_w, Cfg::W_BYTES);
    tma_load_3d(sm + Cfg::WP, &tm_wp, &bar_w, 0, 0, z);
    tma_load_3d(sm + Cfg::W1, &tm_w1, &bar_w, 0, 0, z);
#pragma unroll
    for (int kb = 0; kb < Cfg::KB2; ++kb) tma_load_3d(sm + Cfg::W2 + kb * C * 128, &tm_w2, &bar_w, kb * 64, 0, z);
    if (cta < p.tiles) {
      mbar_arrive_expect_tx(&bar_ld[0], 2 * Cfg::TILE);
      tma_load_3d(sm + Cfg::A0, &tm_att, &bar_ld[0], 0, cta * 128, z);
      tma_load_3d(sm + Cfg::RS, &tm_res, &bar_ld[0], 0, cta * 128, rz);
    }
  }

  const uint32_t idesc_c = umma_idesc_bf16(128, C), idesc_h = umma_idesc_bf16(128, HID);
  uint32_t n_mma = 0;
  int it = 0;
  for (int tile = cta; tile < p.tiles; tile += ncta, ++it) {
    const int buf = it & 1;
    mbar_wait(&bar_ld[buf], (it >> 1) & 1);              // att + x1 tiles landed (every thread reads the residual tile)
    if (warp == 0 && elect_one()) {
      const int nxt = tile + ncta;
      if (nxt < p.tiles) {                               // prefetch the next tile into the other buffer (its readers are long done)
        mbar_arrive_expect_tx(&bar_ld[buf ^ 1], 2 * Cfg::TILE);
        tma_load_3d(sm + Cfg::A0 + (buf ^ 1) * Cfg::TILE, &tm_att, &bar_ld[buf ^ 1], 0, nxt * 128, z);
        tma_load_3d(sm + Cfg::RS + (buf ^ 1) * Cfg::TILE, &tm_res, &bar_ld[buf ^ 1], 0, nxt * 128, rz);
      }
      if (it == 0) mbar_wait(&bar_w, 0);
      tc_fence_after();
      const uint64_t ad = umma_smem_desc(base + Cfg::A0 + buf * Cfg::TILE, ROWB), bd = umma_smem_desc(base + Cfg::WP, ROWB);
#pragma unroll
      for (int k = 0; k < C / 16; ++k) umma_bf16_ss(tmem, ad + 2 * k, bd + 2 * k, idesc_c, k ? 1u : 0u);
      umma_commit(&bar_mma);
      if (it > 0) tma_store_wait_read();                 // previous tile's output left the staging tile (A1)
    }
    __syncwarp();

    // ---------------- epilogue 1: y = acc + bp + x1, row statistics, LN_hat(y) -> A1 ----------------
    mbar_wait(&bar_mma, n_mma & 1); ++n_mma;
    tc_fence_after();
    float y[16];
    {
      uint32_t raw[16];
      __syncwarp();
      tmem_ld_x16(lane_addr + t * 16, raw);
      tmem_ld_wait();
      const uint8_t* rs = sm + Cfg::RS + buf * Cfg::TILE;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t off = r * ROWB + (t * 2 + h) * 16;
        off ^= ((off >> 7) & Cfg::SWZ) << 4;
        float res[8];
        load8(reinterpret_cast<const bf16*>(rs + off), res);
#pragma unroll
        for (int j = 0; j < 8; ++j) y[h * 8 + j] = __uint_as_float(raw[h * 8 + j]) + s_bp[t * 16 + h * 8 + j] + res[j];
      }
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

    // ---------------- MMA2: hidden = LN_hat(y) W1'^T ----------------
    if (warp == 0 && elect_one()) {
      tc_fence_after();
      const uint64_t ad = umma_smem_desc(base + Cfg::A1, ROWB), bd = umma_smem_desc(base + Cfg::W1, ROWB);
#pragma unroll
      for (int k = 0; k < C / 16; ++k) umma_bf16_ss(tmem + Cfg::H_COL, ad + 2 * k, bd + 2 * k, idesc_h, k ? 1u : 0u);
      umma_commit(&bar_mma);
    }
    __syncwarp();
    mbar_wait(&bar_mma, n_mma & 1); ++n_mma;
    tc_fence_after();

    // ---------------- epilogue 2: gelu(acc + b1') -> A2 (my 64 columns = one 128-byte row of K block t) ----------------
    {
      uint8_t* a2 = sm + Cfg::A2 + t * 16384;
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t raw[32];
        __syncwarp();
        tmem_ld_x32(lane_addr + Cfg::H_COL + t * 64 + ch * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint32_t v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 bb = *reinterpret_cast<const float2*>(&s_b1[t * 64 + ch * 32 + q4 * 8 + 2 * j]);
            v[j] = gelu_hidden_bf16x2(__uint_as_float(raw[q4 * 8 + 2 * j]), __uint_as_float(raw[q4 * 8 + 2 * j + 1]), bb.x, bb.y);
          }
          uint32_t off = r * 128 + (ch * 4 + q4) * 16;
          off ^= ((off >> 7) & 7u) << 4;
          *reinterpret_cast<uint4*>(a2 + off) = make_uint4(v[0], v[1], v[2], v[3]);
        }
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();

    // ---------------- MMA3: z_acc = hidden W2^T (into the y accumulator columns, long since consumed) ----------------
    if (warp == 0 && elect_one()) {
      tc_fence_after();
#pragma unroll
      for (int kb = 0; kb < Cfg::KB2; ++kb) {
        const uint64_t ad = umma_smem_desc(base + Cfg::A2 + kb * 16384, 128), bd = umma_smem_desc(base + Cfg::W2 + kb * C * 128, 128);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem, ad + 2 * k, bd + 2 * k, idesc_c, (kb | k) ? 1u : 0u);
      }
      umma_commit(&bar_mma);
    }
    __syncwarp();
    mbar_wait(&bar_mma, n_mma & 1); ++n_mma;
    tc_fence_after();

    // ---------------- epilogue 3: z = acc + b2 + y -> staging tile (A1) -> TMA store ----------------
    {
      uint32_t raw[16];
      __syncwarp();
      tmem_ld_x16(lane_addr + t * 16, raw);
      tmem_ld_wait();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(raw[h * 8 + j]) + s_b2[t * 16 + h * 8 + j] + y[h * 8 + j];
        uint32_t off = r * ROWB + (t * 2 + h) * 16;
        off ^= ((off >> 7) & Cfg::SWZ) << 4;
        store8(reinterpret_cast<bf16*>(sm + Cfg::A1 + off), v);
      }
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    if (warp == 0 && elect_one()) {
      tma_store_3d(&tm_out, sm + Cfg::A1, 0, tile * 128, z);     // rows beyond M are clipped by the tensor map
      tma_store_commit();
    }
    __syncwarp();
  }
  if (warp == 0 && elect_one()) tma_store_wait_all();
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, Cfg::TMEM_COLS);
}

template <int C>
static int launch_tail(const tfswa_tail_args* a, cudaStream_t st) {
  using Cfg = TailCfg<C>;
  const int nb = a->batch;
  CUtensorMap tm_att, tm_res, tm_wp, tm_w1, tm_w2, tm_out;
  int rc = make_tmap_bf16_3d(&tm_att, a->att, C, a->M, nb, a->lda, a->att_bs, C, 128);
  if (rc) return rc;
  const bool res_batched = a->res_bs != 0 && nb > 1;
  rc = make_tmap_bf16_3d(&tm_res, a->res, C, a->M, res_batched ? nb : 1, a->ldr, a->res_bs, C, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tm_wp, a->wp, C, C, nb, C, (uint64_t)C * C, C, C);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tm_w1, a->w1, C, Cfg::HID, nb, C, (uint64_t)Cfg::HID * C, C, Cfg::HID);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tm_w2, a->w2, Cfg::HID, C, nb, Cfg::HID, (uint64_t)Cfg::HID * C, 64, C);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tm_out, a->out, C, a->M, nb, a->ldo, a->out_bs, C, 128);
  if (rc) return rc;
  static int sms = 0;
  static DeviceOnce attr_once;
  const size_t smem = Cfg::BYTES + 1024;
  if (attr_once.needed()) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(tc_tail_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess || sms <= 0) { set_error("branch_tail_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return TFSWA_ECUDA; }
    attr_once.done();
  }
  TailParams p = {};
  p.bp = a->bp; p.b1 = a->b1; p.b2 = a->b2; p.M = a->M; p.nb = nb; p.res_batched = res_batched ? 1 : 0; p.eps = a->eps;
  p.tiles = (int)ceil_div64(a->M, 128);
  int per_branch = (sms * (C == 32 ? 2 : 1)) / nb;
  if (per_branch < 1) per_branch = 1;
  if (per_branch > p.tiles) per_branch = p.tiles;
  tc_tail_kernel<C><<<per_branch * nb, Cfg::THREADS, smem, st>>>(tm_att, tm_res, tm_wp, tm_w1, tm_w2, tm_out, p);
  return check_launch("branch_tail_tc");
}

int launch_tail128(const tfswa_tail_args* a, cudaStream_t st);   // tc_tail128.cu: weights streamed through a TMA ring

}  // namespace tfswa

using namespace tfswa;

extern "C" int tfswa_branch_tail_tc_fwd(const tfswa_tail_args* a, void* stream) {
  TFSWA_REQUIRE(a && a->att && a->res && a->wp && a->w1 && a->w2 && a->bp && a->b1 && a->b2 && a->out, "branch_tail_tc: null pointer");
  TFSWA_REQUIRE(a->M > 0 && a->M / 128 < (1 << 24) && a->batch > 0 && a->batch <= 64, "branch_tail_tc: bad M / batch");
  TFSWA_REQUIRE(a->C == 32 || a->C == 64 || a->C == 128, "branch_tail_tc: C=%d not in {32, 64, 128} (use the unfused tfswa_linear_tc_fwd sequence)", a->C);
  TFSWA_REQUIRE(a->hidden == 4 * a->C, "branch_tail_tc: hidden=%d must be 4*C", a->hidden);
  TFSWA_REQUIRE(a->lda % 8 == 0 && a->att_bs % 8 == 0 && a->ldr % 8 == 0 && a->res_bs % 8 == 0 && a->ldo % 8 == 0 && a->out_bs % 8 == 0,
                "branch_tail_tc: 16-byte alignment of ld/strides");
  TFSWA_REQUIRE(a->eps > 0.f, "branch_tail_tc: eps must be positive");
  if (a->C == 128) return launch_tail128(a, (cudaStream_t)stream);
  return a->C == 32 ? launch_tail<32>(a, (cudaStream_t)stream) : launch_tail<64>(a, (cudaStream_t)stream);
}
