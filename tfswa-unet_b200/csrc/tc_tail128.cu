// Fused branch tail (see tc_tail.cu for the math) for C = 128, hidden = 512: the three weight matrices of a branch are
// 288 KB of bf16 and do not fit in shared memory, so they are STREAMED from L2 by TMA through a 3-deep ring of 32 KB
// slots while the token tile, y, LN_hat(y) and the hidden activations stay on the SM.
//
// One persistent CTA per SM: 16 epilogue warps (one token row per TMEM lane, 4 threads per row) and a driver warp whose
// lane 0 runs the TMA ring and issues every MMA; the two sides only meet on mbarriers (no CTA-wide barrier in the tile
// loop: with thread 0 of an epilogue warp doing the issue work, 27 % of all samples sat in __syncthreads waiting for it).  The hidden dimension is processed in 8 chunks of 64 columns:
//     MMA1   Y  = att  Wp^T            (N=128, K=128)   ring pair A = {Wp k-block 0, Wp k-block 1}
//     MMA2_c Hc = LN_hat(y) W1c^T      (N=64,  K=128)   W1c = rows [64c, 64c+64) of W1'
//     MMA3_c Z += gelu(Hc + b1) W2c^T  (N=128, K=64)    W2c = columns [64c, 64c+64) of W2
// gelu(Hc) goes back to TMEM as bf16 (tcgen05.st) and is the A operand of MMA3_c straight from there - no shared-
// memory tile, no generic->async proxy fence per chunk.  The hidden accumulator and operand are double buffered, and MMA2_{c+2} / MMA3_c are issued right
// after epilogue c, so the tensor pipe (49 % busy) runs entirely underneath the GELU epilogues, which bound the kernel.
// Ring schedule (pairs consumed in this order, each prefetched two epilogues ahead):
//     A = {Wp}, B = {W1_0, W1_1}, P_c = {W2_c, W1_{c+2}} for c = 0..7 (W1_8, W1_9 do not exist).
#include "common.cuh"
#include "sm100.cuh"

namespace tfswa {

using namespace sm100;

struct Tail128Params {
  const float* bp; const float* b1; const float* b2;     // (nb,128), (nb,512), (nb,128)
  const bf16* res; int64_t ldr, res_bs;                  // residual x1 (M,128), shared by the branches when res_bs == 0
  int64_t M;
  int nb, tiles;
  float eps;
};

constexpr int T8_C = 128, T8_HID = 512, T8_HC = 64, T8_NCH = 8;
constexpr int T8_THREADS = 512;
constexpr int T8_SLOT = 32768;                           // one ring slot = one pair
constexpr int T8_RING = 0;                               // 3 slots
constexpr int T8_AT = 3 * T8_SLOT;                       // att tile: 2 k-blocks x (128 rows x 128 B)
constexpr int T8_A1 = T8_AT + 32768;                     // LN_hat(y) operand, later the output staging tile
constexpr int T8_BYTES = T8_A1 + 32768;                  // 160 KB (the hidden operand tiles live in TMEM)
constexpr uint32_t T8_TMEM = 512;                        // Y/Z [0,128), H0 [128,192), H1 [192,256), gelu(H) as bf16 A operand: [256,288), [288,320)
constexpr uint32_t T8_G_COL = 256;
constexpr int T8_PAIRS = 10;                             // ring pairs per tile

__device__ __forceinline__ void t8_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void t8_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void t8_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void __launch_bounds__(T8_THREADS + 32, 1)
tc_tail128_kernel(const __grid_constant__ CUtensorMap tm_att, const __grid_constant__ CUtensorMap tm_wp,
                  const __grid_constant__ CUtensorMap tm_w1, const __grid_constant__ CUtensorMap tm_w2,
                  const __grid_constant__ CUtensorMap tm_out, const Tail128Params p) {
  extern __shared__ uint8_t smem_raw[];
  // driver -> epilogue: bar_m1, bar_h[2], bar_a2[2], bar_z (count 1, tcgen05.commit);  epilogue -> driver: bar_e1,
  // bar_e2[2], bar_e3 (count 16, one arrival per epilogue warp);  TMA: bar_full[3], bar_at;  MMA -> ring: bar_empty[3]
  __shared__ __align__(8) uint64_t bar_full[3], bar_empty[3], bar_at, bar_m1, bar_h[2], bar_a2[2], bar_z, bar_e1, bar_e2[2], bar_e3;
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) float s_bp[T8_C], s_b1[T8_HID], s_b2[T8_C];
  __shared__ float s_sum[4][128], s_sq[4][128];

  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sm = smem_raw + (base - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool driver = warp == 16;                        // warps 0-15: epilogue; warp 16: TMA ring + every MMA (lane 0)
  const int quad = warp & 3, t = (warp >> 2) & 3;        // TMEM lane quadrant, column quarter
  const int r = quad * 32 + lane;
  const int z = blockIdx.x % p.nb;
  const int cta = blockIdx.x / p.nb, ncta = gridDim.x / p.nb;

  if (warp == 0) {
    if (lane == 0) {
      prefetch_tmap(&tm_att); prefetch_tmap(&tm_wp); prefetch_tmap(&tm_w1); prefetch_tmap(&tm_w2); prefetch_tmap(&tm_out);
      for (int i = 0; i < 3; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
      mbar_init(&bar_at, 1); mbar_init(&bar_m1, 1); mbar_init(&bar_z, 1); mbar_init(&bar_e1, 16); mbar_init(&bar_e3, 16);
      for (int i = 0; i < 2; ++i) { mbar_init(&bar_h[i], 1); mbar_init(&bar_a2[i], 1); mbar_init(&bar_e2[i], 16); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&s_tmem, T8_TMEM);
  }
  for (int i = tid; i < T8_C; i += T8_THREADS + 32) { s_bp[i] = p.bp[z * T8_C + i]; s_b2[i] = p.b2[z * T8_C + i]; }
  for (int i = tid; i < T8_HID; i += T8_THREADS + 32) s_b1[i] = p.b1[z * T8_HID + i];
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = s_tmem;
  const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);

  if (driver) {
    if (elect_one()) {
      // ---- the weight ring.  Pair G (global counter) lives in slot G % 3; q = G % 10 selects its contents. ----
      auto produce = [&](int G) {
        const int slot = G % 3, q = G % T8_PAIRS;
        if (G >= 3) mbar_wait(&bar_empty[slot], ((G / 3) - 1) & 1);    // the MMAs that read pair G-3 have completed
        uint8_t* dst = sm + T8_RING + slot * T8_SLOT;
        if (q == 0) {                                    // A: Wp as two k-blocks of [128 rows][128 B]
          mbar_arrive_expect_tx(&bar_full[slot], 32768);
          tma_load_3d(dst, &tm_wp, &bar_full[slot], 0, 0, z);
          tma_load_3d(dst + 16384, &tm_wp, &bar_full[slot], 64, 0, z);
        } else if (q == 1) {                             // B: W1 chunks 0 and 1, each two k-blocks of [64 rows][128 B]
          mbar_arrive_expect_tx(&bar_full[slot], 32768);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            tma_load_3d(dst + cc * 16384, &tm_w1, &bar_full[slot], 0, cc * T8_HC, z);
            tma_load_3d(dst + cc * 16384 + 8192, &tm_w1, &bar_full[slot], 64, cc * T8_HC, z);
          }
        } else {                                         // P_c: W2 chunk c [128 rows][128 B]  |  W1 chunk c+2
          const int c = q - 2;
          mbar_arrive_expect_tx(&bar_full[slot], c + 2 < T8_NCH ? 32768 : 16384);
          tma_load_3d(dst, &tm_w2, &bar_full[slot], c * T8_HC, 0, z);
          if (c + 2 < T8_NCH) {
            tma_load_3d(dst + 16384, &tm_w1, &bar_full[slot], 0, (c + 2) * T8_HC, z);
            tma_load_3d(dst + 16384 + 8192, &tm_w1, &bar_full[slot], 64, (c + 2) * T8_HC, z);
          }
        }
      };
      auto wait_full = [&](int G) { mbar_wait(&bar_full[G % 3], (G / 3) & 1); };
      const uint32_t idesc_y = umma_idesc_bf16(128, T8_C), idesc_h = umma_idesc_bf16(128, T8_HC);
      // D[tmem_d] (+)= A (k-blocks of [128 rows][128 B] at a_addr, stride 16 KB) * B^T (k-blocks at b_addr, stride b_kb_bytes)
      auto mma_k = [&](uint32_t tmem_d, uint32_t a_addr, uint32_t b_addr, uint32_t b_kb_bytes, int kblocks, uint32_t idesc, bool acc) {
        for (int kb = 0; kb < kblocks; ++kb) {
          const uint64_t ad = umma_smem_desc(a_addr + kb * 16384, 128), bd = umma_smem_desc(b_addr + kb * b_kb_bytes, 128);
#pragma unroll
          for (int k = 0; k < 4; ++k) umma_bf16_ss(tmem_d, ad + 2 * k, bd + 2 * k, idesc, (acc || kb || k) ? 1u : 0u);
        }
      };
      auto load_att = [&](int tile) {
        mbar_arrive_expect_tx(&bar_at, 32768);
        tma_load_3d(sm + T8_AT, &tm_att, &bar_at, 0, tile * 128, z);
        tma_load_3d(sm + T8_AT + 16384, &tm_att, &bar_at, 64, tile * 128, z);
      };
      int G = 0;
      uint32_t n_e1 = 0, n_e20 = 0, n_e21 = 0, n_e3 = 0, n_m1 = 0;
      produce(0);
      produce(1);
      if (cta < p.tiles) load_att(cta);
      int it = 0;
      for (int tile = cta; tile < p.tiles; tile += ncta, ++it) {
        // MMA1 (pair A); the commit that releases epilogue 1 is issued only after the previous tile's output store has
        // read the staging tile (A1), which epilogue 1 overwrites
        mbar_wait(&bar_at, it & 1);
        wait_full(G);
        tc_fence_after();
        mma_k(tmem, base + T8_AT, base + T8_RING + (G % 3) * T8_SLOT, 16384, 2, idesc_y, false);
        umma_commit(&bar_empty[G % 3]);
        if (it > 0) t8_store_wait_read();
        umma_commit(&bar_m1);
        produce(G + 2);
        ++G;
        mbar_wait(&bar_m1, n_m1 & 1); ++n_m1;            // att tile consumed: prefetch the next one
        if (tile + ncta < p.tiles) load_att(tile + ncta);
        // MMA2_0, MMA2_1 (pair B) once LN_hat(y) is staged
        mbar_wait(&bar_e1, n_e1 & 1); ++n_e1;
        wait_full(G);
        tc_fence_after();
        {
          const uint32_t w = base + T8_RING + (G % 3) * T8_SLOT;
          mma_k(tmem + 128, base + T8_A1, w, 8192, 2, idesc_h, false);
          umma_commit(&bar_h[0]);
          mma_k(tmem + 192, base + T8_A1, w + 16384, 8192, 2, idesc_h, false);
          umma_commit(&bar_h[1]);
          umma_commit(&bar_empty[G % 3]);
          produce(G + 2);
          ++G;
        }
        for (int c = 0; c < T8_NCH; ++c) {               // pair P_c: MMA3_c, then MMA2_{c+2} into the hidden buffer just drained
          const int b = c & 1;
          if (b == 0) { mbar_wait(&bar_e2[0], n_e20 & 1); ++n_e20; } else { mbar_wait(&bar_e2[1], n_e21 & 1); ++n_e21; }
          wait_full(G);
          tc_fence_after();
          const uint32_t w = base + T8_RING + (G % 3) * T8_SLOT;
#pragma unroll
          for (int k = 0; k < 4; ++k)                      // A = gelu(H_c) from TMEM (8 columns per K=16 step), B = W2_c
            umma_bf16_ts(tmem, tmem + T8_G_COL + b * 32 + k * 8, umma_smem_desc(w, 128) + 2 * k, idesc_y, (c | k) ? 1u : 0u);
          umma_commit(&bar_a2[b]);
          if (c == T8_NCH - 1) umma_commit(&bar_z);
          if (c + 2 < T8_NCH) {
            mma_k(tmem + 128 + b * 64, base + T8_A1, w + 16384, 8192, 2, idesc_h, false);
            umma_commit(&bar_h[b]);
          }
          umma_commit(&bar_empty[G % 3]);
          produce(G + 2);
          ++G;
        }
        mbar_wait(&bar_e3, n_e3 & 1); ++n_e3;            // output staged
        tma_store_3d(&tm_out, sm + T8_A1, 0, tile * 128, z);
        tma_store_3d(&tm_out, sm + T8_A1 + 16384, 64, tile * 128, z);
        t8_store_commit();
      }
      t8_store_wait_all();
    }
  } else {
    uint32_t n_m1 = 0, n_h0 = 0, n_h1 = 0, n_a20 = 0, n_a21 = 0, n_z = 0;
    int it = 0;
    for (int tile = cta; tile < p.tiles; tile += ncta, ++it) {
      const int64_t m = (int64_t)tile * 128 + r;
      // ---------------- epilogue 1: y = acc + bp + x1, statistics, LN_hat(y) -> A1 ----------------
      // the residual does not depend on the MMA: its loads are issued before the wait so their latency hides behind it
      uint4 rraw[4];
      {
        const bf16* rrow = p.res + (int64_t)z * p.res_bs + (m < p.M ? m : 0) * p.ldr + t * 32;
#pragma unroll
        for (int h = 0; h < 4; ++h) rraw[h] = *reinterpret_cast<const uint4*>(rrow + h * 8);
      }
      mbar_wait(&bar_m1, n_m1 & 1); ++n_m1;
      tc_fence_after();
      float y[32];
      {
        uint32_t raw[32];
        __syncwarp();
        tmem_ld_x32(lane_addr + t * 32, raw);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const __nv_bfloat162* hh = reinterpret_cast<const __nv_bfloat162*>(&rraw[h]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            y[h * 8 + 2 * j] = __uint_as_float(raw[h * 8 + 2 * j]) + s_bp[t * 32 + h * 8 + 2 * j] + __low2float(hh[j]);
            y[h * 8 + 2 * j + 1] = __uint_as_float(raw[h * 8 + 2 * j + 1]) + s_bp[t * 32 + h * 8 + 2 * j + 1] + __high2float(hh[j]);
          }
        }
      }
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) s += y[j];
      s_sum[t][r] = s;
      asm volatile("bar.sync 1, 512;" ::: "memory");     // the 16 epilogue warps only
      const float mean = (s_sum[0][r] + s_sum[1][r] + s_sum[2][r] + s_sum[3][r]) * (1.0f / T8_C);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j) { const float d = y[j] - mean; q = fmaf(d, d, q); }
      s_sq[t][r] = q;
      asm volatile("bar.sync 1, 512;" ::: "memory");
      const float rstd = rsqrtf((s_sq[0][r] + s_sq[1][r] + s_sq[2][r] + s_sq[3][r]) * (1.0f / T8_C) + p.eps);
      {
        uint8_t* a1 = sm + T8_A1 + (t >> 1) * 16384;     // k-block of my 32 columns
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (y[h * 8 + j] - mean) * rstd;
          uint32_t off = r * 128 + ((t & 1) * 4 + h) * 16;
          off ^= ((off >> 7) & 7u) << 4;
          store8(reinterpret_cast<bf16*>(a1 + off), v);
        }
      }
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_e1);

      // ---------------- hidden chunks: gelu(acc + b1') -> A2[b] ----------------
#pragma unroll 1
      for (int c = 0; c < T8_NCH; ++c) {
        const int b = c & 1;
        if (b == 0) { mbar_wait(&bar_h[0], n_h0 & 1); ++n_h0; } else { mbar_wait(&bar_h[1], n_h1 & 1); ++n_h1; }
        if (it > 0 || c >= 2) {                          // MMA3 of the chunk that last read operand buffer b has completed
          if (b == 0) { mbar_wait(&bar_a2[0], n_a20 & 1); ++n_a20; } else { mbar_wait(&bar_a2[1], n_a21 & 1); ++n_a21; }
        }
        tc_fence_after();
        uint32_t raw[16];
        __syncwarp();
        tmem_ld_x16(lane_addr + 128 + b * 64 + t * 16, raw);
        tmem_ld_wait();
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 bb = *reinterpret_cast<const float2*>(&s_b1[c * T8_HC + t * 16 + 2 * j]);
          pk[j] = gelu_hidden_bf16x2(__uint_as_float(raw[2 * j]), __uint_as_float(raw[2 * j + 1]), bb.x, bb.y);
        }
        __syncwarp();
        tmem_st_x8(lane_addr + T8_G_COL + b * 32 + t * 8, pk);      // hidden columns (2j, 2j+1) -> 32-bit cell j: A operand of MMA3_c
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar_e2[b]);
      }

      // ---------------- epilogue 3: z = acc + b2 + y -> staging (A1) ----------------
      mbar_wait(&bar_z, n_z & 1); ++n_z;
      tc_fence_after();
      {
        uint32_t raw[32];
        __syncwarp();
        tmem_ld_x32(lane_addr + t * 32, raw);
        tmem_ld_wait();
        uint8_t* a1 = sm + T8_A1 + (t >> 1) * 16384;
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(raw[h * 8 + j]) + s_b2[t * 32 + h * 8 + j] + y[h * 8 + j];
          uint32_t off = r * 128 + ((t & 1) * 4 + h) * 16;
          off ^= ((off >> 7) & 7u) << 4;
          store8(reinterpret_cast<bf16*>(a1 + off), v);
        }
      }
      fence_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar_e3);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, T8_TMEM);
}

int launch_tail128(const tfswa_tail_args* a, cudaStream_t st) {
  const int nb = a->batch;
  CUtensorMap tm_att, tm_wp, tm_w1, tm_w2, tm_out;
  int rc = make_tmap_bf16_3d(&tm_att, a->att, T8_C, a->M, nb, a->lda, a->att_bs, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tm_wp, a->wp, T8_C, T8_C, nb, T8_C, (uint64_t)T8_C * T8_C, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tm_w1, a->w1, T8_C, T8_HID, nb, T8_C, (uint64_t)T8_HID * T8_C, 64, T8_HC);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tm_w2, a->w2, T8_HID, T8_C, nb, T8_HID, (uint64_t)T8_HID * T8_C, 64, 128);
  if (rc) return rc;
  rc = make_tmap_bf16_3d(&tm_out, a->out, T8_C, a->M, nb, a->ldo, a->out_bs, 64, 128);
  if (rc) return rc;
  static int sms = 0;
  static DeviceOnce attr_once;
  const size_t smem = T8_BYTES + 1024;
  if (attr_once.needed()) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaFuncSetAttribute(tc_tail128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess || sms <= 0) { set_error("branch_tail_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return TFSWA_ECUDA; }
    attr_once.done();
  }
  Tail128Params p = {};
  p.bp = a->bp; p.b1 = a->b1; p.b2 = a->b2; p.res = (const bf16*)a->res; p.ldr = a->ldr;
  p.res_bs = (a->res_bs != 0 && nb > 1) ? a->res_bs : 0;
  p.M = a->M; p.nb = nb; p.eps = a->eps;
  p.tiles = (int)ceil_div64(a->M, 128);
  int per_branch = sms / nb;
  if (per_branch < 1) per_branch = 1;
  if (per_branch > p.tiles) per_branch = p.tiles;
  tc_tail128_kernel<<<per_branch * nb, T8_THREADS + 32, smem, st>>>(tm_att, tm_wp, tm_w1, tm_w2, tm_out, p);
  return check_launch("branch_tail_tc");
}

}  // namespace tfswa
