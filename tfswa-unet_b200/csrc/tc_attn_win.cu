// SW-MSA window attention (attention.py:347-403) on tcgen05 with TMA-fed operands (head_dim 4 and 8, bf16).
//
// The pad / roll / window_partition copies of attention.py:358-375 are a TMA tensor map here: the q|k|v token matrix
// (M, ldq) is described as a 4-D tensor (channel, w, h, b) and ONE box (8 channels, 8, 8, 1) at (c, 8 ww + shift,
// 8 wh + shift, b) is a window's 64 tokens of an 8-channel group, landing in shared memory as [token][16 B] in the
// reference's window order (token = 8 i + j).  That image is at once the canonical no-swizzle K-MAJOR operand of
// S = Q K^T (keys = N) and the MN-MAJOR operand of O = P V (channels = N, keys = K): K and V reach the tensor core
// without a thread touching them; the window's raw q rows arrive the same way (boxes at channel 0) and are only masked
// on their way into the tile's A operand.  window_reverse / roll back (:390-401) is the same index map on the output
// store.
//
// A window is 64 queries x 64 keys per head - half of the 128-row UMMA tile.  The tile is filled with TWO HEADS of the
// same window: rows 0-63 carry head 2p, rows 64-127 head 2p+1, through a copy of Q whose rows are masked to their own
// head's channels of the 16-channel "quad" (the same trick that lets the heads of the axial kernel share a K slab):
//   S_tile (128 x 64 keys) = Q_masked (128 x 16) K^T (16 x 64):  one MMA, no wasted column,
//   O_tile (128 x 16)      = P (128 x 64 keys, bf16 in TMEM) [V_group (8 channels) | ones (8)]:  4 MMAs of 16 keys;
// the ones columns make the tensor core accumulate the softmax denominator from the same bf16-rounded P as the numerator.
// head_dim 4: an item is (window, quad) = two tiles (head pairs 0 and 1 of the quad, sharing the window's K|V box);
// head_dim 8: an item is (two neighbouring windows, quad) = two tiles (one head pair each), and rows 64-127 take their
// 8 channels from a second PV accumulator ([V_hi | ones]).
//
// One CTA = 11 warps: two softmax warpgroups (warps 0-3 -> tile 0, warps 4-7 -> tile 1; ONE THREAD PER (query, head)
// ROW: all 64 scores of a row pass through one thread, so the row maximum is exact and needs no exchange - no bound,
// no second pass), one MMA-issuer warp per warpgroup (8, 9) and the TMA warp (10).  TMEM per warpgroup: two buffers of
// 64 columns; item k uses buffer k & 1 for S, then P (bf16 pairs written in place over S columns 0-31), then O (columns
// 32-63, free once S has been read): 256 columns per CTA, two CTAs per SM.  CTAs are persistent: items are claimed
// dynamically by the TMA warp (which runs up to NSTAGE items ahead) and published, already decomposed into
// (b, h0, w0, quad), through a shared-memory ring.  Per warpgroup g (b = item parity):
//   bar_q[g]       (4)  masked Q copy of the next item is in shared memory     softmax -> issuer
//   bar_s[g][b]    (1)  S complete                                            issuer commit -> softmax
//   bar_p[g][b]    (4)  P written over S                                      softmax -> issuer
//   bar_o[g][b]    (1)  O complete                                            issuer commit -> softmax (epilogue)
//   bar_free[g][b] (4)  O read: the buffer may take S of item k + 2           softmax -> issuer
//   bar_full / bar_empty per q|k|v stage (empty: both issuers commit)
// The issuer runs S two items ahead (S(k+1) is computed while the softmax of item k runs), and the epilogue of item k-1
// sits in the MIDDLE of the softmax of item k: PV(k-1) has completed by then, and the buffer it frees receives S(k+1)
// before the softmax of item k ends - no MMA round trip is exposed to a warpgroup.
//
// Measured (B200, C3 stage 1: 67 080 windows x 8 heads per launch, 2.2e9 exponentials): 0.735 ms against 0.935 ms of the
// warp-MMA kernel (stage 2: 0.244 against 0.294).  The kernel is issue-bound (ncu: 70 % issue utilisation at 4-5 warps
// per sub-partition, XU 42 %): what counted was instructions per item - barriers addressed as [base + immediate] in
// dynamic shared memory instead of `uint64_t*` to static __shared__ (each use paid a generic->shared conversion), q rows
// from the TMA stage instead of per-thread global address arithmetic, buffer parity unrolled (705 -> ~450 executed
// instructions per item and warp), and no MMA round trip on a warpgroup's critical path (0.997 -> 0.902 ms).  One
// exponential pair in 3 or 4 on the FMA-pipe polynomial measures the same; MUFU only is 4 % slower.
//
// Only INTERIOR windows - those whose 64 tokens exist and do not wrap around the rolled frame - take this path; the
// bottom / right fringe (two window rows and columns at shift 4: 4.6 % of the stage-1 windows) stays on the warp-MMA
// kernel (attention_win.cu), which folds roll and zero-padding into per-token index arithmetic.
#include "attn_common.cuh"
#include "attn_tc_math.cuh"
#include "sm100.cuh"
#include <stdlib.h>
#include <type_traits>

namespace tfswa {

using namespace sm100;

namespace win_tc {

using namespace tcmath;

constexpr int NTHREADS = 352;          // warps 0-3 / 4-7 softmax warpgroups, 8 / 9 MMA issuers, 10 TMA producer
constexpr int BOX = 1024;              // one TMA box: 64 tokens x 8 channels bf16
constexpr int WIN_BYTES = 6 * BOX;     // K lo, K hi, V lo, V hi, Q lo, Q hi of one window and quad
constexpr int NSTAGE = 4;              // q|k|v stages: items k .. k+2 are in use (issued MMAs, Q copies), one is prefetch
constexpr int NRING = 16;              // published-item ring (the TMA warp is at most NSTAGE + 2 items ahead of the slowest reader)
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t WG_COLS = 128, BUF_COLS = 64, O_COL = 32;   // per warpgroup: two buffers; O sits in the upper half of its buffer
constexpr int Q_BYTES = 4096;          // masked Q copy of one tile: 128 rows x 16 channels
template <int D> __host__ __device__ constexpr int n_win() { return D == 4 ? 1 : 2; }
template <int D> __host__ __device__ constexpr int stage_bytes() { return n_win<D>() * WIN_BYTES; }
template <int D> __host__ __device__ constexpr int ones_off() { return 2 * Q_BYTES + NSTAGE * stage_bytes<D>(); }
template <int D> __host__ __device__ constexpr int smem_bytes();   // 34 / 58 KB (below: needs the barrier table)

#ifndef TFSWA_WINTC_POLY_K
#define TFSWA_WINTC_POLY_K 3
#endif
constexpr int POLY_K = TFSWA_WINTC_POLY_K;   // every POLY_K-th element pair on the FMA-pipe polynomial (0 = MUFU only)

// (a __nanosleep in the retry path of the issuer / producer waits measured no difference: 0.735 ms either way)

// barrier table (8 bytes each), item ring and the TMEM base pointer sit behind the ones tile
constexpr int B_FULL = 0, B_EMPTY = NSTAGE, B_Q = 2 * NSTAGE, B_S = B_Q + 2, B_P = B_S + 4, B_O = B_P + 4, B_FREE = B_O + 4,
              B_ITEM = B_FREE + 4, NBAR = B_ITEM + NRING;
template <int D> __host__ __device__ constexpr int bar_off() { return ones_off<D>() + BOX; }
template <int D> __host__ __device__ constexpr int slot_off() { return (bar_off<D>() + NBAR * 8 + 15) / 16 * 16; }
template <int D> __host__ __device__ constexpr int tptr_off() { return slot_off<D>() + NRING * 16; }

template <int D> __host__ __device__ constexpr int smem_bytes() { return tptr_off<D>() + 16; }

struct Items {
  int n;                 // items [0, n): item = ((b * nWh_int + wh) * nP + wp) * nquads + quad
  int nWh_int, nWw_int;  // interior window rows / columns
  int nP, nquads;        // window groups per row (head_dim 4: windows; 8: pairs of windows), 16-channel quads
  int* next;             // dynamic claims: a CTA's first item is blockIdx.x, the following ones gridDim.x + atomicAdd(next, 1)
};

template <int D>
__global__ void __launch_bounds__(NTHREADS, 2) tc_attn_win_kernel(const __grid_constant__ CUtensorMap tm, const AttnParams p, const Items items) {
  constexpr int NWIN = n_win<D>();
  constexpr int HPQ = 16 / D;            // heads per quad (4 / 2)
  constexpr int STAGE_BYTES = stage_bytes<D>();
  constexpr int Q_OFF = 0, ST_OFF = 2 * Q_BYTES, ONES_OFF = ones_off<D>();
  constexpr int BAR_OFF = bar_off<D>(), SLOT_OFF = slot_off<D>(), TPTR_OFF = tptr_off<D>();
  extern __shared__ __align__(128) uint8_t smem[];
  int4* s_slot = reinterpret_cast<int4*>(smem + SLOT_OFF);   // (b or -1: no more items, first token row, first token column, quad | second window valid << 8)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + TPTR_OFF);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float c = p.qscale;              // head_dim^-0.5 * log2(e)
  const uint32_t sbase = smem_u32(smem);
  auto bar = [&](int idx) { return sbase + BAR_OFF + idx * 8; };

  // ---- set-up, once per CTA ----
  if (warp == 0) {
    if (lane == 0) {
      uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BAR_OFF);
#pragma unroll
      for (int i = 0; i < NSTAGE; ++i) { mbar_init(&bars[B_FULL + i], 1); mbar_init(&bars[B_EMPTY + i], 2); }
#pragma unroll
      for (int i = 0; i < 2; ++i) mbar_init(&bars[B_Q + i], 4);
#pragma unroll
      for (int i = 0; i < 4; ++i) { mbar_init(&bars[B_S + i], 1); mbar_init(&bars[B_P + i], 4); mbar_init(&bars[B_O + i], 1); mbar_init(&bars[B_FREE + i], 4); }
#pragma unroll
      for (int i = 0; i < NRING; ++i) mbar_init(&bars[B_ITEM + i], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(s_tmem, TMEM_COLS);
  }
  if (warp == 10 && elect_one()) prefetch_tmap(&tm);
  if (tid < BOX / 16) reinterpret_cast<uint4*>(smem + ONES_OFF)[tid] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
  fence_async_smem();                    // generic-proxy writes above are read by tcgen05.mma through the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  if (warp == 10) {
    // ---- TMA producer (one lane): claims items, publishes them two ahead of the one it loads, loads K|V boxes up to
    // NSTAGE items ahead ----
    if (elect_one()) {
      auto publish = [&](int k, int it) {
        int4 v = make_int4(-1, 0, 0, 0);
        if (it >= 0) {
          const int quad = it % items.nquads;
          int rest = it / items.nquads;
          const int wp = rest % items.nP; rest /= items.nP;
          const int wh = rest % items.nWh_int;
          const int b = rest / items.nWh_int;
          const int ww = wp * NWIN;
          const int second = (NWIN == 2 && ww + 1 < items.nWw_int) ? 1 : 0;
          v = make_int4(b, wh * 8 + p.shift, ww * 8 + p.shift, quad | (second << 8));
        }
        s_slot[k % NRING] = v;
        arrive_a(bar(B_ITEM + k % NRING));
        return v;
      };
      auto claim = [&]() {
        const int it = (int)gridDim.x + atomicAdd(items.next, 1);
        return it < items.n ? it : -1;
      };
      int4 cur = publish(0, (int)blockIdx.x < items.n ? (int)blockIdx.x : -1);
      int4 nx = publish(1, cur.x >= 0 ? claim() : -1);
      for (int k = 0; cur.x >= 0; ++k) {
        const int4 nn = publish(k + 2, nx.x >= 0 ? claim() : -1);
        const int st = k % NSTAGE;
        if (k >= NSTAGE) wait_a(bar(B_EMPTY + st), ((k / NSTAGE) - 1) & 1);   // both PV MMAs of item k - NSTAGE have completed
        const uint32_t dst = sbase + ST_OFF + st * STAGE_BYTES, full = bar(B_FULL + st);
        expect_tx_a(full, STAGE_BYTES);
        const int quad = cur.w & 0xff;
        const int ck = p.C + quad * 16, cv = 2 * p.C + quad * 16;
#pragma unroll
        for (int t = 0; t < NWIN; ++t) {
          const int w0 = cur.z + ((t == 1 && (cur.w >> 8)) ? 8 : 0);   // an absent second window re-reads the first (rows dropped)
          const uint32_t d = dst + t * WIN_BYTES;
          tma_load_4d_a(d, &tm, full, ck, w0, cur.y, cur.x);
          tma_load_4d_a(d + BOX, &tm, full, ck + 8, w0, cur.y, cur.x);
          tma_load_4d_a(d + 2 * BOX, &tm, full, cv, w0, cur.y, cur.x);
          tma_load_4d_a(d + 3 * BOX, &tm, full, cv + 8, w0, cur.y, cur.x);
          tma_load_4d_a(d + 4 * BOX, &tm, full, quad * 16, w0, cur.y, cur.x);        // raw q rows: masked per head by the softmax threads
          tma_load_4d_a(d + 5 * BOX, &tm, full, quad * 16 + 8, w0, cur.y, cur.x);
        }
        cur = nx; nx = nn;
      }
    }
  } else if (warp >= 8) {
    // ---- MMA issuer of warpgroup g: S runs two items ahead of PV.  Everything is issued by one elected thread, so the
    // tensor core executes it in issue order: ... S(k+1), PV(k-1), S(k+2), PV(k) ... ----
    const int g = warp - 8;
    const uint32_t idesc_s = umma_idesc_bf16(128, 64);
    const uint32_t idesc_pv = idesc_bf16_bmn(128, 16);
    const uint32_t s_col = tmem + g * WG_COLS;
    const uint64_t qdesc = umma_smem_desc_ns(sbase + Q_OFF + g * Q_BYTES, 128, 256);
    const uint32_t ones = sbase + ONES_OFF;
    auto exists = [&](int k) {
      wait_a(bar(B_ITEM + k % NRING), (uint32_t)(k / NRING) & 1u);
      return s_slot[k % NRING].x >= 0;
    };
    auto win_of = [&](int k) {      // my tile's window of item k: K lo | K hi | V lo | V hi
      return sbase + ST_OFF + (k % NSTAGE) * STAGE_BYTES + (NWIN == 2 ? g * WIN_BYTES : 0);
    };
    auto issue_S = [&](int k) {     // all lanes
      wait_a(bar(B_FULL + k % NSTAGE), (uint32_t)(k / NSTAGE) & 1u);
      wait_a(bar(B_Q + g), (uint32_t)k & 1u);
      if (k >= 2) wait_a(bar(B_FREE + 2 * g + (k & 1)), (uint32_t)((k >> 1) - 1) & 1u);   // O(k-2) has been read
      if (elect_one()) {
        tc_fence_after();
        umma_bf16_ss(s_col + (k & 1) * BUF_COLS, qdesc, umma_smem_desc_ns(win_of(k), BOX, 128), idesc_s, 0u);
        commit_a(bar(B_S + 2 * g + (k & 1)));
      }
      __syncwarp();
    };
    if (exists(0)) {
      issue_S(0);
      if (exists(1)) issue_S(1);
      for (int k = 0; exists(k); ++k) {
        wait_a(bar(B_P + 2 * g + (k & 1)), (uint32_t)(k >> 1) & 1u);
        if (elect_one()) {
          tc_fence_after();
          const uint32_t p_col = s_col + (k & 1) * BUF_COLS, o_col = p_col + O_COL;
          const uint32_t win = win_of(k);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {         // 16 keys per MMA = 8 P columns
            if (D == 4) {                          // head pair g lives in 8-channel group g of the quad
              const uint32_t va = win + 2 * BOX + g * BOX + kk * 256;
              umma_bf16_ts(o_col, p_col + 8 * kk, umma_smem_desc_ns(va, 128, ones + kk * 256 - va), idesc_pv, kk ? 1u : 0u);
            } else {                               // rows 0-63 (head 0) read [V lo | 1], rows 64-127 (head 1) [V hi | 1]
              const uint32_t va = win + 2 * BOX + kk * 256;
              umma_bf16_ts(o_col, p_col + 8 * kk, umma_smem_desc_ns(va, 128, ones + kk * 256 - va), idesc_pv, kk ? 1u : 0u);
              umma_bf16_ts(o_col + 16, p_col + 8 * kk, umma_smem_desc_ns(va + BOX, 128, ones + kk * 256 - (va + BOX)), idesc_pv, kk ? 1u : 0u);
            }
          }
          commit_a(bar(B_O + 2 * g + (k & 1)));
          commit_a(bar(B_EMPTY + k % NSTAGE));
        }
        __syncwarp();
        if (exists(k + 2)) issue_S(k + 2);
      }
    }
  } else {
    // ---- softmax warpgroup g: thread = one (query, head) row of the tile ----
    const int g = warp >> 2;
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;               // tile row == TMEM lane
    const int slot = quarter >> 1;                   // rows 0-63: first head of the pair, 64-127: second (warp-uniform)
    const int n = r & 63;                            // token of the window: row n / 8, column n % 8
    const int t_win = NWIN == 2 ? g : 0;             // which window of the item my tile attends
    const int head = D == 4 ? 2 * g + slot : slot;   // head within the quad
    const int w_first = head * (D / 2);              // first of the D / 2 32-bit words (2 channels each) of my head in the quad's 8 words
    const uint32_t t_base = tmem + ((uint32_t)(quarter * 32) << 16) + g * WG_COLS;
    uint8_t* q_dst = smem + Q_OFF + g * Q_BYTES + (r >> 3) * 256 + (r & 7) * 16;

    const uint32_t q_src = sbase + ST_OFF + t_win * WIN_BYTES + 4 * BOX + n * 16;   // my token's raw q row inside a stage (lo; hi = + BOX)
    const uint32_t q_dst_a = smem_u32(q_dst);

    // item k: does it exist (slot published by the TMA warp)?
    auto exists = [&](int k) {
      wait_a(bar(B_ITEM + k % NRING), (uint32_t)(k / NRING) & 1u);
      return s_slot[k % NRING].x >= 0;
    };
    // masked copy of my q row of item k from its stage (TMA-loaded raw rows) into the tile's A operand
    auto write_qm = [&](int k) {
      wait_a(bar(B_FULL + k % NSTAGE), (uint32_t)(k / NSTAGE) & 1u);
      const uint32_t src = q_src + (k % NSTAGE) * STAGE_BYTES;
      uint32_t w[8];
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(src));
      asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(src + BOX));
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = (i >= w_first && i < w_first + D / 2) ? w[i] : 0u;
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(q_dst_a), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(q_dst_a + 128), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
      fence_async_smem();
      __syncwarp();
      if (lane == 0) arrive_a(bar(B_Q + g));
    };
    // O / l of item j (buffer PAR = j & 1) -> out rows of my token; frees the buffer for S(j + 2)
    auto epilogue = [&](auto PAR, int j, float m) {
      constexpr int B = decltype(PAR)::value;
      wait_a(bar(B_O + 2 * g + B), (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      uint32_t o[16];
      tmem_ld_x16(t_base + B * BUF_COLS + O_COL + ((D == 8 && slot) ? 16 : 0), o);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_a(bar(B_FREE + 2 * g + B));        // the buffer may take S(j + 2)
      const int4 v = s_slot[j % NRING];                        // (still there: the ring is NRING - NSTAGE - 2 items deeper than any reader lags)
      const bool second = (v.w >> 8) != 0;
      if (t_win == 0 || second) {
        const int quad = v.w & 0xff;
        const int hp = v.y + (n >> 3), wp = v.z + (t_win == 1 ? 8 : 0) + (n & 7);
        const int64_t tok = (int64_t)((v.x * p.H + hp) * p.W + wp);
        const float l = __uint_as_float(o[8]);
        const float inv = rcp_approx(l);
        bf16* op = (bf16*)p.out + tok * p.ldo + quad * 16 + head * D;
        if (D == 4) {
          float r4[4];
#pragma unroll
          for (int d = 0; d < 4; ++d) r4[d] = __uint_as_float(slot ? o[4 + d] : o[d]) * inv;
          store4(op, r4);
        } else {
          float r8[8];
#pragma unroll
          for (int d = 0; d < 8; ++d) r8[d] = __uint_as_float(o[d]) * inv;
          store8(op, r8);
        }
        if (p.lse) p.lse[tok * p.heads + quad * HPQ + head] = m * c + log2f(l);
      }
    };
    // one item: buffer PAR = k & 1 (compile time: TMEM and barrier addresses are immediates)
    float prev_m = 0.f;
    auto item = [&](auto PAR, int k, bool more) {
      constexpr int B = decltype(PAR)::value;
      wait_a(bar(B_S + 2 * g + B), (uint32_t)(k >> 1) & 1u);   // S(k) complete: the Q copy it read may be replaced
      tc_fence_after();
      if (more) write_qm(k + 1);

      // ---- exact row maximum over the 64 keys, P = ex2(S c - m c) written over S ----
      const uint32_t buf = t_base + B * BUF_COLS;
      uint32_t sc[32], pka[16], pkb[16];
      float m = -CUDART_INF_F;
      tmem_ld_x32(buf, sc);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(sc[i]));
      tmem_ld_x32(buf + 32, sc);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) m = fmaxf(m, __uint_as_float(sc[i]));
      const float mc = m * c;
      softmax_half<true, 0, POLY_K>(sc, pkb, c, mc);
      softmax_half<true, 1, POLY_K>(sc, pkb, c, mc);
      // previous item's epilogue: its PV MMAs have completed meanwhile, and the buffer it frees receives S(k+1) before this
      // item's softmax ends
      if (k > 0) epilogue(std::integral_constant<int, 1 - B>{}, k - 1, prev_m);
      tmem_ld_x32(buf, sc);                          // keys 0-31 again (cheaper than keeping 64 scores live at 80 registers)
      tmem_ld_wait();
      softmax_half<true, 0, POLY_K>(sc, pka, c, mc);
      softmax_half<true, 1, POLY_K>(sc, pka, c, mc);
      tmem_st_x16(buf, pka);                         // key pair (2i, 2i+1) -> 32-bit cell i
      tmem_st_x16(buf + 16, pkb);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) arrive_a(bar(B_P + 2 * g + B));
      prev_m = m;
    };

    if (exists(0)) {
      write_qm(0);
      for (int k = 0;; k += 2) {
        const bool more1 = exists(k + 1);
        item(std::integral_constant<int, 0>{}, k, more1);
        if (!more1) { epilogue(std::integral_constant<int, 0>{}, k, prev_m); break; }
        const bool more2 = exists(k + 2);
        item(std::integral_constant<int, 1>{}, k + 1, more2);
        if (!more2) { epilogue(std::integral_constant<int, 1>{}, k + 1, prev_m); break; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace win_tc

// 4-D view of the q|k|v token matrix: (channel, w, h, b); box = 8 channels x 8 x 8 tokens = one window of a channel group
static int make_tmap_win(CUtensorMap* out, const AttnParams& p) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return TFSWA_ECUDA; }
  cuuint64_t dims[4] = {(cuuint64_t)(3 * p.C), (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
  cuuint64_t strides[3] = {(cuuint64_t)p.ldq * 2, (cuuint64_t)p.W * p.ldq * 2, (cuuint64_t)p.H * p.W * p.ldq * 2};
  cuuint32_t box[4] = {8, 8, 8, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.qkv), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("attn_win_tc: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return TFSWA_ECUDA; }
  return TFSWA_OK;
}

// claim counters of the persistent launches: a small per-device pool of zero-initialised slots used round-robin (each
// launch clears its slot on its stream first), so launches in flight on different streams do not share a counter
static int* claim_counter(cudaStream_t st) {
  constexpr int NSLOT = 256, MAXDEV = 64;
  static int* pool[MAXDEV] = {};
  static unsigned turn[MAXDEV] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAXDEV) return nullptr;
  int* base = __atomic_load_n(&pool[dev], __ATOMIC_ACQUIRE);
  if (!base) {
    int* fresh = nullptr;
    if (cudaMalloc(&fresh, NSLOT * sizeof(int)) != cudaSuccess) return nullptr;
    int* expected = nullptr;
    if (__atomic_compare_exchange_n(&pool[dev], &expected, fresh, false, __ATOMIC_ACQ_REL, __ATOMIC_ACQUIRE)) base = fresh;
    else { cudaFree(fresh); base = expected; }
  }
  int* slot = base + (__atomic_fetch_add(&turn[dev], 1u, __ATOMIC_RELAXED) % NSLOT);
  if (cudaMemsetAsync(slot, 0, sizeof(int), st) != cudaSuccess) return nullptr;
  return slot;
}

static long long g_win_tc_launches = 0;   // diagnostics: launches of the tcgen05 window kernel so far (tests check the dispatch)

// interior windows (rows [0, nWh_int) x columns [0, nWw_int) of the window grid) of an SW-MSA call; the caller runs the
// remaining fringe on attention_win.cu.  Returns 1 when the shape is not covered (nothing launched).
int attn_win_tc_bf16(const AttnParams& p, int nWh_int, int nWw_int, cudaStream_t st) {
  using namespace win_tc;
  const int D = p.C / p.heads;
  if ((D != 4 && D != 8) || p.C % 16 != 0 || nWh_int <= 0 || nWw_int <= 0) return 1;
  if ((int64_t)p.B * p.H * p.W > 0x7fffffffll) return 1;
  if (((uintptr_t)p.qkv & 15) || (p.ldq % 8) || ((uintptr_t)p.out & 15) || (p.ldo % 8)) return 1;
  CUtensorMap tm;
  int rc = make_tmap_win(&tm, p);
  if (rc) return rc;
  static int sms = 0;
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    cudaError_t e1 = cudaFuncSetAttribute(tc_attn_win_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<4>());
    cudaError_t e2 = cudaFuncSetAttribute(tc_attn_win_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<8>());
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, attr_once.dev);
    if (e1 != cudaSuccess || e2 != cudaSuccess || sms <= 0) { set_error("attn_win_tc: cudaFuncSetAttribute failed"); return TFSWA_ECUDA; }
    attr_once.done();
  }
  Items items = {};
  items.nWh_int = nWh_int; items.nWw_int = nWw_int;
  items.nP = D == 4 ? nWw_int : (nWw_int + 1) / 2;
  items.nquads = p.C / 16;
  const int64_t n = (int64_t)p.B * nWh_int * items.nP * items.nquads;
  if (n > 0x7fffffff) return 1;
  items.n = (int)n;
  items.next = claim_counter(st);
  if (!items.next) { set_error("attn_win_tc: no claim counter (cudaMalloc / cudaMemsetAsync failed)"); return TFSWA_ECUDA; }
  const int grid = items.n < 2 * sms ? items.n : 2 * sms;   // two resident CTAs per SM
  if (D == 4) tc_attn_win_kernel<4><<<grid, NTHREADS, smem_bytes<4>(), st>>>(tm, p, items);
  else tc_attn_win_kernel<8><<<grid, NTHREADS, smem_bytes<8>(), st>>>(tm, p, items);
  __atomic_fetch_add(&g_win_tc_launches, 1ll, __ATOMIC_RELAXED);
  return check_launch("attn_win_tc(tcgen05)");
}

}  // namespace tfswa

extern "C" long long tfswa_attn_win_tc_interior_launches(void) { return __atomic_load_n(&tfswa::g_win_tc_launches, __ATOMIC_RELAXED); }
