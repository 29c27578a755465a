// Axial (TSA / FSA) attention on tcgen05 with TMA-fed operands (round 2; head_dim 4, 8, 16, bf16).
//
// The token regrouping of attention.py:143 / :217 (permute(0,3,2,1) / permute(0,2,3,1)) is a TMA tensor map here: the
// q|k|v token matrix (M, ldq) is described once per launch as a 4-D tensor (channel, w, h, b); a key tile of a TSA
// sequence is the box (8 channels, 1, 128 keys, 1), of an FSA sequence (8, 128, 1, 1).  A box lands in shared memory
// as [key][16 B] - which is at the same time
//   * the canonical no-swizzle K-MAJOR operand of S = Q K^T (keys = N, 8 keys x 16 B per core matrix; the second
//     8-channel chunk of the quad is a second box, LBO apart), and
//   * the canonical no-swizzle MN-MAJOR operand of O = P V (channels = N, keys = K).
// So K and V go from HBM / L2 to the tensor core without a single thread touching them; rows beyond the sequence end
// are zero-filled by the TMA unit.  What round 1 did by staging an "expanded K" and per-head V' with a producer warp
// (a third of one sub-partition's issue slots) is moved to the operands that are constant over the key loop:
//   * heads share the 16-channel K slab through HPQ = 16/d MASKED COPIES OF Q (head h's copy keeps channels
//     [d h, d h + d) and zeros the rest): S_h = Q_h K^T, one MMA (M=128, N=KT, K=16) per head writes its 128/HPQ keys
//     into its own TMEM columns - same MMA work as the expanded-K form, operand prepared once per CTA;
//   * the softmax denominator comes out of the PV MMA: its B operand (N = 16) is [8 v channels | 8 x ones], the second
//     N-group being a constant all-ones tile that the descriptor's stride field (SBO) points at.  D_h = [P_h V | l_h..]:
//     the row sum is accumulated by the tensor core from the same bf16-rounded P as the numerator; the last key block
//     uses a second tile with zeros at the absent keys.
// Softmax warps (0-7; a query row is shared by two threads, 32 of a tile's 64 columns each): S (TMEM) -> registers -> P
// (bf16, written back IN PLACE over the thread's own S columns, where the PV MMA reads it as its A operand) with
//   * packed fp32 arithmetic (fma/add.f32x2 -> FFMA2/FADD2: one issue slot per two elements),
//   * MUFU.EX2 for 10 of 16 element pairs and a degree-3 Cody-Waite polynomial on the FMA pipe for the other 6:
//     tools/mufu_bench.cu measures 16.0 ex2/clk/SM for MUFU alone (4.64e12/s) and 21.8/clk/SM (6.3e12/s) for this mix
//     at 4 warps per sub-partition,
//   * no running maximum (row bound from the per-channel extrema of k, see tc_attention.cu); the polynomial's exponent
//     clamp is compiled out for warps whose rows cannot reach 2^-120 (bound minus lower bound, checked once per CTA).
// TMEM (256 columns, two CTAs per SM): three S/P buffers of 64 columns (tile = HPQ heads x 64/HPQ keys) + O (64).
// Warp 8 issues every MMA (one elected lane: `elect.sync`, so ptxas moves descriptors to uniform registers without
// waterfall loops), warp 9 every TMA; the key loop has no CTA-wide barrier and is unrolled three-fold so that buffer
// and barrier addresses are immediates:
//   bar_full[s]  (tx)  stage s (128 keys: K lo|hi, V lo|hi = 8 KB) landed            TMA           -> issuer
//   bar_empty[s] (1)   every MMA reading stage s has completed                       issuer commit -> producer
//   bar_s[b]     (1)   S(t) complete in buffer b = t % 3                             issuer commit -> softmax
//   bar_p[b]     (8)   P(t) written over S(t) by all softmax warps                   softmax       -> issuer
//   bar_done     (1)   the last PV MMA has completed                                 issuer commit -> softmax (epilogue)
// On bar_p[b] the issuer issues PV(t) and, back to back, S(t+3) into the same buffer: the tensor core executes one thread's
// MMAs in issue order, so S(t+3) cannot overtake the MMA that reads the columns it overwrites, and a softmax warp can run
// up to two tiles ahead of the slowest one.  What was measured and dropped on the way (tools/debug/tma_trace.py timelines,
// ncu source pages, profiles/r2_attn_history.md): 128-column tiles with a separate P buffer and three barrier pairs
// (10.4 ms per stage-1 TSA launch at B=8: the MMA issuer's ~270-instruction loop under `if (lane == 0)` was the critical
// path), separate S- and PV-issuer warps (9.7), this design (9.4; round 1: 10.6); software-pipelined tcgen05.ld of the
// next tile (11.0: one tile less slack between warps), 16 softmax warps x 16 columns at 56 registers (11.3: twice the
// per-tile bookkeeping per exponential), MUFU only (10.4), polynomial share 1/2 (9.9) and 1/4 (9.5).
// Replaces attention.py:70-85 (+ permutes :143,:162,:217,:236), bf16 activations.
#include "attn_common.cuh"
#include "attn_tc_math.cuh"
#include "sm100.cuh"
#include <stdlib.h>

namespace tfswa {

using namespace sm100;

namespace tma_attn {

#ifdef TFSWA_TMA_TRACE
__device__ long long g_trace[8][128];   // debug: clock64 per (event, tile) of one CTA
#ifndef TFSWA_TMA_TRACE_BLOCK
#define TFSWA_TMA_TRACE_BLOCK 0
#endif
#define TRACE(ev, t) do { if (blockIdx.x == TFSWA_TMA_TRACE_BLOCK && (t) < 128) g_trace[ev][t] = clock64(); } while (0)
#else
#define TRACE(ev, t) do { } while (0)
#endif

constexpr int NTHREADS = 320;          // warps 0-7 softmax, 8 MMA issuer, 9 TMA producer
constexpr int QTILE = 128;             // queries per CTA
constexpr int SKEYS = 128;             // keys per shared-memory stage
constexpr int NSTAGE = 4;             // TMA runs two stages ahead of the stage in use; a fourth covers the one being drained
constexpr uint32_t TMEM_COLS = 256;
constexpr int BOX_BYTES = SKEYS * 16;  // one TMA box: 128 keys x 8 channels bf16
constexpr int STAGE_BYTES = 4 * BOX_BYTES;            // K lo, K hi, V lo, V hi
template <int D> __host__ __device__ constexpr int q_bytes() { return (16 / D) * 4096; }
template <int D> __host__ __device__ constexpr int stage_off() { return q_bytes<D>(); }
template <int D> __host__ __device__ constexpr int ones_off() { return stage_off<D>() + NSTAGE * STAGE_BYTES; }
template <int D> __host__ __device__ constexpr int tail_off() { return ones_off<D>() + 2 * BOX_BYTES; }
template <int D> __host__ __device__ constexpr int smem_bytes() { return tail_off<D>() + 2 * BOX_BYTES; }   // 48 / 40 / 36 KB

// which of the 16 element PAIRS of a thread's 32 scores take the FMA-pipe polynomial instead of MUFU.EX2 (bit i = pair i).
// Measured per stage-1 TSA launch (B=8) after the v14 barrier forms: 4 of 16: 8.55 ms, 5 of 16 (every third pair, the share
// tools/mufu_bench.cu found best in isolation): 8.21, 6 of 16: 8.11, 7 of 16: 8.33, 8 of 16: 8.68 - 6 of 16 it is.
#ifndef TFSWA_TMA_POLY_MASK
#define TFSWA_TMA_POLY_MASK 0x5252u    // pairs 1, 4, 6 | 9, 12, 14; -D...=0 for MUFU only, 0x4924u = every third pair
#endif

using namespace tcmath;                  // TMA box load, MN-major descriptor, packed fp32 helpers, FMA-pipe exponential

// 16 scores sc[16*HALF ..] -> 8 packed bf16x2 probabilities pk[8*HALF ..]: p = 2^(s*c - mc)
template <bool CLAMP, int HALF>
__device__ __forceinline__ void softmax_half(const uint32_t (&sc)[32], uint32_t (&pk)[16], float c, float mc) {
  tcmath::softmax_half_mask<CLAMP, HALF, TFSWA_TMA_POLY_MASK>(sc, pk, c, mc);
}

// TMEM: three S/P buffers of 64 columns (S tile = HPQ heads x KT keys fp32; P overwrites the thread's own S columns as
// bf16 pairs) + O.  A softmax warp may run up to two tiles ahead of the slowest one.
constexpr uint32_t NBUF = 3, BUF_COLS = 64;
constexpr int NRING = 8;                              // claimed-item ring (the TMA warp is at most NSTAGE items ahead of the slowest role)
constexpr uint32_t O_COL2 = NBUF * BUF_COLS;          // 192: O accumulators, HPQ x 16 columns (d = 16: 32)

// PERSISTENT CTAs.  A work item is (sequence, 128-query tile, 16-channel quad).  A launch-per-item version of this kernel
// measured 10.9 k cycles of fixed cost per CTA next to 1.1 k cycles per key tile (TSA 65 tiles, FSA 33: 13 % / 23 % of the
// time; stage 3 has 4-5 tiles per item): TMEM allocation, barrier init, the constant operand tiles, the first TMA and MMA
// latencies.  Here 2 x #SM CTAs loop over the items (item = blockIdx.x + k * gridDim.x, so CTAs that run together share a
// sequence's K|V in L2), the set-up is done once, the TMA warp runs up to four stages ahead ACROSS item boundaries, and
// the roles only meet on mbarriers:
//   bar_q (8)  the item's masked Q copies are in shared memory                      softmax -> issuer
// Rows whose bound was too loose (denominator underflow) are not repeated in place - that would need a CTA-wide decision
// per item - but appended to a device list that a second, normally empty launch of the same kernel processes with exact
// row maxima (`exact`: pass 0 streams S and reduces the maxima, pass 1 is the normal one).
struct Items {
  int n;                 // implicit item range [0, n) ...
  const int* list;       // ... or (list != nullptr) the *count entries of a device list
  const int* count;
  int nqt, nquads;       // item -> (row, query tile, quad): item = (row * nqt + qt) * nquads + quad
  int* redo_count;       // non-exact launches: items with an underflowed row are appended here (flag = first writer wins)
  int* redo_flag;
  int* redo_list;
  int exact;
  int* next;              // dynamic item claims: a CTA's first item is blockIdx.x, the following ones gridDim.x + atomicAdd(next, 1)
};

template <int D>
__global__ void __launch_bounds__(NTHREADS, 2) tc_attn_tma_kernel(const __grid_constant__ CUtensorMap tm, const AttnParams p, const Items items) {
  constexpr int HPQ = 16 / D;            // heads per CTA (one 16-channel "quad")
  constexpr int KT = 64 / HPQ;           // keys per S tile and head: HPQ heads x KT keys = 64 TMEM columns (16 / 32 / 64)
  constexpr int TPS = SKEYS / KT;        // S tiles per shared-memory stage (8 / 4 / 2)
  constexpr int HPT = D == 4 ? 2 : 1;    // head slots per softmax thread (32 columns = 2 heads x 16 keys at d = 4)
  constexpr int NCH = HPT * D;           // channels of my head slots (8 / 8 / 16), starting at channel c0 of the quad
  constexpr int Q_OFF = 0, ST_OFF = stage_off<D>(), ONES_OFF = ones_off<D>(), TAIL_OFF = tail_off<D>();
  extern __shared__ __align__(128) uint8_t smem[];
  // one table: every barrier is [table base register + immediate] for the 32-bit-address forms of attn_tc_math.cuh
  constexpr int B_FULL = 0, B_EMPTY = NSTAGE, B_S = 2 * NSTAGE, B_P = B_S + (int)NBUF, B_DONE = B_P + (int)NBUF, B_Q = B_DONE + 1,
                B_ITEM = B_Q + 1, NBAR = B_ITEM + NRING;
  __shared__ __align__(8) uint64_t bars[NBAR];
  __shared__ uint32_t s_tmem;
  __shared__ int s_ring[NRING];                       // claimed items: id (-1: none left) ...
  __shared__ int4 s_slot[NRING];                      // ... and where it is (row, b, column / row inside b, q0 | quad)
  __shared__ float s_xch[HPQ == 1 ? 256 : 1];   // d = 16, exact pass: the two threads of a row exchange their maxima

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool issuer = warp == 8, producer = warp == 9;       // warps 0-7: softmax
  const bool tsa = p.geom == TFSWA_GEOM_TSA;
  const int N = tsa ? p.H : p.W;
  const int quarter = warp & 3, half = (warp >> 2) & 1;
  const int r = quarter * 32 + lane;                 // my query row == my TMEM lane
  const float c = p.qscale;                          // head_dim^-0.5 * log2(e)
  const int T = (N + KT - 1) / KT;                   // S tiles
  const int NST = (N + SKEYS - 1) / SKEYS;           // stages (TMA loads) per pass
  const int tail_keys = N - (NST - 1) * SKEYS;       // valid keys in the last stage (1..128)
  const bool exact = items.exact != 0;
  const int n_items = items.list ? *items.count : items.n;
  if ((int)blockIdx.x >= n_items) return;             // (the exact launch over the redo list is normally empty: no set-up at all)

  // ---- set-up, once per CTA ----
  if (warp == 0) {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < NSTAGE; ++i) { mbar_init(&bars[B_FULL + i], 1); mbar_init(&bars[B_EMPTY + i], 1); }
#pragma unroll
      for (int i = 0; i < (int)NBUF; ++i) { mbar_init(&bars[B_S + i], 1); mbar_init(&bars[B_P + i], 8); }
      mbar_init(&bars[B_DONE], 1); mbar_init(&bars[B_Q], 8);
#pragma unroll
      for (int i = 0; i < NRING; ++i) mbar_init(&bars[B_ITEM + i], 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(&s_tmem, TMEM_COLS);
  }
  if (producer && elect_one()) prefetch_tmap(&tm);
  // constant PV operands: all-ones tile (2 N-groups) and the last stage's tile with zeros at the absent keys
  for (int i = tid; i < 2 * BOX_BYTES / 16; i += NTHREADS) {
    reinterpret_cast<uint4*>(smem + ONES_OFF)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    const uint32_t v = ((i & (SKEYS - 1)) < tail_keys) ? 0x3F803F80u : 0u;
    reinterpret_cast<uint4*>(smem + TAIL_OFF)[i] = make_uint4(v, v, v, v);
  }
  fence_async_smem();                    // generic-proxy writes above are read by tcgen05.mma through the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t sbase = smem_u32(smem);
  const uint32_t tmem = s_tmem;
  // the two barriers of the key loop as 32-bit shared addresses kept in registers: wait / arrive are [register + immediate]
  // (through `uint64_t*` the lane-0 arrive recomputed the generic->shared conversion - S2R SR_CgaCtaId, MOV, LEA - per tile)
  const uint32_t bar0 = smem_u32(bars);
  auto bar = [&](int idx) { return bar0 + 8u * (uint32_t)idx; };

  // item -> sequence `row`, first query q0, quad; TSA row = b * W + w (keys along h), FSA row = b * H + h (keys along w)
  struct Where { int row, q0, quad, cb, cf, item; };
  auto locate = [&](int it) {             // three integer divisions: done once per item by the TMA warp, published in s_slot
    const int item = items.list ? items.list[it] : it;
    const int per_row = items.nqt * items.nquads;
    Where w;
    w.item = item;
    w.row = item / per_row;
    const int rem = item - w.row * per_row;
    const int qt = rem / items.nquads;
    w.quad = rem - qt * items.nquads;
    w.q0 = qt * QTILE;
    w.cb = tsa ? w.row / p.W : w.row / p.H;
    w.cf = tsa ? w.row - w.cb * p.W : w.row - w.cb * p.H;
    return w;
  };

  uint32_t n_stage = 0;                  // stages issued / consumed so far (monotonic over passes and items)
  // tile t of a pass uses TMEM buffer t % 3; every role keeps the parity of the next phase it will wait for per buffer
  // (softmax: bar_s, issuer: bar_p) in three scalars, so the 3-way unrolled loops address buffers with immediates
  uint32_t ph0 = 0, ph1 = 0, ph2 = 0;
  auto wait_buf = [&](int first, int b) {          // runtime b (rare paths)
    const uint32_t ph = b == 0 ? ph0 : (b == 1 ? ph1 : ph2);
    wait_a(bar(first + b), ph);
    if (b == 0) ph0 ^= 1; else if (b == 1) ph1 ^= 1; else ph2 ^= 1;
  };

  // k-th item of this CTA, published by the TMA warp (-1: no more)
  auto item_at = [&](int k) {
    wait_fast(bar(B_ITEM + k % NRING), (uint32_t)(k / NRING) & 1u);
    return s_ring[k % NRING];
  };
  auto where_at = [&](int k, int item) {  // after item_at(k) returned `item` >= 0
    const int4 v = s_slot[k % NRING];
    Where w;
    w.item = item; w.row = v.x; w.cb = v.y; w.cf = v.z; w.q0 = v.w & ~(QTILE - 1); w.quad = v.w & (QTILE - 1);
    return w;
  };

  if (producer) {
    // ---- warp 9, one lane: TMA loads, up to NSTAGE stages ahead (also across items); a stage is reused once every MMA that
    // read it has completed ----
    if (elect_one()) {
      Where w, wn;
      auto publish = [&](int k, int it, Where& ww) {
        if (it >= 0) {
          ww = locate(it);
          s_slot[k % NRING] = make_int4(ww.row, ww.cb, ww.cf, ww.q0 | ww.quad);
        }
        s_ring[k % NRING] = it >= 0 ? ww.item : -1;
        arrive_a(bar(B_ITEM + k % NRING));
      };
      int it = (int)blockIdx.x < n_items ? (int)blockIdx.x : -1;
      publish(0, it, w);
      for (int k = 0; it >= 0; ++k) {
        // claim the next item before loading this one: the softmax warps prefetch its q row during this item's key loop
        int nxt = (int)gridDim.x + atomicAdd(items.next, 1);
        if (nxt >= n_items) nxt = -1;
        publish(k + 1, nxt, wn);
        const int ck = p.C + w.quad * 16, cv = 2 * p.C + w.quad * 16;
        for (int pass = exact ? 0 : 1; pass < 2; ++pass) {
          for (int i = 0; i < NST; ++i, ++n_stage) {
            const uint32_t g = n_stage, st = g % NSTAGE;
            if (g >= NSTAGE) wait_a(bar(B_EMPTY + st), ((g / NSTAGE) - 1) & 1);   // stage g - NSTAGE released
            const uint32_t dst = sbase + ST_OFF + st * STAGE_BYTES, full = bar(B_FULL + st);
            expect_tx_a(full, STAGE_BYTES);
            const int k0 = i * SKEYS;
            if (tsa) {
              tma_load_4d_a(dst, &tm, full, ck, w.cf, k0, w.cb);
              tma_load_4d_a(dst + BOX_BYTES, &tm, full, ck + 8, w.cf, k0, w.cb);
              tma_load_4d_a(dst + 2 * BOX_BYTES, &tm, full, cv, w.cf, k0, w.cb);
              tma_load_4d_a(dst + 3 * BOX_BYTES, &tm, full, cv + 8, w.cf, k0, w.cb);
            } else {
              tma_load_4d_a(dst, &tm, full, ck, k0, w.cf, w.cb);
              tma_load_4d_a(dst + BOX_BYTES, &tm, full, ck + 8, k0, w.cf, w.cb);
              tma_load_4d_a(dst + 2 * BOX_BYTES, &tm, full, cv, k0, w.cf, w.cb);
              tma_load_4d_a(dst + 3 * BOX_BYTES, &tm, full, cv + 8, k0, w.cf, w.cb);
            }
          }
        }
        it = nxt; w = wn;
      }
    }
  } else if (issuer) {
    // ---- warp 8: every MMA.  P(t) written (bar_p) -> O += P(t) [V | 1], then S(t+3) into the same TMEM buffer: the two are
    // issued back to back by one thread and the tensor core executes a thread's MMAs in issue order, so S(t+3) cannot
    // overtake the PV MMA that reads the columns it overwrites ----
    const uint32_t idesc_s = umma_idesc_bf16(128, KT);
    const uint32_t idesc_pv = idesc_bf16_bmn(128, 16);
    uint32_t n_q = 0;
    auto issue_S = [&](int u, uint32_t b) {        // one elected lane; S(u) -> buffer b
      const uint32_t st = (n_stage + u / TPS) % NSTAGE;
      const uint32_t kaddr = sbase + ST_OFF + st * STAGE_BYTES + (u % TPS) * KT * 16;
      const uint64_t kdesc = umma_smem_desc_ns(kaddr, BOX_BYTES, 128);
#pragma unroll
      for (int h = 0; h < HPQ; ++h)
        umma_bf16_ss(tmem + b * BUF_COLS + h * KT, umma_smem_desc_ns(sbase + Q_OFF + h * 4096, 128, 256), kdesc, idesc_s, 0u);
      commit_a(bar(B_S + b));
      TRACE(2, u);
    };
    auto wait_stage = [&](int u) {                 // all lanes: the stage holding tile u has landed
      const uint32_t g = n_stage + u / TPS, st = g % NSTAGE;
      wait_fast(bar(B_FULL + st), (g / NSTAGE) & 1);
    };
    for (int k = 0; item_at(k) >= 0; ++k) {
      wait_fast(bar(B_Q), n_q & 1); ++n_q;           // this item's masked Q copies are in shared memory
      for (int pass = exact ? 0 : 1; pass < 2; ++pass) {
        const bool maxpass = pass == 0;
        for (int u = 0; u < (int)NBUF && u < T; ++u) {
          if (u % TPS == 0) wait_stage(u);
          if (elect_one()) { tc_fence_after(); issue_S(u, u); }
          __syncwarp();
        }
        for (int t0 = 0; t0 < T; t0 += (int)NBUF) {
#pragma unroll
          for (int b = 0; b < (int)NBUF; ++b) {    // compile-time buffer index
            const int t = t0 + b;
            if (t >= T) break;
            const uint32_t st = (n_stage + t / TPS) % NSTAGE;
            const int u = t + NBUF;
            if (u < T && u % TPS == 0) wait_stage(u);
            uint32_t& ph = b == 0 ? ph0 : (b == 1 ? ph1 : ph2);
            wait_fast(bar(B_P + b), ph);            // P(t) written over S(t) (max pass: S(t) consumed)
            ph ^= 1;
            if (elect_one()) {
              tc_fence_after();
              TRACE(3, t);
              if (!maxpass) {
                const uint32_t vaddr = sbase + ST_OFF + st * STAGE_BYTES + 2 * BOX_BYTES + (t % TPS) * KT * 16;
                const bool last_stage = t / TPS == NST - 1;
                const uint32_t oaddr = sbase + (last_stage ? TAIL_OFF : ONES_OFF) + (t % TPS) * KT * 16;
#pragma unroll
                for (int h = 0; h < HPQ; ++h) {
#pragma unroll
                  for (int kk = 0; kk < KT / 16; ++kk) {   // 16 keys per MMA = 8 P columns inside the S columns they came from
                    const uint32_t acc = (t | kk) ? 1u : 0u;
                    const int scol = h * KT + kk * 16;     // first S column of these 16 keys
                    const uint32_t pcol = tmem + b * BUF_COLS + (scol / 32) * 32 + (scol % 32) / 2;
                    if (D == 16) {
                      // [v channels 0-7 | 8-15] and, as a second accumulator, [ones | ones]
                      umma_bf16_ts(tmem + O_COL2, pcol, umma_smem_desc_ns(vaddr + kk * 256, 128, BOX_BYTES), idesc_pv, acc);
                      umma_bf16_ts(tmem + O_COL2 + 16, pcol, umma_smem_desc_ns(oaddr + kk * 256, 128, BOX_BYTES), idesc_pv, acc);
                    } else {
                      const uint32_t va = vaddr + ((h * D) / 8) * BOX_BYTES + kk * 256;   // the 8-channel group holding head h
                      umma_bf16_ts(tmem + O_COL2 + 16 * h, pcol, umma_smem_desc_ns(va, 128, oaddr + kk * 256 - va), idesc_pv, acc);
                    }
                  }
                }
              }
              if (u < T) issue_S(u, b);
              if (t % TPS == TPS - 1 || t == T - 1) commit_a(bar(B_EMPTY + st));   // K rows (S) and V rows (PV) of the stage are done with
              if (!maxpass && t == T - 1) commit_a(bar(B_DONE));
            }
            __syncwarp();
          }
        }
        n_stage += NST;
      }
    }
  } else {
    // ---- warps 0-7: softmax.  Item boundaries are software-pipelined: the next item's q row is fetched before the key loop
    // of the current one, and the epilogue of item i (wait for its last PV MMA, O / l, stores) runs AFTER the masked Q copies
    // of item i+1 have been handed to the issuer, i.e. underneath the first S MMAs of item i+1. ----
    const uint32_t my_taddr = tmem + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(half * 32);   // my 32 columns of a buffer
    const uint32_t o_taddr = tmem + ((uint32_t)(quarter * 32) << 16) + O_COL2;
    const int c0 = D == 16 ? 0 : half * 8;         // first channel (within the quad) of my head slots
    uint32_t n_done = 0;
    auto q_token = [&](const Where& w, bool& valid) {
      int64_t tok_base, tok_stride;
      if (tsa) { tok_base = (int64_t)w.cb * p.H * p.W + w.cf; tok_stride = p.W; }
      else { tok_base = (int64_t)w.row * p.W; tok_stride = 1; }
      valid = w.q0 + r < (p.q_end ? p.q_end : N);
      return tok_base + (int64_t)(valid ? w.q0 + r : 0) * tok_stride;
    };
    auto load_q = [&](const Where& w, uint4& qa, uint4& qb) {
      bool valid;
      const int64_t tok = q_token(w, valid);
      qa = make_uint4(0, 0, 0, 0); qb = qa;
      if (valid) {
        const uint4* src = reinterpret_cast<const uint4*>((const bf16*)p.qkv + tok * p.ldq + w.quad * 16);
        qa = src[0]; qb = src[1];
      }
    };
    // deferred epilogue of the previous item
    bool e_pending = false, e_valid = false;
    int64_t e_tok = 0;
    int e_quad = 0, e_item = 0;
    float e_m[HPT];
#pragma unroll
    for (int i = 0; i < HPT; ++i) e_m[i] = 0.f;
    auto epilogue = [&]() {
      wait_fast(bar(B_DONE), n_done & 1); ++n_done;    // every PV MMA of that item has completed
      tc_fence_after();
      // O / l.  D_h = [P_h V over the 8-channel group holding head h | l_h x 8]
      uint32_t o[32];
      __syncwarp();
      if (D == 8) {
        uint32_t t16[16];
        tmem_ld_x16(o_taddr + 16 * half, t16);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 16; ++i) o[i] = t16[i];
      } else {
        tmem_ld_x32(o_taddr + (D == 4 ? 32 * half : 0), o);
        tmem_ld_wait();
      }
      tc_fence_before();                             // (the next item's first PV MMA overwrites O only after all 8 warps arrived on bar_p)
      // the bound keeps every exponent <= 0; if it was so loose that a whole row underflowed (denominator ~ 0) the item is
      // queued for the exact launch, which overwrites all of its rows
      bool bad = false;
      if (D == 16) bad = e_valid && !(__uint_as_float(o[16]) > 1e-30f);
      else {
#pragma unroll
        for (int hh = 0; hh < HPT; ++hh) bad = bad || (e_valid && !(__uint_as_float(o[hh * 16 + 8]) > 1e-30f));
      }
      if (!exact && items.redo_list != nullptr && __any_sync(0xffffffffu, bad) && lane == 0) {
        if (atomicExch(&items.redo_flag[e_item], 1) == 0) items.redo_list[atomicAdd(items.redo_count, 1)] = e_item;
      }
      if (e_valid) {
        if (D == 16) {                               // both threads of the row hold the same 32 columns: split the 16 dims
          const float l = __uint_as_float(o[16]);
          const float inv = rcp_approx(l);              // (one MUFU.RCP; `1.0f / l` is a guarded Newton sequence of ~10 instructions)
          float v[8];
#pragma unroll
          for (int d = 0; d < 8; ++d) v[d] = __uint_as_float(half ? o[8 + d] : o[d]) * inv;
          store8((bf16*)p.out + e_tok * p.ldo + e_quad * 16 + half * 8, v);
          if (p.lse && half == 0) p.lse[e_tok * p.heads + e_quad] = e_m[0] * c + log2f(l);
        } else {
#pragma unroll
          for (int hh = 0; hh < HPT; ++hh) {
            const int head = half * HPT + hh;        // head within the quad
            const int off = hh * 16 + ((D == 4) ? (hh & 1) * 4 : 0);   // head h's dims start at (h*D) % 8 inside its group
            const float l = __uint_as_float(o[hh * 16 + 8]);
            const float inv = rcp_approx(l);              // (one MUFU.RCP; `1.0f / l` is a guarded Newton sequence of ~10 instructions)
            bf16* op = (bf16*)p.out + e_tok * p.ldo + e_quad * 16 + head * D;
            if (D == 4) {
              float v[4];
#pragma unroll
              for (int d = 0; d < 4; ++d) v[d] = __uint_as_float(o[off + d]) * inv;
              store4(op, v);
            } else {
              float v[8];
#pragma unroll
              for (int d = 0; d < 8; ++d) v[d] = __uint_as_float(o[off + d]) * inv;
              store8(op, v);
            }
            if (p.lse) p.lse[e_tok * p.heads + e_quad * HPQ + head] = e_m[hh] * c + log2f(l);
          }
        }
      }
    };

    uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;      // my row's 16 q channels of the current item
    int it = item_at(0);
    if (it >= 0) { const Where w0 = where_at(0, it); load_q(w0, qa, qb); }
    for (int k = 0; it >= 0; ++k) {
      if (tid == 0) TRACE(0, k);                   // item top
      const Where w = where_at(k, it);
      bool q_valid;
      const int64_t q_tok = q_token(w, q_valid);
      // per-channel extrema of k over the sequence (attn_kext_kernel), channels of my head slots only
      float kmin[NCH], kmax[NCH];
      {
        const float* ke = p.kext + (int64_t)w.row * 2 * p.C + w.quad * 16 + c0;
#pragma unroll
        for (int i = 0; i < NCH; i += 4) {
          const float4 a = *reinterpret_cast<const float4*>(ke + i), b = *reinterpret_cast<const float4*>(ke + p.C + i);
          kmin[i] = a.x; kmin[i + 1] = a.y; kmin[i + 2] = a.z; kmin[i + 3] = a.w;
          kmax[i] = b.x; kmax[i + 1] = b.y; kmax[i + 2] = b.z; kmax[i + 3] = b.w;
        }
      }
      const uint32_t qw[8] = {qa.x, qa.y, qa.z, qa.w, qb.x, qb.y, qb.z, qb.w};   // 2 channels per word
      // masked copies of Q (K-major, 8-row x 16-byte core matrices): copy h keeps head h's channels.  The two threads of a
      // row split the copies.  (Every S MMA of the previous item has completed: this warp saw its last bar_s.)
#pragma unroll
      for (int h = 0; h < HPQ; ++h) {              // compile-time h: the masks are immediates, no local array
        if ((HPQ >= 2 ? h / (HPQ / 2) : 0) != half) continue;
        uint32_t m8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) m8[i] = (2 * i >= D * h && 2 * i < D * h + D) ? qw[i] : 0u;
        uint8_t* dst = smem + Q_OFF + h * 4096 + (r >> 3) * 256 + (r & 7) * 16;
        *reinterpret_cast<uint4*>(dst) = make_uint4(m8[0], m8[1], m8[2], m8[3]);
        *reinterpret_cast<uint4*>(dst + 128) = make_uint4(m8[4], m8[5], m8[6], m8[7]);
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) arrive_a(bar(B_Q));
      if (tid == 0) TRACE(1, k);      // Q handed over
      // row bounds per head slot: sum_d min/max(q_d kmax_d, q_d kmin_d) <= s_ij <= ... (raw score units)
      float m[HPT];
      bool wide = false;                 // some row of this warp spans more than 120 binades: polynomial needs its clamp
#pragma unroll
      for (int hh = 0; hh < HPT; ++hh) {
        float hi = 0.f, lo = 0.f;
#pragma unroll
        for (int d = 0; d < D; ++d) {
          const int i = hh * D + d;                // index into my channel range; quad channel = c0 + i
          // q channel c0 + i lives in word (c0 + i) / 2; c0 is 0 or 8, so both candidates have compile-time indices
          const uint32_t word = (D != 16 && half) ? qw[(4 + (i >> 1)) & 7] : qw[(i >> 1) & 7];
          const float qv = __uint_as_float((i & 1) ? (word & 0xFFFF0000u) : (word << 16));   // bf16 -> fp32
          const float a = qv * kmax[i], b = qv * kmin[i];
          hi += fmaxf(a, b); lo += fminf(a, b);
        }
        m[hh] = hi;
        wide = wide || !((hi - lo) * c < 120.0f);
      }
      wide = __any_sync(0xffffffffu, wide);

      if (e_pending) epilogue();                     // previous item: underneath this item's first S MMAs
      if (tid == 0) TRACE(4, k);      // previous epilogue done
      const int it_next = item_at(k + 1);            // claimed by the TMA warp before it loaded this item's first stage
      if (it_next >= 0) { const Where wn = where_at(k + 1, it_next); load_q(wn, qa, qb); }   // next item's q row

      if (exact) {
        // ---- pass 0: exact row maxima ----
#pragma unroll
        for (int i = 0; i < HPT; ++i) m[i] = -CUDART_INF_F;
        for (int t = 0; t < T; ++t) {
          const int b = t % (int)NBUF;
          wait_buf(B_S, b);
          tc_fence_after();
          const int kcount = min(KT, N - t * KT);  // valid keys of this tile (per head)
          uint32_t s32[32];
          __syncwarp();
          tmem_ld_x32(my_taddr + b * BUF_COLS, s32);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int key = D == 4 ? (i & 15) : (D == 8 ? i : half * 32 + i);
            if (key < kcount) m[D == 4 ? i / 16 : 0] = fmaxf(m[D == 4 ? i / 16 : 0], __uint_as_float(s32[i]));
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_a(bar(B_P + b));    // S(t) consumed
        }
        if (HPQ == 1) {                            // the two threads of a row each saw half of the keys
          s_xch[half * 128 + r] = m[0];
          asm volatile("bar.sync 1, 256;" ::: "memory");
          m[0] = fmaxf(s_xch[r], s_xch[128 + r]);
          asm volatile("bar.sync 1, 256;" ::: "memory");
        }
      }

      // ---- P = ex2(S c - m c), written over S ----
      float mc[HPT];
#pragma unroll
      for (int i = 0; i < HPT; ++i) mc[i] = m[i] * c;
      const float mc0 = mc[0], mc1 = mc[HPT - 1];
      uint32_t sc[32], pk[16];
      for (int t0 = 0; t0 < T; t0 += (int)NBUF) {
#pragma unroll
        for (int b = 0; b < (int)NBUF; ++b) {      // compile-time buffer index: TMEM / barrier addresses are immediates
          const int t = t0 + b;
          if (t >= T) break;
          {
            uint32_t& ph = b == 0 ? ph0 : (b == 1 ? ph1 : ph2);
            wait_fast(bar(B_S + b), ph);
            ph ^= 1;
            tc_fence_after();
            __syncwarp();
            tmem_ld_x32(my_taddr + b * BUF_COLS, sc);
          }
          tmem_ld_wait();
          if (tid == 0 && t == 0) TRACE(5, k);   // first S tile in registers
          const bool tail = t == T - 1 && T * KT > N;   // last tile: absent keys score 0, which may exceed the bound
          if (tail) {
#pragma unroll
            for (int i = 0; i < 32; ++i) sc[i] = __float_as_uint(fminf(__uint_as_float(sc[i]), m[D == 4 ? i / 16 : 0]));
          }
          if (tail || wide) { softmax_half<true, 0>(sc, pk, c, mc0); softmax_half<true, 1>(sc, pk, c, mc1); }
          else { softmax_half<false, 0>(sc, pk, c, mc0); softmax_half<false, 1>(sc, pk, c, mc1); }
          tmem_st_x16(my_taddr + b * BUF_COLS, pk);  // score pair (2i, 2i+1) -> 32-bit cell i of my own columns
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) arrive_a(bar(B_P + b));     // P(t) in TMEM
        }
      }
      if (tid == 0) TRACE(6, k);      // last P published
      e_pending = true; e_valid = q_valid; e_tok = q_tok; e_quad = w.quad; e_item = it;
#pragma unroll
      for (int i = 0; i < HPT; ++i) e_m[i] = m[i];
      it = it_next;
    }
    if (e_pending) epilogue();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace tma_attn

// 4-D view of the q|k|v token matrix: (channel, w, h, b); box = 8 channels x 128 keys along the attended axis
static int make_tmap_qkv(CUtensorMap* out, const AttnParams& p) {
  PFN_encodeTiled enc = get_encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled is not available from the driver"); return TFSWA_ECUDA; }
  const bool tsa = p.geom == TFSWA_GEOM_TSA;
  cuuint64_t dims[4] = {(cuuint64_t)(3 * p.C), (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
  cuuint64_t strides[3] = {(cuuint64_t)p.ldq * 2, (cuuint64_t)p.W * p.ldq * 2, (cuuint64_t)p.H * p.W * p.ldq * 2};
  cuuint32_t box[4] = {8, tsa ? 1u : (cuuint32_t)tma_attn::SKEYS, tsa ? (cuuint32_t)tma_attn::SKEYS : 1u, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (((uintptr_t)p.qkv & 15) || (strides[0] & 15)) { set_error("attn_tc: q|k|v base / row stride must be 16-byte aligned"); return TFSWA_EINVAL; }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p.qkv), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("attn_tc: cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return TFSWA_ECUDA; }
  return TFSWA_OK;
}

// bytes of work space after the k extrema: [redo count (16 B)] [redo flag per item] [redo list]
int64_t attn_axial_tma_work_bytes(const AttnParams& p) {
  const int N = p.geom == TFSWA_GEOM_TSA ? p.H : p.W;
  const int64_t rows = p.geom == TFSWA_GEOM_TSA ? (int64_t)p.B * p.W : (int64_t)p.B * p.H;
  const int64_t n_items = rows * ((N + tma_attn::QTILE - 1) / tma_attn::QTILE) * (p.C / 16);
  return 16 + 2 * n_items * (int64_t)sizeof(int);
}

// queries [0, p.q_end) of every sequence (q_end = 0: all); needs p.kext filled by attn_kext_kernel; `work` = device buffer
// of attn_axial_tma_work_bytes(p) bytes
int attn_axial_tma_bf16(const AttnParams& p, void* work, cudaStream_t st) {
  using namespace tma_attn;
  const int D = p.C / p.heads;
  const int N = p.geom == TFSWA_GEOM_TSA ? p.H : p.W;
  const int rows = p.geom == TFSWA_GEOM_TSA ? p.B * p.W : p.B * p.H;
  const int q_end = p.q_end ? p.q_end : N;
  CUtensorMap tm;
  int rc = make_tmap_qkv(&tm, p);
  if (rc) return rc;
  static int sms = 0;
  static DeviceOnce attr_once;
  if (attr_once.needed()) {
    cudaError_t e1 = cudaFuncSetAttribute(tc_attn_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<4>());
    cudaError_t e2 = cudaFuncSetAttribute(tc_attn_tma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<8>());
    cudaError_t e3 = cudaFuncSetAttribute(tc_attn_tma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<16>());
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, attr_once.dev);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || sms <= 0) { set_error("attn_tc: cudaFuncSetAttribute failed"); return TFSWA_ECUDA; }
    attr_once.done();
  }
  Items items = {};
  items.nqt = (q_end + QTILE - 1) / QTILE;
  items.nquads = p.C / 16;
  items.n = rows * items.nqt * items.nquads;
  items.exact = p.force_exact ? 1 : 0;
  int* wk = (int*)work;                 // header: [redo count, claim counter of the main launch, claim counter of the exact launch, -]
  if (cudaMemsetAsync(wk, 0, 16 + (items.exact ? 0 : (size_t)items.n * sizeof(int)), st) != cudaSuccess) return check_launch("attn_tc(tma) memset");
  items.next = wk + 1;
  if (!items.exact) { items.redo_count = wk; items.redo_flag = wk + 4; items.redo_list = wk + 4 + items.n; }
  static int pad = -1;                  // TFSWA_TMA_SMEM_PAD=<bytes>: occupancy experiment (e.g. 70000 -> one CTA per SM)
  if (pad < 0) { const char* e = getenv("TFSWA_TMA_SMEM_PAD"); pad = e ? atoi(e) : 0; }
  auto launch = [&](const Items& its, int grid) {
    if (D == 4) {
      if (pad > 0) cudaFuncSetAttribute(tc_attn_tma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<4>() + pad);
      tc_attn_tma_kernel<4><<<grid, NTHREADS, smem_bytes<4>() + (pad > 0 ? pad : 0), st>>>(tm, p, its);
    } else if (D == 8) tc_attn_tma_kernel<8><<<grid, NTHREADS, smem_bytes<8>(), st>>>(tm, p, its);
    else tc_attn_tma_kernel<16><<<grid, NTHREADS, smem_bytes<16>(), st>>>(tm, p, its);
  };
  const int full = 2 * sms;             // two resident CTAs per SM
  launch(items, items.n < full ? items.n : full);
  rc = check_launch("attn_tc(tma)");
  if (rc || items.exact) return rc;
  // the (normally empty) exact launch over the items whose bound underflowed
  Items redo = items;
  redo.list = items.redo_list; redo.count = items.redo_count; redo.exact = 1;
  redo.redo_count = nullptr; redo.redo_flag = nullptr; redo.redo_list = nullptr;
  redo.next = wk + 2;
  launch(redo, items.n < sms ? items.n : sms);
  return check_launch("attn_tc(tma, exact pass)");
}

}  // namespace tfswa

#ifdef TFSWA_TMA_TRACE
// debug builds only (tools/build_variant.sh trace tc_attn_tma.cu -DTFSWA_TMA_TRACE): copy the clock samples of CTA 0 out
extern "C" int tfswa_dbg_tma_trace(long long* host) {
  return (int)cudaMemcpyFromSymbol(host, tfswa::tma_attn::g_trace, sizeof(long long) * 8 * 128);
}
#endif
