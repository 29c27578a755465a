// Helpers shared by the tcgen05 attention kernels (tc_attn_tma.cu: axial TSA / FSA, tc_attn_win.cu: SW-MSA windows):
// the 4-D TMA load of a q|k|v box, the MN-major instruction descriptor, packed fp32 arithmetic (FFMA2 / FADD2) and the
// FMA-pipe exponential that shares the softmax with MUFU.EX2 (tools/mufu_bench.cu: 16.0 results/clk/SM for MUFU alone,
// 21.8 with one element pair in three on the polynomial).
#pragma once
#include "sm100.cuh"

namespace tfswa {
namespace tcmath {

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(sm100::smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(sm100::smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// kind::f16 instruction descriptor with a K-major A and an MN-major B operand (bit 16)
__device__ __forceinline__ uint32_t idesc_bf16_bmn(uint32_t M, uint32_t N) { return sm100::umma_idesc_bf16(M, N) | (1u << 16); }

__device__ __forceinline__ float ex2_f32(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { uint32_t y; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }
__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ uint64_t pk2u(uint32_t a, uint32_t b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "r"(a), "r"(b)); return r; }
__device__ __forceinline__ void up2(uint64_t r, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// mbarrier / TMA / commit on 32-bit shared-window addresses: the barriers live in dynamic shared memory at constant
// offsets from the CTA's base, so every use is [base register + immediate] (a `uint64_t*` to a __shared__ barrier costs a
// generic->shared conversion - S2UR SR_CgaCtaId, UMOV, ULEA - at each use: ~30 instructions per item and warp)
__device__ __forceinline__ bool try_wait_a(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void wait_a(uint32_t bar, uint32_t parity) {   // bounded like sm100::mbar_wait: a protocol bug traps
  uint32_t spins = 0;
  uint64_t t0 = 0;
  while (!try_wait_a(bar, parity)) {
    if ((++spins & 255u) == 0) {
      uint64_t now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t0 == 0) t0 = now;
      else if (now - t0 > 4000000000ull) __trap();
    }
  }
}
// hot-loop form: the first probe costs YIELD + SYNCS + BRA; the retry loop with its timeout bookkeeping is off the fast path
__device__ __forceinline__ void wait_fast(uint32_t bar, uint32_t parity) {
  if (!try_wait_a(bar, parity)) wait_a(bar, parity);
}
__device__ __forceinline__ void arrive_a(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void expect_tx_a(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void commit_a(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_4d_a(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 2^x for a pair of x <= 0 on the FMA / ALU pipes: x = n + r (round to nearest through the 1.5 * 2^23 trick), degree-3
// minimax polynomial of 2^r on [-0.5, 0.5] (7.5e-5 relative - P is rounded to bf16, 3.9e-3), exponent patched in with an
// integer shift-add.  3 FADD2 + 3 FFMA2 + 2 LEA for two results.  CLAMP bounds n at -125 (needed only when the row's
// bound is more than 120 binades above its smallest possible score).
template <bool CLAMP>
__device__ __forceinline__ void ex2_poly2(uint64_t x, float& e0, float& e1) {
  constexpr float MAGIC = 12582912.0f;
  if (CLAMP) { float a, b; up2(x, a, b); x = pk2(fmaxf(a, -125.0f), fmaxf(b, -125.0f)); }
  const uint64_t t = add2(x, pk2(MAGIC, MAGIC));
  const uint64_t u = add2(t, pk2(-MAGIC, -MAGIC));
  const uint64_t r = sub2(x, u);
  uint64_t p = fma2(pk2(0.0551716685f, 0.0551716685f), r, pk2(0.2426111251f, 0.2426111251f));
  p = fma2(p, r, pk2(0.6932609677f, 0.6932609677f));
  p = fma2(p, r, pk2(0.9999280572f, 0.9999280572f));
  float p0, p1, t0, t1;
  up2(p, p0, p1); up2(t, t0, t1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

// 16 scores sc[16*HALF ..] -> 8 packed bf16x2 probabilities pk[8*HALF ..]: p = 2^(s*c - mc); every POLY_K-th element
// pair takes the polynomial (0 = MUFU only)
template <bool CLAMP, int HALF, int POLY_K>
__device__ __forceinline__ void softmax_half(const uint32_t (&sc)[32], uint32_t (&pk)[16], float c, float mc) {
  const uint64_t c2 = pk2(c, c), nm = pk2(-mc, -mc);
#pragma unroll
  for (int i = 8 * HALF; i < 8 * HALF + 8; ++i) {
    const uint64_t x = fma2(pk2u(sc[2 * i], sc[2 * i + 1]), c2, nm);
    float e0, e1;
    if (POLY_K > 0 && (i % POLY_K) == POLY_K - 1) ex2_poly2<CLAMP>(x, e0, e1);
    else { float x0, x1; up2(x, x0, x1); e0 = ex2_f32(x0); e1 = ex2_f32(x1); }
    pk[i] = pack_bf16x2(e0, e1);
  }
}

// the same with an explicit choice of the polynomial pairs: bit i of MASK (i = pair index 0..15 within the 32 scores)
template <bool CLAMP, int HALF, uint32_t MASK>
__device__ __forceinline__ void softmax_half_mask(const uint32_t (&sc)[32], uint32_t (&pk)[16], float c, float mc) {
  const uint64_t c2 = pk2(c, c), nm = pk2(-mc, -mc);
#pragma unroll
  for (int i = 8 * HALF; i < 8 * HALF + 8; ++i) {
    const uint64_t x = fma2(pk2u(sc[2 * i], sc[2 * i + 1]), c2, nm);
    float e0, e1;
    if ((MASK >> i) & 1u) ex2_poly2<CLAMP>(x, e0, e1);
    else { float x0, x1; up2(x, x0, x1); e0 = ex2_f32(x0); e1 = ex2_f32(x1); }
    pk[i] = pack_bf16x2(e0, e1);
  }
}

}  // namespace tcmath
}  // namespace tfswa
