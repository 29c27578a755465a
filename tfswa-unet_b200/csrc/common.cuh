// Shared device/host helpers for libtfswa_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include "../../include/tfswa_b200.h"

namespace tfswa {

// ---- error plumbing (thread-local message, never throws across the ABI) ----
void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define TFSWA_REQUIRE(cond, ...)                 \
  do {                                           \
    if (!(cond)) {                               \
      ::tfswa::set_error(__VA_ARGS__);           \
      return TFSWA_EINVAL;                       \
    }                                            \
  } while (0)

typedef __nv_bfloat16 bf16;

// ---- element access: everything computes in fp32, storage is float or bf16 ----
__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(bf16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

// load 4 consecutive elements (16 B fp32 / 8 B bf16), pointer must be aligned to that
__device__ __forceinline__ void load4(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load4(const bf16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
  __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  v[0] = __low2float(a); v[1] = __high2float(a); v[2] = __low2float(b); v[3] = __high2float(b);
}
__device__ __forceinline__ void store4(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
  __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}
// 8 consecutive elements (32 B fp32 / 16 B bf16)
__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  float4 a = *reinterpret_cast<const float4*>(p);
  float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* p, float (&v)[8]) {
  uint4 t = *reinterpret_cast<const uint4*>(p);
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) { v[2 * i] = __low2float(h[i]); v[2 * i + 1] = __high2float(h[i]); }
}
__device__ __forceinline__ void store8(float* p, const float (&v)[8]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ void store8(bf16* p, const float (&v)[8]) {
  uint4 t;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&t);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = t;
}

// exact (erf) GELU, as nn.GELU() default (attention.py:122, blocks.py:88)
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
// erf-GELU with erf from Abramowitz-Stegun 7.1.26 (|err| < 1.5e-7, far below bf16 resolution) on the raw approximate
// MUFU ops (rcp.approx / ex2.approx: no slow-path branches, no denormal fix-ups): ~17 instructions incl. two MUFU,
// instead of erff's ~40 with __frcp_rn/exp2f.  Used by the bf16 tensor-core epilogues only; the fp32 parity path keeps erff.
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erfv = copysignf(fmaf(-p * t, e, 1.0f), x);          // erf(x/sqrt2) = sign(x) * (1 - p t e)
  const float hx = 0.5f * x;
  return fmaf(hx, erfv, hx);
}
// GELU of the MLP hidden activations in the fused bf16 tail kernels (the value is rounded to bf16 and consumed by fc2):
// x * Phi(x) with Phi(x) = 0.5 (1 + tanh(u)), u = x (c1 + c3 x^2 + c5 x^4) FITTED TO THE ERF FORM (not the classic
// "tanh GELU": max |error| of the fit against erf-GELU is 3.0e-5 over the reals, tests/test_gelu_fit.py pins it) and
// one MUFU.TANH (2^-11 relative).  9 instructions / 1 MUFU instead of 17 / 2: the tail kernels are bound by exactly this.
// Total deviation from erf-GELU <= 3e-5 + 2.5e-4 min(|x|, 6) under the worst-case tanh error model: below the bf16
// rounding of the result for x > -1.15 and below 1.5e-3 absolute everywhere; measured end to end (tail output against the
// fp32 restatement, 20000 tokens) the relative L2 error is 1.9199e-3 with this form and 1.9196e-3 with the erf form.  -DTFSWA_GELU_HIDDEN_EXACT selects the A&S erf instead.
__device__ __forceinline__ float gelu_hidden(float x) {
#ifdef TFSWA_GELU_HIDDEN_EXACT
  return gelu_erf_fast(x);
#else
  x = fmaxf(x, -6.0f);                               // gelu(x <= -6) = -0.0 to 8 digits; bounds the 0.5|x| factor on the tanh error
  const float x2 = fminf(x * x, 51.6f);              // beyond |x| = 7.2 the fitted polynomial would turn over; tanh is saturated there
  float g = fmaf(-3.58732361e-4f, x2, 0.0370503451f);
  g = fmaf(g, x2, 0.797458471f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x * g));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
#endif
}
// The same arithmetic for two activations at once on the packed fp32 pipe (add / mul / fma.rn.f32x2: one issue slot per two
// elements; the clamps and MUFU.TANH stay scalar), bias added first, result rounded to a bf16 pair (lo = first element):
// 15 instructions per pair instead of 21 - the tail kernels are bound by the issue rate of exactly this sequence.  Bit-identical
// to gelu_hidden(a + b): every packed op rounds like its scalar form.
__device__ __forceinline__ uint32_t gelu_hidden_bf16x2(float a0, float a1, float b0, float b1) {
#ifdef TFSWA_GELU_HIDDEN_EXACT
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(gelu_hidden(a1 + b1)), "f"(gelu_hidden(a0 + b0)));
  return r;
#else
  uint64_t x, x2, g, u, hx, y;
  float x0, x1, q0, q1, u0, u1, t0, t1, y0, y1;
  asm("{ .reg .b64 a, b; mov.b64 a, {%1, %2}; mov.b64 b, {%3, %4}; add.rn.f32x2 %0, a, b; }" : "=l"(x) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x0), "=f"(x1) : "l"(x));
  x0 = fmaxf(x0, -6.0f); x1 = fmaxf(x1, -6.0f);
  asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(x0), "f"(x1));
  asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(x2) : "l"(x));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(q0), "=f"(q1) : "l"(x2));
  q0 = fminf(q0, 51.6f); q1 = fminf(q1, 51.6f);
  asm("mov.b64 %0, {%1, %2};" : "=l"(x2) : "f"(q0), "f"(q1));
  asm("{ .reg .b64 c5, c3; mov.b64 c5, {%2, %2}; mov.b64 c3, {%3, %3}; fma.rn.f32x2 %0, c5, %1, c3; }"
      : "=l"(g) : "l"(x2), "f"(-3.58732361e-4f), "f"(0.0370503451f));
  asm("{ .reg .b64 c1; mov.b64 c1, {%3, %3}; fma.rn.f32x2 %0, %1, %2, c1; }" : "=l"(g) : "l"(g), "l"(x2), "f"(0.797458471f));
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(u) : "l"(x), "l"(g));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(u0), "=f"(u1) : "l"(u));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(u0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(u1));
  asm("{ .reg .b64 h; mov.b64 h, {%2, %2}; mul.rn.f32x2 %0, %1, h; }" : "=l"(hx) : "l"(x), "f"(0.5f));
  asm("{ .reg .b64 t; mov.b64 t, {%2, %3}; fma.rn.f32x2 %0, %1, t, %1; }" : "=l"(y) : "l"(hx), "f"(t0), "f"(t1));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(y0), "=f"(y1) : "l"(y));
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(y1), "f"(y0));
  return r;
#endif
}
// d/dx GELU
__device__ __forceinline__ float gelu_erf_grad(float x) {
  const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752f));
  const float pdf = 0.39894228040143268f * __expf(-0.5f * x * x);
  return cdf + x * pdf;
}
// the same on the raw approximate MUFU ops, sharing ONE exponential between the A&S 7.1.26 erf and the Gaussian density
// (exp(-z^2) with z = |x|/sqrt2 is exp(-x^2/2)): ~20 instructions instead of ~55; |error| < 3e-7, used for bf16 storage only
__device__ __forceinline__ float gelu_erf_grad_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float erfv = copysignf(fmaf(-p * t, e, 1.0f), x);
  return fmaf(0.5f, erfv, 0.5f) + x * (0.39894228040143268f * e);
}
// storage-type dispatch of the elementwise kernels: bf16 tensors take the fast forms (their error is 4 orders of magnitude
// below the rounding of the result), fp32 tensors (the parity path) the libm forms
template <typename T> __device__ __forceinline__ float gelu_for(float x) { return gelu_erf(x); }
template <> __device__ __forceinline__ float gelu_for<bf16>(float x) { return gelu_erf_fast(x); }
template <typename T> __device__ __forceinline__ float gelu_grad_for(float x) { return gelu_erf_grad(x); }
template <> __device__ __forceinline__ float gelu_grad_for<bf16>(float x) { return gelu_erf_grad_fast(x); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// One-time-per-DEVICE set-up guard for launchers (cudaFuncSetAttribute is a per-device property; a process-wide bool
// would leave the second GPU of a multi-device process without the attribute).  The guarded body must be idempotent:
// two host threads may both run it, which is harmless.
struct DeviceOnce {
  unsigned long long mask = 0;          // bit d = done on device d (devices >= 64: always redo)
  int dev = 0;
  bool needed() {
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    return dev >= 64 || !((__atomic_load_n(&mask, __ATOMIC_ACQUIRE) >> dev) & 1ull);
  }
  void done() { if (dev < 64) __atomic_fetch_or(&mask, 1ull << dev, __ATOMIC_RELEASE); }
};

}  // namespace tfswa
