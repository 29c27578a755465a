// Optimiser step over the flat parameter arena: global gradient norm, clip and AdamW in two HBM-bound launches
// (reference: clip_grad_norm_ + optimizer.step() at src/training/trainer.py:214-219, AdamW built at scripts/train.py:251-255).
// The arena (train_step.py) keeps every parameter, gradient and moment in one contiguous fp32 buffer each, so the
// step is read p,g,m,v once + write p,m,v once = 28 B per parameter, 128-bit accesses, no per-tensor launches and no
// host synchronisation: the clip coefficient is computed on the device from the norm the first launch left behind.
#include "common.cuh"

namespace tfswa {

__global__ void __launch_bounds__(256) grad_sumsq_kernel(const float* __restrict__ g, int64_t nvec, double* __restrict__ out) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 t = reinterpret_cast<const float4*>(g)[i];
    acc = fmaf(t.x, t.x, acc);
    acc = fmaf(t.y, t.y, acc);
    acc = fmaf(t.z, t.z, acc);
    acc = fmaf(t.w, t.w, acc);
  }
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? part[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) atomicAdd(out, (double)v);      // <= 592 double atomics per step: order-insensitive at fp32 resolution
  }
}

struct AdamwParams {
  float* p; const float* g; float* m; float* v;
  int64_t nvec;
  const double* sumsq;
  float* norm_out;
  float grad_scale;      // 1/world: the arena holds the SUM over ranks after the all-reduce
  float max_norm;        // <= 0: no clipping
  float lr, beta1, beta2, eps, weight_decay;
  long long step;        // optimiser steps taken including this one (host count)
  long long* skipped;    // optional device counter of skipped (non-finite) updates: effective step = step - *skipped
};

__global__ void __launch_bounds__(256) adamw_clip_kernel(AdamwParams a) {
  // total_norm of the averaged gradient; clip_grad_norm_ semantics: coef = min(1, max_norm / (norm + 1e-6))
  const float norm = (float)sqrt(*a.sumsq) * a.grad_scale;
  if (blockIdx.x == 0 && threadIdx.x == 0 && a.norm_out) *a.norm_out = norm;
  if (!isfinite(norm)) {                          // inf/nan gradients: skip the update (what GradScaler.step does) ...
    if (blockIdx.x == 0 && threadIdx.x == 0 && a.skipped) *a.skipped += 1;   // ... and do not advance the bias corrections
    return;
  }
  float gs = a.grad_scale;
  if (a.max_norm > 0.f) gs *= fminf(1.f, a.max_norm / (norm + 1e-6f));
  // (*skipped only changes in launches that return above, so every thread of an updating launch reads the same value)
  const double t_eff = (double)(a.step - (a.skipped ? *a.skipped : 0));
  const float bc1 = (float)(1.0 - pow((double)a.beta1, t_eff));
  const float rsqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)a.beta2, t_eff)));
  const float decay = 1.f - a.lr * a.weight_decay, step = a.lr / bc1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float4 p = reinterpret_cast<float4*>(a.p)[i];
    const float4 g4 = reinterpret_cast<const float4*>(a.g)[i];
    float4 m = reinterpret_cast<float4*>(a.m)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    float* pp = &p.x; const float* gp = &g4.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float g = gp[j] * gs;
      mp[j] = fmaf(a.beta1, mp[j], (1.f - a.beta1) * g);
      vp[j] = fmaf(a.beta2, vp[j], (1.f - a.beta2) * g * g);
      const float denom = fmaf(sqrtf(vp[j]), rsqrt_bc2, a.eps);
      pp[j] = fmaf(-step, mp[j] / denom, pp[j] * decay);
    }
    reinterpret_cast<float4*>(a.p)[i] = p;
    reinterpret_cast<float4*>(a.m)[i] = m;
    reinterpret_cast<float4*>(a.v)[i] = v;
  }
}

static inline unsigned flat_grid(int64_t nvec) {
  const int64_t want = ceil_div64(nvec, 256);
  return (unsigned)(want < 148 * 4 ? (want > 0 ? want : 1) : 148 * 4);
}

}  // namespace tfswa

using namespace tfswa;

extern "C" {

int tfswa_grad_sumsq(const float* g, int64_t n, double* sumsq, void* stream) {
  TFSWA_REQUIRE(g && sumsq && n > 0 && n % 4 == 0, "grad_sumsq: n=%lld must be a positive multiple of 4", (long long)n);
  TFSWA_REQUIRE(((uintptr_t)g & 15) == 0, "grad_sumsq: g must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(sumsq, 0, sizeof(double), st) != cudaSuccess) {
    check_launch("grad_sumsq memset");
    return TFSWA_ECUDA;
  }
  grad_sumsq_kernel<<<flat_grid(n / 4), 256, 0, st>>>(g, n / 4, sumsq);
  return check_launch("grad_sumsq");
}

int tfswa_adamw_clip_step(float* p, const float* g, float* m, float* v, int64_t n, const double* sumsq, float* norm_out,
                          float grad_scale, float max_norm, float lr, float beta1, float beta2, float eps, float weight_decay,
                          int64_t step, long long* skipped, void* stream) {
  TFSWA_REQUIRE(p && g && m && v && sumsq && n > 0 && n % 4 == 0, "adamw_clip_step: bad arguments (n=%lld)", (long long)n);
  TFSWA_REQUIRE((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0, "adamw_clip_step: buffers must be 16-byte aligned");
  TFSWA_REQUIRE(step >= 1 && beta1 >= 0.f && beta1 < 1.f && beta2 >= 0.f && beta2 < 1.f && lr >= 0.f && eps > 0.f && grad_scale > 0.f,
                "adamw_clip_step: bad hyper-parameters");
  AdamwParams a;
  a.p = p; a.g = g; a.m = m; a.v = v; a.nvec = n / 4; a.sumsq = sumsq; a.norm_out = norm_out;
  a.grad_scale = grad_scale; a.max_norm = max_norm; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps;
  a.weight_decay = weight_decay;
  a.step = step; a.skipped = skipped;
  adamw_clip_kernel<<<flat_grid(a.nvec), 256, 0, (cudaStream_t)stream>>>(a);
  return check_launch("adamw_clip_step");
}

}  // extern "C"
