// Bandwidth-bound backward kernels: activation / BatchNorm-apply gradients, the LayerNorm (LN-hat) input
// gradient, batch (branch) reduction of a broadcast residual's gradient, bilinear resize gradient.
// All use 128-bit accesses and fp32 math; per-channel reductions go registers -> shared -> one atomic per CTA.
#include "common.cuh"

namespace tfswa {

// ---------------------------------------------------------------------------------------------
// out = g * gelu'(pre)                       (mode 0: gradient through an exact-erf GELU epilogue)
// out = g + ds1[c] + 2*pre*ds2[c]            (mode 1: add the gradient that reaches `pre` through the
//                                             train-mode BatchNorm column sums  S1 = sum pre, S2 = sum pre^2)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ g, const T* __restrict__ pre, const float* __restrict__ ds,
                               T* __restrict__ out, int64_t nvec, int C, int mode) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i * 8;
    float a[8], b[8];
    load8(g + e, a);
    load8(pre + e, b);
    if (mode == 0) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] *= gelu_grad_for<T>(b[j]);
    } else {
      const int c = (int)(e % C);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += ds[c + j] + 2.f * b[j] * ds[C + c + j];
    }
    store8(out + e, a);
  }
}

// ---------------------------------------------------------------------------------------------
// backward of y = act(v*scale + shift) (+r1) (+r2):  dv = dy*act'(t)*scale,  dscale[c] = sum dy*act'(t)*v,
// dshift[c] = sum dy*act'(t)   (t = v*scale+shift).  Each thread keeps the same 8 channels for its whole
// grid-stride loop (blockDim*8 is a multiple of C), so channel sums accumulate in registers.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) affine_act_bwd_kernel(const T* __restrict__ dy, const T* __restrict__ v,
                                                             const float* __restrict__ scale, const float* __restrict__ shift,
                                                             T* __restrict__ dv, float* __restrict__ dscale,
                                                             float* __restrict__ dshift, int64_t nvec, int C, int epilogue) {
  __shared__ float s_sc[256], s_sh[256];
  const int c = (threadIdx.x * 8) % C;
  float sc[8], sh[8], asc[8], ash[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = scale ? scale[c + j] : 1.f; sh[j] = shift ? shift[c + j] : 0.f; asc[j] = 0.f; ash[j] = 0.f; }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i * 8;
    float g[8], x[8];
    load8(dy + e, g);
    load8(v + e, x);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float t = x[j] * sc[j] + sh[j];
      const float gt = (epilogue == TFSWA_EPI_GELU) ? g[j] * gelu_grad_for<T>(t) : g[j];
      asc[j] += gt * x[j];
      ash[j] += gt;
      g[j] = gt * sc[j];
    }
    store8(dv + e, g);
  }
  if (dscale) {
    for (int i = threadIdx.x; i < C; i += blockDim.x) { s_sc[i] = 0.f; s_sh[i] = 0.f; }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 8; ++j) { atomicAdd(&s_sc[c + j], asc[j]); atomicAdd(&s_sh[c + j], ash[j]); }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) { atomicAdd(dscale + i, s_sc[i]); atomicAdd(dshift + i, s_sh[i]); }
  }
}

// ---------------------------------------------------------------------------------------------
// LN-hat input gradient, per row over K channels:
//   xh = (x-mean)*rstd,  dx = rstd * (da - mean_k(da) - xh * mean_k(da*xh))
// LPR lanes per row as in row_stats_kernel.
// ---------------------------------------------------------------------------------------------
template <typename T, int LPR>
__global__ void lnhat_bwd_kernel(const T* __restrict__ da, int64_t ldd, int64_t d_bs, const T* __restrict__ x, int64_t ldx,
                                 int64_t x_bs, const float* __restrict__ st, int64_t st_bs, T* __restrict__ dx, int64_t ldo,
                                 int64_t o_bs, int64_t M, int K) {
  constexpr int RPW = 32 / LPR;
  constexpr int VPL = LPR == 32 ? 4 : 1;
  const int warps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const int64_t row = ((int64_t)blockIdx.x * warps + (threadIdx.x >> 5)) * RPW + sub;
  const bool ok = row < M;
  const int z = blockIdx.y;
  const int64_t r = ok ? row : 0;
  const T* pd = da + (int64_t)z * d_bs + r * ldd;
  const T* px = x + (int64_t)z * x_bs + r * ldx;
  const float mean = ok ? st[(int64_t)z * st_bs + r * 2] : 0.f;
  const float rstd = ok ? st[(int64_t)z * st_bs + r * 2 + 1] : 0.f;
  const int nvec = K >> 3;
  float d[VPL][8], xh[VPL][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = l + i * LPR;
    if (ok && vi < nvec) {
      load8(pd + vi * 8, d[i]);
      load8(px + vi * 8, xh[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) { xh[i][j] = (xh[i][j] - mean) * rstd; s1 += d[i][j]; s2 += d[i][j] * xh[i][j]; }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
  const float c1 = s1 / (float)K, c2 = s2 / (float)K;
  T* po = dx + (int64_t)z * o_bs + r * ldo;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = l + i * LPR;
    if (ok && vi < nvec) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rstd * (d[i][j] - c1 - xh[i][j] * c2);
      store8(po + vi * 8, o);
    }
  }
}

// out[m][:] = sum_b g[m][b][:]   (gradient of a residual that was broadcast over the branch dimension)
template <typename T>
__global__ void sum_batch_kernel(const T* __restrict__ g, T* __restrict__ out, int64_t M, int nb, int N) {
  const int nv = N >> 3;
  const int64_t total = M * nv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / nv;
    const int c = (int)(i % nv) * 8;
    float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int b = 0; b < nb; ++b) {
      float t[8];
      load8(g + (m * nb + b) * N + c, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += t[j];
    }
    store8(out + m * N + c, a);
  }
}

// ---------------------------------------------------------------------------------------------
// bilinear (align_corners=False) gradient as a gather: for every input pixel, loop over the output
// rows/cols whose two source taps can include it and re-evaluate the forward weights (no atomics).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_src_b(int o, int in_size, float scale, int& i0, int& i1, float& l1) {
  float s = ((float)o + 0.5f) * scale - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - (float)i0;
}

template <typename T>
__global__ void bilinear_bwd_kernel(const T* __restrict__ dy, T* __restrict__ dx, int B, int Hin, int Win, int Hout, int Wout,
                                    int C, float sh, float sw) {
  const int cv = C >> 3;
  const int64_t total = (int64_t)B * Hin * Win * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    int64_t t = i / cv;
    const int ix = (int)(t % Win); t /= Win;
    const int iy = (int)(t % Hin);
    const int b = (int)(t / Hin);
    // candidate outputs: source coordinate s(o) in (i-1, i+1)  ->  o in ((i-0.5)/scale - 0.5 - 1, (i+1.5)/scale - 0.5 + 1)
    int oy_lo = (int)floorf(((float)iy - 0.5f) / sh - 1.5f), oy_hi = (int)ceilf(((float)iy + 1.5f) / sh + 0.5f);
    int ox_lo = (int)floorf(((float)ix - 0.5f) / sw - 1.5f), ox_hi = (int)ceilf(((float)ix + 1.5f) / sw + 0.5f);
    if (oy_lo < 0) oy_lo = 0;
    if (ox_lo < 0) ox_lo = 0;
    if (oy_hi > Hout - 1) oy_hi = Hout - 1;
    if (ox_hi > Wout - 1) ox_hi = Wout - 1;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int oy = oy_lo; oy <= oy_hi; ++oy) {
      int y0, y1; float ly;
      bilinear_src_b(oy, Hin, sh, y0, y1, ly);
      float wy = 0.f;
      if (y0 == iy) wy += 1.f - ly;
      if (y1 == iy) wy += ly;
      if (wy == 0.f) continue;
      for (int ox = ox_lo; ox <= ox_hi; ++ox) {
        int x0, x1; float lx;
        bilinear_src_b(ox, Win, sw, x0, x1, lx);
        float wx = 0.f;
        if (x0 == ix) wx += 1.f - lx;
        if (x1 == ix) wx += lx;
        if (wx == 0.f) continue;
        float g[8];
        load8(dy + (((int64_t)b * Hout + oy) * Wout + ox) * C + c8 * 8, g);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += wy * wx * g[j];
      }
    }
    store8(dx + (((int64_t)b * Hin + iy) * Win + ix) * C + c8 * 8, acc);
  }
}

static unsigned grid_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  return (unsigned)(b < 148 * 16 ? (b > 0 ? b : 1) : 148 * 16);
}

}  // namespace tfswa

using namespace tfswa;

extern "C" {

int tfswa_act_bwd(const void* g, const void* pre, const float* ds, void* out, int64_t M, int32_t C, int32_t mode, int32_t dtype,
                  void* stream) {
  TFSWA_REQUIRE(g && pre && out && M > 0 && C % 8 == 0, "act_bwd: bad arguments");
  TFSWA_REQUIRE(mode == 0 || (mode == 1 && ds), "act_bwd: mode 1 needs ds");
  const int64_t nvec = M * C / 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TFSWA_F32) act_bwd_kernel<float><<<grid_for(nvec, 256), 256, 0, st>>>((const float*)g, (const float*)pre, ds, (float*)out, nvec, C, mode);
  else if (dtype == TFSWA_BF16) act_bwd_kernel<bf16><<<grid_for(nvec, 256), 256, 0, st>>>((const bf16*)g, (const bf16*)pre, ds, (bf16*)out, nvec, C, mode);
  else TFSWA_REQUIRE(false, "act_bwd: bad dtype");
  return check_launch("act_bwd");
}

int tfswa_affine_act_bwd(const void* dy, const void* v, const float* scale, const float* shift, void* dv, float* dscale,
                         float* dshift, int64_t M, int32_t C, int32_t epilogue, int32_t dtype, void* stream) {
  TFSWA_REQUIRE(dy && v && dv && M > 0, "affine_act_bwd: bad arguments");
  TFSWA_REQUIRE(C % 8 == 0 && C <= 256 && 2048 % C == 0, "affine_act_bwd: C=%d must divide 2048 and be a multiple of 8", C);
  TFSWA_REQUIRE((dscale == nullptr) == (dshift == nullptr), "affine_act_bwd: dscale/dshift both or none");
  const int64_t nvec = M * C / 8;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = grid_for(nvec, 256) < 148 * 4 ? grid_for(nvec, 256) : 148 * 4;
  if (dtype == TFSWA_F32)
    affine_act_bwd_kernel<float><<<grid, 256, 0, st>>>((const float*)dy, (const float*)v, scale, shift, (float*)dv, dscale, dshift, nvec, C, epilogue);
  else if (dtype == TFSWA_BF16)
    affine_act_bwd_kernel<bf16><<<grid, 256, 0, st>>>((const bf16*)dy, (const bf16*)v, scale, shift, (bf16*)dv, dscale, dshift, nvec, C, epilogue);
  else TFSWA_REQUIRE(false, "affine_act_bwd: bad dtype");
  return check_launch("affine_act_bwd");
}

int tfswa_lnhat_bwd(const void* da, int64_t ldd, int64_t d_bs, const void* x, int64_t ldx, int64_t x_bs, const float* stats,
                    int64_t st_bs, void* dx, int64_t ldo, int64_t o_bs, int64_t M, int32_t K, int32_t batch, int32_t dtype,
                    void* stream) {
  TFSWA_REQUIRE(da && x && stats && dx && M > 0 && batch > 0, "lnhat_bwd: bad arguments");
  TFSWA_REQUIRE(K % 8 == 0 && K <= 1024 && ldd % 8 == 0 && ldx % 8 == 0 && ldo % 8 == 0 && d_bs % 8 == 0 && x_bs % 8 == 0 && o_bs % 8 == 0,
                "lnhat_bwd: K/ld alignment");
  const int warps = 8, nvec = K / 8;
  cudaStream_t st = (cudaStream_t)stream;
#define TFSWA_LN(T, LPR)                                                                                               \
  {                                                                                                                    \
    dim3 grid((unsigned)ceil_div64(M, (int64_t)warps * (32 / LPR)), batch);                                            \
    lnhat_bwd_kernel<T, LPR><<<grid, warps * 32, 0, st>>>((const T*)da, ldd, d_bs, (const T*)x, ldx, x_bs, stats, st_bs, \
                                                          (T*)dx, ldo, o_bs, M, K);                                    \
  }
#define TFSWA_LN_T(T)            \
  if (nvec <= 4) TFSWA_LN(T, 4)  \
  else if (nvec <= 8) TFSWA_LN(T, 8) \
  else if (nvec <= 16) TFSWA_LN(T, 16) \
  else TFSWA_LN(T, 32)
  if (dtype == TFSWA_F32) { TFSWA_LN_T(float) }
  else if (dtype == TFSWA_BF16) { TFSWA_LN_T(bf16) }
  else TFSWA_REQUIRE(false, "lnhat_bwd: bad dtype");
#undef TFSWA_LN_T
#undef TFSWA_LN
  return check_launch("lnhat_bwd");
}

int tfswa_sum_batch(const void* g, void* out, int64_t M, int32_t nb, int32_t N, int32_t dtype, void* stream) {
  TFSWA_REQUIRE(g && out && M > 0 && nb > 0 && N % 8 == 0, "sum_batch: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = M * (N / 8);
  if (dtype == TFSWA_F32) sum_batch_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((const float*)g, (float*)out, M, nb, N);
  else if (dtype == TFSWA_BF16) sum_batch_kernel<bf16><<<grid_for(total, 256), 256, 0, st>>>((const bf16*)g, (bf16*)out, M, nb, N);
  else TFSWA_REQUIRE(false, "sum_batch: bad dtype");
  return check_launch("sum_batch");
}

int tfswa_bilinear_bwd(const void* dy, void* dx, int32_t B, int32_t Hin, int32_t Win, int32_t Hout, int32_t Wout, int32_t C,
                       int32_t dtype, void* stream) {
  TFSWA_REQUIRE(dy && dx && B > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0 && C % 8 == 0, "bilinear_bwd: bad arguments");
  const int64_t total = (int64_t)B * Hin * Win * (C / 8);
  const float sh = (float)Hin / (float)Hout, sw = (float)Win / (float)Wout;
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TFSWA_F32) bilinear_bwd_kernel<float><<<grid_for(total, 256), 256, 0, st>>>((const float*)dy, (float*)dx, B, Hin, Win, Hout, Wout, C, sh, sw);
  else if (dtype == TFSWA_BF16) bilinear_bwd_kernel<bf16><<<grid_for(total, 256), 256, 0, st>>>((const bf16*)dy, (bf16*)dx, B, Hin, Win, Hout, Wout, C, sh, sw);
  else TFSWA_REQUIRE(false, "bilinear_bwd: bad dtype");
  return check_launch("bilinear_bwd");
}

}  // extern "C"
