// Network boundary kernels (bandwidth-bound): the 7x7 stem conv reads the NCHW fp32 model input and
// writes NHWC activations (tfswa_unet.py:58-62); the head tail reads NHWC activations and writes the
// NCHW fp32 sigmoid masks (tfswa_unet.py:141-144).  The NCHW<->NHWC layout change therefore never
// costs a pass of its own.
#include "common.cuh"

namespace tfswa {

constexpr int ST_TH = 8, ST_TW = 32;        // output pixels per CTA (one per thread)
constexpr int ST_CI = 4;                    // input channels staged per pass

template <typename T>
__global__ void __launch_bounds__(ST_TH * ST_TW) stem_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                             const float* __restrict__ bias, T* __restrict__ y,
                                                             T* __restrict__ pre, float* __restrict__ col_stats,
                                                             int B, int Cin, int H, int W, int Cout, int epilogue) {
  __shared__ float tile[ST_CI][ST_TH + 6][ST_TW + 6 + 1];
  __shared__ __align__(16) float sw[ST_CI][49][32];
  __shared__ float s_stats[2][32];
  const int tx = threadIdx.x % ST_TW, ty = threadIdx.x / ST_TW;
  const int tiles_w = (W + ST_TW - 1) / ST_TW;
  const int ox0 = (blockIdx.x % tiles_w) * ST_TW, oy0 = (blockIdx.x / tiles_w) * ST_TH;
  const int b = blockIdx.y;
  const int co0 = blockIdx.z * 32;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += ST_CI) {
    const int nc = min(ST_CI, Cin - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < nc * (ST_TH + 6) * (ST_TW + 6); i += blockDim.x) {
      const int c = i / ((ST_TH + 6) * (ST_TW + 6));
      const int r = i % ((ST_TH + 6) * (ST_TW + 6));
      const int yy = r / (ST_TW + 6), xx = r % (ST_TW + 6);
      const int iy = oy0 + yy - 3, ix = ox0 + xx - 3;
      float v = 0.f;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[(((int64_t)b * Cin + c0 + c) * H + iy) * W + ix];
      tile[c][yy][xx] = v;
    }
    for (int i = threadIdx.x; i < nc * 49 * 32; i += blockDim.x) {
      const int c = i / (49 * 32);
      const int r = i % (49 * 32);
      const int tap = r / 32, co = r % 32;
      float v = 0.f;
      if (co0 + co < Cout) v = w[(((int64_t)(co0 + co)) * Cin + c0 + c) * 49 + tap];
      sw[c][tap][co] = v;
    }
    __syncthreads();
    for (int c = 0; c < nc; ++c) {
#pragma unroll
      for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
        for (int kx = 0; kx < 7; ++kx) {
          const float v = tile[c][ty + ky][tx + kx];
          const float4* wp = reinterpret_cast<const float4*>(&sw[c][ky * 7 + kx][0]);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 ww = wp[q];
            acc[q * 4 + 0] = fmaf(v, ww.x, acc[q * 4 + 0]); acc[q * 4 + 1] = fmaf(v, ww.y, acc[q * 4 + 1]);
            acc[q * 4 + 2] = fmaf(v, ww.z, acc[q * 4 + 2]); acc[q * 4 + 3] = fmaf(v, ww.w, acc[q * 4 + 3]);
          }
        }
      }
    }
  }

  const int oy = oy0 + ty, ox = ox0 + tx;
  const bool ok = oy < H && ox < W;
  const int nco = min(32, Cout - co0);
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] += (i < nco ? bias[co0 + i] : 0.f);
  if (col_stats) {
    if (threadIdx.x < 64) (&s_stats[0][0])[threadIdx.x] = 0.f;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float v = ok ? acc[i] : 0.f;
      const float s = warp_sum(v), q = warp_sum(v * v);
      if ((threadIdx.x & 31) == 0) { atomicAdd(&s_stats[0][i], s); atomicAdd(&s_stats[1][i], q); }
    }
    __syncthreads();
    if (threadIdx.x < 64) {
      const int which = threadIdx.x / 32, c = threadIdx.x % 32;
      if (c < nco) atomicAdd(col_stats + which * Cout + co0 + c, s_stats[which][c]);
    }
  }
  if (!ok) return;
  const int64_t o = (((int64_t)b * H + oy) * W + ox) * Cout + co0;
  if (pre) {
#pragma unroll
    for (int i = 0; i < 32; i += 8) if (i < nco) { float t[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] = acc[i + j];
      store8(pre + o + i, t); }
  }
  if (epilogue == TFSWA_EPI_GELU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i] = gelu_erf(acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 32; i += 8) if (i < nco) { float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = acc[i + j];
    store8(y + o + i, t); }
}

// head tail: one thread per pixel; C <= 256 channels read as 8-element vectors; Cout <= 8
template <typename T>
__global__ void __launch_bounds__(256) head_tail_kernel(const T* __restrict__ v, const float* __restrict__ scale,
                                                        const float* __restrict__ shift, const float* __restrict__ w3,
                                                        const float* __restrict__ b3, float* __restrict__ masks,
                                                        float* __restrict__ logits, int64_t M, int HW, int C, int Cout) {
  extern __shared__ float sm[];                 // w3 (Cout*C) | scale (C) | shift (C)
  float* s_w = sm; float* s_sc = sm + Cout * C; float* s_sh = s_sc + C;
  for (int i = threadIdx.x; i < Cout * C; i += blockDim.x) s_w[i] = w3[i];
  for (int i = threadIdx.x; i < C; i += blockDim.x) { s_sc[i] = scale ? scale[i] : 1.f; s_sh[i] = shift ? shift[i] : 0.f; }
  __syncthreads();
  const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float acc[8];
#pragma unroll
  for (int o = 0; o < 8; ++o) acc[o] = (o < Cout) ? b3[o] : 0.f;
  const T* row = v + m * C;
  for (int c = 0; c < C; c += 8) {
    float t[8]; load8(row + c, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float u = gelu_erf(t[j] * s_sc[c + j] + s_sh[c + j]);
#pragma unroll
      for (int o = 0; o < 8; ++o) if (o < Cout) acc[o] = fmaf(u, s_w[o * C + c + j], acc[o]);
    }
  }
  const int64_t b = m / HW, pix = m % HW;
#pragma unroll
  for (int o = 0; o < 8; ++o) {
    if (o < Cout) {
      const int64_t off = (b * Cout + o) * HW + pix;
      if (logits) logits[off] = acc[o];
      masks[off] = 1.0f / (1.0f + __expf(-acc[o]));
    }
  }
}

}  // namespace tfswa

using namespace tfswa;

extern "C" {

int tfswa_stem_fwd(const float* x_nchw, const float* w, const float* bias, void* y, void* pre, float* col_stats, int32_t B,
                   int32_t Cin, int32_t H, int32_t W, int32_t Cout, int32_t epilogue, int32_t dtype, void* stream) {
  TFSWA_REQUIRE(x_nchw && w && bias && y, "stem: null pointer");
  TFSWA_REQUIRE(B > 0 && Cin > 0 && H > 0 && W > 0 && Cout > 0 && Cout % 8 == 0, "stem: bad shape (Cout must be a multiple of 8)");
  TFSWA_REQUIRE(B <= 65535, "stem: batch too large");
  dim3 grid((unsigned)(((W + ST_TW - 1) / ST_TW) * ((H + ST_TH - 1) / ST_TH)), B, (Cout + 31) / 32);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TFSWA_F32)
    stem_kernel<float><<<grid, ST_TH * ST_TW, 0, st>>>(x_nchw, w, bias, (float*)y, (float*)pre, col_stats, B, Cin, H, W, Cout, epilogue);
  else if (dtype == TFSWA_BF16)
    stem_kernel<bf16><<<grid, ST_TH * ST_TW, 0, st>>>(x_nchw, w, bias, (bf16*)y, (bf16*)pre, col_stats, B, Cin, H, W, Cout, epilogue);
  else TFSWA_REQUIRE(false, "stem: bad dtype %d", dtype);
  return check_launch("stem");
}

int tfswa_head_tail_fwd(const void* v, const float* scale, const float* shift, const float* w3, const float* b3,
                        float* masks_nchw, float* logits_nchw, int32_t B, int32_t H, int32_t W, int32_t C, int32_t Cout,
                        int32_t dtype, void* stream) {
  TFSWA_REQUIRE(v && w3 && b3 && masks_nchw, "head_tail: null pointer");
  TFSWA_REQUIRE(C % 8 == 0 && C <= 256 && Cout >= 1 && Cout <= 8, "head_tail: need C%%8==0, C<=256, 1<=Cout<=8");
  TFSWA_REQUIRE((scale == nullptr) == (shift == nullptr), "head_tail: scale/shift must both be given or both NULL");
  const int64_t M = (int64_t)B * H * W;
  const size_t smem = (size_t)(Cout * C + 2 * C) * sizeof(float);
  dim3 grid((unsigned)ceil_div64(M, 256));
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TFSWA_F32)
    head_tail_kernel<float><<<grid, 256, smem, st>>>((const float*)v, scale, shift, w3, b3, masks_nchw, logits_nchw, M, H * W, C, Cout);
  else if (dtype == TFSWA_BF16)
    head_tail_kernel<bf16><<<grid, 256, smem, st>>>((const bf16*)v, scale, shift, w3, b3, masks_nchw, logits_nchw, M, H * W, C, Cout);
  else TFSWA_REQUIRE(false, "head_tail: bad dtype %d", dtype);
  return check_launch("head_tail");
}

}  // extern "C"
