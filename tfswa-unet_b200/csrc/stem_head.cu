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
    for (int i = 0; i < 32; ++i) acc[i] = gelu_for<T>(acc[i]);
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
      const float u = gelu_for<T>(t[j] * s_sc[c + j] + s_sh[c + j]);
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


// ------------------------------------------------------------------------------------------------
// stem backward.  Data gradient (only needed when the network input requires grad): one thread per input pixel,
//   dx[b,ci,y,x] = sum_{co,ky,kx} g[b, y-ky+3, x-kx+3, co] * w[co,ci,ky,kx]      (NHWC g -> NCHW fp32 dx)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) stem_dgrad_kernel(const T* __restrict__ g, const float* __restrict__ w,
                                                         float* __restrict__ dx, int B, int Cin, int H, int W, int Cout) {
  extern __shared__ float s_w[];              // [Cin][49][Cout]
  for (int i = threadIdx.x; i < Cin * 49 * Cout; i += blockDim.x) {
    const int co = i % Cout, tap = (i / Cout) % 49, ci = i / (Cout * 49);
    s_w[i] = w[((int64_t)co * Cin + ci) * 49 + tap];
  }
  __syncthreads();
  const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (int64_t)B * H * W) return;
  const int x = (int)(pix % W), y = (int)((pix / W) % H), b = (int)(pix / ((int64_t)W * H));
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int ky = 0; ky < 7; ++ky) {
    const int oy = y - ky + 3;
    if (oy < 0 || oy >= H) continue;
    for (int kx = 0; kx < 7; ++kx) {
      const int ox = x - kx + 3;
      if (ox < 0 || ox >= W) continue;
      const T* gp = g + (((int64_t)b * H + oy) * W + ox) * Cout;
      const int tap = ky * 7 + kx;
      for (int c8 = 0; c8 < Cout; c8 += 8) {
        float t[8]; load8(gp + c8, t);
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int ci = 0; ci < 4; ++ci)
            if (ci < Cin) acc[ci] = fmaf(t[j], s_w[(ci * 49 + tap) * Cout + c8 + j], acc[ci]);
      }
    }
  }
  for (int ci = 0; ci < Cin && ci < 4; ++ci) dx[(((int64_t)b * Cin + ci) * H + y) * W + x] = acc[ci];
}

// stem weight gradient: persistent CTAs walk 8x32 output tiles; thread t owns outputs o = t, t+256, ... with
// o -> (co = o % 32, (ci, tap) = o / 32), accumulated in registers across tiles, one atomic per output at the end.
constexpr int SW_MAXO = 25;   // ceil(32 * 4 * 49 / 256)
template <typename T>
__global__ void __launch_bounds__(ST_TH * ST_TW) stem_wgrad_kernel(const float* __restrict__ x, const T* __restrict__ g,
                                                                   float* __restrict__ dw, float* __restrict__ dbias,
                                                                   int B, int Cin, int H, int W, int Cout, int co0) {
  __shared__ float tile[ST_CI][ST_TH + 6][ST_TW + 6 + 1];
  __shared__ float gs[ST_TH * ST_TW][33];
  const int tiles_w = (W + ST_TW - 1) / ST_TW, tiles_h = (H + ST_TH - 1) / ST_TH;
  const int64_t ntiles = (int64_t)B * tiles_w * tiles_h;
  const int nout = 32 * Cin * 49;
  float acc[SW_MAXO];
#pragma unroll
  for (int i = 0; i < SW_MAXO; ++i) acc[i] = 0.f;
  float bacc = 0.f;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int b = (int)(t / (tiles_w * tiles_h));
    const int r = (int)(t % (tiles_w * tiles_h));
    const int ox0 = (r % tiles_w) * ST_TW, oy0 = (r / tiles_w) * ST_TH;
    __syncthreads();
    for (int i = threadIdx.x; i < Cin * (ST_TH + 6) * (ST_TW + 6); i += blockDim.x) {
      const int c = i / ((ST_TH + 6) * (ST_TW + 6));
      const int rr = i % ((ST_TH + 6) * (ST_TW + 6));
      const int yy = rr / (ST_TW + 6), xx = rr % (ST_TW + 6);
      const int iy = oy0 + yy - 3, ix = ox0 + xx - 3;
      tile[c][yy][xx] = (iy >= 0 && iy < H && ix >= 0 && ix < W) ? x[(((int64_t)b * Cin + c) * H + iy) * W + ix] : 0.f;
    }
    {
      const int tx = threadIdx.x % ST_TW, ty = threadIdx.x / ST_TW;
      const int oy = oy0 + ty, ox = ox0 + tx;
      const bool ok = oy < H && ox < W;
      for (int c8 = 0; c8 < 32; c8 += 8) {
        float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (ok && co0 + c8 < Cout) load8(g + (((int64_t)b * H + oy) * W + ox) * Cout + co0 + c8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) gs[threadIdx.x][c8 + j] = v[j];
      }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < SW_MAXO; ++i) {
      const int o = threadIdx.x + i * 256;
      if (o < nout) {
        const int co = o & 31, ct = o >> 5, ci = ct / 49, tap = ct % 49, ky = tap / 7, kx = tap % 7;
        float a = 0.f;
        for (int pp = 0; pp < ST_TH * ST_TW; ++pp) a = fmaf(gs[pp][co], tile[ci][pp / ST_TW + ky][pp % ST_TW + kx], a);
        acc[i] += a;
      }
    }
    if (threadIdx.x < 32) {
      float a = 0.f;
      for (int pp = 0; pp < ST_TH * ST_TW; ++pp) a += gs[pp][threadIdx.x];
      bacc += a;
    }
  }
#pragma unroll
  for (int i = 0; i < SW_MAXO; ++i) {
    const int o = threadIdx.x + i * 256;
    if (o < nout) {
      const int co = o & 31, ct = o >> 5;
      if (co0 + co < Cout) atomicAdd(dw + (int64_t)(co0 + co) * Cin * 49 + ct, acc[i]);
    }
  }
  if (threadIdx.x < 32 && co0 + threadIdx.x < Cout) atomicAdd(dbias + co0 + threadIdx.x, bacc);
}

// ------------------------------------------------------------------------------------------------
// head tail backward: per pixel, recompute u = GELU(v*scale+shift) and the logits, then
//   dl = dmask*sigma*(1-sigma) (+dlogit);  du = W3^T dl;  dv = du*gelu'(t)*scale;
//   dW3 += dl u^T, db3 += dl, dscale += du*gelu'(t)*v, dshift += du*gelu'(t)
// Parameter sums: warp shuffle -> shared accumulators -> one global atomic per value per CTA.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) head_tail_bwd_kernel(const T* __restrict__ v, const float* __restrict__ scale,
                                                            const float* __restrict__ shift, const float* __restrict__ w3,
                                                            const float* __restrict__ b3, const float* __restrict__ dmask,
                                                            const float* __restrict__ dlogit, T* __restrict__ dv,
                                                            float* __restrict__ dw3, float* __restrict__ db3,
                                                            float* __restrict__ dscale, float* __restrict__ dshift,
                                                            int64_t M, int HW, int C, int Cout) {
  extern __shared__ float sm[];      // w3 (Cout*C) | scale (C) | shift (C) | acc_w3 (Cout*C) | acc_sc (C) | acc_sh (C) | acc_b (8)
  float* s_w = sm; float* s_sc = s_w + Cout * C; float* s_sh = s_sc + C;
  float* a_w = s_sh + C; float* a_sc = a_w + Cout * C; float* a_sh = a_sc + C; float* a_b = a_sh + C;
  for (int i = threadIdx.x; i < Cout * C; i += blockDim.x) { s_w[i] = w3[i]; a_w[i] = 0.f; }
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s_sc[i] = scale ? scale[i] : 1.f; s_sh[i] = shift ? shift[i] : 0.f; a_sc[i] = 0.f; a_sh[i] = 0.f;
  }
  if (threadIdx.x < 8) a_b[threadIdx.x] = 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t iters = (M + stride - 1) / stride;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t m = it * stride + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool ok = m < M;
    const T* row = v + (ok ? m : 0) * C;
    float logit[8];
#pragma unroll
    for (int o = 0; o < 8; ++o) logit[o] = (o < Cout) ? b3[o] : 0.f;
    for (int c = 0; c < C; c += 8) {
      float t[8]; load8(row + c, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float u = gelu_for<T>(t[j] * s_sc[c + j] + s_sh[c + j]);
#pragma unroll
        for (int o = 0; o < 8; ++o) if (o < Cout) logit[o] = fmaf(u, s_w[o * C + c + j], logit[o]);
      }
    }
    float dl[8];
    const int64_t b = (ok ? m : 0) / HW, pix = (ok ? m : 0) % HW;
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      dl[o] = 0.f;
      if (o < Cout && ok) {
        const int64_t off = (b * Cout + o) * HW + pix;
        const float sg = 1.0f / (1.0f + __expf(-logit[o]));
        dl[o] = (dmask ? dmask[off] * sg * (1.f - sg) : 0.f) + (dlogit ? dlogit[off] : 0.f);
      }
    }
#pragma unroll
    for (int o = 0; o < 8; ++o) {
      if (o < Cout) { const float s = warp_sum(dl[o]); if (lane == 0) atomicAdd(&a_b[o], s); }
    }
    for (int c = 0; c < C; c += 8) {
      float t[8], out[8]; load8(row + c, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float tt = t[j] * s_sc[c + j] + s_sh[c + j];
        const float u = gelu_for<T>(tt);
        float du = 0.f;
#pragma unroll
        for (int o = 0; o < 8; ++o) if (o < Cout) du = fmaf(dl[o], s_w[o * C + c + j], du);
        const float gt = ok ? du * gelu_grad_for<T>(tt) : 0.f;
        out[j] = gt * s_sc[c + j];
        const float r1 = warp_sum(gt * t[j]), r2 = warp_sum(gt);
        if (lane == 0) { atomicAdd(&a_sc[c + j], r1); atomicAdd(&a_sh[c + j], r2); }
#pragma unroll
        for (int o = 0; o < 8; ++o) {
          if (o < Cout) { const float r = warp_sum(dl[o] * u); if (lane == 0) atomicAdd(&a_w[o * C + c + j], r); }
        }
      }
      if (ok) store8(dv + m * C + c, out);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Cout * C; i += blockDim.x) atomicAdd(dw3 + i, a_w[i]);
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    if (dscale) { atomicAdd(dscale + i, a_sc[i]); atomicAdd(dshift + i, a_sh[i]); }
  }
  if (threadIdx.x < Cout) atomicAdd(db3 + threadIdx.x, a_b[threadIdx.x]);
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


int tfswa_stem_bwd(const float* x_nchw, const float* w, const void* g, float* dx_nchw, float* dw, float* dbias, int32_t B,
                   int32_t Cin, int32_t H, int32_t W, int32_t Cout, int32_t dtype, void* stream) {
  TFSWA_REQUIRE(x_nchw && w && g && dw && dbias, "stem_bwd: null pointer");
  TFSWA_REQUIRE(Cin >= 1 && Cin <= 4 && Cout % 8 == 0 && Cout <= 64, "stem_bwd: need 1<=Cin<=4, Cout%%8==0, Cout<=64");
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = B * ((W + ST_TW - 1) / ST_TW) * ((H + ST_TH - 1) / ST_TH);
  const unsigned grid_w = (unsigned)(tiles < 148 * 4 ? tiles : 148 * 4);
  const int64_t npix = (int64_t)B * H * W;
  const size_t smem_d = (size_t)Cin * 49 * Cout * sizeof(float);
#define TFSWA_STEM_BWD(T)                                                                                                   \
  for (int co0 = 0; co0 < Cout; co0 += 32)                                                                                  \
    stem_wgrad_kernel<T><<<grid_w, ST_TH * ST_TW, 0, st>>>(x_nchw, (const T*)g, dw, dbias, B, Cin, H, W, Cout, co0);         \
  if (dx_nchw) stem_dgrad_kernel<T><<<(unsigned)ceil_div64(npix, 256), 256, smem_d, st>>>((const T*)g, w, dx_nchw, B, Cin, H, W, Cout);
  if (dtype == TFSWA_F32) { TFSWA_STEM_BWD(float) }
  else if (dtype == TFSWA_BF16) { TFSWA_STEM_BWD(bf16) }
  else TFSWA_REQUIRE(false, "stem_bwd: bad dtype %d", dtype);
#undef TFSWA_STEM_BWD
  return check_launch("stem_bwd");
}

int tfswa_head_tail_bwd(const void* v, const float* scale, const float* shift, const float* w3, const float* b3,
                        const float* dmasks_nchw, const float* dlogits_nchw, void* dv, float* dw3, float* db3, float* dscale,
                        float* dshift, int32_t B, int32_t H, int32_t W, int32_t C, int32_t Cout, int32_t dtype, void* stream) {
  TFSWA_REQUIRE(v && w3 && b3 && dv && dw3 && db3 && (dmasks_nchw || dlogits_nchw), "head_tail_bwd: null pointer");
  TFSWA_REQUIRE(C % 8 == 0 && C <= 256 && Cout >= 1 && Cout <= 8, "head_tail_bwd: need C%%8==0, C<=256, 1<=Cout<=8");
  TFSWA_REQUIRE((scale == nullptr) == (shift == nullptr) && (dscale == nullptr) == (dshift == nullptr), "head_tail_bwd: scale/shift pairing");
  const int64_t M = (int64_t)B * H * W;
  const size_t smem = (size_t)(2 * Cout * C + 4 * C + 8) * sizeof(float);
  const int64_t want = ceil_div64(M, 256);
  const unsigned grid = (unsigned)(want < 148 * 8 ? want : 148 * 8);
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == TFSWA_F32)
    head_tail_bwd_kernel<float><<<grid, 256, smem, st>>>((const float*)v, scale, shift, w3, b3, dmasks_nchw, dlogits_nchw, (float*)dv,
                                                         dw3, db3, dscale, dshift, M, H * W, C, Cout);
  else if (dtype == TFSWA_BF16)
    head_tail_bwd_kernel<bf16><<<grid, 256, smem, st>>>((const bf16*)v, scale, shift, w3, b3, dmasks_nchw, dlogits_nchw, (bf16*)dv,
                                                        dw3, db3, dscale, dshift, M, H * W, C, Cout);
  else TFSWA_REQUIRE(false, "head_tail_bwd: bad dtype %d", dtype);
  return check_launch("head_tail_bwd");
}

}  // extern "C"
