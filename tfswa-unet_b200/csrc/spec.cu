// The spectrogram steps either side of the model in overlap-add separation, HBM-bound:
//   * tfswa_spec_pack_norm: complex STFT (B, F, T) -> model input (B, 2, F, T) = [real | imag] planes
//     (stft_processor.py:186-204 `to_model_input`) with the instance normalisation over time of
//     `SpectrogramNormalizer` (stft_processor.py:283-297: mean and UNBIASED std per (b, channel, frequency) row, + eps) in the
//     same pass: one warp per frequency row, the row is read from HBM once (4 KB at T = 517: the second and third sweep hit L1);
//   * tfswa_spec_mask_apply: masks (B, S, F, T) -> complex stems (B, S, F, T) = spec * (mask * std + mean)
//     (inference.py:132-145: the reference "denormalises" the masks with the INPUT's statistics, channel s for stem s);
//   * tfswa_ola_add: Hann-weighted overlap-add of a batch of reconstructed segments into the output and weight rows
//     (inference.py:209-216), as a gather over output samples so that the per-sample summation order is the reference's
//     (ascending segment index) and no atomics are needed.
//   * tfswa_mrstft_mag_loss (row f4): magnitude + log-magnitude L1 of one resolution of the multi-resolution STFT loss
//     (losses.py:125-141, 171-183) with the gradient with respect to the predicted spectrogram from the same pass.
// STFT / ISTFT themselves stay batched cuFFT (torch.stft / torch.istft).
#include "common.cuh"

namespace tfswa {

// one warp per (b, f) row of T complex values
__global__ void __launch_bounds__(256) spec_pack_norm_kernel(const float2* __restrict__ spec, float* __restrict__ x,
                                                             float* __restrict__ stats, int64_t rows, int F, int T, float eps,
                                                             int normalize) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int64_t b = row / F;
  const int f = (int)(row - b * F);
  const float2* src = spec + row * T;
  float mean_r = 0.f, mean_i = 0.f, std_r = 1.f, std_i = 1.f;
  if (normalize) {
    float sr = 0.f, si = 0.f;
    for (int t = lane; t < T; t += 32) { const float2 v = src[t]; sr += v.x; si += v.y; }
    mean_r = warp_sum(sr) / (float)T; mean_i = warp_sum(si) / (float)T;
    float qr = 0.f, qi = 0.f;
    for (int t = lane; t < T; t += 32) {
      const float2 v = src[t];
      const float dr = v.x - mean_r, di = v.y - mean_i;
      qr += dr * dr; qi += di * di;
    }
    std_r = sqrtf(warp_sum(qr) / (float)(T - 1)) + eps;     // torch.std: unbiased
    std_i = sqrtf(warp_sum(qi) / (float)(T - 1)) + eps;
    if (lane == 0) {
      float* st_r = stats + ((b * 2 + 0) * F + f) * 2;
      float* st_i = stats + ((b * 2 + 1) * F + f) * 2;
      st_r[0] = mean_r; st_r[1] = std_r; st_i[0] = mean_i; st_i[1] = std_i;
    }
  }
  float* xr = x + ((b * 2 + 0) * F + f) * (int64_t)T;
  float* xi = x + ((b * 2 + 1) * F + f) * (int64_t)T;
  for (int t = lane; t < T; t += 32) {
    const float2 v = src[t];
    xr[t] = normalize ? (v.x - mean_r) / std_r : v.x;
    xi[t] = normalize ? (v.y - mean_i) / std_i : v.y;
  }
}

// one warp per (b, s, f) row
__global__ void __launch_bounds__(256) spec_mask_apply_kernel(const float* __restrict__ masks, const float2* __restrict__ spec,
                                                              const float* __restrict__ stats, float2* __restrict__ out,
                                                              int64_t rows, int S, int F, int T, int normalize) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int f = (int)(row % F);
  const int64_t bs = row / F;
  const int s = (int)(bs % S);
  const int64_t b = bs / S;
  float mean = 0.f, sd = 1.f;
  if (normalize) {
    const float* st = stats + ((b * 2 + s) * F + f) * 2;
    mean = st[0]; sd = st[1];
  }
  const float* m = masks + row * T;
  const float2* src = spec + (b * F + f) * (int64_t)T;
  float2* dst = out + row * T;
  for (int t = lane; t < T; t += 32) {
    const float w = normalize ? __fadd_rn(__fmul_rn(m[t], sd), mean) : m[t];   // (mask * std + mean) rounded like the eager ops
    const float2 v = src[t];
    dst[t] = make_float2(v.x * w, v.y * w);
  }
}

// acc rows 0..S-1 += window-weighted samples of the segments covering each output sample, row S += the window weights
__global__ void __launch_bounds__(256) ola_add_kernel(const float* __restrict__ wav, const int64_t* __restrict__ starts,
                                                      const float* __restrict__ win, float* __restrict__ acc, int nseg, int S,
                                                      int64_t L, int64_t seg_len, int64_t total, int64_t t_lo, int64_t t_hi) {
  const int64_t t = t_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y;                 // 0..S
  if (t >= t_hi) return;
  float a = acc[row * total + t];
  bool touched = false;
  for (int j = 0; j < nseg; ++j) {
    const int64_t s0 = starts[j];
    int64_t n = seg_len < total - s0 ? seg_len : total - s0;
    n = n < L ? n : L;
    const int64_t o = t - s0;
    if (o >= 0 && o < n) {
      const float w = win[o];
      a = row < S ? __fadd_rn(a, __fmul_rn(wav[((int64_t)j * S + row) * L + o], w)) : __fadd_rn(a, w);
      touched = true;
    }
  }
  if (touched) acc[row * total + t] = a;
}

// magnitude + log-magnitude L1 of one STFT resolution (losses.py:125-141, 171-183) and its gradient with respect to the
// predicted spectrogram, in one pass over the two complex tensors: |P| and |T| never reach HBM
template <bool GRAD>
__global__ void __launch_bounds__(256) mrstft_mag_loss_kernel(const float2* __restrict__ pred, const float2* __restrict__ target,
                                                              int64_t n, float a, float b, float eps, double* __restrict__ loss,
                                                              float2* __restrict__ grad) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 P = pred[i], T = target[i];
    const float pm = sqrtf(P.x * P.x + P.y * P.y), tm = sqrtf(T.x * T.x + T.y * T.y);
    const float dm = pm - tm;
    const float dl = logf(pm + eps) - logf(tm + eps);
    acc += a * fabsf(dm) + b * fabsf(dl);
    if (GRAD) {
      // d|dm|/dpm = sign(dm), d|dl|/dpm = sign(dl) / (pm + eps); d pm / dP = P / pm (0 at P = 0, like torch.abs)
      const float sm = dm > 0.f ? 1.f : (dm < 0.f ? -1.f : 0.f), sl = dl > 0.f ? 1.f : (dl < 0.f ? -1.f : 0.f);
      const float g = a * sm + b * sl / (pm + eps);
      const float r = pm > 0.f ? g / pm : 0.f;
      grad[i] = make_float2(P.x * r, P.y * r);
    }
  }
  acc = warp_sum(acc);
  __shared__ float part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += (double)part[w];
    atomicAdd(loss, t);
  }
}

}  // namespace tfswa

using namespace tfswa;

extern "C" int tfswa_mrstft_mag_loss(const void* pred_c64, const void* target_c64, int64_t n, float w_mag, float w_log, float eps,
                                     double* loss, void* grad_c64, void* stream) {
  TFSWA_REQUIRE(pred_c64 && target_c64 && loss && n > 0, "mrstft_mag_loss: bad arguments");
  TFSWA_REQUIRE(((((uintptr_t)pred_c64) | ((uintptr_t)target_c64) | ((uintptr_t)grad_c64)) & 7) == 0, "mrstft_mag_loss: complex64 buffers must be 8-byte aligned");
  const float a = w_mag / (float)n, b = w_log / (float)n;        // F.l1_loss: mean over the elements
  int sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t blocks = ceil_div64(n, 256 * 4);
  if (blocks > (int64_t)sms * 8) blocks = (int64_t)sms * 8;    // grid-stride: a multiple of the SM count, 8 CTAs of 256 threads per SM
  cudaStream_t st = (cudaStream_t)stream;
  if (grad_c64) mrstft_mag_loss_kernel<true><<<(unsigned)blocks, 256, 0, st>>>((const float2*)pred_c64, (const float2*)target_c64, n, a, b, eps, loss, (float2*)grad_c64);
  else mrstft_mag_loss_kernel<false><<<(unsigned)blocks, 256, 0, st>>>((const float2*)pred_c64, (const float2*)target_c64, n, a, b, eps, loss, nullptr);
  return check_launch("mrstft_mag_loss");
}

extern "C" int tfswa_spec_pack_norm(const void* spec_c64, float* x, float* stats, int32_t B, int32_t F, int32_t T, float eps,
                                    int32_t normalize, void* stream) {
  TFSWA_REQUIRE(spec_c64 && x && B > 0 && F > 0 && T > 0, "spec_pack_norm: bad arguments");
  TFSWA_REQUIRE(!normalize || stats, "spec_pack_norm: normalisation needs the stats buffer (B, 2, F, 2)");
  TFSWA_REQUIRE((((uintptr_t)spec_c64) & 7) == 0, "spec_pack_norm: complex64 input must be 8-byte aligned");
  const int64_t rows = (int64_t)B * F;
  spec_pack_norm_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, (cudaStream_t)stream>>>((const float2*)spec_c64, x, stats, rows, F, T,
                                                                                        eps, normalize);
  return check_launch("spec_pack_norm");
}

extern "C" int tfswa_spec_mask_apply(const float* masks, const void* spec_c64, const float* stats, void* out_c64, int32_t B,
                                     int32_t S, int32_t F, int32_t T, int32_t normalize, void* stream) {
  TFSWA_REQUIRE(masks && spec_c64 && out_c64 && B > 0 && S > 0 && F > 0 && T > 0, "spec_mask_apply: bad arguments");
  TFSWA_REQUIRE(!normalize || (stats && S <= 2), "spec_mask_apply: denormalisation takes the statistics of input channel s for stem s (S <= 2)");
  TFSWA_REQUIRE(((((uintptr_t)spec_c64) | ((uintptr_t)out_c64)) & 7) == 0, "spec_mask_apply: complex64 buffers must be 8-byte aligned");
  const int64_t rows = (int64_t)B * S * F;
  spec_mask_apply_kernel<<<(unsigned)ceil_div64(rows, 8), 256, 0, (cudaStream_t)stream>>>(masks, (const float2*)spec_c64, stats,
                                                                                         (float2*)out_c64, rows, S, F, T, normalize);
  return check_launch("spec_mask_apply");
}

extern "C" int tfswa_ola_add(const float* wav, const int64_t* starts, int64_t first_start, int64_t last_start, const float* win,
                             float* acc, int32_t nseg, int32_t S, int64_t L, int64_t seg_len, int64_t total, void* stream) {
  TFSWA_REQUIRE(wav && starts && win && acc && nseg > 0 && S > 0 && L > 0 && seg_len > 0 && total > 0, "ola_add: bad arguments");
  TFSWA_REQUIRE(first_start >= 0 && first_start <= last_start && last_start < total, "ola_add: segment starts outside the output");
  const int64_t t_lo = first_start;
  int64_t t_hi = last_start + (seg_len < L ? seg_len : L);
  if (t_hi > total) t_hi = total;
  if (t_hi <= t_lo) return TFSWA_OK;
  dim3 grid((unsigned)ceil_div64(t_hi - t_lo, 256), (unsigned)(S + 1), 1);
  ola_add_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(wav, starts, win, acc, nseg, S, L, seg_len, total, t_lo, t_hi);
  return check_launch("ola_add");
}
