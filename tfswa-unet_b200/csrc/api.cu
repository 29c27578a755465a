// Error plumbing, version, and the small bandwidth-bound kernels (row statistics, BN finalize,
// affine+activation, bilinear resize).
#include "common.cuh"
#include <string.h>

namespace tfswa {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return TFSWA_ECUDA;
  }
  return TFSWA_OK;
}

// ---------------------------------------------------------------------------------------------
// row statistics: one warp per row, K <= 1024, 128-bit loads, shuffle reduction (HBM-bound)
// ---------------------------------------------------------------------------------------------
// LPR lanes cooperate on one row (LPR = K/8 rounded up to a power of two, <= 32), so a warp covers 32/LPR rows and
// every lane issues 128-bit loads even at K = 32.
template <typename T, int LPR>
__global__ void row_stats_kernel(const T* __restrict__ x, int64_t ldx, int64_t x_bs, float* __restrict__ st,
                                 int64_t st_bs, int64_t M, int K) {
  constexpr int RPW = 32 / LPR;                 // rows per warp
  constexpr int VPL = LPR == 32 ? 4 : 1;        // 8-element vectors per lane (K <= 1024)
  const int warps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, l = lane % LPR;
  const int64_t row = ((int64_t)blockIdx.x * warps + (threadIdx.x >> 5)) * RPW + sub;
  const bool ok = row < M;
  const T* p = x + (int64_t)blockIdx.y * x_bs + (ok ? row : 0) * ldx;
  float v[VPL][8];
  float s = 0.f;
  const int nvec = K >> 3;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = l + i * LPR;
    if (ok && vi < nvec) {
      load8(p + vi * 8, v[i]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[i][j];
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)K;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int vi = l + i * LPR;
    if (ok && vi < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { const float d = v[i][j] - mean; q += d * d; }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  if (ok && l == 0) {
    float* o = st + (int64_t)blockIdx.y * st_bs + row * 2;
    *reinterpret_cast<float2*>(o) = make_float2(mean, rsqrtf(q / (float)K + 1e-5f));
  }
}

template <typename T>
static void launch_row_stats(const T* x, int64_t ldx, int64_t x_bs, float* st, int64_t st_bs, int64_t M, int K, int batch,
                             cudaStream_t s) {
  const int warps = 8;
  const int nvec = K / 8;
#define TFSWA_RS(LPR)                                                                                   \
  {                                                                                                     \
    dim3 grid((unsigned)ceil_div64(M, (int64_t)warps * (32 / LPR)), batch);                             \
    row_stats_kernel<T, LPR><<<grid, warps * 32, 0, s>>>(x, ldx, x_bs, st, st_bs, M, K);                \
  }
  if (nvec <= 4) TFSWA_RS(4)
  else if (nvec <= 8) TFSWA_RS(8)
  else if (nvec <= 16) TFSWA_RS(16)
  else TFSWA_RS(32)
#undef TFSWA_RS
}

// ---------------------------------------------------------------------------------------------
// BN finalize (C <= 1024 threads in one block)
// ---------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const float* __restrict__ cs, double inv_count, double unbias,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ rmean, float* __restrict__ rvar, float momentum, float eps,
                                   float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ save, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double mean = (double)cs[c] * inv_count;
  double var = (double)cs[C + c] * inv_count - mean * mean;
  if (var < 0.0) var = 0.0;
  const float rstd = (float)(1.0 / sqrt(var + (double)eps));
  const float sc = gamma[c] * rstd;
  scale[c] = sc;
  shift[c] = beta[c] - (float)mean * sc;
  if (save) { save[c] = (float)mean; save[C + c] = rstd; }
  if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)mean;
  if (rvar) rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)(var * unbias);
}

// ---------------------------------------------------------------------------------------------
// y = act(v*scale + shift) + r1 + r2   (8 elements per thread, 128-bit accesses for bf16)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void affine_act_kernel(const T* __restrict__ v, const float* __restrict__ scale, const float* __restrict__ shift,
                                  const T* __restrict__ r1, const T* __restrict__ r2, T* __restrict__ y,
                                  int64_t nvec, int C, int epilogue) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = i * 8;
    const int c = (int)(e % C);
    float a[8];
    load8(v + e, a);
    if (scale) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = a[j] * scale[c + j] + shift[c + j];
    }
    if (epilogue == TFSWA_EPI_GELU) {
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] = gelu_for<T>(a[j]);
    }
    if (r1) { float t[8]; load8(r1 + e, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += t[j]; }
    if (r2) { float t[8]; load8(r2 + e, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) a[j] += t[j]; }
    store8(y + e, a);
  }
}

// ---------------------------------------------------------------------------------------------
// bilinear resize (align_corners=False), NHWC; one thread per 8 channels of one output pixel
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void bilinear_src(int o, int in_size, float scale, int& i0, int& i1, float& l1) {
  // PyTorch area_pixel_compute_source_index, align_corners=False, clamped at 0
  float s = ((float)o + 0.5f) * scale - 0.5f;
  if (s < 0.f) s = 0.f;
  i0 = (int)s;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = s - (float)i0;
}

template <typename T>
__global__ void bilinear_kernel(const T* __restrict__ x, T* __restrict__ y, int B, int Hin, int Win, int Hout, int Wout,
                                int C, float sh, float sw) {
  const int cv = C >> 3;
  const int64_t total = (int64_t)B * Hout * Wout * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    int64_t t = i / cv;
    const int ox = (int)(t % Wout); t /= Wout;
    const int oy = (int)(t % Hout);
    const int b = (int)(t / Hout);
    int y0, y1, x0, x1; float ly, lx;
    bilinear_src(oy, Hin, sh, y0, y1, ly);
    bilinear_src(ox, Win, sw, x0, x1, lx);
    const T* base = x + (int64_t)b * Hin * Win * C + c8 * 8;
    float a[8], bb[8], c_[8], d[8], o[8];
    load8(base + ((int64_t)y0 * Win + x0) * C, a);
    load8(base + ((int64_t)y0 * Win + x1) * C, bb);
    load8(base + ((int64_t)y1 * Win + x0) * C, c_);
    load8(base + ((int64_t)y1 * Win + x1) * C, d);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = w00 * a[j] + w01 * bb[j] + w10 * c_[j] + w11 * d[j];
    store8(y + (((int64_t)b * Hout + oy) * Wout + ox) * C + c8 * 8, o);
  }
}

}  // namespace tfswa

using namespace tfswa;

extern "C" {

const char* tfswa_last_error(void) { return g_err; }
const char* tfswa_version(void) { return "tfswa_b200 0.1 (sm_100a)"; }

int tfswa_device_supported(void) {
  cudaFuncAttributes fa;
  cudaError_t e = cudaFuncGetAttributes(&fa, (const void*)bn_finalize_kernel);
  if (e != cudaSuccess) {
    set_error("no kernel image for this device: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return 0;
  }
  return 1;
}

int tfswa_row_stats(const void* x, int64_t ldx, int64_t x_bs, float* stats, int64_t st_bs, int64_t M, int32_t K,
                    int32_t batch, int32_t dtype, void* stream) {
  TFSWA_REQUIRE(x && stats && M > 0 && batch > 0, "row_stats: null pointer or empty problem");
  TFSWA_REQUIRE(K % 8 == 0 && K >= 8 && K <= 1024, "row_stats: K=%d must be a multiple of 8 in [8,1024]", K);
  TFSWA_REQUIRE(ldx % 8 == 0 && x_bs % 8 == 0, "row_stats: ldx/x_bs must be multiples of 8 elements");
  if (dtype == TFSWA_F32) launch_row_stats<float>((const float*)x, ldx, x_bs, stats, st_bs, M, K, batch, (cudaStream_t)stream);
  else if (dtype == TFSWA_BF16) launch_row_stats<bf16>((const bf16*)x, ldx, x_bs, stats, st_bs, M, K, batch, (cudaStream_t)stream);
  else TFSWA_REQUIRE(false, "row_stats: bad dtype %d", dtype);
  return check_launch("row_stats");
}

int tfswa_bn_finalize(const float* col_stats, int64_t count, const float* gamma, const float* beta, float* running_mean,
                      float* running_var, float momentum, float eps, float* scale, float* shift, float* save_mean_rstd,
                      int32_t C, void* stream) {
  TFSWA_REQUIRE(col_stats && gamma && beta && scale && shift && count > 0 && C > 0, "bn_finalize: bad arguments");
  const double unbias = count > 1 ? (double)count / (double)(count - 1) : 1.0;
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, (cudaStream_t)stream>>>(col_stats, 1.0 / (double)count, unbias, gamma, beta,
                                                                      running_mean, running_var, momentum, eps, scale, shift,
                                                                      save_mean_rstd, C);
  return check_launch("bn_finalize");
}

int tfswa_affine_act(const void* v, const float* scale, const float* shift, const void* r1, const void* r2, void* y,
                     int64_t M, int32_t C, int32_t epilogue, int32_t dtype, void* stream) {
  TFSWA_REQUIRE(v && y && M > 0, "affine_act: bad arguments");
  TFSWA_REQUIRE(C % 8 == 0, "affine_act: C=%d must be a multiple of 8", C);
  TFSWA_REQUIRE((scale == nullptr) == (shift == nullptr), "affine_act: scale and shift must both be given or both NULL");
  const int64_t nvec = M * C / 8;
  const int threads = 256;
  const unsigned blocks = (unsigned)((nvec + threads - 1) / threads < 148 * 16 ? (nvec + threads - 1) / threads : 148 * 16);
  if (dtype == TFSWA_F32)
    affine_act_kernel<float><<<blocks, threads, 0, (cudaStream_t)stream>>>((const float*)v, scale, shift, (const float*)r1,
                                                                          (const float*)r2, (float*)y, nvec, C, epilogue);
  else if (dtype == TFSWA_BF16)
    affine_act_kernel<bf16><<<blocks, threads, 0, (cudaStream_t)stream>>>((const bf16*)v, scale, shift, (const bf16*)r1,
                                                                         (const bf16*)r2, (bf16*)y, nvec, C, epilogue);
  else TFSWA_REQUIRE(false, "affine_act: bad dtype %d", dtype);
  return check_launch("affine_act");
}

int tfswa_bilinear_fwd(const void* x, void* y, int32_t B, int32_t Hin, int32_t Win, int32_t Hout, int32_t Wout, int32_t C,
                       int32_t dtype, void* stream) {
  TFSWA_REQUIRE(x && y && B > 0 && Hin > 0 && Win > 0 && Hout > 0 && Wout > 0, "bilinear: bad arguments");
  TFSWA_REQUIRE(C % 8 == 0, "bilinear: C=%d must be a multiple of 8", C);
  const int64_t total = (int64_t)B * Hout * Wout * (C / 8);
  const int threads = 256;
  const unsigned blocks = (unsigned)((total + threads - 1) / threads < 148 * 16 ? (total + threads - 1) / threads : 148 * 16);
  const float sh = (float)Hin / (float)Hout, sw = (float)Win / (float)Wout;
  if (dtype == TFSWA_F32)
    bilinear_kernel<float><<<blocks, threads, 0, (cudaStream_t)stream>>>((const float*)x, (float*)y, B, Hin, Win, Hout, Wout, C, sh, sw);
  else if (dtype == TFSWA_BF16)
    bilinear_kernel<bf16><<<blocks, threads, 0, (cudaStream_t)stream>>>((const bf16*)x, (bf16*)y, B, Hin, Win, Hout, Wout, C, sh, sw);
  else TFSWA_REQUIRE(false, "bilinear: bad dtype %d", dtype);
  return check_launch("bilinear");
}

}  // extern "C"
