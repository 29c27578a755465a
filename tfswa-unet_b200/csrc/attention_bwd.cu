// Attention backward for the three TFSWA geometries (autograd of attention.py:70-85 through the same index maps
// as the forward).  Flash-style: the N x N weights are never stored; they are recomputed from q, k and the saved
// log-sum-exp.  Two kernels, no atomics on the activation gradients:
//   dq kernel  : one thread per (query, head), streams K/V tiles;   also writes D_i = dO_i . O_i per head
//   dkv kernel : one thread per (key, head),  streams Q/dO tiles;   zero-padded SW-MSA keys reduce into dpad
//                (the gradient of the folded qkv bias that stands in for LN(0)=beta tokens).
// With s = scale * q.k, p = softmax(s):  ds = p * (dO.v - D),  dq = scale * sum_j ds k_j,
// dk = scale * sum_i ds q_i,  dv = sum_i p dO_i.
#include "attn_common.cuh"
#include <stdlib.h>

namespace tfswa {

template <typename T, int D>
__device__ __forceinline__ void load_vec(const T* p, float (&v)[D]) {
  if (D == 4) { float t[4]; load4(p, t);
#pragma unroll
    for (int d = 0; d < 4; ++d) v[d] = t[d];
  } else {
#pragma unroll
    for (int d8 = 0; d8 < D / 8; ++d8) { float t[8]; load8(p + d8 * 8, t);
#pragma unroll
      for (int d = 0; d < 8; ++d) v[d8 * 8 + d] = t[d]; }
  }
}
template <typename T, int D>
__device__ __forceinline__ void store_vec(T* p, const float (&v)[D]) {
  if (D == 4) { float t[4] = {v[0], v[1], v[2], v[3]}; store4(p, t); }
  else {
#pragma unroll
    for (int d8 = 0; d8 < D / 8; ++d8) { float t[8];
#pragma unroll
      for (int d = 0; d < 8; ++d) t[d] = v[d8 * 8 + d];
      store8(p + d8 * 8, t); }
  }
}

// ------------------------------------------------------------------------------------------------
template <typename T, int D, bool WINDOW>
__global__ void __launch_bounds__(QT * (32 / D)) attn_bwd_dq_kernel(const AttnParams p) {
  constexpr int HG = 32 / D;
  constexpr int NT = QT * HG;
  __shared__ __align__(16) float Ks[KT][32];
  __shared__ __align__(16) float Vs[KT][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int head_local = warp % HG, qsub = warp / HG;
  const int row = blockIdx.x, q0 = blockIdx.y * QT, hg = blockIdx.z;
  const int N = WINDOW ? p.ws * p.ws : (p.geom == TFSWA_GEOM_TSA ? p.H : p.W);
  const int ch0 = hg * 32, head = hg * HG + head_local;
  const T* qkv = (const T*)p.qkv;

  const int qn = q0 + qsub * 32 + lane;
  bool q_valid = qn < N;
  int64_t q_tok = 0;
  if (q_valid) { bool v; q_tok = token_of<WINDOW>(p, row, qn, v); q_valid = v; }
  float q[D], g[D], acc[D];
  float dsum = 0.f, lse = CUDART_INF_F;
#pragma unroll
  for (int d = 0; d < D; ++d) { q[d] = 0.f; g[d] = 0.f; acc[d] = 0.f; }
  if (q_valid) {
    float o[D];
    load_vec<T, D>(qkv + q_tok * p.ldq + ch0 + head_local * D, q);
    load_vec<T, D>((const T*)p.dout + q_tok * p.ldo + ch0 + head_local * D, g);
    load_vec<T, D>((const T*)p.o + q_tok * p.ldo + ch0 + head_local * D, o);
#pragma unroll
    for (int d = 0; d < D; ++d) { dsum += g[d] * o[d]; q[d] *= p.qscale; }
    lse = p.lse[q_tok * p.heads + head];
    p.dsum[q_tok * p.heads + head] = dsum;
  }

  for (int k0 = 0; k0 < N; k0 += KT) {
    const int kcount = min(KT, N - k0);
    if (k0 > 0) __syncthreads();
    for (int v = tid; v < kcount * 8; v += NT) {
      const int j = v >> 3, part = v & 7;
      bool valid; const int64_t tok = token_of<WINDOW>(p, row, k0 + j, valid);
      const int col = (part & 3) * 8;
      float t[8];
      if (valid) load8(qkv + tok * p.ldq + (part < 4 ? p.C : 2 * p.C) + ch0 + col, t);
      else {
        const float* pk = p.pad_kv + (part < 4 ? 0 : p.C) + ch0 + col;
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = pk[e];
      }
      float* dst = (part < 4 ? &Ks[j][col] : &Vs[j][col]);
      *reinterpret_cast<float4*>(dst) = make_float4(t[0], t[1], t[2], t[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(t[4], t[5], t[6], t[7]);
    }
    __syncthreads();
    const int c0 = head_local * D;
    for (int j = 0; j < kcount; ++j) {
      float s = 0.f, dp = 0.f;
      float kk[D];
#pragma unroll
      for (int d4 = 0; d4 < D / 4; ++d4) {
        const float4 k4 = *reinterpret_cast<const float4*>(&Ks[j][c0 + d4 * 4]);
        const float4 v4 = *reinterpret_cast<const float4*>(&Vs[j][c0 + d4 * 4]);
        kk[d4 * 4] = k4.x; kk[d4 * 4 + 1] = k4.y; kk[d4 * 4 + 2] = k4.z; kk[d4 * 4 + 3] = k4.w;
        s = fmaf(q[d4 * 4], k4.x, s); s = fmaf(q[d4 * 4 + 1], k4.y, s); s = fmaf(q[d4 * 4 + 2], k4.z, s); s = fmaf(q[d4 * 4 + 3], k4.w, s);
        dp = fmaf(g[d4 * 4], v4.x, dp); dp = fmaf(g[d4 * 4 + 1], v4.y, dp); dp = fmaf(g[d4 * 4 + 2], v4.z, dp); dp = fmaf(g[d4 * 4 + 3], v4.w, dp);
      }
      const float pw = fast_exp2(s - lse);
      const float ds = pw * (dp - dsum);
#pragma unroll
      for (int d = 0; d < D; ++d) acc[d] = fmaf(ds, kk[d], acc[d]);
    }
  }
  if (q_valid) {
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] *= p.scale;
    store_vec<T, D>((T*)p.dqkv + q_tok * p.ldq + ch0 + head_local * D, acc);
  }
}

// ------------------------------------------------------------------------------------------------
constexpr int BQT = 128;   // queries per shared-memory tile in the dkv kernel

template <typename T, int D, bool WINDOW>
__global__ void __launch_bounds__(QT * (32 / D)) attn_bwd_dkv_kernel(const AttnParams p) {
  constexpr int HG = 32 / D;
  constexpr int NT = QT * HG;
  __shared__ __align__(16) float Qs[BQT][32];
  __shared__ __align__(16) float Gs[BQT][32];
  __shared__ float Ls[BQT][HG], Ds[BQT][HG];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int head_local = warp % HG, ksub = warp / HG;
  const int row = blockIdx.x, kbase = blockIdx.y * QT, hg = blockIdx.z;
  const int N = WINDOW ? p.ws * p.ws : (p.geom == TFSWA_GEOM_TSA ? p.H : p.W);
  const int ch0 = hg * 32, c0 = head_local * D;
  const T* qkv = (const T*)p.qkv;

  const int kn = kbase + ksub * 32 + lane;
  const bool k_in = kn < N;
  bool k_real = false;
  int64_t k_tok = 0;
  if (k_in) k_tok = token_of<WINDOW>(p, row, kn, k_real);
  float k[D], v[D], dk[D], dv[D];
#pragma unroll
  for (int d = 0; d < D; ++d) { k[d] = 0.f; v[d] = 0.f; dk[d] = 0.f; dv[d] = 0.f; }
  if (k_in) {
    if (k_real) {
      load_vec<T, D>(qkv + k_tok * p.ldq + p.C + ch0 + c0, k);
      load_vec<T, D>(qkv + k_tok * p.ldq + 2 * p.C + ch0 + c0, v);
    } else {
#pragma unroll
      for (int d = 0; d < D; ++d) { k[d] = p.pad_kv[ch0 + c0 + d]; v[d] = p.pad_kv[p.C + ch0 + c0 + d]; }
    }
  }

  for (int q0 = 0; q0 < N; q0 += BQT) {
    const int qcount = min(BQT, N - q0);
    if (q0 > 0) __syncthreads();
    for (int e = tid; e < BQT * 8; e += NT) {
      const int i = e >> 3, part = e & 7;
      bool valid = false; int64_t tok = 0;
      if (i < qcount) tok = token_of<WINDOW>(p, row, q0 + i, valid);
      const int col = (part & 3) * 8;
      float t[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (valid) {
        if (part < 4) load8(qkv + tok * p.ldq + ch0 + col, t);
        else load8((const T*)p.dout + tok * p.ldo + ch0 + col, t);
      }
      float* dst = (part < 4 ? &Qs[i][col] : &Gs[i][col]);
      *reinterpret_cast<float4*>(dst) = make_float4(t[0], t[1], t[2], t[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(t[4], t[5], t[6], t[7]);
    }
    for (int e = tid; e < BQT * HG; e += NT) {
      const int i = e / HG, h = e % HG;
      bool valid = false; int64_t tok = 0;
      if (i < qcount) tok = token_of<WINDOW>(p, row, q0 + i, valid);
      Ls[i][h] = valid ? p.lse[tok * p.heads + hg * HG + h] : CUDART_INF_F;    // +inf -> p = 0 for absent queries
      Ds[i][h] = valid ? p.dsum[tok * p.heads + hg * HG + h] : 0.f;
    }
    __syncthreads();
    for (int i = 0; i < qcount; ++i) {
      float s = 0.f, dp = 0.f;
      float qq[D], gg[D];
#pragma unroll
      for (int d4 = 0; d4 < D / 4; ++d4) {
        const float4 q4 = *reinterpret_cast<const float4*>(&Qs[i][c0 + d4 * 4]);
        const float4 g4 = *reinterpret_cast<const float4*>(&Gs[i][c0 + d4 * 4]);
        qq[d4 * 4] = q4.x; qq[d4 * 4 + 1] = q4.y; qq[d4 * 4 + 2] = q4.z; qq[d4 * 4 + 3] = q4.w;
        gg[d4 * 4] = g4.x; gg[d4 * 4 + 1] = g4.y; gg[d4 * 4 + 2] = g4.z; gg[d4 * 4 + 3] = g4.w;
        s = fmaf(q4.x, k[d4 * 4], s); s = fmaf(q4.y, k[d4 * 4 + 1], s); s = fmaf(q4.z, k[d4 * 4 + 2], s); s = fmaf(q4.w, k[d4 * 4 + 3], s);
        dp = fmaf(g4.x, v[d4 * 4], dp); dp = fmaf(g4.y, v[d4 * 4 + 1], dp); dp = fmaf(g4.z, v[d4 * 4 + 2], dp); dp = fmaf(g4.w, v[d4 * 4 + 3], dp);
      }
      const float pw = fast_exp2(s * p.qscale - Ls[i][head_local]);
      const float ds = pw * (dp - Ds[i][head_local]);
#pragma unroll
      for (int d = 0; d < D; ++d) { dv[d] = fmaf(pw, gg[d], dv[d]); dk[d] = fmaf(ds, qq[d], dk[d]); }
    }
  }
  if (k_in) {
#pragma unroll
    for (int d = 0; d < D; ++d) dk[d] *= p.scale;
    if (k_real) {
      store_vec<T, D>((T*)p.dqkv + k_tok * p.ldq + p.C + ch0 + c0, dk);
      store_vec<T, D>((T*)p.dqkv + k_tok * p.ldq + 2 * p.C + ch0 + c0, dv);
    } else if (p.dpad) {
#pragma unroll
      for (int d = 0; d < D; ++d) { atomicAdd(p.dpad + ch0 + c0 + d, dk[d]); atomicAdd(p.dpad + p.C + ch0 + c0 + d, dv[d]); }
    }
  }
}

template <typename T, int D>
static int launch_attn_bwd(const AttnParams& p, cudaStream_t st) {
  constexpr int HG = 32 / D;
  const int hgs = p.C / 32;
  if (p.geom == TFSWA_GEOM_SWA) {
    dim3 grid((unsigned)(p.B * p.nWh * p.nWw), 1, hgs);
    attn_bwd_dq_kernel<T, D, true><<<grid, QT * HG, 0, st>>>(p);
    attn_bwd_dkv_kernel<T, D, true><<<grid, QT * HG, 0, st>>>(p);
  } else {
    const int N = p.geom == TFSWA_GEOM_TSA ? p.H : p.W;
    const int rows = p.geom == TFSWA_GEOM_TSA ? p.B * p.W : p.B * p.H;
    dim3 grid(rows, (N + QT - 1) / QT, hgs);
    attn_bwd_dq_kernel<T, D, false><<<grid, QT * HG, 0, st>>>(p);
    attn_bwd_dkv_kernel<T, D, false><<<grid, QT * HG, 0, st>>>(p);
  }
  return check_launch("attn_bwd");
}

}  // namespace tfswa

using namespace tfswa;

extern "C" int tfswa_attn_bwd(const tfswa_attn_args* a, const void* dout, void* dqkv, float* dsum, float* dpad, void* stream) {
  TFSWA_REQUIRE(a && a->qkv && a->out && a->lse && dout && dqkv && dsum, "attn_bwd: null pointer");
  TFSWA_REQUIRE(a->C % 32 == 0 && a->heads > 0 && a->C % a->heads == 0, "attn_bwd: bad C/heads");
  const int D = a->C / a->heads;
  TFSWA_REQUIRE(D == 4 || D == 8 || D == 16 || D == 32, "attn_bwd: head_dim %d not in {4,8,16,32}", D);
  TFSWA_REQUIRE(a->ldq % 8 == 0 && a->ldo % 4 == 0, "attn_bwd: ldq/ldo alignment");
  TFSWA_REQUIRE(!a->rel_bias && !a->use_shift_mask, "attn_bwd: the optional Swin mask / relative-position bias are forward-only");
  AttnParams p = {};
  p.qkv = a->qkv; p.ldq = a->ldq; p.o = a->out; p.ldo = a->ldo; p.lse = a->lse; p.pad_kv = a->pad_kv;
  p.B = a->B; p.H = a->H; p.W = a->W; p.C = a->C; p.heads = a->heads;
  p.geom = a->geom; p.ws = a->ws; p.shift = a->shift;
  p.scale = (float)(1.0 / sqrt((double)D));
  p.qscale = (float)(1.4426950408889634 / sqrt((double)D));
  p.dout = dout; p.dqkv = dqkv; p.dsum = dsum; p.dpad = dpad;
  if (a->geom == TFSWA_GEOM_SWA) {
    TFSWA_REQUIRE(a->ws == 8 && a->shift >= 0 && a->shift < a->ws, "attn_bwd: window 8 only");
    attn_fill_geometry(p);
    TFSWA_REQUIRE((p.Hp == a->H && p.Wp == a->W) || a->pad_kv, "attn_bwd: padded windows need pad_kv");
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (a->dtype == TFSWA_BF16 && !getenv("TFSWA_ATTN_BWD_SIMT")) {  // tensor-core path (attention_bwd_mma.cu); 1 = not covered
    const int rc = attn_bwd_mma_bf16(p, st);
    if (rc != 1) return rc;
  }
#define TFSWA_ATTN_D(T)                                       \
  switch (D) {                                                \
    case 4: return launch_attn_bwd<T, 4>(p, st);              \
    case 8: return launch_attn_bwd<T, 8>(p, st);              \
    case 16: return launch_attn_bwd<T, 16>(p, st);            \
    default: return launch_attn_bwd<T, 32>(p, st);            \
  }
  if (a->dtype == TFSWA_F32) { TFSWA_ATTN_D(float) }
  if (a->dtype == TFSWA_BF16) { TFSWA_ATTN_D(bf16) }
#undef TFSWA_ATTN_D
  TFSWA_REQUIRE(false, "attn_bwd: bad dtype %d", a->dtype);
}
