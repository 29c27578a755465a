// Shared definitions of the attention forward and backward kernels: geometry parameters and the token index maps
// that replace the reference's permute / pad / roll / window_partition copies (attention.py:143,217,358-375).
#pragma once
#include "common.cuh"
#include <math_constants.h>

namespace tfswa {

constexpr int QT = 64;    // queries per CTA
constexpr int KT = 128;   // keys per shared-memory tile

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

struct AttnParams {
  const void* qkv; int64_t ldq;
  void* out; int64_t ldo;
  float* lse;
  const float* pad_kv;
  const float* rel_bias;
  int B, H, W, C, heads;
  int geom, ws, shift, use_shift_mask;
  int Hp, Wp, nWh, nWw;
  int win_edge, nWh_int, nWw_int;   // SW-MSA: blockIdx.x enumerates only the windows outside the interior [0, nWh_int) x [0, nWw_int)
  int q_begin, q_end;   // axial kernels: queries [q_begin, q_end) of every sequence (q_end = 0 means the whole sequence)
  float qscale;   // head_dim^-0.5 * log2(e)
  float* kext;     // tc attention: per-sequence per-channel min/max of k (scratch, (rows, 2, C) fp32)
  int force_exact; // tc attention: skip the row-max bound and run the exact two-pass path (tests)
  // backward
  const void* dout; const void* o; void* dqkv; float* dsum; float* dpad; float scale;
};

// fill the geometry-derived fields shared by forward and backward launchers
// SIMT flash kernel on a query range (attention.cu); used by the tensor-core launcher for ragged remainders
int attn_simt_axial_bf16(const AttnParams& p, cudaStream_t st);
// register-resident warp-MMA flash kernel on queries [0, q_end) (attention_axial_mma.cu); needs p.kext
int attn_axial_mma_bf16(const AttnParams& p, cudaStream_t st);
// tcgen05 + TMA flash kernel on queries [0, q_end) (tc_attn_tma.cu); needs p.kext and a work buffer of
// attn_axial_tma_work_bytes(p) bytes (redo list of the exact pass)
int64_t attn_axial_tma_work_bytes(const AttnParams& p);
int attn_axial_tma_bf16(const AttnParams& p, void* work, cudaStream_t st);
// SW-MSA on tcgen05 + TMA for the interior windows [0, nWh_int) x [0, nWw_int) (tc_attn_win.cu, head_dim 4 / 8); returns 1
// when the shape is not covered (nothing launched)
int attn_win_tc_bf16(const AttnParams& p, int nWh_int, int nWw_int, cudaStream_t st);
// bf16 axial attention backward on warp-level MMAs (attention_bwd_mma.cu); returns 1 when the shape is not covered
int attn_bwd_mma_bf16(const AttnParams& p, cudaStream_t st);

inline void attn_fill_geometry(AttnParams& p) {
  if (p.geom == TFSWA_GEOM_SWA) {
    p.Hp = (p.H + p.ws - 1) / p.ws * p.ws; p.Wp = (p.W + p.ws - 1) / p.ws * p.ws;
    p.nWh = p.Hp / p.ws; p.nWw = p.Wp / p.ws;
  }
}

// token index of element n of the sequence/window `row`; valid=false for zero-padded window tokens
template <bool WINDOW>
__device__ __forceinline__ int64_t token_of(const AttnParams& p, int row, int n, bool& valid) {
  valid = true;
  if (!WINDOW) {
    if (p.geom == TFSWA_GEOM_TSA) {
      const int b = row / p.W, w = row - b * p.W;
      return ((int64_t)b * p.H + n) * p.W + w;
    }
    return (int64_t)row * p.W + n;
  }
  const int per_img = p.nWh * p.nWw;
  const int b = row / per_img;
  const int r = row - b * per_img;
  const int wh = r / p.nWw, ww = r - wh * p.nWw;
  int hp = wh * p.ws + n / p.ws + p.shift;
  int wp = ww * p.ws + n % p.ws + p.shift;
  if (hp >= p.Hp) hp -= p.Hp;
  if (wp >= p.Wp) wp -= p.Wp;
  valid = (hp < p.H) && (wp < p.W);
  return ((int64_t)b * p.H + hp) * p.W + wp;
}

// Swin region id of shifted-frame coordinate (ys, xs): 3x3 regions split at (size-ws) and (size-shift)
__device__ __forceinline__ int swin_region(const AttnParams& p, int ys, int xs) {
  const int rh = ys < p.Hp - p.ws ? 0 : (ys < p.Hp - p.shift ? 1 : 2);
  const int rw = xs < p.Wp - p.ws ? 0 : (xs < p.Wp - p.shift ? 1 : 2);
  return rh * 3 + rw;
}

}  // namespace tfswa
