// Attention core for the three TFSWA geometries, fp32 math, streaming ("flash") softmax.
//
// Replaces attention.py:70-85 plus the token regrouping around it: TSA permute(0,3,2,1) (:143,:162),
// FSA permute(0,2,3,1) (:217,:236) and the SW-MSA pad / roll / window_partition / window_reverse /
// roll-back / crop chain (:358-375, :390-401).  None of those regroupings is materialised: a
// sequence is just (base token, stride) for the axial cases and a modular coordinate map for the
// windows.  Zero-padded window tokens are REAL keys whose k|v equal the folded qkv bias (LN(0)=beta),
// passed as pad_kv (attention.py:358-365 semantics).
//
// Work decomposition: one thread owns one (query, head); a CTA owns 64 queries x (32/D) heads that
// share one 32-channel slab of K and V, staged through shared memory as fp32 in tiles of 128 keys.
// All lanes of a warp have the same head, so every K/V shared-memory read is a broadcast.
// The softmax runs in the exp2 domain (scale*log2e folded into q); at head_dim 4..8 this kernel is
// bound by MUFU.EX2 + the online-softmax bookkeeping, not by the dot products.
#include "attn_common.cuh"

namespace tfswa {

template <typename T, int D, bool WINDOW, bool EXTRAS>
__global__ void __launch_bounds__(QT * (32 / D)) attn_fwd_kernel(const AttnParams p) {
  constexpr int HG = 32 / D;               // heads per CTA (one 32-channel slab)
  constexpr int NT = QT * HG;
  __shared__ __align__(16) float Ks[KT][32];
  __shared__ __align__(16) float Vs[KT][32];
  __shared__ int s_region[EXTRAS ? 64 : 1];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int head_local = warp % HG;
  const int qsub = warp / HG;
  const int row = blockIdx.x;
  const int q0 = (WINDOW ? 0 : p.q_begin) + blockIdx.y * QT;
  const int hg = blockIdx.z;
  const int N = WINDOW ? p.ws * p.ws : (p.geom == TFSWA_GEOM_TSA ? p.H : p.W);
  const int ch0 = hg * 32;                 // first channel of this CTA's slab
  const int head = hg * HG + head_local;
  const T* qkv = (const T*)p.qkv;

  // ---- my query ----
  const int qn = q0 + qsub * 32 + lane;
  bool q_valid = qn < (WINDOW || p.q_end == 0 ? N : p.q_end);
  int64_t q_tok = 0;
  if (q_valid) { bool v; q_tok = token_of<WINDOW>(p, row, qn, v); q_valid = v; }
  float q[D];
  if (q_valid) {
    const T* qp = qkv + q_tok * p.ldq + ch0 + head_local * D;
    if (D == 4) { float t[4]; load4(qp, t);
#pragma unroll
      for (int d = 0; d < 4; ++d) q[d] = t[d] * p.qscale;
    } else {
#pragma unroll
      for (int d8 = 0; d8 < D / 8; ++d8) { float t[8]; load8(qp + d8 * 8, t);
#pragma unroll
        for (int d = 0; d < 8; ++d) q[d8 * 8 + d] = t[d] * p.qscale; }
    }
  } else {
#pragma unroll
    for (int d = 0; d < D; ++d) q[d] = 0.f;
  }
  int q_region = 0;
  const float* bias_row = nullptr;
  if (EXTRAS) {
    if (p.use_shift_mask && p.shift > 0 && qn < N) {
      const int per_img = p.nWh * p.nWw; const int r = row % per_img;
      q_region = swin_region(p, (r / p.nWw) * p.ws + qn / p.ws, (r % p.nWw) * p.ws + qn % p.ws);
    }
    if (p.rel_bias && qn < N) bias_row = p.rel_bias + ((int64_t)head * N + qn) * N;
  }

  float m = -CUDART_INF_F, l = 0.f;
  float acc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) acc[d] = 0.f;

  for (int k0 = 0; k0 < N; k0 += KT) {
    const int kcount = min(KT, N - k0);
    if (k0 > 0) __syncthreads();
    // ---- stage K|V slab of this tile: kcount keys x 32 channels, 8 elements per load ----
    for (int v = tid; v < kcount * 8; v += NT) {
      const int j = v >> 3, part = v & 7;          // part 0-3: K cols part*8.., 4-7: V cols (part-4)*8..
      bool valid; const int64_t tok = token_of<WINDOW>(p, row, k0 + j, valid);
      const int col = (part & 3) * 8;
      float t[8];
      if (valid) {
        load8(qkv + tok * p.ldq + (part < 4 ? p.C : 2 * p.C) + ch0 + col, t);
      } else {
        const float* pk = p.pad_kv + (part < 4 ? 0 : p.C) + ch0 + col;
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = pk[e];
      }
      float* dst = (part < 4 ? &Ks[j][col] : &Vs[j][col]);
      *reinterpret_cast<float4*>(dst) = make_float4(t[0], t[1], t[2], t[3]);
      *reinterpret_cast<float4*>(dst + 4) = make_float4(t[4], t[5], t[6], t[7]);
    }
    if (EXTRAS && WINDOW && k0 == 0 && p.use_shift_mask && p.shift > 0 && tid < N) {
      const int per_img = p.nWh * p.nWw; const int r = row % per_img;
      s_region[tid] = swin_region(p, (r / p.nWw) * p.ws + tid / p.ws, (r % p.nWw) * p.ws + tid % p.ws);
    }
    __syncthreads();

    const int c0 = head_local * D;
    for (int j = 0; j < kcount; j += 4) {
      float s[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int jj = min(j + u, kcount - 1);     // clamp keeps the smem read in range; masked below
        float a = 0.f;
#pragma unroll
        for (int d4 = 0; d4 < D / 4; ++d4) {
          const float4 kk = *reinterpret_cast<const float4*>(&Ks[jj][c0 + d4 * 4]);
          a = fmaf(q[d4 * 4 + 0], kk.x, a); a = fmaf(q[d4 * 4 + 1], kk.y, a);
          a = fmaf(q[d4 * 4 + 2], kk.z, a); a = fmaf(q[d4 * 4 + 3], kk.w, a);
        }
        if (EXTRAS) {
          if (bias_row) a += bias_row[k0 + jj] * 1.4426950408889634f;
          if (p.use_shift_mask && p.shift > 0 && s_region[jj] != q_region) a += -100.0f * 1.4426950408889634f;
        }
        s[u] = (j + u < kcount) ? a : -CUDART_INF_F;
      }
      const float mx = fmaxf(fmaxf(s[0], s[1]), fmaxf(s[2], s[3]));
      if (mx > m) {
        const float c = fast_exp2(m - mx);
        l *= c;
#pragma unroll
        for (int d = 0; d < D; ++d) acc[d] *= c;
        m = mx;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int jj = min(j + u, kcount - 1);
        const float pu = fast_exp2(s[u] - m);
        l += pu;
#pragma unroll
        for (int d4 = 0; d4 < D / 4; ++d4) {
          const float4 vv = *reinterpret_cast<const float4*>(&Vs[jj][c0 + d4 * 4]);
          acc[d4 * 4 + 0] = fmaf(pu, vv.x, acc[d4 * 4 + 0]); acc[d4 * 4 + 1] = fmaf(pu, vv.y, acc[d4 * 4 + 1]);
          acc[d4 * 4 + 2] = fmaf(pu, vv.z, acc[d4 * 4 + 2]); acc[d4 * 4 + 3] = fmaf(pu, vv.w, acc[d4 * 4 + 3]);
        }
      }
    }
  }

  if (q_valid) {
    const float inv = 1.0f / l;
    T* op = (T*)p.out + q_tok * p.ldo + ch0 + head_local * D;
    if (D == 4) {
      float t[4] = {acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv};
      store4(op, t);
    } else {
#pragma unroll
      for (int d8 = 0; d8 < D / 8; ++d8) {
        float t[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) t[d] = acc[d8 * 8 + d] * inv;
        store8(op + d8 * 8, t);
      }
    }
    if (p.lse) p.lse[q_tok * p.heads + head] = m + log2f(l);
  }
}

template <typename T, int D>
static int launch_attn(const AttnParams& p, cudaStream_t st) {
  constexpr int HG = 32 / D;
  const bool window = p.geom == TFSWA_GEOM_SWA;
  const bool extras = window && (p.rel_bias || (p.use_shift_mask && p.shift > 0));
  const int hgs = p.C / 32;
  if (window) {
    dim3 grid((unsigned)(p.B * p.nWh * p.nWw), 1, hgs);
    if (extras) attn_fwd_kernel<T, D, true, true><<<grid, QT * HG, 0, st>>>(p);
    else attn_fwd_kernel<T, D, true, false><<<grid, QT * HG, 0, st>>>(p);
  } else {
    const int N = p.geom == TFSWA_GEOM_TSA ? p.H : p.W;
    const int rows = p.geom == TFSWA_GEOM_TSA ? p.B * p.W : p.B * p.H;
    const int nq = (p.q_end ? p.q_end : N) - p.q_begin;
    dim3 grid(rows, (nq + QT - 1) / QT, hgs);
    attn_fwd_kernel<T, D, false, false><<<grid, QT * HG, 0, st>>>(p);
  }
  return check_launch("attn_fwd");
}

// Ragged-remainder kernel for the tensor-core path: a handful of queries per sequence (N mod 128 < 32).  One warp per
// (sequence, query, 32-channel slab); the 32 lanes split the KEYS (lane l takes keys l, l+32, ...) and each lane
// handles all 32/D heads of the slab from one 64-byte k row and one 64-byte v row (full-sector loads), keeps its own
// online-softmax state per head, and the warp merges the 32 partial states with shuffles.
template <int D>
__global__ void __launch_bounds__(256) attn_rem_kernel(const AttnParams p) {
  constexpr int HG = 32 / D;
  const int row = blockIdx.x, lane = threadIdx.x & 31;
  const int N = p.geom == TFSWA_GEOM_TSA ? p.H : p.W;
  const int qn = p.q_begin + blockIdx.y * 8 + (threadIdx.x >> 5);
  if (qn >= (p.q_end ? p.q_end : N)) return;
  const int ch0 = blockIdx.z * 32;
  const bf16* qkv = (const bf16*)p.qkv;
  bool vld; const int64_t q_tok = token_of<false>(p, row, qn, vld);
  float q[32];
#pragma unroll
  for (int i = 0; i < 4; ++i) { float t[8]; load8(qkv + q_tok * p.ldq + ch0 + i * 8, t);
#pragma unroll
    for (int d = 0; d < 8; ++d) q[i * 8 + d] = t[d] * p.qscale; }
  float m[HG], l[HG], acc[32];
#pragma unroll
  for (int h = 0; h < HG; ++h) { m[h] = -CUDART_INF_F; l[h] = 0.f; }
#pragma unroll
  for (int d = 0; d < 32; ++d) acc[d] = 0.f;
  for (int j = lane; j < N; j += 32) {
    const int64_t tok = token_of<false>(p, row, j, vld);
    float kk[32], vv[32];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      load8(qkv + tok * p.ldq + p.C + ch0 + i * 8, *reinterpret_cast<float(*)[8]>(kk + i * 8));
      load8(qkv + tok * p.ldq + 2 * p.C + ch0 + i * 8, *reinterpret_cast<float(*)[8]>(vv + i * 8));
    }
#pragma unroll
    for (int h = 0; h < HG; ++h) {
      float s = 0.f;
#pragma unroll
      for (int d = 0; d < D; ++d) s = fmaf(q[h * D + d], kk[h * D + d], s);
      if (s > m[h]) {
        const float c = fast_exp2(m[h] - s);
        l[h] *= c;
#pragma unroll
        for (int d = 0; d < D; ++d) acc[h * D + d] *= c;
        m[h] = s;
      }
      const float pw = fast_exp2(s - m[h]);
      l[h] += pw;
#pragma unroll
      for (int d = 0; d < D; ++d) acc[h * D + d] = fmaf(pw, vv[h * D + d], acc[h * D + d]);
    }
  }
  // merge the 32 partial softmax states of every head
  float out[32], lse[HG];
#pragma unroll
  for (int h = 0; h < HG; ++h) {
    float mg = m[h];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, o));
    const float sc = (m[h] == -CUDART_INF_F) ? 0.f : fast_exp2(m[h] - mg);
    const float lt = warp_sum(l[h] * sc);
    const float inv = 1.0f / lt;
#pragma unroll
    for (int d = 0; d < D; ++d) out[h * D + d] = warp_sum(acc[h * D + d] * sc) * inv;
    lse[h] = mg + log2f(lt);
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) store8((bf16*)p.out + q_tok * p.ldo + ch0 + i * 8, *reinterpret_cast<float(*)[8]>(out + i * 8));
    if (p.lse) {
#pragma unroll
      for (int h = 0; h < HG; ++h) p.lse[q_tok * p.heads + blockIdx.z * HG + h] = lse[h];
    }
  }
}

// variant with one warp per (sequence, query, head): more warps in flight when only one or two queries are left over
template <int D>
__global__ void __launch_bounds__(256) attn_rem_head_kernel(const AttnParams p) {
  const int row = blockIdx.x, qn = p.q_begin + blockIdx.y;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.z * 8 + (threadIdx.x >> 5);
  const int N = p.geom == TFSWA_GEOM_TSA ? p.H : p.W;
  if (head >= p.heads) return;
  const bf16* qkv = (const bf16*)p.qkv;
  bool v; const int64_t q_tok = token_of<false>(p, row, qn, v);
  float q[D];
  {
    const bf16* qp = qkv + q_tok * p.ldq + head * D;
    if (D == 4) { float t[4]; load4(qp, t);
#pragma unroll
      for (int d = 0; d < 4; ++d) q[d] = t[d] * p.qscale;
    } else {
#pragma unroll
      for (int d8 = 0; d8 < D / 8; ++d8) { float t[8]; load8(qp + d8 * 8, t);
#pragma unroll
        for (int d = 0; d < 8; ++d) q[d8 * 8 + d] = t[d] * p.qscale; }
    }
  }
  float m = -CUDART_INF_F, l = 0.f, acc[D];
#pragma unroll
  for (int d = 0; d < D; ++d) acc[d] = 0.f;
  for (int j = lane; j < N; j += 32) {
    const int64_t tok = token_of<false>(p, row, j, v);
    float kk[D], vv[D];
    const bf16* kp = qkv + tok * p.ldq + p.C + head * D;
    const bf16* vp = qkv + tok * p.ldq + 2 * p.C + head * D;
    if (D == 4) { load4(kp, *reinterpret_cast<float(*)[4]>(kk)); load4(vp, *reinterpret_cast<float(*)[4]>(vv)); }
    else {
#pragma unroll
      for (int d8 = 0; d8 < D / 8; ++d8) {
        load8(kp + d8 * 8, *reinterpret_cast<float(*)[8]>(kk + d8 * 8));
        load8(vp + d8 * 8, *reinterpret_cast<float(*)[8]>(vv + d8 * 8));
      }
    }
    float s = 0.f;
#pragma unroll
    for (int d = 0; d < D; ++d) s = fmaf(q[d], kk[d], s);
    if (s > m) {
      const float c = fast_exp2(m - s);
      l *= c;
#pragma unroll
      for (int d = 0; d < D; ++d) acc[d] *= c;
      m = s;
    }
    const float pw = fast_exp2(s - m);
    l += pw;
#pragma unroll
    for (int d = 0; d < D; ++d) acc[d] = fmaf(pw, vv[d], acc[d]);
  }
  // merge the 32 partial softmax states
  float mg = m;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mg = fmaxf(mg, __shfl_xor_sync(0xffffffffu, mg, o));
  const float sc = (m == -CUDART_INF_F) ? 0.f : fast_exp2(m - mg);
  l *= sc;
#pragma unroll
  for (int d = 0; d < D; ++d) acc[d] *= sc;
  l = warp_sum(l);
#pragma unroll
  for (int d = 0; d < D; ++d) acc[d] = warp_sum(acc[d]);
  if (lane == 0) {
    const float inv = 1.0f / l;
    bf16* op = (bf16*)p.out + q_tok * p.ldo + head * D;
    if (D == 4) { float t[4] = {acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv}; store4(op, t); }
    else {
#pragma unroll
      for (int d8 = 0; d8 < D / 8; ++d8) { float t[8];
#pragma unroll
        for (int d = 0; d < 8; ++d) t[d] = acc[d8 * 8 + d] * inv;
        store8(op + d8 * 8, t); }
    }
    if (p.lse) p.lse[q_tok * p.heads + head] = mg + log2f(l);
  }
}

int attn_simt_axial_bf16(const AttnParams& p, cudaStream_t st) {
  const int D = p.C / p.heads;
  const int N = p.geom == TFSWA_GEOM_TSA ? p.H : p.W;
  const int rows = p.geom == TFSWA_GEOM_TSA ? p.B * p.W : p.B * p.H;
  const int nq = (p.q_end ? p.q_end : N) - p.q_begin;
  if (nq <= 2) {
    dim3 gridh(rows, nq, (p.heads + 7) / 8);
    if (D == 4) attn_rem_head_kernel<4><<<gridh, 256, 0, st>>>(p);
    else if (D == 8) attn_rem_head_kernel<8><<<gridh, 256, 0, st>>>(p);
    else if (D == 16) attn_rem_head_kernel<16><<<gridh, 256, 0, st>>>(p);
    else attn_rem_head_kernel<32><<<gridh, 256, 0, st>>>(p);
    return check_launch("attn_rem");
  }
  dim3 grid(rows, (nq + 7) / 8, p.C / 32);
  if (D == 4) attn_rem_kernel<4><<<grid, 256, 0, st>>>(p);
  else if (D == 8) attn_rem_kernel<8><<<grid, 256, 0, st>>>(p);
  else if (D == 16) attn_rem_kernel<16><<<grid, 256, 0, st>>>(p);
  else attn_rem_kernel<32><<<grid, 256, 0, st>>>(p);
  return check_launch("attn_rem");
}

}  // namespace tfswa

using namespace tfswa;

extern "C" int tfswa_attn_fwd(const tfswa_attn_args* a, void* stream) {
  TFSWA_REQUIRE(a && a->qkv && a->out, "attn: null pointer");
  TFSWA_REQUIRE(a->B > 0 && a->H > 0 && a->W > 0, "attn: empty problem");
  TFSWA_REQUIRE(a->C % 32 == 0 && a->heads > 0 && a->C % a->heads == 0, "attn: C=%d must be a multiple of 32 and of heads=%d", a->C, a->heads);
  const int D = a->C / a->heads;
  TFSWA_REQUIRE(D == 4 || D == 8 || D == 16 || D == 32, "attn: head_dim %d not in {4,8,16,32}", D);
  TFSWA_REQUIRE(a->ldq % 8 == 0 && a->ldo % 4 == 0, "attn: ldq/ldo alignment");
  TFSWA_REQUIRE(a->geom >= 0 && a->geom <= 2, "attn: bad geometry %d", a->geom);
  AttnParams p = {};
  p.qkv = a->qkv; p.ldq = a->ldq; p.out = a->out; p.ldo = a->ldo; p.lse = a->lse; p.pad_kv = a->pad_kv; p.rel_bias = a->rel_bias;
  p.B = a->B; p.H = a->H; p.W = a->W; p.C = a->C; p.heads = a->heads;
  p.geom = a->geom; p.ws = a->ws; p.shift = a->shift; p.use_shift_mask = a->use_shift_mask;
  p.qscale = (float)(1.4426950408889634 / sqrt((double)D));
  if (a->geom == TFSWA_GEOM_SWA) {
    TFSWA_REQUIRE(a->ws == 8, "attn: window size %d unsupported (8 only)", a->ws);
    TFSWA_REQUIRE(a->shift >= 0 && a->shift < a->ws, "attn: bad shift %d", a->shift);
    p.Hp = (a->H + a->ws - 1) / a->ws * a->ws; p.Wp = (a->W + a->ws - 1) / a->ws * a->ws;
    p.nWh = p.Hp / a->ws; p.nWw = p.Wp / a->ws;
    TFSWA_REQUIRE((p.Hp == a->H && p.Wp == a->W) || a->pad_kv, "attn: padded windows need pad_kv");
  }
  cudaStream_t st = (cudaStream_t)stream;
#define TFSWA_ATTN_D(T)                                   \
  switch (D) {                                            \
    case 4: return launch_attn<T, 4>(p, st);              \
    case 8: return launch_attn<T, 8>(p, st);              \
    case 16: return launch_attn<T, 16>(p, st);            \
    default: return launch_attn<T, 32>(p, st);            \
  }
  if (a->dtype == TFSWA_F32) { TFSWA_ATTN_D(float) }
  if (a->dtype == TFSWA_BF16) { TFSWA_ATTN_D(bf16) }
#undef TFSWA_ATTN_D
  TFSWA_REQUIRE(false, "attn: bad dtype %d", a->dtype);
}
