// Exponential-roofline microbenchmark for B200 (sm_100a): sustained results/s of the instruction mixes a softmax inner
// loop can be built from.  The attention kernels at head_dim 4/8 are bound by exponentials, not by MMA rate or HBM
// (DESIGN 4.3), so this is the roofline denominator bench.py reports their `exp_rate` against.
//
//   mufu_f32      ex2.approx.ftz.f32 only (32 independent chains per thread)
//   mufu_f16x2    ex2.approx.f16x2 (2 results per instruction)
//   mufu_bf16x2   ex2.approx.ftz.bf16x2
//   poly_scalar   degree-3 Cody-Waite 2^x on the FMA/ALU pipes, scalar instructions (no MUFU)
//   poly_packed   the same with fma/add .f32x2 (FFMA2 / FADD2: one issue slot per two elements)
//   sm_scalar     the softmax element as round 1 shipped it: FFMA (scale, shift) + MUFU.EX2 + cvt.bf16x2 pack
//   sm_packed     FFMA2 + 2 MUFU.EX2 + pack
//   sm_mixK       sm_packed with every K-th PAIR on poly_packed instead of MUFU (K = 2, 3, 4, 6, 8)
//   sm_mixK_c     ... with the -125 clamp (FMNMX per element) on the polynomial pairs
// at 1, 2, 4, 8, 16 warps per SM sub-partition.  One CTA per SM, grid = #SMs, every thread keeps 32 "scores" in
// registers (like one tcgen05.ld.x32) and re-evaluates them ITER times with a loop-variant shift.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo tools/mufu_bench.cu -o tools/bin/mufu_bench
// Run  :  tools/bin/mufu_bench [iters]     -> one JSON object on stdout
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_h2(uint32_t x) { uint32_t y; asm("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_b2(uint32_t x) { uint32_t y; asm("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { uint32_t y; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo)); return y; }
__device__ __forceinline__ uint64_t pk2(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void up2(uint64_t r, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(r)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { uint64_t d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

constexpr float MAGIC = 12582912.0f;   // 1.5 * 2^23
constexpr float P3 = 0.0551716685f, P2 = 0.2426111251f, P1 = 0.6932609677f, P0 = 0.9999280572f;

__device__ __forceinline__ float poly_scalar(float x) {
  x = fmaxf(x, -125.0f);
  const float t = x + MAGIC;
  const float r = x - (t - MAGIC);
  float p = fmaf(P3, r, P2);
  p = fmaf(p, r, P1);
  p = fmaf(p, r, P0);
  return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}
// two elements: 3 FADD2 + 3 FFMA2 + 2 integer shift-adds
template <bool CLAMP>
__device__ __forceinline__ void poly_packed(uint64_t x, float& e0, float& e1) {
  if (CLAMP) { float a, b; up2(x, a, b); x = pk2(fmaxf(a, -125.0f), fmaxf(b, -125.0f)); }
  const uint64_t t = add2(x, pk2(MAGIC, MAGIC));
  const uint64_t u = add2(t, pk2(-MAGIC, -MAGIC));
  const uint64_t r = sub2(x, u);
  uint64_t p = fma2(pk2(P3, P3), r, pk2(P2, P2));
  p = fma2(p, r, pk2(P1, P1));
  p = fma2(p, r, pk2(P0, P0));
  float p0, p1, t0, t1;
  up2(p, p0, p1); up2(t, t0, t1);
  e0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  e1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

enum Mode { MUFU_F32, MUFU_F16X2, MUFU_BF16X2, POLY_SCALAR, POLY_PACKED, SM_SCALAR, SM_PACKED, SM_MIX, SM_MIX_CLAMP };

// results per thread per iteration = 32 (f16x2 / bf16x2: 32 instructions = 64 results, reported as such by the host)
template <int MODE, int K, int WPS>
__global__ void __launch_bounds__(WPS * 128, 1) bench_kernel(float* sink, long long* cycles, int iters, float c, float mc0, float dmc) {
  float s[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) s[i] = -0.01f * (float)((threadIdx.x * 7 + i * 13) % 97) - 0.003f * i;
  uint32_t acc = 0;
  float mc = mc0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == MUFU_F32) {
#pragma unroll
      for (int i = 0; i < 32; ++i) s[i] = ex2f(s[i]);
    } else if (MODE == MUFU_F16X2) {
#pragma unroll
      for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(ex2_h2(__float_as_uint(s[i])));
    } else if (MODE == MUFU_BF16X2) {
#pragma unroll
      for (int i = 0; i < 32; ++i) s[i] = __uint_as_float(ex2_b2(__float_as_uint(s[i])));
    } else if (MODE == POLY_SCALAR) {
#pragma unroll
      for (int i = 0; i < 32; ++i) s[i] = poly_scalar(s[i]) - 1.0f;
    } else if (MODE == POLY_PACKED) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float e0, e1;
        poly_packed<false>(pk2(s[2 * i], s[2 * i + 1]), e0, e1);
        s[2 * i] = e0 - 1.0f; s[2 * i + 1] = e1 - 1.0f;
      }
    } else if (MODE == SM_SCALAR) {
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float e0 = ex2f(fmaf(s[2 * i], c, -mc)), e1 = ex2f(fmaf(s[2 * i + 1], c, -mc));
        pk[i] = pack_bf16x2(e0, e1);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc ^= pk[2 * i] ^ pk[2 * i + 1];   // sink: one LOP3 per 4 results (stands in for tcgen05.st)
    } else {
      uint32_t pk[16];
      const uint64_t c2 = pk2(c, c), m2 = pk2(-mc, -mc);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const uint64_t x = fma2(pk2(s[2 * i], s[2 * i + 1]), c2, m2);
        float e0, e1;
        const bool poly = (MODE == SM_MIX || MODE == SM_MIX_CLAMP) && K > 0 && (i % K) == K - 1;
        if (poly) poly_packed<MODE == SM_MIX_CLAMP>(x, e0, e1);
        else { float x0, x1; up2(x, x0, x1); e0 = ex2f(x0); e1 = ex2f(x1); }
        pk[i] = pack_bf16x2(e0, e1);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc ^= pk[2 * i] ^ pk[2 * i + 1];   // sink: one LOP3 per 4 results (stands in for tcgen05.st)
    }
    mc += dmc;
  }
  const long long t1 = clock64();
#pragma unroll
  for (int i = 0; i < 32; ++i) acc ^= __float_as_uint(s[i]);
  if (acc == 0x12345678u) sink[threadIdx.x] = mc;      // never true in practice; keeps the chains alive
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct Variant { const char* name; int mode, k; double results_per_instr; };

template <int MODE, int K, int WPS>
static void run_wps(const char* name, double rpi, int nsm, int iters, float* sink, long long* dcycles, bool& first);

template <int MODE, int K>
static void run(const char* name, double rpi, int nsm, int iters, float* sink, long long* dcycles, bool& first) {
  run_wps<MODE, K, 1>(name, rpi, nsm, iters, sink, dcycles, first);
  run_wps<MODE, K, 2>(name, rpi, nsm, iters, sink, dcycles, first);
  run_wps<MODE, K, 4>(name, rpi, nsm, iters, sink, dcycles, first);
  run_wps<MODE, K, 8>(name, rpi, nsm, iters, sink, dcycles, first);
}

template <int MODE, int K, int WPS>
static void run_wps(const char* name, double rpi, int nsm, int iters, float* sink, long long* dcycles, bool& first) {
  {
    const int wps = WPS, threads = WPS * 4 * 32;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    bench_kernel<MODE, K, WPS><<<nsm, threads>>>(sink, dcycles, iters / 8 + 1, 0.7213f, 3.0f, 1e-6f);   // warm-up
    CK(cudaDeviceSynchronize());
    float best_ms = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
      CK(cudaEventRecord(e0));
      bench_kernel<MODE, K, WPS><<<nsm, threads>>>(sink, dcycles, iters, 0.7213f, 3.0f, 1e-6f);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (ms < best_ms) best_ms = ms;
    }
    long long* h = (long long*)malloc(sizeof(long long) * nsm);
    CK(cudaMemcpy(h, dcycles, sizeof(long long) * nsm, cudaMemcpyDeviceToHost));
    double cyc = 0; for (int i = 0; i < nsm; ++i) cyc += (double)h[i]; cyc /= nsm;
    free(h);
    const double results = (double)nsm * threads * 32.0 * rpi * iters;
    const double per_clk_sm = (double)threads * 32.0 * rpi * iters / cyc;
    printf("%s\n  {\"variant\": \"%s\", \"warps_per_smsp\": %d, \"results_per_s\": %.4e, \"results_per_clk_per_sm\": %.3f, "
           "\"cycles_per_warp_result_per_smsp\": %.3f, \"ms\": %.4f, \"sm_cycles\": %.0f}",
           first ? "" : ",", name, wps, results / (best_ms * 1e-3), per_clk_sm, 128.0 / per_clk_sm, best_ms, cyc);
    first = false;
    CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
  }
}

int main(int argc, char** argv) {
  const int iters = argc > 1 ? atoi(argv[1]) : 4000;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int nsm = prop.multiProcessorCount;
  float* sink; long long* dcycles;
  CK(cudaMalloc(&sink, 4096 * sizeof(float)));
  CK(cudaMalloc(&dcycles, sizeof(long long) * nsm));
  int clock_khz = 0;
  CK(cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0));
  printf("{\"device\": \"%s\", \"sms\": %d, \"max_sm_khz\": %d, \"iters\": %d, \"rows\": [", prop.name, nsm, clock_khz, iters);
  bool first = true;
  run<MUFU_F32, 0>("mufu_f32", 1.0, nsm, iters, sink, dcycles, first);
  run<MUFU_F16X2, 0>("mufu_f16x2", 2.0, nsm, iters, sink, dcycles, first);
  run<MUFU_BF16X2, 0>("mufu_bf16x2", 2.0, nsm, iters, sink, dcycles, first);
  run<POLY_SCALAR, 0>("poly_scalar", 1.0, nsm, iters, sink, dcycles, first);
  run<POLY_PACKED, 0>("poly_packed", 1.0, nsm, iters, sink, dcycles, first);
  run<SM_SCALAR, 0>("sm_scalar", 1.0, nsm, iters, sink, dcycles, first);
  run<SM_PACKED, 0>("sm_packed", 1.0, nsm, iters, sink, dcycles, first);
  run<SM_MIX, 8>("sm_mix8", 1.0, nsm, iters, sink, dcycles, first);
  run<SM_MIX, 6>("sm_mix6", 1.0, nsm, iters, sink, dcycles, first);
  run<SM_MIX, 4>("sm_mix4", 1.0, nsm, iters, sink, dcycles, first);
  run<SM_MIX, 3>("sm_mix3", 1.0, nsm, iters, sink, dcycles, first);
  run<SM_MIX, 2>("sm_mix2", 1.0, nsm, iters, sink, dcycles, first);
  run<SM_MIX_CLAMP, 4>("sm_mix4_c", 1.0, nsm, iters, sink, dcycles, first);
  run<SM_MIX_CLAMP, 3>("sm_mix3_c", 1.0, nsm, iters, sink, dcycles, first);
  run<SM_MIX_CLAMP, 2>("sm_mix2_c", 1.0, nsm, iters, sink, dcycles, first);
  printf("\n]}\n");
  return 0;
}
