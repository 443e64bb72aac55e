// rb_probe.cu -- FP32-pipe peak probe used as the roofline denominator of the FIR-bank kernel.
// MEASURED_PEAKS.json carries HBM and bf16-tensor peaks only; a direct convolution is bounded by the
// non-tensor FP32 pipe, so bench.py measures that pipe's attainable rate with this register-resident chain
// (same instruction the hot loop uses: FFMA2 / fma.rn.f32x2, or scalar FFMA for comparison).
#include "rb_common.cuh"

namespace rb {
namespace {

constexpr int kProbeThreads = 256;
constexpr int kProbeAcc = 16;      // independent accumulator pairs per thread
constexpr int kProbeInner = 64;    // unrolled FMAs per accumulator per outer iteration

// The multiplier `a` is uniform across the grid; built with -Xptxas --register-usage-level=3 (csrc/build.sh) ptxas keeps it in
// a uniform register (`FFMA2 R, R, UR, R`), so each FFMA2 reads two register pairs instead of three and issues every 2 cycles
// without depending on the operand-reuse cache: 73.9 TFLOP/s = 99 % of 148 SMs x 128 lanes x 2 x 1.965 GHz. At the default
// level the multiplier sits in a vector register, every sixth FFMA2 goes without a reuse flag, and the same chain measures
// 72.3 (profiles/r02m_register_usage_level.log). The higher figure is the pipe's attainable rate, hence the roofline
// denominator; tests/test_host_logic.py checks that the built probe has the uniform-register form.
template <bool PACKED>
__global__ void __launch_bounds__(kProbeThreads)
fp32_probe_kernel(int iters, float seed, float* __restrict__ sink) {
  float2 acc[kProbeAcc];
#pragma unroll
  for (int i = 0; i < kProbeAcc; ++i) acc[i] = make_float2(seed + i, seed - i);
  float2 a = make_float2(1.0f + seed * 1e-7f, 1.0f - seed * 1e-7f);
  float2 b = make_float2(seed * 1e-3f + threadIdx.x * 1e-9f, seed * -1e-3f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < kProbeInner; ++k) {
#pragma unroll
      for (int i = 0; i < kProbeAcc; ++i) {
        if (PACKED) {
          acc[i] = __ffma2_rn(a, acc[i], b);
        } else {
          acc[i].x = __fmaf_rn(a.x, acc[i].x, b.x);
          acc[i].y = __fmaf_rn(a.y, acc[i].y, b.y);
        }
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kProbeAcc; ++i) s += acc[i].x + acc[i].y;
  if (s == 123.456f) sink[0] = s;  // keeps the chain alive without a store in the common case
}

// Operand-pattern probes (packed = 2, 3): how fast FFMA2 runs when its operands come from the register file the way
// the FIR loop's do. Mode 2: one operand shared by runs of 10 (the tap), one distinct per instruction (the window),
// one accumulator -- exactly the hot loop's pattern without the shared-memory loads. Mode 3: three distinct pairs.
template <int MODE>
__global__ void __launch_bounds__(128)
fp32_pattern_kernel(int iters, float seed, float* __restrict__ sink) {
  float2 acc[40], w[12], t[4];
#pragma unroll
  for (int i = 0; i < 40; ++i) acc[i] = make_float2(seed + i, seed - i);
#pragma unroll
  for (int i = 0; i < 12; ++i) w[i] = make_float2(1.0f + (seed + i + threadIdx.x) * 1e-7f, 1.0f - (seed + i) * 1e-7f);
#pragma unroll
  for (int i = 0; i < 4; ++i) t[i] = make_float2(seed * 1e-3f + i, seed * -1e-3f - i);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
      for (int q = 0; q < 4; ++q) {
#pragma unroll
        for (int r = 0; r < 10; ++r) {
          if (MODE == 2) acc[(q & 1) * 20 + 2 * r + (q >> 1)] = __ffma2_rn(t[q], w[(r + (q >> 1) + k) % 12], acc[(q & 1) * 20 + 2 * r + (q >> 1)]);
          else acc[q * 10 + r] = __ffma2_rn(acc[(q * 10 + r + 7) % 40], w[(r + q + k) % 12], acc[q * 10 + r]);
        }
      }
    }
    // rotate the window so the compiler cannot hoist anything
    const float2 w0 = w[0];
#pragma unroll
    for (int i = 0; i < 11; ++i) w[i] = w[i + 1];
    w[11] = w0;
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 40; ++i) s += acc[i].x + acc[i].y;
  if (s == 123.456f) sink[0] = s;
}

}  // namespace
}  // namespace rb

extern "C" int rb_probe_fp32(int packed, int iters, float* sink, double* flops, void* stream) {
  using namespace rb;
  if (!sink || iters <= 0) return RB_ERR_INVALID_ARG;
  int dev = 0, sms = 0;
  RB_CUDA(cudaGetDevice(&dev));
  RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8;  // 2048 threads per SM
  if (packed >= 2) {
    const int pblocks = sms * 4;  // 4 CTAs x 128 threads per SM, like the FIR kernel
    if (packed == 2) fp32_pattern_kernel<2><<<pblocks, 128, 0, (cudaStream_t)stream>>>(iters, 0.5f, sink);
    else fp32_pattern_kernel<3><<<pblocks, 128, 0, (cudaStream_t)stream>>>(iters, 0.5f, sink);
    RB_LAUNCH_CHECK();
    if (flops) *flops = 2.0 * 2.0 * 160.0 * (double)iters * 128.0 * (double)pblocks;
    return RB_OK;
  }
  if (packed)
    fp32_probe_kernel<true><<<blocks, kProbeThreads, 0, (cudaStream_t)stream>>>(iters, 0.5f, sink);
  else
    fp32_probe_kernel<false><<<blocks, kProbeThreads, 0, (cudaStream_t)stream>>>(iters, 0.5f, sink);
  RB_LAUNCH_CHECK();
  if (flops) *flops = 2.0 * 2.0 * kProbeAcc * kProbeInner * (double)iters * kProbeThreads * (double)blocks;
  return RB_OK;
}
