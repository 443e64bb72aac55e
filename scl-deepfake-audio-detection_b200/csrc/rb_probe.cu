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

}  // namespace
}  // namespace rb

extern "C" int rb_probe_fp32(int packed, int iters, float* sink, double* flops, void* stream) {
  using namespace rb;
  if (!sink || iters <= 0) return RB_ERR_INVALID_ARG;
  int dev = 0, sms = 0;
  RB_CUDA(cudaGetDevice(&dev));
  RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int blocks = sms * 8;  // 2048 threads per SM
  if (packed)
    fp32_probe_kernel<true><<<blocks, kProbeThreads, 0, (cudaStream_t)stream>>>(iters, 0.5f, sink);
  else
    fp32_probe_kernel<false><<<blocks, kProbeThreads, 0, (cudaStream_t)stream>>>(iters, 0.5f, sink);
  RB_LAUNCH_CHECK();
  if (flops) *flops = 2.0 * 2.0 * kProbeAcc * kProbeInner * (double)iters * kProbeThreads * (double)blocks;
  return RB_OK;
}
