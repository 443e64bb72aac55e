"""Generate FFMA2 operand-pattern microbenchmarks to learn the register-file model behind the FIR loop's rate.

Registers are steered through 128-bit values: a float4 that is loaded by LDG.128 and stored by STG.128 lives in an
aligned register quad, so its .xy half sits in registers 4k,4k+1 ("class 0") and its .zw half in 4k+2,4k+3 ("class 1").
Each variant runs 160 FFMA2 per iteration in 16 runs of 10 that share the tap operand; variants differ in whether the
window operand and the accumulator operand of an instruction come from the same class."""
import sys
variants = {
    0: "diff",    # window .xy (class 0) with accumulator .zw (class 1) and vice versa: never the same class
    1: "same",    # window .xy with accumulator .xy: always the same class
    2: "half",    # alternate
    3: "lds3",    # "half" plus 3 LDS.128 per 40 FFMA2 (the FIR loop's ratio), results consumed as windows/taps
    4: "lds6",    # 6 LDS.128 per 40 FFMA2
    5: "lds1",    # 1 LDS.128 per 40 FFMA2
}
out = ['#include <cuda_runtime.h>', '#include <stdio.h>']
for v, kind in variants.items():
    out.append(f'__global__ void __launch_bounds__(128) k{v}(int iters, const float4* __restrict__ init, float4* __restrict__ sink) {{')
    out.append('  float4 A[20], W[6], T[2];')
    out.append('  for (int i = 0; i < 20; ++i) A[i] = init[threadIdx.x + 128 * i];')
    out.append('  for (int i = 0; i < 6; ++i) W[i] = init[threadIdx.x + 128 * (20 + i)];')
    out.append('  for (int i = 0; i < 2; ++i) T[i] = init[threadIdx.x + 128 * (26 + i)];')
    out.append('  float2 a[40], w[12], t[4];')
    out.append('  for (int i = 0; i < 20; ++i) { a[2*i] = make_float2(A[i].x, A[i].y); a[2*i+1] = make_float2(A[i].z, A[i].w); }')
    out.append('  for (int i = 0; i < 6; ++i) { w[2*i] = make_float2(W[i].x, W[i].y); w[2*i+1] = make_float2(W[i].z, W[i].w); }')
    out.append('  for (int i = 0; i < 2; ++i) { t[2*i] = make_float2(T[i].x, T[i].y); t[2*i+1] = make_float2(T[i].z, T[i].w); }')
    nl = {"lds3": 3, "lds6": 6, "lds1": 1}.get(kind, 0)
    if nl:
        out.append('  __shared__ float4 S[128 * 8];')
        out.append('  for (int i = 0; i < 8; ++i) S[threadIdx.x + 128 * i] = init[threadIdx.x + 128 * i];')
        out.append('  __syncthreads();')
    out.append('  for (int it = 0; it < iters; ++it) {')
    n = 0
    for run in range(16):
        if nl and run % 4 == 0:
            for j in range(nl):
                tgt = (run // 4 * nl + j)
                if tgt % 4 == 3:
                    out.append(f'    {{ float4 q = S[threadIdx.x + 128 * ((it + {tgt}) & 7)]; t[0] = make_float2(q.x, q.y); t[1] = make_float2(q.z, q.w); }}')
                else:
                    wq = tgt % 6
                    out.append(f'    {{ float4 q = S[threadIdx.x + 128 * ((it + {tgt}) & 7)]; w[{2*wq}] = make_float2(q.x, q.y); w[{2*wq+1}] = make_float2(q.z, q.w); }}')
        for r in range(10):
            acc = (run % 4) * 10 + r                      # pair index 0..39; class = acc % 2
            base = (r + run) % 6                          # window quad
            if kind == "diff":
                wi = 2 * base + (1 - acc % 2)
            elif kind == "same":
                wi = 2 * base + (acc % 2)
            else:
                wi = 2 * base + ((acc + r) % 2)
                if nl: wi = (wi + 2 * (run // 4)) % 12
            out.append(f'    a[{acc}] = __ffma2_rn(t[{run % 4}], w[{wi}], a[{acc}]);')
    out.append('  }')
    out.append('  for (int i = 0; i < 20; ++i) sink[threadIdx.x + 128 * i] = make_float4(a[2*i].x, a[2*i].y, a[2*i+1].x, a[2*i+1].y);')
    out.append('  if (iters < 0) { for (int i = 0; i < 6; ++i) sink[threadIdx.x + 128 * (20 + i)] = make_float4(w[2*i].x, w[2*i].y, w[2*i+1].x, w[2*i+1].y); }')
    out.append('}')
out.append('int main() {')
out.append('  float4 *init, *sink; cudaMalloc(&init, 128 * 32 * 16); cudaMemset(init, 0, 128 * 32 * 16); cudaMalloc(&sink, 128 * 32 * 16);')
out.append('  int sms = 148; cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); const int iters = 20000;')
for v in variants:
    out.append(f'  {{ float best = 1e9; for (int rep = 0; rep < 4; ++rep) {{ cudaEventRecord(e0); k{v}<<<sms * 4, 128>>>(iters, init, sink); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms; }}')
    out.append(f'    printf("k{v} ({variants[v]}) %.3f ms %.2f TFLOP/s\\n", best, 2.0 * 2.0 * 160.0 * iters * 128.0 * sms * 4 / best / 1e9); }}')
out.append('  return 0; }')
open('rfprobe.cu', 'w').write('\n'.join(out))
