// rb_dense.cu -- the HBM-bound passes around the FIR bank: ISD impulse mask, per-tile statistics, the
// per-utterance finalisers (mean / max-abs / norms via warp shuffles), the dense apply pass and the sparse
// impulsive-noise scatter.
//
// Reference arithmetic being restated (file:line under /root/reference/datautils/RawBoost.py):
//   normWav 20-25, LnL tail (mean removal + normWav) 67-68, ISD 76-84, SSI tail 93-96.
#include "rb_common.cuh"
#include "rb_dense.cuh"
#include "rb_finalize.cuh"

namespace rb {

namespace {

constexpr int kSparseChunks = 8;  // CTAs per utterance for the impulse kernels (grid-stride inside)

// ---- ISD impulse bit mask: bit p of row u set iff p is an impulse position of utterance u ---------
__global__ void __launch_bounds__(256)
mask_build_kernel(const int32_t* __restrict__ isd_off, const int32_t* __restrict__ isd_idx, const int32_t* __restrict__ len_arr,
                  uint32_t* __restrict__ mask, int mask_ld) {
  const int u = blockIdx.y;
  const int beg = isd_off[u], end = isd_off[u + 1], len = len_arr[u];
  uint32_t* mrow = mask + (size_t)u * mask_ld;
  for (int i = beg + blockIdx.x * 256 + threadIdx.x; i < end; i += gridDim.x * 256) {
    const int p = isd_idx[i];
    if (p >= 0 && p < len) atomicOr(mrow + (p >> 5), 1u << (p & 31));
  }
}

// ---- per-tile statistics of a waveform batch (same layout as the FIR-bank epilogue) ---------------
__global__ void __launch_bounds__(kThreads)
dense_stats_kernel(const float* __restrict__ x, const int32_t* __restrict__ len_arr, int ld, float* __restrict__ stats,
                   const uint32_t* __restrict__ mask, int mask_ld) {
  __shared__ float red[4][kStatN];
  const int u = blockIdx.y, tile = blockIdx.x, tid = threadIdx.x;
  const int len = len_arr[u];
  const int tile0 = tile * kTile;
  float* st_out = stats + ((size_t)u * gridDim.x + tile) * kStatN;
  float s_sum = 0.f, s_sq = 0.f, s_min = INFINITY, s_max = -INFINITY, s_minu = INFINITY, s_maxu = -INFINITY;
  if (tile0 < len) {
    const float* row = x + (size_t)u * ld;
    const uint32_t* mrow = mask ? mask + (size_t)u * mask_ld : nullptr;
    float4 v[kR / 4];
    uint32_t hit[kR / 4];
#pragma unroll
    for (int k = 0; k < kR / 4; ++k) {  // issue all loads first
      const int p = tile0 + 4 * (k * kThreads + tid);
      if (p + 3 < len) {
        v[k] = __ldg(reinterpret_cast<const float4*>(row + p));
      } else {
        v[k].x = (p + 0 < len) ? __ldg(row + p + 0) : 0.f;
        v[k].y = (p + 1 < len) ? __ldg(row + p + 1) : 0.f;
        v[k].z = (p + 2 < len) ? __ldg(row + p + 2) : 0.f;
        v[k].w = (p + 3 < len) ? __ldg(row + p + 3) : 0.f;
      }
      hit[k] = (mrow && p < len) ? ((__ldg(mrow + (p >> 5)) >> (p & 31)) & 0xFu) : 0u;
    }
#pragma unroll
    for (int k = 0; k < kR / 4; ++k) {
      const int p = tile0 + 4 * (k * kThreads + tid);
      const float e[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (p + q < len) {
          s_sum += e[q];
          s_sq = fmaf(e[q], e[q], s_sq);
          s_min = fminf(s_min, e[q]);
          s_max = fmaxf(s_max, e[q]);
          if (!((hit[k] >> q) & 1u)) {
            s_minu = fminf(s_minu, e[q]);
            s_maxu = fmaxf(s_maxu, e[q]);
          }
        }
      }
    }
  }
  s_sum = warp_sum(s_sum);
  s_sq = warp_sum(s_sq);
  s_min = warp_min(s_min);
  s_max = warp_max(s_max);
  s_minu = warp_min(s_minu);
  s_maxu = warp_max(s_maxu);
  const int warp = tid >> 5, lane = tid & 31;
  if (lane == 0) {
    red[warp][S_SUM] = s_sum;
    red[warp][S_SUMSQ] = s_sq;
    red[warp][S_MIN] = s_min;
    red[warp][S_MAX] = s_max;
    red[warp][S_MINU] = s_minu;
    red[warp][S_MAXU] = s_maxu;
  }
  __syncthreads();
  if (tid < kStatN) {
    float r;
    if (tid == S_SUM || tid == S_SUMSQ) r = (red[0][tid] + red[1][tid]) + (red[2][tid] + red[3][tid]);
    else if (tid == S_MIN || tid == S_MINU) r = fminf(fminf(red[0][tid], red[1][tid]), fminf(red[2][tid], red[3][tid]));
    else if (tid == S_MAX || tid == S_MAXU) r = fmaxf(fmaxf(red[0][tid], red[1][tid]), fmaxf(red[2][tid], red[3][tid]));
    else r = 0.f;
    st_out[tid] = r;
  }
}

// ---- per-utterance finaliser: one CTA per utterance (arithmetic in rb_finalize.cuh) ----------------------------------------
__global__ void __launch_bounds__(kThreads)
finalize_kernel(FinalizeArgs a) {
  const int u = blockIdx.x;
  const int n = a.len[u];
  const bool with_isd = a.isd_off != nullptr;
  const UttParams p = finalize_block(a.stats + (size_t)u * a.ntiles * kStatN, a.ntiles, n, a.center, a.always,
                                     a.raw + (size_t)u * a.ld, a.isd_idx, a.isd_fr, with_isd ? a.isd_off[u] : 0,
                                     with_isd ? a.isd_off[u + 1] : 0, with_isd, a.g_sd);
  if (threadIdx.x == 0) a.out[u] = p;
}

// ---- SSI scale (RawBoost.py:95) -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32)
ssi_finalize_kernel(const float* __restrict__ stats_x, const float* __restrict__ stats_n, int ntiles,
                    const float* __restrict__ snr_db, UttParams* __restrict__ out) {
  const int u = blockIdx.x;
  const float scale = ssi_scale_block(stats_x + (size_t)u * ntiles * kStatN, S_SUMSQ, stats_n + (size_t)u * ntiles * kStatN, ntiles,
                                      snr_db[u]);
  if (threadIdx.x == 0) {
    UttParams p;
    p.sub = 0.f;
    p.div1 = 1.f;
    p.div2 = 1.f;
    p.scale = scale;
    out[u] = p;
  }
}

// ---- dense elementwise passes ---------------------------------------------------------------------
enum { APPLY_AFFINE = 0, APPLY_SSI = 1, APPLY_SUM = 2 };

template <int MODE>
__global__ void __launch_bounds__(kThreads)
apply_kernel(const float* __restrict__ a, const float* __restrict__ b, const int32_t* __restrict__ len_arr, int ld,
             const UttParams* __restrict__ params, float* __restrict__ out) {
  const int u = blockIdx.y, tid = threadIdx.x;
  const int len = len_arr[u];
  const int tile0 = blockIdx.x * kTile;
  if (tile0 >= len) return;
  UttParams pr;
  if (MODE != APPLY_SUM) pr = params[u];
  const float* ra = a + (size_t)u * ld;
  const float* rb_ = (MODE != APPLY_AFFINE) ? b + (size_t)u * ld : nullptr;
  float* ro = out + (size_t)u * ld;
  float4 va[kR / 4], vb[kR / 4];
#pragma unroll
  for (int k = 0; k < kR / 4; ++k) {
    const int p = tile0 + 4 * (k * kThreads + tid);
    if (p + 3 < len) {
      va[k] = __ldg(reinterpret_cast<const float4*>(ra + p));
      if (MODE != APPLY_AFFINE) vb[k] = __ldg(reinterpret_cast<const float4*>(rb_ + p));
    } else {
      float ta[4] = {0.f, 0.f, 0.f, 0.f}, tb[4] = {0.f, 0.f, 0.f, 0.f};
      for (int q = 0; q < 4; ++q)
        if (p + q < len) {
          ta[q] = __ldg(ra + p + q);
          if (MODE != APPLY_AFFINE) tb[q] = __ldg(rb_ + p + q);
        }
      va[k] = make_float4(ta[0], ta[1], ta[2], ta[3]);
      vb[k] = make_float4(tb[0], tb[1], tb[2], tb[3]);
    }
  }
#pragma unroll
  for (int k = 0; k < kR / 4; ++k) {
    const int p = tile0 + 4 * (k * kThreads + tid);
    float ea[4] = {va[k].x, va[k].y, va[k].z, va[k].w};
    float eb[4] = {vb[k].x, vb[k].y, vb[k].z, vb[k].w};
    float r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (MODE == APPLY_AFFINE) r[q] = affine_value(ea[q], pr);
      else if (MODE == APPLY_SSI) r[q] = __fmaf_rn(eb[q], pr.scale, ea[q]);
      else r[q] = __fadd_rn(ea[q], eb[q]);
    }
    if (p + 3 < len) {
      reinterpret_cast<float4*>(ro + p)[0] = make_float4(r[0], r[1], r[2], r[3]);
    } else {
      for (int q = 0; q < 4; ++q)
        if (p + q < len) ro[p + q] = r[q];
    }
  }
}

// ---- sparse impulsive-noise scatter: overwrites out[p] at the impulse positions ---------------------
__global__ void __launch_bounds__(256)
isd_scatter_kernel(const float* __restrict__ raw, const int32_t* __restrict__ len_arr, int ld, const int32_t* __restrict__ isd_off,
                   const int32_t* __restrict__ isd_idx, const double* __restrict__ isd_fr, float g_sd,
                   const UttParams* __restrict__ params, float* __restrict__ out) {
  const int u = blockIdx.y;
  const int len = len_arr[u];
  const UttParams pr = params[u];
  const float* row = raw + (size_t)u * ld;
  float* ro = out + (size_t)u * ld;
  for (int i = isd_off[u] + blockIdx.x * 256 + threadIdx.x; i < isd_off[u + 1]; i += gridDim.x * 256) {
    const int p = isd_idx[i];
    if (p >= 0 && p < len) {
      const float v = __fdiv_rn(__fsub_rn(row[p], pr.sub), pr.div1);
      ro[p] = __fdiv_rn(isd_value(v, g_sd, isd_fr[i]), pr.div2);
    }
  }
}

// ---- ISD / normWav in ONE pass over HBM: one CTA per utterance -----------------------------------------------------------
// normWav(x, always) followed (optionally) by the impulse scatter and its normWav(., 0) (RawBoost.py:20-25, 76-84). Every
// quantity is a max / min, so the result does not depend on the reduction order and equals the multi-pass path bit for bit.
// The waveform is read from HBM once (statistics), read again while still L2-resident (apply) and written once; the impulse
// bit mask lives in shared memory.
constexpr int kFusedThreads = 512;

__global__ void __launch_bounds__(kFusedThreads)
isd_fused_kernel(const float* __restrict__ x, const int32_t* __restrict__ len_arr, int ld, int always,
                 const int32_t* __restrict__ isd_off, const int32_t* __restrict__ isd_idx, const double* __restrict__ isd_fr,
                 float g_sd, float* __restrict__ out) {
  extern __shared__ uint32_t smask[];  // [ceil(len/32)] impulse bit mask (unused without impulses)
  __shared__ float red[kFusedThreads / 32][2];
  __shared__ float bc[3];
  const int u = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int len = len_arr[u];
  if (len <= 0) return;
  const float* row = x + (size_t)u * ld;
  float* orow = out + (size_t)u * ld;
  const bool with_isd = isd_off != nullptr;
  const int ibeg = with_isd ? isd_off[u] : 0, iend = with_isd ? isd_off[u + 1] : 0;
  const int nwords = (len + 31) >> 5;
  if (with_isd) {
    for (int w = tid; w < nwords; w += kFusedThreads) smask[w] = 0u;
    __syncthreads();
    for (int i = ibeg + tid; i < iend; i += kFusedThreads) {
      const int p = isd_idx[i];
      if (p >= 0 && p < len) atomicOr(smask + (p >> 5), 1u << (p & 31));
    }
    __syncthreads();
  }
  // pass 1: peak over all samples and over the samples no impulse touches (NaN-propagating like numpy's amax)
  const int nchunk = (len + 3) >> 2;
  float m_all = 0.f, m_unt = 0.f;
  bool nan_all = false, nan_unt = false;
  constexpr int kU = 8;  // float4 loads in flight per thread: the pass is latency-bound otherwise
  for (int c0 = tid; c0 < nchunk; c0 += kU * kFusedThreads) {
    float4 v[kU];
#pragma unroll
    for (int k = 0; k < kU; ++k) {
      const int p = 4 * (c0 + k * kFusedThreads);
      if (p + 3 < len) {
        v[k] = __ldg(reinterpret_cast<const float4*>(row + p));
      } else {
        v[k].x = (p + 0 < len) ? __ldg(row + p + 0) : 0.f;
        v[k].y = (p + 1 < len) ? __ldg(row + p + 1) : 0.f;
        v[k].z = (p + 2 < len) ? __ldg(row + p + 2) : 0.f;
        v[k].w = 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < kU; ++k) {
      const int p = 4 * (c0 + k * kFusedThreads);
      const float e[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
      const uint32_t hit = (with_isd && p < len) ? ((smask[p >> 5] >> (p & 31)) & 0xFu) : 0u;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float a = fabsf(e[q]);  // padding lanes hold 0: neutral for a peak
        nan_all |= (a != a);
        m_all = fmaxf(m_all, a);
        if (!((hit >> q) & 1u)) {
          nan_unt |= (a != a);
          m_unt = fmaxf(m_unt, a);
        }
      }
    }
  }
  if (nan_all) m_all = NAN;
  if (nan_unt) m_unt = NAN;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m_all = nan_max(m_all, __shfl_xor_sync(0xffffffffu, m_all, o));
    m_unt = nan_max(m_unt, __shfl_xor_sync(0xffffffffu, m_unt, o));
  }
  if (lane == 0) {
    red[warp][0] = m_all;
    red[warp][1] = m_unt;
  }
  __syncthreads();
  if (tid == 0) {
    float a = red[0][0], b = red[0][1];
    for (int w = 1; w < kFusedThreads / 32; ++w) {
      a = nan_max(a, red[w][0]);
      b = nan_max(b, red[w][1]);
    }
    const float div1 = (always || a > 1.f) ? a : 1.f;  // a NaN peak: "NaN > 1" is false, like the reference
    bc[0] = div1;
    bc[1] = b / div1;  // peak of the untouched samples after the first normalisation
  }
  __syncthreads();
  const float div1 = bc[0];
  float div2 = 1.f;
  if (with_isd) {
    float mt = 0.f;
    for (int i = ibeg + tid; i < iend; i += kFusedThreads) {
      const int p = isd_idx[i];
      if (p >= 0 && p < len) mt = fmaxf(mt, fabsf(isd_value(__fdiv_rn(__ldg(row + p), div1), g_sd, isd_fr[i])));
    }
    mt = warp_max(mt);
    __syncthreads();
    if (lane == 0) red[warp][0] = mt;
    __syncthreads();
    if (tid == 0) {
      float m2 = bc[1];
      for (int w = 0; w < kFusedThreads / 32; ++w) m2 = fmaxf(m2, red[w][0]);
      bc[2] = (m2 > 1.f) ? m2 : 1.f;
    }
    __syncthreads();
    div2 = bc[2];
  }
  // pass 2: out = (x / div1) / div2 (division by 1 skipped: identity), then the impulses
  const int ndiv = (div1 != 1.f) + (div2 != 1.f);
  const float dv = (div1 != 1.f) ? div1 : div2;
  auto nrm = [&](float e) {
    if (ndiv == 0) return e;
    if (ndiv == 1) return __fdiv_rn(e, dv);
    return __fdiv_rn(__fdiv_rn(e, div1), div2);
  };
  for (int c0 = tid; c0 < nchunk; c0 += kU * kFusedThreads) {
    float4 v[kU];
#pragma unroll
    for (int k = 0; k < kU; ++k) {
      const int p = 4 * (c0 + k * kFusedThreads);
      if (p + 3 < len) {
        v[k] = __ldg(reinterpret_cast<const float4*>(row + p));
      } else {
        v[k].x = (p + 0 < len) ? __ldg(row + p + 0) : 0.f;
        v[k].y = (p + 1 < len) ? __ldg(row + p + 1) : 0.f;
        v[k].z = (p + 2 < len) ? __ldg(row + p + 2) : 0.f;
        v[k].w = 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < kU; ++k) {
      const int p = 4 * (c0 + k * kFusedThreads);
      const float4 r = make_float4(nrm(v[k].x), nrm(v[k].y), nrm(v[k].z), nrm(v[k].w));
      if (p + 3 < len) {
        *reinterpret_cast<float4*>(orow + p) = r;
      } else {
        if (p + 0 < len) orow[p + 0] = r.x;
        if (p + 1 < len) orow[p + 1] = r.y;
        if (p + 2 < len) orow[p + 2] = r.z;
      }
    }
  }
  if (with_isd) {
    __syncthreads();  // impulse positions overwrite what the dense pass just stored
    for (int i = ibeg + tid; i < iend; i += kFusedThreads) {
      const int p = isd_idx[i];
      if (p >= 0 && p < len) orow[p] = __fdiv_rn(isd_value(__fdiv_rn(__ldg(row + p), div1), g_sd, isd_fr[i]), div2);
    }
  }
}

}  // namespace

int launch_mask_build(const int32_t* isd_off, const int32_t* isd_idx, const int32_t* len, int B, uint32_t* mask,
                      int mask_ld, cudaStream_t st) {
  if (B <= 0) return RB_OK;
  RB_CUDA(cudaMemsetAsync(mask, 0, (size_t)B * mask_ld * sizeof(uint32_t), st));
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = min(65535, B - b0);
    mask_build_kernel<<<dim3(kSparseChunks, nb), 256, 0, st>>>(isd_off + b0, isd_idx, len + b0, mask + (size_t)b0 * mask_ld, mask_ld);
    RB_LAUNCH_CHECK();
  }
  return RB_OK;
}

int launch_dense_stats(const float* x, const int32_t* len, int B, int ld, float* stats, const uint32_t* mask, int mask_ld,
                       cudaStream_t st) {
  if (B <= 0) return RB_OK;
  const int ntiles = tiles_for(ld);
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = min(65535, B - b0);
    dense_stats_kernel<<<dim3(ntiles, nb), kThreads, 0, st>>>(x + (size_t)b0 * ld, len + b0, ld, stats + (size_t)b0 * ntiles * kStatN,
                                                            mask ? mask + (size_t)b0 * mask_ld : nullptr, mask_ld);
    RB_LAUNCH_CHECK();
  }
  return RB_OK;
}

int launch_finalize(const FinalizeArgs& args, int B, cudaStream_t st) {
  if (B <= 0) return RB_OK;
  finalize_kernel<<<B, kThreads, 0, st>>>(args);
  RB_LAUNCH_CHECK();
  return RB_OK;
}

int launch_ssi_finalize(const float* stats_x, const float* stats_n, int ntiles, const float* snr_db, UttParams* out, int B,
                        cudaStream_t st) {
  if (B <= 0) return RB_OK;
  ssi_finalize_kernel<<<B, 32, 0, st>>>(stats_x, stats_n, ntiles, snr_db, out);
  RB_LAUNCH_CHECK();
  return RB_OK;
}

template <int MODE>
static int launch_apply_mode(const float* a, const float* b, const int32_t* len, int B, int ld, const UttParams* params,
                             float* out, cudaStream_t st) {
  if (B <= 0) return RB_OK;
  const int ntiles = tiles_for(ld);
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = min(65535, B - b0);
    apply_kernel<MODE><<<dim3(ntiles, nb), kThreads, 0, st>>>(a + (size_t)b0 * ld, b ? b + (size_t)b0 * ld : nullptr, len + b0, ld,
                                                            params ? params + b0 : nullptr, out + (size_t)b0 * ld);
    RB_LAUNCH_CHECK();
  }
  return RB_OK;
}

int launch_apply_affine(const float* in, const int32_t* len, int B, int ld, const UttParams* params, float* out, cudaStream_t st) {
  return launch_apply_mode<APPLY_AFFINE>(in, nullptr, len, B, ld, params, out, st);
}
int launch_apply_ssi(const float* x, const float* noise, const int32_t* len, int B, int ld, const UttParams* params, float* out,
                     cudaStream_t st) {
  return launch_apply_mode<APPLY_SSI>(x, noise, len, B, ld, params, out, st);
}
int launch_apply_sum(const float* a, const float* b, const int32_t* len, int B, int ld, float* out, cudaStream_t st) {
  return launch_apply_mode<APPLY_SUM>(a, b, len, B, ld, nullptr, out, st);
}

int launch_isd_scatter(const float* raw, const int32_t* len, int B, int ld, const int32_t* isd_off, const int32_t* isd_idx,
                       const double* isd_fr, float g_sd, const UttParams* params, float* out, cudaStream_t st) {
  if (B <= 0) return RB_OK;
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = min(65535, B - b0);
    isd_scatter_kernel<<<dim3(kSparseChunks, nb), 256, 0, st>>>(raw + (size_t)b0 * ld, len + b0, ld, isd_off + b0, isd_idx, isd_fr, g_sd,
                                                                params + b0, out + (size_t)b0 * ld);
    RB_LAUNCH_CHECK();
  }
  return RB_OK;
}

}  // namespace rb

namespace rb {
// normWav(x, always) [+ ISD] in one kernel. Returns RB_ERR_UNSUPPORTED when the impulse mask does not fit shared memory
// (utterances beyond ~1.8 M samples); the caller then takes the multi-pass path.
int launch_isd_fused(const float* x, const int32_t* len, int B, int ld, int always, const int32_t* isd_off, const int32_t* isd_idx,
                     const double* isd_fr, float g_sd, float* out, cudaStream_t st) {
  if (B <= 0) return RB_OK;
  const size_t smem = isd_off ? ((size_t)(ld + 31) / 32) * 4 : 0;
  if (smem > 200 * 1024) return RB_ERR_UNSUPPORTED;
  if (smem > 48 * 1024) RB_CUDA(cudaFuncSetAttribute(isd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  isd_fused_kernel<<<B, kFusedThreads, smem, st>>>(x, len, ld, always, isd_off, isd_idx, isd_fr, g_sd, out);
  RB_LAUNCH_CHECK();
  return RB_OK;
}
}  // namespace rb
