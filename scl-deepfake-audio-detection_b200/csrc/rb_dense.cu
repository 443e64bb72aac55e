// rb_dense.cu -- the HBM-bound passes around the FIR bank: ISD impulse mask, per-tile statistics, the
// per-utterance finalisers (mean / max-abs / norms via warp shuffles), the dense apply pass and the sparse
// impulsive-noise scatter.
//
// Reference arithmetic being restated (file:line under /root/reference/datautils/RawBoost.py):
//   normWav 20-25, LnL tail (mean removal + normWav) 67-68, ISD 76-84, SSI tail 93-96.
#include "rb_common.cuh"
#include "rb_dense.cuh"

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

// The impulsive-noise value at one position, with the reference's exact operation order and precisions
// (RawBoost.py:81-82 on float32 input): t = fl32(g_sd*x), r = fl64(t*f_r), y = fl32(fl64(x + r)).
__device__ __forceinline__ float isd_value(float v, float g_sd, double fr) {
  const float t = __fmul_rn(g_sd, v);
  const double r = __dmul_rn((double)t, fr);
  return (float)__dadd_rn((double)v, r);
}

// ---- per-utterance finaliser: one CTA per utterance ----------------------------------------------
__global__ void __launch_bounds__(kThreads)
finalize_kernel(FinalizeArgs a) {
  __shared__ double red_sum[4];
  __shared__ float red_f[4][4];
  __shared__ float bc[4];
  const int u = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = a.len[u];
  const float* st = a.stats + (size_t)u * a.ntiles * kStatN;
  double sum = 0.0;
  float mn = INFINITY, mx = -INFINITY, mnu = INFINITY, mxu = -INFINITY;
  for (int t = tid; t < a.ntiles; t += kThreads) {
    sum += (double)st[t * kStatN + S_SUM];
    mn = fminf(mn, st[t * kStatN + S_MIN]);
    mx = fmaxf(mx, st[t * kStatN + S_MAX]);
    mnu = fminf(mnu, st[t * kStatN + S_MINU]);
    mxu = fmaxf(mxu, st[t * kStatN + S_MAXU]);
  }
  sum = warp_sum(sum);
  mn = warp_min(mn);
  mx = warp_max(mx);
  mnu = warp_min(mnu);
  mxu = warp_max(mxu);
  if (lane == 0) {
    red_sum[warp] = sum;
    red_f[warp][0] = mn;
    red_f[warp][1] = mx;
    red_f[warp][2] = mnu;
    red_f[warp][3] = mxu;
  }
  __syncthreads();
  if (tid == 0) {
    const double s = (red_sum[0] + red_sum[1]) + (red_sum[2] + red_sum[3]);
    const float fmn = fminf(fminf(red_f[0][0], red_f[1][0]), fminf(red_f[2][0], red_f[3][0]));
    const float fmx = fmaxf(fmaxf(red_f[0][1], red_f[1][1]), fmaxf(red_f[2][1], red_f[3][1]));
    const float fmnu = fminf(fminf(red_f[0][2], red_f[1][2]), fminf(red_f[2][2], red_f[3][2]));
    const float fmxu = fmaxf(fmaxf(red_f[0][3], red_f[1][3]), fmaxf(red_f[2][3], red_f[3][3]));
    const float sub = (a.center && n > 0) ? (float)(s / (double)n) : 0.f;
    float m1 = (n > 0) ? fmaxf(fabsf(fmx - sub), fabsf(fmn - sub)) : 0.f;
    const float div1 = (n > 0 && (a.always || m1 > 1.f)) ? m1 : 1.f;
    // peak of the untouched samples after the first normalisation (fp32 division is monotone, so the peak of
    // the quotients is the quotient of the peak)
    float mu = 0.f;
    if (fmnu <= fmxu) mu = fmaxf(fabsf(fmxu - sub), fabsf(fmnu - sub)) / div1;
    bc[0] = sub;
    bc[1] = div1;
    bc[2] = mu;
  }
  __syncthreads();
  const float sub = bc[0], div1 = bc[1];
  float div2 = 1.f;
  if (a.isd_off) {
    const float* row = a.raw + (size_t)u * a.ld;
    float mt = 0.f;
    for (int i = a.isd_off[u] + tid; i < a.isd_off[u + 1]; i += kThreads) {
      const int p = a.isd_idx[i];
      if (p >= 0 && p < n) {
        const float v = (row[p] - sub) / div1;
        mt = fmaxf(mt, fabsf(isd_value(v, a.g_sd, a.isd_fr[i])));
      }
    }
    mt = warp_max(mt);
    __syncthreads();
    if (lane == 0) red_f[warp][0] = mt;
    __syncthreads();
    if (tid == 0) {
      const float m2 = fmaxf(bc[2], fmaxf(fmaxf(red_f[0][0], red_f[1][0]), fmaxf(red_f[2][0], red_f[3][0])));
      div2 = (m2 > 1.f) ? m2 : 1.f;
    }
  }
  if (tid == 0) {
    UttParams p;
    p.sub = sub;
    p.div1 = div1;
    p.div2 = div2;
    p.scale = 0.f;
    a.out[u] = p;
  }
}

// ---- SSI scale: ||x||_2 / (||coloured noise||_2 * 10^(snr/20))  (RawBoost.py:95) ------------------
__global__ void __launch_bounds__(32)
ssi_finalize_kernel(const float* __restrict__ stats_x, const float* __restrict__ stats_n, int ntiles,
                    const float* __restrict__ snr_db, UttParams* __restrict__ out) {
  const int u = blockIdx.x, lane = threadIdx.x;
  double sx = 0.0, sn = 0.0;
  for (int t = lane; t < ntiles; t += 32) {
    sx += (double)stats_x[((size_t)u * ntiles + t) * kStatN + S_SUMSQ];
    sn += (double)stats_n[((size_t)u * ntiles + t) * kStatN + S_SUMSQ];
  }
  sx = warp_sum(sx);
  sn = warp_sum(sn);
  if (lane == 0) {
    UttParams p;
    p.sub = 0.f;
    p.div1 = 1.f;
    p.div2 = 1.f;
    p.scale = (float)(sqrt(sx) / (sqrt(sn) * pow(10.0, 0.05 * (double)snr_db[u])));
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
      if (MODE == APPLY_AFFINE) r[q] = __fdiv_rn(__fdiv_rn(__fsub_rn(ea[q], pr.sub), pr.div1), pr.div2);
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
