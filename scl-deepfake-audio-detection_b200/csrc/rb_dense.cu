// rb_dense.cu -- the HBM-bound operators: normWav, the stand-alone impulsive noise (ISD) and the two-branch sum of algo 8.
//
// Reference arithmetic being restated (file:line under /root/reference/datautils/RawBoost.py):
//   normWav 20-25:  m = max|x|;  x / m  if (always or m > 1) else x
//   ISD     76-84:  y = copy(x); y[p] = x[p] + g_sd * x[p] * f_r;  normWav(y, 0)          (x itself is NOT normalised first)
//   algo 8  (asvspoof_2019_augall_3.py:425-432):  normWav(LnL(x) + ISD(x), 0)
//
// Design: normWav is the identity unless the peak exceeds 1 (or `always`), so nothing has to wait for the peak. One streaming
// pass over (utterance, 4096-sample tile) CTAs copies every tile to the output at copy speed while it takes the tile's peak;
// a per-utterance arrival counter elects the CTA that finishes an utterance's last tile, and only that CTA
//   * applies the utterance's impulses (gathering x[p] and scattering y[p] through L2, where the row was streamed a moment ago),
//   * folds their magnitudes into the peak, and
//   * only if the reference would divide (peak > 1 or `always`) rescales the row in place while it is still L2-resident.
// HBM therefore sees each sample once in and once out; many small CTAs per SM (8 x 256 threads, 16 KB in flight each) keep the
// memory pipeline full, where the previous one-CTA-per-utterance kernel (load everything -> reduce -> store) sat at 35-40 %.
// All reductions are maxima of |.|, taken on the bit patterns (for non-negative floats the unsigned order is the numeric
// order and every NaN sorts above +inf), so the result is independent of the reduction order, bit-exact against numpy, and
// NaN propagates like np.amax.
#include "rb_common.cuh"
#include "rb_dense.cuh"
#include "rb_finalize.cuh"

namespace rb {

namespace {

// ---- ISD impulse bit mask: bit p of row u set iff p is an impulse position of utterance u ---------
// One CTA per utterance assembles the row in shared memory (no global atomics, no memset) and writes it out once.
constexpr int kMaskThreads = 256;
constexpr int kMaskSmemWords = 12 * 1024;  // rows of up to 393216 samples; longer ones take the global-atomic kernel

__global__ void __launch_bounds__(kMaskThreads)
mask_build_kernel(const int32_t* __restrict__ isd_off, const int32_t* __restrict__ isd_idx, const int32_t* __restrict__ len_arr,
                  uint32_t* __restrict__ mask, int mask_ld) {
  extern __shared__ uint32_t smask[];
  const int u = blockIdx.x, tid = threadIdx.x;
  const int beg = isd_off[u], end = isd_off[u + 1], len = len_arr[u];
  for (int w = tid; w < mask_ld; w += kMaskThreads) smask[w] = 0u;
  __syncthreads();
  constexpr int kU = 4;  // positions in flight per thread
  for (int i0 = beg + tid; i0 < end; i0 += kU * kMaskThreads) {
    int p[kU];
#pragma unroll
    for (int k = 0; k < kU; ++k) {
      const int i = i0 + k * kMaskThreads;
      p[k] = (i < end) ? __ldg(isd_idx + i) : -1;
    }
#pragma unroll
    for (int k = 0; k < kU; ++k)
      if (p[k] >= 0 && p[k] < len) atomicOr(smask + (p[k] >> 5), 1u << (p[k] & 31));
  }
  __syncthreads();
  uint32_t* mrow = mask + (size_t)u * mask_ld;
  for (int w = tid; w < mask_ld; w += kMaskThreads) mrow[w] = smask[w];
}

__global__ void __launch_bounds__(256)
mask_build_global_kernel(const int32_t* __restrict__ isd_off, const int32_t* __restrict__ isd_idx, const int32_t* __restrict__ len_arr,
                         uint32_t* __restrict__ mask, int mask_ld) {
  const int u = blockIdx.y;
  const int beg = isd_off[u], end = isd_off[u + 1], len = len_arr[u];
  uint32_t* mrow = mask + (size_t)u * mask_ld;
  for (int i = beg + blockIdx.x * 256 + threadIdx.x; i < end; i += gridDim.x * 256) {
    const int p = isd_idx[i];
    if (p >= 0 && p < len) atomicOr(mrow + (p >> 5), 1u << (p & 31));
  }
}

// ---- the streaming pass ---------------------------------------------------------------------------------------------------
#ifndef RB_STREAM_THREADS
#define RB_STREAM_THREADS 256
#endif
#ifndef RB_STREAM_CHUNKS
#define RB_STREAM_CHUNKS 4
#endif
constexpr int kSThreads = RB_STREAM_THREADS;     // threads per CTA
constexpr int kSU = RB_STREAM_CHUNKS;            // float4 chunks per thread, all in flight at once
constexpr int kSTile = kSThreads * kSU * 4;      // samples per CTA (4096)

__device__ __forceinline__ uint32_t abs_bits(float v) { return __float_as_uint(v) & 0x7fffffffu; }

// max over the CTA of a uint32 held by every thread; returned in every thread. `scratch` holds one word per warp.
__device__ __forceinline__ uint32_t block_umax(uint32_t v, uint32_t* scratch) {
  const int tid = threadIdx.x;
  v = __reduce_max_sync(0xffffffffu, v);
  __syncthreads();  // scratch may still be read by a previous call
  if ((tid & 31) == 0) scratch[tid >> 5] = v;
  __syncthreads();
  uint32_t r = scratch[0];
#pragma unroll
  for (int w = 1; w < kSThreads / 32; ++w) r = max(r, scratch[w]);
  return r;
}

struct StreamArgs {
  const float* a;            // [B][ld] input
  const float* b;            // [B][ld] second addend (kSum) or nullptr
  const int32_t* len;        // [B]
  int ld;
  int ntiles;                // tiles per row = ceil(ld / kSTile)
  int always;                // normWav(., 1)
  const uint32_t* mask;      // [B][mask_ld] impulse bit mask (kIsd)
  int mask_ld;
  const int32_t* isd_off;    // kIsd: impulses of utterance u are [off[u], off[u+1])
  const int32_t* isd_idx;
  const double* isd_fr;
  float g_sd;
  float* out;                // [B][ld]; may equal a when !kSum (in place: the copy is skipped)
  uint32_t* tile_peak;       // [B][ntiles] bit patterns of the per-tile peaks
  uint32_t* counters;        // [B] arrival counters, zero on entry, left zero
};

#ifndef RB_STREAM_MIN_BLOCKS
#define RB_STREAM_MIN_BLOCKS 6
#endif
template <bool kIsd, bool kSum>
__global__ void __launch_bounds__(kSThreads, RB_STREAM_MIN_BLOCKS)
norm_stream_kernel(const StreamArgs s) {
  __shared__ uint32_t scratch[kSThreads / 32];
  __shared__ int s_last;
  const int u = blockIdx.x / s.ntiles, tile = blockIdx.x - u * s.ntiles;
  const int tid = threadIdx.x;
  const int len = s.len[u];
  const int tile0 = tile * kSTile;
  if (tile0 >= len) return;
  const float* ra = s.a + (size_t)u * s.ld;
  const float* rb_ = kSum ? s.b + (size_t)u * s.ld : nullptr;
  float* ro = s.out + (size_t)u * s.ld;
  const bool copy = kSum || (ro != ra);

  float4 v[kSU];
  uint32_t hit[kSU];
#pragma unroll
  for (int k = 0; k < kSU; ++k) {  // every load of the tile is issued before the first use
    const int p = tile0 + 4 * (k * kSThreads + tid);
    v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    hit[k] = 0u;
    if (p + 3 < len) {
      v[k] = __ldg(reinterpret_cast<const float4*>(ra + p));
      if (kSum) {
        const float4 w = __ldg(reinterpret_cast<const float4*>(rb_ + p));
        v[k] = make_float4(__fadd_rn(v[k].x, w.x), __fadd_rn(v[k].y, w.y), __fadd_rn(v[k].z, w.z), __fadd_rn(v[k].w, w.w));
      }
    } else if (p < len) {  // the ragged last chunk; zeros beyond the end are neutral for the peak
      float e[4] = {0.f, 0.f, 0.f, 0.f};
      for (int q = 0; q < 4; ++q)
        if (p + q < len) e[q] = kSum ? __fadd_rn(__ldg(ra + p + q), __ldg(rb_ + p + q)) : __ldg(ra + p + q);
      v[k] = make_float4(e[0], e[1], e[2], e[3]);
    }
    if (kIsd && p < len) hit[k] = (__ldg(s.mask + (size_t)u * s.mask_ld + (p >> 5)) >> (p & 31)) & 0xFu;
  }
  uint32_t m = 0u;  // peak of the samples no impulse touches
#pragma unroll
  for (int k = 0; k < kSU; ++k) {
    const int p = tile0 + 4 * (k * kSThreads + tid);
    if (!(hit[k] & 1u)) m = max(m, abs_bits(v[k].x));
    if (!(hit[k] & 2u)) m = max(m, abs_bits(v[k].y));
    if (!(hit[k] & 4u)) m = max(m, abs_bits(v[k].z));
    if (!(hit[k] & 8u)) m = max(m, abs_bits(v[k].w));
    if (copy) {
      if (p + 3 < len) {
        *reinterpret_cast<float4*>(ro + p) = v[k];
      } else if (p < len) {
        ro[p] = v[k].x;
        if (p + 1 < len) ro[p + 1] = v[k].y;
        if (p + 2 < len) ro[p + 2] = v[k].z;
      }
    }
  }
  m = block_umax(m, scratch);
  const int nact = (len + kSTile - 1) / kSTile;  // tiles of this utterance that do work
  if (tid == 0) s.tile_peak[(size_t)u * s.ntiles + tile] = m;
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned prev = atomicAdd(s.counters + u, 1u);
    s_last = (prev == (unsigned)(nact - 1));
    if (s_last) s.counters[u] = 0u;  // ready for the next launch
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();

  // ---- the utterance is complete in `out`: impulses, peak, conditional rescale (all through L2) -----------------------------
  uint32_t pk = 0u;
  for (int t = tid; t < nact; t += kSThreads) pk = max(pk, __ldcg(s.tile_peak + (size_t)u * s.ntiles + t));
  if (kIsd) {
    const int ibeg = s.isd_off[u], iend = s.isd_off[u + 1];
    constexpr int kU = 4;  // impulses in flight per thread
    for (int i0 = ibeg + tid; i0 < iend; i0 += kU * kSThreads) {
      int p[kU];
      double fr[kU];
      float xv[kU];
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        const int i = i0 + k * kSThreads;
        p[k] = (i < iend) ? __ldg(s.isd_idx + i) : -1;
        if (p[k] >= len) p[k] = -1;
      }
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        fr[k] = (p[k] >= 0) ? __ldg(s.isd_fr + i0 + k * kSThreads) : 0.0;
        xv[k] = (p[k] >= 0) ? __ldcg(ra + p[k]) : 0.f;
      }
#pragma unroll
      for (int k = 0; k < kU; ++k) {
        if (p[k] >= 0) {
          const float t = isd_value(xv[k], s.g_sd, fr[k]);
          pk = max(pk, abs_bits(t));
          __stcg(ro + p[k], t);
        }
      }
    }
  }
  pk = block_umax(pk, scratch);  // (its barriers also order the impulse stores before the rescale below)
  const float peak = __uint_as_float(pk);
  if (!(s.always || peak > 1.f)) return;  // a NaN peak: "NaN > 1" is false, like the reference
  const int nchunk = (len + 3) >> 2;
  constexpr int kRU = 4;  // chunks in flight per thread: the reads come from L2
  for (int c0 = tid; c0 < nchunk; c0 += kRU * kSThreads) {
    float4 w[kRU];
#pragma unroll
    for (int k = 0; k < kRU; ++k) {
      const int p = 4 * (c0 + k * kSThreads);
      w[k] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p + 3 < len) {
        w[k] = __ldcg(reinterpret_cast<const float4*>(ro + p));
      } else if (p < len) {
        w[k].x = __ldcg(ro + p);
        if (p + 1 < len) w[k].y = __ldcg(ro + p + 1);
        if (p + 2 < len) w[k].z = __ldcg(ro + p + 2);
      }
    }
#pragma unroll
    for (int k = 0; k < kRU; ++k) {
      const int p = 4 * (c0 + k * kSThreads);
      const float4 r = make_float4(__fdiv_rn(w[k].x, peak), __fdiv_rn(w[k].y, peak), __fdiv_rn(w[k].z, peak), __fdiv_rn(w[k].w, peak));
      if (p + 3 < len) {
        *reinterpret_cast<float4*>(ro + p) = r;
      } else if (p < len) {
        ro[p] = r.x;
        if (p + 1 < len) ro[p + 1] = r.y;
        if (p + 2 < len) ro[p + 2] = r.z;
      }
    }
  }
}

}  // namespace

int stream_tiles_for(int ld) { return (ld + kSTile - 1) / kSTile; }

int launch_mask_build(const int32_t* isd_off, const int32_t* isd_idx, const int32_t* len, int B, uint32_t* mask,
                      int mask_ld, cudaStream_t st) {
  if (B <= 0) return RB_OK;
  if (mask_ld <= kMaskSmemWords) {
    mask_build_kernel<<<B, kMaskThreads, (size_t)mask_ld * sizeof(uint32_t), st>>>(isd_off, isd_idx, len, mask, mask_ld);
    RB_LAUNCH_CHECK();
    return RB_OK;
  }
  RB_CUDA(cudaMemsetAsync(mask, 0, (size_t)B * mask_ld * sizeof(uint32_t), st));
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = min(65535, B - b0);
    mask_build_global_kernel<<<dim3(8, nb), 256, 0, st>>>(isd_off + b0, isd_idx, len + b0, mask + (size_t)b0 * mask_ld, mask_ld);
    RB_LAUNCH_CHECK();
  }
  return RB_OK;
}

int launch_norm_stream(const float* a, const float* b, const int32_t* len, int B, int ld, int always, const uint32_t* mask,
                       int mask_ld, const int32_t* isd_off, const int32_t* isd_idx, const double* isd_fr, float g_sd, float* out,
                       uint32_t* tile_peak, uint32_t* counters, cudaStream_t st) {
  if (B <= 0 || ld <= 0) return RB_OK;
  const bool isd = isd_off != nullptr;
  if (isd && (!mask || !isd_idx || !isd_fr || b)) return RB_ERR_INVALID_ARG;  // impulses apply to a single input
  if (!a || !len || !out || !tile_peak || !counters) return RB_ERR_INVALID_ARG;
  RB_CUDA(cudaMemsetAsync(counters, 0, (size_t)B * sizeof(uint32_t), st));
  StreamArgs s;
  s.ld = ld;
  s.ntiles = stream_tiles_for(ld);
  s.always = always;
  s.mask_ld = mask_ld;
  s.isd_idx = isd_idx;
  s.isd_fr = isd_fr;
  s.g_sd = g_sd;
  const int per = max(1, (int)(0x7fffffff / (long long)s.ntiles));  // utterances per launch (grid.x < 2^31)
  for (int b0 = 0; b0 < B; b0 += per) {
    const int nb = min(per, B - b0);
    s.a = a + (size_t)b0 * ld;
    s.b = b ? b + (size_t)b0 * ld : nullptr;
    s.len = len + b0;
    s.mask = mask ? mask + (size_t)b0 * mask_ld : nullptr;
    s.isd_off = isd ? isd_off + b0 : nullptr;
    s.out = out + (size_t)b0 * ld;
    s.tile_peak = tile_peak + (size_t)b0 * s.ntiles;
    s.counters = counters + b0;
    const unsigned grid = (unsigned)nb * (unsigned)s.ntiles;
    if (isd) norm_stream_kernel<true, false><<<grid, kSThreads, 0, st>>>(s);
    else if (b) norm_stream_kernel<false, true><<<grid, kSThreads, 0, st>>>(s);
    else norm_stream_kernel<false, false><<<grid, kSThreads, 0, st>>>(s);
    RB_LAUNCH_CHECK();
  }
  return RB_OK;
}

}  // namespace rb
