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
#include <stdlib.h>

#include <algorithm>

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
// Grid: per utterance `ntiles` tile CTAs followed by ONE finisher CTA (block index u * (ntiles + 1) + j). Tile CTAs never wait:
// load, peak, store, release-increment the utterance's arrival counter, exit. The finisher spins (one thread, acquire loads)
// until all tiles of its utterance have arrived -- they were dispatched before it, so they always make progress -- and then
// owns the row: impulses, exact peak, conditional rescale, all through L2.
#ifndef RB_STREAM_THREADS
#define RB_STREAM_THREADS 512
#endif
#ifndef RB_STREAM_CHUNKS
#define RB_STREAM_CHUNKS 4
#endif
#ifndef RB_STREAM_SUBTILES
#define RB_STREAM_SUBTILES 4
#endif
#ifndef RB_STREAM_MIN_BLOCKS
#define RB_STREAM_MIN_BLOCKS 2
#endif
#ifndef RB_STREAM_IMP_U
#define RB_STREAM_IMP_U 8
#endif
#ifndef RB_STREAM_RESCALE_U
#define RB_STREAM_RESCALE_U 8
#endif
constexpr int kSThreads = RB_STREAM_THREADS;     // threads per CTA
constexpr int kSU = RB_STREAM_CHUNKS;            // float4 chunks per thread and sub-tile
constexpr int kSub = RB_STREAM_SUBTILES;         // sub-tiles per tile CTA (two of them in flight at any time)
constexpr int kSTile = kSThreads * kSU * kSub * 4;  // samples per tile CTA (32768)
constexpr uint32_t kInfBits = 0x7f800000u;

__device__ __forceinline__ uint32_t abs_bits(float v) { return __float_as_uint(v) & 0x7fffffffu; }

// Barrier over a team of kT threads: the whole CTA (kBar == 0, __syncthreads) or a named barrier shared by the warps of a team.
template <int kT, int kBar>
__device__ __forceinline__ void team_sync() {
  if (kBar == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"n"(kBar), "n"(kT) : "memory");
}

// max over a team (kT threads, team-local thread index tid) of a uint32 held by every thread; returned in every thread.
// `scratch`: one word per warp of the team.
template <int kT, int kBar>
__device__ __forceinline__ uint32_t block_umax(uint32_t v, uint32_t* scratch, int tid) {
  v = __reduce_max_sync(0xffffffffu, v);
  team_sync<kT, kBar>();  // scratch may still be read by a previous call
  if ((tid & 31) == 0) scratch[tid >> 5] = v;
  team_sync<kT, kBar>();
  uint32_t r = scratch[0];
#pragma unroll
  for (int w = 1; w < kT / 32; ++w) r = max(r, scratch[w]);
  return r;
}

__device__ __forceinline__ void red_release_add(uint32_t* addr, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}
// Polling load: relaxed (straight to L2, no cache maintenance). An acquire load invalidates the SM's whole L1 every time it
// is issued (CCTL.IVALL) -- hundreds of thousands of times per launch when many finishers wait; the acquire is done once,
// by a fence, after the loop.
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* addr) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}

struct StreamArgs {
  const float* a;            // [B][ld] input
  const float* b;            // [B][ld] second addend (kSum) or nullptr
  const int32_t* len;        // [B]
  int ld;
  int ntiles;                // tiles per row = ceil(ld / kSTile)
  int always;                // normWav(., 1)
  const int32_t* isd_off;    // kIsd: impulses of utterance u are [off[u], off[u+1])
  const int32_t* isd_idx;
  const double* isd_fr;
  float g_sd;
  float* out;                // [B][ld]; may equal a when !kSum (in place: the copy is skipped)
  uint2* state;              // [B] {peak bits, tiles arrived}; zero on entry, left zero
  int dbg_B;                 // (RB_STREAM_DEBUG) batch size: timestamps are written behind the state array
};

// ---- the finisher of one utterance: impulses, exact peak, conditional rescale (all through L2) ---------------------------------
// Called by every thread of a team of kT threads (team-local index `tid`) once `nact` tiles of utterance u are (or are about to be) in `out`.
template <bool kIsd, int kT, int kBar, bool kEarlyGather>
__device__ __forceinline__ void finish_row(const StreamArgs& s, int u, int len, int nact, uint32_t* scratch, int tid) {
  const float* ra = s.a + (size_t)u * s.ld;
  float* ro = s.out + (size_t)u * s.ld;
  uint32_t* st_peak = &s.state[u].x;
  uint32_t* st_count = &s.state[u].y;
  // Impulses, first round (all of them for a typical utterance). What can be fetched before the row is complete is fetched
  // before: with kEarlyGather (a finisher dispatched right behind its row's tiles: the input row is in flight or in L2) the
  // input samples too, so that only the stores wait for the tiles; otherwise (a team that claimed its row long before it is
  // streamed) only the coalesced position / gain streams -- gathering the samples now would be 32-byte random reads from
  // HBM, which the streaming pass pays for.
  constexpr int kU = RB_STREAM_IMP_U;  // impulses in flight per thread
  const int ibeg = kIsd ? s.isd_off[u] : 0, iend = kIsd ? s.isd_off[u + 1] : 0;
  uint32_t mt = 0u, mx = 0u;  // largest new magnitude / largest magnitude an impulse replaced
  int p0[kU];
  double fr0[kU];
  float t0[kU];
  auto load_plan = [&](int i0, int (&p)[kU], double (&fr)[kU]) {
#pragma unroll
    for (int k = 0; k < kU; ++k) {
      const int i = i0 + k * kT;
      p[k] = (i < iend) ? __ldg(s.isd_idx + i) : -1;
      if (p[k] >= len) p[k] = -1;
    }
#pragma unroll
    for (int k = 0; k < kU; ++k) fr[k] = (p[k] >= 0) ? __ldg(s.isd_fr + i0 + k * kT) : 0.0;
  };
  auto evaluate = [&](const int (&p)[kU], const double (&fr)[kU], float (&t)[kU]) {
    float xv[kU];
#pragma unroll
    for (int k = 0; k < kU; ++k) xv[k] = (p[k] >= 0) ? __ldcg(ra + p[k]) : 0.f;
#pragma unroll
    for (int k = 0; k < kU; ++k) {
      t[k] = 0.f;
      if (p[k] >= 0) {
        t[k] = isd_value(xv[k], s.g_sd, fr[k]);  // y[p] = x[p] + g_sd * x[p] * f_r  (RawBoost.py:81-82)
        mx = max(mx, abs_bits(xv[k]));
        mt = max(mt, abs_bits(t[k]));
      }
    }
  };
  if (kIsd) {
    load_plan(ibeg + tid, p0, fr0);
    if (kEarlyGather) evaluate(p0, fr0, t0);
  }
  if (tid == 0) {
    uint32_t spins = 0;
    while (ld_relaxed(st_count) < (uint32_t)nact) {
      __nanosleep(100);
      if (++spins > (1u << 25)) __trap();  // a row that never completes becomes a launch error, not a hung GPU
    }
    __threadfence();  // acquire: everything the arriving tiles released is visible from here on (cumulative through the barrier)
#ifdef RB_STREAM_DEBUG
    {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
      reinterpret_cast<unsigned long long*>((uint2*)s.state + s.dbg_B + 1)[4 * (size_t)u + 1] = t1;
    }
#endif
  }
  team_sync<kT, kBar>();
  const uint32_t M = __ldcg(st_peak);  // max |v| over the whole row
  team_sync<kT, kBar>();
  if (tid == 0) {  // ready for the next launch
    *st_peak = 0u;
    *st_count = 0u;
  }
  uint32_t pk = M;
  const int nchunk = (len + 3) >> 2;
  if (kIsd) {
    if (!kEarlyGather) evaluate(p0, fr0, t0);
#pragma unroll
    for (int k = 0; k < kU; ++k)
      if (p0[k] >= 0) __stcg(ro + p0[k], t0[k]);
    for (int i0 = ibeg + tid + kU * kT; i0 < iend; i0 += kU * kT) {  // utterances with more impulses
      int p[kU];
      double fr[kU];
      float t[kU];
      load_plan(i0, p, fr);
      evaluate(p, fr, t);
#pragma unroll
      for (int k = 0; k < kU; ++k)
        if (p[k] >= 0) __stcg(ro + p[k], t[k]);
    }
    mt = block_umax<kT, kBar>(mt, scratch, tid);
    mx = block_umax<kT, kBar>(mx, scratch, tid);  // (these barriers also order the impulse stores before any read of the row below)
    // The peak of y = max(peak of the untouched samples, mt). The untouched peak is M unless the row's largest sample was
    // itself replaced (mx == M); even then nothing more is needed when a new value reaches M, or when nothing can exceed 1.
    if (mx < M || mt >= M) {
      pk = max(M, mt);
    } else if (M <= __float_as_uint(1.f)) {
      pk = M;  // some value <= M <= 1: no rescale either way
    } else {   // rare: take the peak of y itself
      uint32_t r = 0u;
      for (int c = tid; c < nchunk; c += kT) {
        const int q = 4 * c;
        if (q + 3 < len) {
          const float4 w = __ldcg(reinterpret_cast<const float4*>(ro + q));
          r = max(max(r, abs_bits(w.x)), max(abs_bits(w.y), max(abs_bits(w.z), abs_bits(w.w))));
        } else {
          for (int e = q; e < len; ++e) r = max(r, abs_bits(__ldcg(ro + e)));
        }
      }
      pk = block_umax<kT, kBar>(r, scratch, tid);
    }
  }
  const float peak = __uint_as_float(pk);
  if (!(s.always || peak > 1.f)) return;  // a NaN peak: "NaN > 1" is false, like the reference
  constexpr int kRU = RB_STREAM_RESCALE_U;  // chunks in flight per thread: the reads come from L2
  for (int c0 = tid; c0 < nchunk; c0 += kRU * kT) {
    float4 w[kRU];
#pragma unroll
    for (int k = 0; k < kRU; ++k) {
      const int p = 4 * (c0 + k * kT);
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
      const int p = 4 * (c0 + k * kT);
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

template <bool kIsd, bool kSum>
__global__ void __launch_bounds__(kSThreads, RB_STREAM_MIN_BLOCKS)
norm_stream_kernel(const StreamArgs s) {
  __shared__ uint32_t scratch[kSThreads / 32];
  const int per = s.ntiles + 1;
  const int u = blockIdx.x / per, tile = blockIdx.x - u * per;
  const int tid = threadIdx.x;
  const int len = min(s.len[u], s.ld);  // (a length beyond the row stride would leave the finisher waiting for tiles that do not exist)
  const float* ra = s.a + (size_t)u * s.ld;
  float* ro = s.out + (size_t)u * s.ld;
  uint32_t* st_peak = &s.state[u].x;
  uint32_t* st_count = &s.state[u].y;

  if (tile < s.ntiles) {
    // ---- tile CTA: copy + peak ------------------------------------------------------------------------------------------
    const int tile0 = tile * kSTile;
    if (tile0 >= len) return;
    const float* rb_ = kSum ? s.b + (size_t)u * s.ld : nullptr;
    const bool copy = kSum || (ro != ra);
    // kSub sub-tiles of kSU chunks per thread, software-pipelined: the loads of sub-tile j+1 are in flight while sub-tile j
    // is reduced and stored, so the CTA keeps requests outstanding for most of its life instead of only at its start.
    auto load_sub = [&](int j, float4 (&v)[kSU]) {
#pragma unroll
      for (int k = 0; k < kSU; ++k) {
        const int p = tile0 + 4 * ((j * kSU + k) * kSThreads + tid);
        v[k] = make_float4(0.f, 0.f, 0.f, 0.f);
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
      }
    };
    uint32_t m = 0u;
    auto consume_sub = [&](int j, const float4 (&v)[kSU]) {
#pragma unroll
      for (int k = 0; k < kSU; ++k) {
        const int p = tile0 + 4 * ((j * kSU + k) * kSThreads + tid);
        m = max(max(m, abs_bits(v[k].x)), max(abs_bits(v[k].y), max(abs_bits(v[k].z), abs_bits(v[k].w))));
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
    };
    float4 va[kSU], vb[kSU];
    load_sub(0, va);
#pragma unroll
    for (int j = 0; j < kSub; j += 2) {
      if (j + 1 < kSub) load_sub(j + 1, vb);
      consume_sub(j, va);
      if (j + 2 < kSub) load_sub(j + 2, va);
      if (j + 1 < kSub) consume_sub(j + 1, vb);
    }
    m = __reduce_max_sync(0xffffffffu, m);
    if ((tid & 31) == 0) scratch[tid >> 5] = m;
    __threadfence();  // this thread's stores are visible device-wide before the arrival below
    __syncthreads();
    if (tid == 0) {
      uint32_t r = scratch[0];
#pragma unroll
      for (int w = 1; w < kSThreads / 32; ++w) r = max(r, scratch[w]);
      if (r) atomicMax(st_peak, r);
      red_release_add(st_count, 1u);  // release: the peak update above is ordered before the arrival
    }
    return;
  }

  // ---- finisher CTA -----------------------------------------------------------------------------------------------------------
  const int nact = (len + kSTile - 1) / kSTile;  // tiles of this utterance that do work
  if (nact <= 0) return;
  finish_row<kIsd, kSThreads, 0, true>(s, u, len, nact, scratch, tid);
}

// ---- the streaming pass, TMA form -------------------------------------------------------------------------------------------
// The tile CTAs above hold their samples in registers, so every CTA that streams also occupies the registers a finisher needs,
// and finishers (latency-bound, ~10 us each) take slots away from the streaming. Here the two roles stop competing. The grid is
// a persistent set of identical CTAs (as many as fit the GPU, four per SM); in each of them
//   * warp 0 STREAMS: lane 0 drives a ring of three 16 KB shared-memory stages with bulk asynchronous copies (cp.async.bulk:
//     global -> shared completing on an mbarrier, shared -> global in bulk groups) over (utterance, 4096-sample tile) work items
//     claimed four at a time from a global counter; the warp's lanes read each landed tile from shared memory to take its
//     peak. No sample passes through a register on its way from input to output, a stage is refilled as soon as its store has
//     read it, and the arrival of a tile is published two iterations later, when its store group has long completed --
//     nothing in the loop waits for memory it has just touched.
//   * warps 1-7 are a FINISHER TEAM: they claim utterances in order from a second counter, wait until all tiles of the claimed
//     row have arrived (they are being streamed by CTAs that are resident by construction -- every resident CTA streams, so
//     no assumption about which CTAs are resident is needed) and finish it exactly as above: impulses, exact peak, conditional
//     rescale through L2, synchronising among themselves with a named barrier.
// Streaming never waits for a team. Rows whose input is also the output (in-place normWav) skip the store; the two-addend
// form (algo 8) stays with the register kernel.
#ifndef RB_TMA_STAGES
#define RB_TMA_STAGES 3
#endif
#ifndef RB_TMA_BATCH
#define RB_TMA_BATCH 4
#endif
#ifndef RB_TMA_STORE
#define RB_TMA_STORE 0   // 1: tiles leave shared memory by bulk stores; 0: by the streamer warp's own st.global (stays in L2)
#endif
constexpr int kTThreads = 256;
constexpr int kTeam = kTThreads - 32;              // finisher team: warps 1..7
constexpr int kTStages = RB_TMA_STAGES;
constexpr int kTTile = 4096;                       // samples per work item (16 KB)
constexpr int kTBatch = RB_TMA_BATCH;              // consecutive tiles of one row claimed per atomic
constexpr uint32_t kSpinLimit = 1u << 22;          // a wait that never ends becomes a trap (a launch error), not a hung GPU

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    if (!ok && ++spins > kSpinLimit) __trap();
  }
}
// L2 policy of the bulk copies: the finisher comes back to both the input row (impulse gathers) and the output row (scatter,
// rescale) a few microseconds after they were streamed, so they must stay in L2 (RB_TMA_L2: 0 no hint, 1 evict_last, 2 evict_normal).
#ifndef RB_TMA_L2
#define RB_TMA_L2 1
#endif
__device__ __forceinline__ uint64_t l2_policy() {
  uint64_t pol = 0;
#if RB_TMA_L2 == 1
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
#elif RB_TMA_L2 == 2
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(pol));
#endif
  return pol;
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t pol) {
#if RB_TMA_L2
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                   smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
               : "memory");
#else
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
#endif
}
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes, uint64_t pol) {
#if RB_TMA_L2
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(smem_u32(smem_src)),
               "r"(bytes), "l"(pol)
               : "memory");
#else
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
#endif
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() {
  asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory");
}

struct TmaSmem {
  float stage[kTStages][kTTile];   // the ring (16-byte aligned: first member of a 128-byte aligned block)
  uint64_t full[kTStages];         // a tile has landed (transaction barriers)
  uint32_t scratch[kTeam / 32];    // the finisher team's reduction scratch
  int row;                         // the row the team leader has claimed
};

struct TmaCounters {               // device scratch right behind the per-utterance state, zeroed by the launcher
  uint32_t next_batch;             // work items of the streamers: (row, group of kTBatch tiles), row-major
  uint32_t next_row;               // rows for the finisher teams
};

template <bool kIsd>
__global__ void __launch_bounds__(kTThreads, 4)
norm_stream_tma_kernel(const StreamArgs s, int B, TmaCounters* __restrict__ ctr) {
  extern __shared__ __align__(128) unsigned char tma_raw[];
  TmaSmem& sm = *reinterpret_cast<TmaSmem*>(tma_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    for (int k = 0; k < kTStages; ++k) mbar_init(&sm.full[k], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();

  if (warp == 0) {
    // ================================ streamer warp =========================================================================
    const bool copy = s.out != s.a;
    const uint64_t pol = l2_policy();
    const int groups = (s.ntiles + kTBatch - 1) / kTBatch;           // batches per row
    const uint32_t total = (uint32_t)B * (uint32_t)groups;
    // the current batch: tiles [bt, bt_end) of row bu (length blen); lane 0 claims, every lane holds a copy
    int bu = 0, bt = 0, bt_end = 0, blen = 0;
    bool exhausted = false;
    // per tile in flight (ring slot i % kTStages): row, first sample, valid samples; -1 = end of work
    int t_u[kTStages], t_tile0[kTStages], t_n[kTStages];
    uint32_t t_max[kTStages];
#pragma unroll
    for (int k = 0; k < kTStages; ++k) {
      t_u[k] = -1;
      t_tile0[k] = t_n[k] = 0;
      t_max[k] = 0u;
    }
    // next tile of the global order that has samples (uniform across the warp); false when the work is exhausted
    auto next_tile = [&](int& u, int& tile0, int& n) -> bool {
      for (;;) {
        if (bt < bt_end) {
          tile0 = bt * kTTile;
          ++bt;
          if (tile0 < blen) {
            u = bu;
            n = min(kTTile, blen - tile0);
            return true;
          }
          bt = bt_end;  // the rest of this batch lies beyond the row's end
          continue;
        }
        if (exhausted) return false;
        uint32_t q = 0;
        if (lane == 0) q = atomicAdd(&ctr->next_batch, 1u);
        q = __shfl_sync(0xffffffffu, q, 0);
        if (q >= total) {
          exhausted = true;
          return false;
        }
        bu = (int)(q / (uint32_t)groups);
        bt = (int)(q - (uint32_t)bu * (uint32_t)groups) * kTBatch;
        bt_end = min(bt + kTBatch, s.ntiles);
        blen = min(__ldg(s.len + bu), s.ld);
      }
    };
    auto issue = [&](int slot) {  // claim the next tile and start its load into ring slot `slot`
      int u = -1, tile0 = 0, n = 0;
      const bool ok = next_tile(u, tile0, n);
#pragma unroll
      for (int k = 0; k < kTStages; ++k)
        if (k == slot) {
          t_u[k] = ok ? u : -1;
          t_tile0[k] = tile0;
          t_n[k] = n;
          t_max[k] = 0u;
        }
      if (ok && lane == 0) {
        const uint32_t bytes = (uint32_t)((n + 3) & ~3) * 4u;  // whole 16-byte chunks; rows are padded to a multiple of 4 samples
        mbar_arrive_expect_tx(&sm.full[slot], bytes);
        bulk_load(sm.stage[slot], s.a + (size_t)u * s.ld + tile0, bytes, &sm.full[slot], pol);
      }
    };
    auto publish = [&](int slot) {  // the tile in `slot`: its store has completed (caller waited) -> peak, then arrival
      if (lane == 0) {
#pragma unroll
        for (int k = 0; k < kTStages; ++k)
          if (k == slot && t_u[k] >= 0) {
            asm volatile("fence.proxy.async;" ::: "memory");  // the tile reached global memory through the async proxy
            if (t_max[k]) atomicMax(&s.state[t_u[k]].x, t_max[k]);
            red_release_add(&s.state[t_u[k]].y, 1u);           // release: peak update and tile are ordered before the arrival
          }
      }
    };
#pragma unroll
    for (int k = 0; k < kTStages - 1; ++k) issue(k);
    uint32_t phase_bits = 0u;  // bit k: parity to wait for on ring slot k
    int i_end = 0;
    for (int i = 0;; ++i) {
      i_end = i;
      const int slot = i % kTStages;
      int u = -1, tile0 = 0, n = 0;
#pragma unroll
      for (int k = 0; k < kTStages; ++k)
        if (k == slot) {
          u = t_u[k];
          tile0 = t_tile0[k];
          n = t_n[k];
        }
      if (u < 0) break;
      mbar_wait(&sm.full[slot], (phase_bits >> slot) & 1u);  // every lane observes the completion itself
      phase_bits ^= 1u << slot;
#if RB_TMA_STORE
      if (lane == 0) {  // the landed tile goes straight back out; the lanes only look at it
        const uint32_t sbytes = (uint32_t)(n & ~3) * 4u;
        if (copy && sbytes) bulk_store(s.out + (size_t)u * s.ld + tile0, sm.stage[slot], sbytes, pol);
        bulk_commit();  // (an empty group when nothing was stored: one group per tile keeps the bookkeeping uniform)
      }
#else
      // tile i-1 was stored one iteration ago: its stores have had a whole iteration to drain, so this fence is short
      const int prev_slot = (slot + kTStages - 1) % kTStages;
      if (i >= 1) {
        __threadfence();
        __syncwarp();
        publish(prev_slot);
      }
#endif
      const float4* src = reinterpret_cast<const float4*>(sm.stage[slot]);
      float* dst = s.out + (size_t)u * s.ld + tile0;
      uint32_t m = 0u;
      const int nfull = n >> 2;
#pragma unroll 8
      for (int c = lane; c < nfull; c += 32) {
        const float4 v = src[c];
        m = max(max(m, abs_bits(v.x)), max(abs_bits(v.y), max(abs_bits(v.z), abs_bits(v.w))));
#if !RB_TMA_STORE
        if (copy) reinterpret_cast<float4*>(dst)[c] = v;
#endif
      }
      if (lane < (n & 3)) {  // the ragged end of the row: the <= 3 samples a 16-byte store cannot carry
        const float v = sm.stage[slot][4 * nfull + lane];
        m = max(m, abs_bits(v));
        if (copy) dst[4 * nfull + lane] = v;
      }
      m = __reduce_max_sync(0xffffffffu, m);
#pragma unroll
      for (int k = 0; k < kTStages; ++k)
        if (k == slot) t_max[k] = m;
      __syncwarp();  // every lane is done reading the stage (and the ragged-end stores are issued) before lane 0 moves on
      // ring slot (i + kTStages - 1) % kTStages held tile i-1: wait for its store, publish it, refill the slot
      const int prev = (slot + kTStages - 1) % kTStages;
#if RB_TMA_STORE
      if (i >= 1) {
        if (lane == 0) bulk_wait<1>();
        publish(prev);
      }
#endif
      issue(prev);
    }
    // the loop stopped at iteration `i_end`; tile i_end - 1 has been stored but not yet published
    if (i_end >= 1) {
#if RB_TMA_STORE
      if (lane == 0) bulk_wait<0>();
#else
      __threadfence();
      __syncwarp();
#endif
      publish((i_end - 1) % kTStages);
    }
    return;
  }

  // ================================== finisher team (warps 1..7) ================================================================
  const int ttid = tid - 32;
  for (;;) {
    if (ttid == 0) sm.row = (int)atomicAdd(&ctr->next_row, 1u);
    team_sync<kTeam, 1>();
    const int u = sm.row;
    team_sync<kTeam, 1>();  // everyone has read the claim before the leader overwrites it
    if (u >= B) return;
    const int len = min(s.len[u], s.ld);
    const int nact = (len + kTTile - 1) / kTTile;
    if (nact <= 0) continue;
#ifdef RB_STREAM_DEBUG
    unsigned long long t0 = 0;
    if (ttid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
#endif
    finish_row<kIsd, kTeam, 1, false>(s, u, len, nact, sm.scratch, ttid);
#ifdef RB_STREAM_DEBUG
    if (ttid == 0) {
      unsigned long long t2;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t2));
      unsigned long long* dbg = reinterpret_cast<unsigned long long*>((uint2*)s.state + B + 1) + 4 * (size_t)u;
      dbg[0] = t0;
      dbg[2] = t2;
      dbg[3] = blockIdx.x;
    }
#endif
  }
}

// 16-bit PCM -> float32, the conversion a wav reader applies (sample / 32768: exact in float32). n8 = groups of 8 samples.
__global__ void __launch_bounds__(256)
pcm16_to_f32_kernel(const int4* __restrict__ in, float4* __restrict__ out, size_t n8) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const int4 v = __ldg(in + i);
    const int w[4] = {v.x, v.y, v.z, v.w};
    float f[8];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      f[2 * k] = (float)(short)(w[k] & 0xffff) * (1.f / 32768.f);
      f[2 * k + 1] = (float)(short)((unsigned)w[k] >> 16) * (1.f / 32768.f);
    }
    out[2 * i] = make_float4(f[0], f[1], f[2], f[3]);
    out[2 * i + 1] = make_float4(f[4], f[5], f[6], f[7]);
  }
}

}  // namespace

int stream_tiles_for(int ld) { return (ld + kSTile - 1) / kSTile; }

int launch_pcm16_to_f32(const int16_t* in, float* out, size_t n, cudaStream_t st) {
  if (n == 0) return RB_OK;
  if (n % 8 != 0 || ((uintptr_t)in & 15u) || ((uintptr_t)out & 15u)) return RB_ERR_ALIGNMENT;
  const size_t n8 = n / 8;
  pcm16_to_f32_kernel<<<(unsigned)min((n8 + 255) / 256, (size_t)148 * 16), 256, 0, st>>>((const int4*)in, (float4*)out, n8);
  RB_LAUNCH_CHECK();
  return RB_OK;
}

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

// which streaming kernel serves the single-input forms: the TMA streamer (default) or the register kernel
// (RAWBOOST_B200_STREAM=reg, for A/B measurements; the two-addend form always uses the register kernel)
static bool use_tma_streamer() {
  static const bool v = [] {
    const char* e = getenv("RAWBOOST_B200_STREAM");
    return !(e && e[0] == 'r');
  }();
  return v;
}

int launch_norm_stream(const float* a, const float* b, const int32_t* len, int B, int ld, int always, const int32_t* isd_off,
                       const int32_t* isd_idx, const double* isd_fr, float g_sd, float* out, void* state, cudaStream_t st) {
  if (B <= 0 || ld <= 0) return RB_OK;
  const bool isd = isd_off != nullptr;
  if (isd && (!isd_idx || !isd_fr || b)) return RB_ERR_INVALID_ARG;  // impulses apply to a single input
  if (!a || !len || !out || !state) return RB_ERR_INVALID_ARG;
  RB_CUDA(cudaMemsetAsync(state, 0, ((size_t)B + 1) * sizeof(uint2), st));  // per-utterance {peak, arrivals} + the work counter
  StreamArgs s;
  s.ld = ld;
  s.always = always;
  s.isd_idx = isd_idx;
  s.isd_fr = isd_fr;
  s.g_sd = g_sd;
  s.a = a;
  s.b = b;
  s.len = len;
  s.isd_off = isd_off;
  s.out = out;
  s.state = (uint2*)state;
  s.dbg_B = B;
  if (!b && use_tma_streamer()) {
    s.ntiles = (ld + kTTile - 1) / kTTile;
    auto kernel = isd ? norm_stream_tma_kernel<true> : norm_stream_tma_kernel<false>;
    const size_t smem = sizeof(TmaSmem);
    RB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    RB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    int dev = 0, sms = 0, per_sm = 0;
    RB_CUDA(cudaGetDevice(&dev));
    RB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    RB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kTThreads, smem));
    const int grid = std::max(1, sms * std::max(1, per_sm));  // persistent: as many CTAs as the GPU holds at once
    kernel<<<grid, kTThreads, smem, st>>>(s, B, reinterpret_cast<TmaCounters*>((uint2*)state + B));
    RB_LAUNCH_CHECK();
    return RB_OK;
  }
  s.ntiles = stream_tiles_for(ld);
  const int per = max(1, (int)(0x7fffffff / (long long)(s.ntiles + 1)));  // utterances per launch (grid.x < 2^31)
  for (int b0 = 0; b0 < B; b0 += per) {
    const int nb = min(per, B - b0);
    s.a = a + (size_t)b0 * ld;
    s.b = b ? b + (size_t)b0 * ld : nullptr;
    s.len = len + b0;
    s.isd_off = isd ? isd_off + b0 : nullptr;
    s.out = out + (size_t)b0 * ld;
    s.state = (uint2*)state + b0;
    const unsigned grid = (unsigned)nb * (unsigned)(s.ntiles + 1);
    if (isd) norm_stream_kernel<true, false><<<grid, kSThreads, 0, st>>>(s);
    else if (b) norm_stream_kernel<false, true><<<grid, kSThreads, 0, st>>>(s);
    else norm_stream_kernel<false, false><<<grid, kSThreads, 0, st>>>(s);
    RB_LAUNCH_CHECK();
  }
  return RB_OK;
}

}  // namespace rb
