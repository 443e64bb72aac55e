// rb_api.cu -- the extern "C" surface declared in include/rawboost_b200.h: argument checks, workspace carving and
// the composition of the kernels into the reference's operators and its 9-way dispatcher
// (/root/reference/datautils/asvspoof_2019_augall_3.py:377-439).
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <new>
#include <vector>

#include "rb_common.cuh"
#include "rb_dense.cuh"

namespace rb {

std::atomic<uint64_t> g_launches{0};

// ---- FIR-bank launch timing (see rb_common.cuh) -------------------------------------------------------
namespace {
std::atomic<int> g_profile_on{0};
std::mutex g_profile_mu;
std::vector<cudaEvent_t> g_profile_events;  // begin/end pairs not yet read
double g_profile_ms = 0.0;
uint64_t g_profile_launches = 0;
}  // namespace

void profile_begin(cudaStream_t st) {
  if (!g_profile_on.load(std::memory_order_relaxed)) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, st);
  std::lock_guard<std::mutex> lk(g_profile_mu);
  g_profile_events.push_back(e);
}

void profile_end(cudaStream_t st) {
  if (!g_profile_on.load(std::memory_order_relaxed)) return;
  std::lock_guard<std::mutex> lk(g_profile_mu);
  if (g_profile_events.size() % 2 == 0) return;  // begin failed
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) {
    cudaEventDestroy(g_profile_events.back());
    g_profile_events.pop_back();
    return;
  }
  cudaEventRecord(e, st);
  g_profile_events.push_back(e);
}

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int mask_ld_for(int ld) { return (int)align_up((size_t)ld / 32 + 2, 4); }

// Device workspace layout for one batch.
struct Workspace {
  float* stats_a;
  float* stats_b;
  UttParams* params;
  uint32_t* mask;
  int mask_ld;
  float* buf0;  // raw FIR-bank output / coloured noise
  float* buf1;  // intermediate waveform of chained algos
  float* buf2;  // second branch of algo 8
  size_t bytes;
};

Workspace carve(void* base, int B, int ld) {
  Workspace w;
  const int ntiles = tiles_for(ld);
  size_t off = 0;
  char* p = (char*)base;
  auto take = [&](size_t n) {
    char* r = p ? p + off : nullptr;
    off += align_up(n, 256);
    return r;
  };
  w.stats_a = (float*)take((size_t)B * ntiles * kStatN * sizeof(float));
  w.stats_b = (float*)take((size_t)B * ntiles * kStatN * sizeof(float));
  w.params = (UttParams*)take((size_t)B * sizeof(UttParams));
  w.mask_ld = mask_ld_for(ld);
  w.mask = (uint32_t*)take((size_t)B * w.mask_ld * sizeof(uint32_t));
  w.buf0 = (float*)take((size_t)B * ld * sizeof(float));
  w.buf1 = (float*)take((size_t)B * ld * sizeof(float));
  w.buf2 = (float*)take((size_t)B * ld * sizeof(float));
  w.bytes = off;
  return w;
}

int check_batch(const void* x, const int32_t* len, int B, int ld, const void* y) {
  if (B < 0 || ld < 0) return RB_ERR_INVALID_ARG;
  if (B == 0 || ld == 0) return RB_OK;
  if (!x || !len || !y) return RB_ERR_INVALID_ARG;
  if (ld % 4 != 0 || ((uintptr_t)x & 15u) || ((uintptr_t)y & 15u)) return RB_ERR_ALIGNMENT;
  return RB_OK;
}

int check_ws(void* ws, size_t ws_bytes, int B, int ld, Workspace* out) {
  if (B == 0 || ld == 0) return RB_OK;
  if (!ws || ((uintptr_t)ws & 255u)) return ws ? RB_ERR_ALIGNMENT : RB_ERR_WORKSPACE;
  *out = carve(ws, B, ld);
  if (out->bytes > ws_bytes) return RB_ERR_WORKSPACE;
  return RB_OK;
}

#define RB_TRY(expr)          \
  do {                        \
    int rc__ = (expr);        \
    if (rc__ != RB_OK) return rc__; \
  } while (0)

// LnL (optionally followed by ISD in the same finalise / apply passes): x -> out
int do_lnl(const float* x, const int32_t* len, int B, int ld, const rb_plan* pl, bool with_isd, float* out, const Workspace& w,
           cudaStream_t st) {
  if (!pl || pl->n_f < 1 || !pl->lnl_taps || !pl->lnl_tap_off) return RB_ERR_PLAN;
  if (with_isd && (!pl->isd_off || !pl->isd_idx || !pl->isd_fr)) return RB_ERR_PLAN;
  if (with_isd) RB_TRY(launch_mask_build(pl->isd_off, pl->isd_idx, len, B, w.mask, w.mask_ld, st));
  RB_TRY(launch_fir_bank(x, len, B, ld, pl->lnl_taps, pl->lnl_tap_off, pl->n_f, 1, 1, w.buf0, w.stats_a,
                         with_isd ? w.mask : nullptr, w.mask_ld, st));
  FinalizeArgs fa{};
  fa.stats = w.stats_a;
  fa.ntiles = tiles_for(ld);
  fa.len = len;
  fa.center = 1;
  fa.always = 0;
  fa.raw = w.buf0;
  fa.ld = ld;
  fa.isd_off = with_isd ? pl->isd_off : nullptr;
  fa.isd_idx = pl->isd_idx;
  fa.isd_fr = pl->isd_fr;
  fa.g_sd = pl->g_sd;
  fa.out = w.params;
  RB_TRY(launch_finalize(fa, B, st));
  RB_TRY(launch_apply_affine(w.buf0, len, B, ld, w.params, out, st));
  if (with_isd)
    RB_TRY(launch_isd_scatter(w.buf0, len, B, ld, pl->isd_off, pl->isd_idx, pl->isd_fr, pl->g_sd, w.params, out, st));
  return RB_OK;
}

// ISD on an existing waveform: x -> out (out != x)
int do_isd(const float* x, const int32_t* len, int B, int ld, const rb_plan* pl, float* out, const Workspace& w, cudaStream_t st) {
  if (!pl || !pl->isd_off || !pl->isd_idx || !pl->isd_fr) return RB_ERR_PLAN;
  RB_TRY(launch_mask_build(pl->isd_off, pl->isd_idx, len, B, w.mask, w.mask_ld, st));
  RB_TRY(launch_dense_stats(x, len, B, ld, w.stats_a, w.mask, w.mask_ld, st));
  FinalizeArgs fa{};
  fa.stats = w.stats_a;
  fa.ntiles = tiles_for(ld);
  fa.len = len;
  fa.raw = x;
  fa.ld = ld;
  fa.isd_off = pl->isd_off;
  fa.isd_idx = pl->isd_idx;
  fa.isd_fr = pl->isd_fr;
  fa.g_sd = pl->g_sd;
  fa.out = w.params;
  RB_TRY(launch_finalize(fa, B, st));
  RB_TRY(launch_apply_affine(x, len, B, ld, w.params, out, st));
  RB_TRY(launch_isd_scatter(x, len, B, ld, pl->isd_off, pl->isd_idx, pl->isd_fr, pl->g_sd, w.params, out, st));
  return RB_OK;
}

// SSI: x -> out (out may alias x). Uses buf0 for the coloured noise.
int do_ssi(const float* x, const int32_t* len, int B, int ld, const rb_plan* pl, float* out, const Workspace& w, cudaStream_t st) {
  if (!pl || !pl->ssi_noise || !pl->ssi_taps || !pl->ssi_tap_off || !pl->ssi_snr_db) return RB_ERR_PLAN;
  if ((uintptr_t)pl->ssi_noise & 15u) return RB_ERR_ALIGNMENT;
  RB_TRY(launch_fir_bank(pl->ssi_noise, len, B, ld, pl->ssi_taps, pl->ssi_tap_off, 1, 1, 0, w.buf0, w.stats_a, nullptr, 0, st));
  RB_TRY(launch_dense_stats(x, len, B, ld, w.stats_b, nullptr, 0, st));
  RB_TRY(launch_ssi_finalize(w.stats_b, w.stats_a, tiles_for(ld), pl->ssi_snr_db, w.params, B, st));
  RB_TRY(launch_apply_ssi(x, w.buf0, len, B, ld, w.params, out, st));
  return RB_OK;
}

// normWav: x -> out
int do_normwav(const float* x, const int32_t* len, int B, int ld, int always, float* out, const Workspace& w, cudaStream_t st) {
  RB_TRY(launch_dense_stats(x, len, B, ld, w.stats_a, nullptr, 0, st));
  FinalizeArgs fa{};
  fa.stats = w.stats_a;
  fa.ntiles = tiles_for(ld);
  fa.len = len;
  fa.always = always ? 1 : 0;
  fa.raw = x;
  fa.ld = ld;
  fa.out = w.params;
  RB_TRY(launch_finalize(fa, B, st));
  RB_TRY(launch_apply_affine(x, len, B, ld, w.params, out, st));
  return RB_OK;
}

}  // namespace
}  // namespace rb

using namespace rb;

extern "C" {

const char* rb_error_string(int code) {
  switch (code) {
    case RB_OK: return "ok";
    case RB_ERR_INVALID_ARG: return "rawboost_b200: invalid argument";
    case RB_ERR_ALIGNMENT: return "rawboost_b200: ld must be a multiple of 4 and waveform/workspace pointers 16/256-byte aligned";
    case RB_ERR_WORKSPACE: return "rawboost_b200: workspace missing or smaller than rb_workspace_bytes()";
    case RB_ERR_NO_DEVICE: return "rawboost_b200: no usable CUDA device (needs compute capability 10.0)";
    case RB_ERR_PLAN: return "rawboost_b200: a plan field required by this algo is NULL";
    default: break;
  }
  if (code > 0) return cudaGetErrorString((cudaError_t)code);
  return "rawboost_b200: unknown error";
}

int rb_abi_version(void) { return RB_ABI_VERSION; }

size_t rb_workspace_bytes(int B, int ld) {
  if (B <= 0 || ld <= 0) return 0;
  return carve(nullptr, B, ld).bytes;
}

uint64_t rb_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int rb_profile_enable(int on) {
  g_profile_on.store(on ? 1 : 0, std::memory_order_relaxed);
  return RB_OK;
}

int rb_profile_read(double* fir_ms, uint64_t* fir_launches, int reset) {
  std::lock_guard<std::mutex> lk(g_profile_mu);
  for (size_t i = 0; i + 1 < g_profile_events.size(); i += 2) {
    float ms = 0.f;
    RB_CUDA(cudaEventSynchronize(g_profile_events[i + 1]));
    RB_CUDA(cudaEventElapsedTime(&ms, g_profile_events[i], g_profile_events[i + 1]));
    g_profile_ms += (double)ms;
    g_profile_launches += 1;
  }
  for (cudaEvent_t e : g_profile_events) cudaEventDestroy(e);
  g_profile_events.clear();
  if (fir_ms) *fir_ms = g_profile_ms;
  if (fir_launches) *fir_launches = g_profile_launches;
  if (reset) {
    g_profile_ms = 0.0;
    g_profile_launches = 0;
  }
  return RB_OK;
}

int rb_filter_fir(const float* x, const int32_t* len, int B, int ld, const float* taps, const int32_t* tap_off, float* y,
                  void* stream) {
  RB_TRY(check_batch(x, len, B, ld, y));
  if (B == 0 || ld == 0) return RB_OK;
  if (!taps || !tap_off) return RB_ERR_INVALID_ARG;
  return launch_fir_bank(x, len, B, ld, taps, tap_off, 1, 1, 0, y, nullptr, nullptr, 0, (cudaStream_t)stream);
}

int rb_normwav(const float* x, const int32_t* len, int B, int ld, int always, float* y, void* workspace, size_t workspace_bytes,
               void* stream) {
  RB_TRY(check_batch(x, len, B, ld, y));
  if (B == 0 || ld == 0) return RB_OK;
  Workspace w;
  RB_TRY(check_ws(workspace, workspace_bytes, B, ld, &w));
  return do_normwav(x, len, B, ld, always, y, w, (cudaStream_t)stream);
}

int rb_lnl(const float* x, const int32_t* len, int B, int ld, const rb_plan* plan, float* y, void* workspace,
           size_t workspace_bytes, void* stream) {
  return rb_process(1, x, len, B, ld, plan, y, workspace, workspace_bytes, stream);
}
int rb_isd(const float* x, const int32_t* len, int B, int ld, const rb_plan* plan, float* y, void* workspace,
           size_t workspace_bytes, void* stream) {
  return rb_process(2, x, len, B, ld, plan, y, workspace, workspace_bytes, stream);
}
int rb_ssi(const float* x, const int32_t* len, int B, int ld, const rb_plan* plan, float* y, void* workspace,
           size_t workspace_bytes, void* stream) {
  return rb_process(3, x, len, B, ld, plan, y, workspace, workspace_bytes, stream);
}

int rb_process(int algo, const float* x, const int32_t* len, int B, int ld, const rb_plan* plan, float* y, void* workspace,
               size_t workspace_bytes, void* stream) {
  RB_TRY(check_batch(x, len, B, ld, y));
  if (B == 0 || ld == 0) return RB_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (algo < 1 || algo > 8) {  // identity (asvspoof_2019_augall_3.py:435-437)
    if (x != y) RB_CUDA(cudaMemcpyAsync(y, x, (size_t)B * ld * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return RB_OK;
  }
  if (x == y) return RB_ERR_INVALID_ARG;
  Workspace w;
  RB_TRY(check_ws(workspace, workspace_bytes, B, ld, &w));
  switch (algo) {
    case 1: return do_lnl(x, len, B, ld, plan, false, y, w, st);
    case 2: return do_isd(x, len, B, ld, plan, y, w, st);
    case 3: return do_ssi(x, len, B, ld, plan, y, w, st);
    case 4:
      RB_TRY(do_lnl(x, len, B, ld, plan, true, w.buf1, w, st));
      return do_ssi(w.buf1, len, B, ld, plan, y, w, st);
    case 5: return do_lnl(x, len, B, ld, plan, true, y, w, st);
    case 6:
      RB_TRY(do_lnl(x, len, B, ld, plan, false, w.buf1, w, st));
      return do_ssi(w.buf1, len, B, ld, plan, y, w, st);
    case 7:
      RB_TRY(do_isd(x, len, B, ld, plan, w.buf1, w, st));
      return do_ssi(w.buf1, len, B, ld, plan, y, w, st);
    case 8:
      RB_TRY(do_lnl(x, len, B, ld, plan, false, w.buf1, w, st));
      RB_TRY(do_isd(x, len, B, ld, plan, w.buf2, w, st));
      RB_TRY(launch_apply_sum(w.buf1, w.buf2, len, B, ld, w.buf0, st));
      return do_normwav(w.buf0, len, B, ld, 0, y, w, st);
  }
  return RB_ERR_INVALID_ARG;
}

// ---------------------------------------------------------------------------------------------------
// host-buffer context
// ---------------------------------------------------------------------------------------------------
struct rb_ctx {
  int device;
  cudaStream_t stream;
  char* dev;          // one device arena, grown on demand
  size_t dev_bytes;
  uint64_t h2d, d2h;
};

int rb_ctx_create(rb_ctx** out, int device) {
  if (!out) return RB_ERR_INVALID_ARG;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return RB_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  RB_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return RB_ERR_NO_DEVICE;
  RB_CUDA(cudaSetDevice(device));
  rb_ctx* c = new (std::nothrow) rb_ctx();
  if (!c) return RB_ERR_INVALID_ARG;
  c->device = device;
  c->dev = nullptr;
  c->dev_bytes = 0;
  c->h2d = c->d2h = 0;
  cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) {
    delete c;
    return (int)e;
  }
  *out = c;
  return RB_OK;
}

int rb_ctx_destroy(rb_ctx* c) {
  if (!c) return RB_OK;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  if (c->dev) cudaFree(c->dev);
  cudaStreamDestroy(c->stream);
  delete c;
  return RB_OK;
}

int rb_ctx_last_traffic(const rb_ctx* c, uint64_t* h2d, uint64_t* d2h) {
  if (!c) return RB_ERR_INVALID_ARG;
  if (h2d) *h2d = c->h2d;
  if (d2h) *d2h = c->d2h;
  return RB_OK;
}

int rb_process_host(rb_ctx* c, int algo, const float* x, const int32_t* len, int B, int ld, const rb_plan* plan, float* y) {
  if (!c) return RB_ERR_INVALID_ARG;
  if (B < 0 || ld < 0) return RB_ERR_INVALID_ARG;
  if (B == 0 || ld == 0) return RB_OK;
  if (!x || !len || !y) return RB_ERR_INVALID_ARG;
  if (ld % 4 != 0) return RB_ERR_ALIGNMENT;
  RB_CUDA(cudaSetDevice(c->device));
  const bool use_lnl = (algo == 1 || algo == 4 || algo == 5 || algo == 6 || algo == 8);
  const bool use_isd = (algo == 2 || algo == 4 || algo == 5 || algo == 7 || algo == 8);
  const bool use_ssi = (algo == 3 || algo == 4 || algo == 6 || algo == 7);
  if ((use_lnl || use_isd || use_ssi) && !plan) return RB_ERR_PLAN;
  if (use_lnl && (plan->n_f < 1 || !plan->lnl_taps || !plan->lnl_tap_off)) return RB_ERR_PLAN;
  if (use_isd && (!plan->isd_off || !plan->isd_idx || !plan->isd_fr)) return RB_ERR_PLAN;
  if (use_ssi && (!plan->ssi_noise || !plan->ssi_taps || !plan->ssi_tap_off || !plan->ssi_snr_db)) return RB_ERR_PLAN;

  const size_t wave = (size_t)B * ld * sizeof(float);
  const size_t n_lnl_off = use_lnl ? (size_t)B * plan->n_f + 1 : 0;
  const size_t n_lnl_taps = use_lnl ? (size_t)plan->lnl_tap_off[n_lnl_off - 1] : 0;
  const size_t n_isd = use_isd ? (size_t)plan->isd_off[B] : 0;
  const size_t n_ssi_taps = use_ssi ? (size_t)plan->ssi_tap_off[B] : 0;

  // arena layout
  size_t off = 0;
  auto take = [&](size_t n) {
    size_t r = off;
    off += align_up(n, 256);
    return r;
  };
  const size_t o_x = take(wave), o_y = take(wave), o_len = take((size_t)B * 4);
  const size_t o_lo = take(n_lnl_off * 4), o_lt = take(n_lnl_taps * 4);
  const size_t o_io = take(use_isd ? (size_t)(B + 1) * 4 : 0), o_ii = take(n_isd * 4), o_if = take(n_isd * 8);
  const size_t o_sn = take(use_ssi ? wave : 0), o_so = take(use_ssi ? (size_t)(B + 1) * 4 : 0), o_st = take(n_ssi_taps * 4),
               o_sr = take(use_ssi ? (size_t)B * 4 : 0);
  const size_t ws_bytes = rb_workspace_bytes(B, ld);
  const size_t o_ws = take(ws_bytes);
  if (off > c->dev_bytes) {
    RB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->dev) RB_CUDA(cudaFree(c->dev));
    c->dev = nullptr;
    c->dev_bytes = 0;
    RB_CUDA(cudaMalloc((void**)&c->dev, off));
    c->dev_bytes = off;
  }
  char* d = c->dev;
  cudaStream_t st = c->stream;
  uint64_t h2d = 0;
  auto up = [&](size_t o, const void* src, size_t n) -> int {
    if (n == 0) return RB_OK;
    h2d += n;
    return (int)cudaMemcpyAsync(d + o, src, n, cudaMemcpyHostToDevice, st);
  };
  RB_TRY(up(o_x, x, wave));
  RB_TRY(up(o_len, len, (size_t)B * 4));
  rb_plan dp;
  memset(&dp, 0, sizeof(dp));
  if (plan) {
    dp.n_f = plan->n_f;
    dp.g_sd = plan->g_sd;
  }
  if (use_lnl) {
    RB_TRY(up(o_lo, plan->lnl_tap_off, n_lnl_off * 4));
    RB_TRY(up(o_lt, plan->lnl_taps, n_lnl_taps * 4));
    dp.lnl_tap_off = (const int32_t*)(d + o_lo);
    dp.lnl_taps = (const float*)(d + o_lt);
  }
  if (use_isd) {
    RB_TRY(up(o_io, plan->isd_off, (size_t)(B + 1) * 4));
    RB_TRY(up(o_ii, plan->isd_idx, n_isd * 4));
    RB_TRY(up(o_if, plan->isd_fr, n_isd * 8));
    dp.isd_off = (const int32_t*)(d + o_io);
    dp.isd_idx = (const int32_t*)(d + o_ii);
    dp.isd_fr = (const double*)(d + o_if);
  }
  if (use_ssi) {
    RB_TRY(up(o_sn, plan->ssi_noise, wave));
    RB_TRY(up(o_so, plan->ssi_tap_off, (size_t)(B + 1) * 4));
    RB_TRY(up(o_st, plan->ssi_taps, n_ssi_taps * 4));
    RB_TRY(up(o_sr, plan->ssi_snr_db, (size_t)B * 4));
    dp.ssi_noise = (const float*)(d + o_sn);
    dp.ssi_tap_off = (const int32_t*)(d + o_so);
    dp.ssi_taps = (const float*)(d + o_st);
    dp.ssi_snr_db = (const float*)(d + o_sr);
  }
  RB_TRY(rb_process(algo, (const float*)(d + o_x), (const int32_t*)(d + o_len), B, ld, &dp, (float*)(d + o_y), d + o_ws, ws_bytes, st));
  RB_CUDA(cudaMemcpyAsync(y, d + o_y, wave, cudaMemcpyDeviceToHost, st));
  RB_CUDA(cudaStreamSynchronize(st));
  c->h2d = h2d;
  c->d2h = wave;
  return RB_OK;
}

}  // extern "C"
