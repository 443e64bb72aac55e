// Launchers of the HBM-bound passes (rb_dense.cu). All asynchronous on `st`, no allocation, no sync.
#pragma once
#include "rb_common.cuh"

namespace rb {

struct FinalizeArgs {
  const float* stats;        // [B][ntiles][kStatN]
  int ntiles;
  const int32_t* len;        // [B]
  int center;                // subtract the mean (LnL tail, RawBoost.py:67)
  int always;                // normWav(x, 1)
  // impulsive noise applied after the first normalisation (NULL isd_off = none)
  const float* raw;          // the waveform the statistics were taken from, [B][ld]
  int ld;
  const int32_t* isd_off;
  const int32_t* isd_idx;
  const double* isd_fr;
  float g_sd;
  UttParams* out;            // [B]
};

int launch_mask_build(const int32_t* isd_off, const int32_t* isd_idx, const int32_t* len, int B, uint32_t* mask,
                      int mask_ld, cudaStream_t st);
int launch_dense_stats(const float* x, const int32_t* len, int B, int ld, float* stats, const uint32_t* mask, int mask_ld,
                       cudaStream_t st);
int launch_finalize(const FinalizeArgs& args, int B, cudaStream_t st);
int launch_ssi_finalize(const float* stats_x, const float* stats_n, int ntiles, const float* snr_db, UttParams* out, int B,
                        cudaStream_t st);
int launch_apply_affine(const float* in, const int32_t* len, int B, int ld, const UttParams* params, float* out, cudaStream_t st);
int launch_apply_ssi(const float* x, const float* noise, const int32_t* len, int B, int ld, const UttParams* params, float* out,
                     cudaStream_t st);
int launch_apply_sum(const float* a, const float* b, const int32_t* len, int B, int ld, float* out, cudaStream_t st);
int launch_isd_scatter(const float* raw, const int32_t* len, int B, int ld, const int32_t* isd_off, const int32_t* isd_idx,
                       const double* isd_fr, float g_sd, const UttParams* params, float* out, cudaStream_t st);

int launch_isd_fused(const float* x, const int32_t* len, int B, int ld, int always, const int32_t* isd_off, const int32_t* isd_idx,
                     const double* isd_fr, float g_sd, float* out, cudaStream_t st);

}  // namespace rb
