// rb_multiview.cu -- the step right after the RawBoost path for every view of an item:
//   * batch_pad_for_multiview (/root/reference/core_scripts/data_io/wav_augmentation.py:209-282): every view is cut or
//     extended (zeros, or repetition with repeat_pad) to the length of view 0, then ONE crop [start, start+length) is
//     applied to all views; if view 0 is shorter than `length`, repeat_pad tiles it up to `length`, otherwise the views
//     stay at view 0's length;
//   * the view assembly of Dataset_for.__getitem__ (/root/reference/datautils/asvspoof_2019_augall_3.py:133-142:
//     np.concatenate(axis=1) -> (length, V) float32) and the [1, length, V] -> [V, length] reshape the training loop does
//     next (/root/reference/main.py:57-60).
// The crop start is an integer the HOST draws (int(np.random.rand() * (new_len - length)), lines 256 / 273) so that the
// global numpy stream stays where the reference leaves it. The kernel is pure index arithmetic and HBM-bound: one
// coalesced read and one coalesced write per output sample; views stay on the device between RawBoost and the model.
#include <algorithm>

#include "rb_common.cuh"

namespace rb {
namespace {

// out_len[g] = (first_len < length && !repeat_pad) ? first_len : length
// sample k of view v of group g:  m = start + k (mod first_len when view 0 is tiled);  value = m < len_v ? view[m]
//                                 : (repeat_pad ? view[m mod len_v] : 0)
// View v of group g lives in row r = view_row[g*V + v] of `views` (r >= 0) or in row -1 - r of `views_b` (r < 0); without a row
// table the views are the rows g*V + v of `views`. Both sources share the row stride and the per-row lengths (RawBoost keeps
// an utterance's length), so an item's original and augmented waveforms are read where they already are.
template <int LAYOUT>  // 0: [G][length][V] (Dataset layout)   1: [G][V][length] (model layout)
__global__ void __launch_bounds__(256)
multiview_kernel(const float* __restrict__ views, const float* __restrict__ views_b, const int32_t* __restrict__ view_row,
                 const int32_t* __restrict__ len_arr, int V, int ld, const int32_t* __restrict__ start_arr, int length, int repeat_pad,
                 float* __restrict__ out, int32_t* __restrict__ out_len, const float* __restrict__ view_label,
                 float* __restrict__ labels) {
  const int g = blockIdx.y;
  auto row_of = [&](int v) { return view_row ? view_row[(size_t)g * V + v] : (int)((size_t)g * V + v); };
  auto len_of = [&](int r) { return len_arr[r >= 0 ? r : -1 - r]; };
  const int first_len = len_of(row_of(0));
  if (labels && blockIdx.x == 0 && threadIdx.x < V) labels[(size_t)g * V + threadIdx.x] = view_label[threadIdx.x];
  const bool tile_first = first_len < length && repeat_pad;
  const int olen = (first_len < length && !repeat_pad) ? first_len : length;
  const int start = (first_len < length) ? 0 : start_arr[g];
  if (blockIdx.x == 0 && threadIdx.x == 0 && out_len) out_len[g] = olen;
  const size_t total = (size_t)olen * V;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    int k, v;
    if (LAYOUT == 0) {  // consecutive threads walk the V columns of one output row, then the next row
      k = (int)(e / V);
      v = (int)(e - (size_t)k * V);
    } else {            // consecutive threads walk consecutive samples of one view: coalesced on both sides
      v = (int)(e / olen);
      k = (int)(e - (size_t)v * olen);
    }
    const int r = row_of(v);
    const int lv = len_of(r);
    const float* src = r >= 0 ? views + (size_t)r * ld : views_b + (size_t)(-1 - r) * ld;
    int m = start + k;
    if (tile_first && first_len > 0) m %= first_len;
    float val = 0.f;
    if (lv > 0) {
      if (m < lv) val = __ldg(src + m);
      else if (repeat_pad) val = __ldg(src + (m % lv));
    }
    if (LAYOUT == 0) out[((size_t)g * length + k) * V + v] = val;
    else out[((size_t)g * V + v) * length + k] = val;
  }
}

}  // namespace
}  // namespace rb

using namespace rb;

extern "C" int rb_multiview_assemble_ex(const float* views, const float* views_b, const int32_t* view_row, const int32_t* len, int G,
                                        int V, int ld, const int32_t* start, int length, int repeat_pad, int layout, float* out,
                                        int32_t* out_len, const float* view_label, float* labels, void* stream) {
  if (G < 0 || V < 0 || ld < 0 || length < 0 || (layout != 0 && layout != 1)) return RB_ERR_INVALID_ARG;
  if (G == 0 || V == 0 || length == 0) return RB_OK;
  if (!views || !len || !start || !out) return RB_ERR_INVALID_ARG;
  if ((labels != nullptr) != (view_label != nullptr) || (labels && V > 256)) return RB_ERR_INVALID_ARG;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t per_group = (size_t)length * V;
  const int bx = (int)std::min<size_t>((per_group + 255) / 256, 1184);  // 8 CTAs per SM cover one group; grid-stride beyond
  for (int g0 = 0; g0 < G; g0 += 65535) {
    const int ng = std::min(65535, G - g0);
    const dim3 grid(bx, ng);
    // without a row table the group's rows are consecutive in `views`; with one, the table holds absolute rows
    const float* vw = view_row ? views : views + (size_t)g0 * V * ld;
    const int32_t* ln = view_row ? len : len + (size_t)g0 * V;
    const int32_t* vr = view_row ? view_row + (size_t)g0 * V : nullptr;
    float* o = out + (size_t)g0 * per_group;
    int32_t* ol = out_len ? out_len + g0 : nullptr;
    float* lb = labels ? labels + (size_t)g0 * V : nullptr;
    if (layout == 0) multiview_kernel<0><<<grid, 256, 0, st>>>(vw, views_b, vr, ln, V, ld, start + g0, length, repeat_pad, o, ol, view_label, lb);
    else multiview_kernel<1><<<grid, 256, 0, st>>>(vw, views_b, vr, ln, V, ld, start + g0, length, repeat_pad, o, ol, view_label, lb);
    RB_LAUNCH_CHECK();
  }
  return RB_OK;
}

extern "C" int rb_multiview_assemble(const float* views, const int32_t* len, int G, int V, int ld, const int32_t* start, int length,
                                     int repeat_pad, int layout, float* out, int32_t* out_len, void* stream) {
  return rb_multiview_assemble_ex(views, nullptr, nullptr, len, G, V, ld, start, length, repeat_pad, layout, out, out_len, nullptr,
                                  nullptr, stream);
}
