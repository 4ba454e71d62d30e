// multibox.cu - target matching of the SSD MultiBox loss (Object_Detection/layers/modules/multibox_loss.py:66-74 calls
// layers/box_utils.py:71-113 `match` once per image from a Python loop, on the CPU, and copies loc_t / conf_t to the device):
// jaccard overlap of every ground-truth box with every prior, best prior per truth, best truth per prior, the "every truth
// keeps its best prior" override, background below the threshold, offset encoding (box_utils.py:115-138).
// One CTA per image, the whole batch in one launch; truths live in shared memory.  The arithmetic is the reference's, operation
// by operation in fp32 without contraction (IoU = inter / (area_a + area_b - inter), point_form(priors) = c -+ wh / 2), so the
// matched indices - and with them conf_t - are the reference's bit for bit; loc_t differs only by logf's last ulp.
#include "common.cuh"

namespace frost {

constexpr int MB_MAX_OBJ = 128;
constexpr int MB_THREADS = 256;

__global__ void __launch_bounds__(MB_THREADS) multibox_match_kernel(const float* truths, const int64_t* labels, const int* num_objs,
                                                                    int max_obj, const float* priors, int P, float threshold, float var0,
                                                                    float var1, float* loc_t, int64_t* conf_t, float* best_ov,
                                                                    int* best_idx) {
  __shared__ float s_t[MB_MAX_OBJ][4];
  __shared__ float s_area[MB_MAX_OBJ];
  __shared__ float s_bp_ov[MB_MAX_OBJ];      // best prior per truth: overlap
  __shared__ int s_bp_idx[MB_MAX_OBJ];       //                       index (first maximum, like torch.max)
  __shared__ float s_red_ov[MB_THREADS / 32];
  __shared__ int s_red_idx[MB_THREADS / 32];
  const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n = num_objs[img];
  float* bo = best_ov + (int64_t)img * P;    // best truth per prior: overlap
  int* bi = best_idx + (int64_t)img * P;     //                       index
  for (int j = tid; j < n; j += MB_THREADS) {
    const float* t = truths + ((int64_t)img * max_obj + j) * 4;
    s_t[j][0] = t[0]; s_t[j][1] = t[1]; s_t[j][2] = t[2]; s_t[j][3] = t[3];
    s_area[j] = __fmul_rn(__fsub_rn(t[2], t[0]), __fsub_rn(t[3], t[1]));
  }
  __syncthreads();
  auto iou = [&](int j, float x0, float y0, float x1, float y1, float area_b) {
    const float w = fmaxf(__fsub_rn(fminf(s_t[j][2], x1), fmaxf(s_t[j][0], x0)), 0.0f);
    const float h = fmaxf(__fsub_rn(fminf(s_t[j][3], y1), fmaxf(s_t[j][1], y0)), 0.0f);
    const float inter = __fmul_rn(w, h);
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(s_area[j], area_b), inter));
  };
  // ---- pass 1: best truth per prior (first maximum over the truths)
  for (int p = tid; p < P; p += MB_THREADS) {
    const float cx = priors[4 * p], cy = priors[4 * p + 1], w = priors[4 * p + 2], h = priors[4 * p + 3];
    const float x0 = __fsub_rn(cx, __fdiv_rn(w, 2.0f)), y0 = __fsub_rn(cy, __fdiv_rn(h, 2.0f));
    const float x1 = __fadd_rn(cx, __fdiv_rn(w, 2.0f)), y1 = __fadd_rn(cy, __fdiv_rn(h, 2.0f));
    const float area_b = __fmul_rn(__fsub_rn(x1, x0), __fsub_rn(y1, y0));
    float best = -1.0f;
    int arg = 0;
    for (int j = 0; j < n; ++j) {
      const float o = iou(j, x0, y0, x1, y1, area_b);
      if (o > best) { best = o; arg = j; }
    }
    bo[p] = best;
    bi[p] = arg;
  }
  // ---- pass 2: best prior per truth (first maximum over the priors): block arg-max per truth
  for (int j = 0; j < n; ++j) {
    float best = -1.0f;
    int arg = 0x7fffffff;
    for (int p = tid; p < P; p += MB_THREADS) {
      const float cx = priors[4 * p], cy = priors[4 * p + 1], w = priors[4 * p + 2], h = priors[4 * p + 3];
      const float x0 = __fsub_rn(cx, __fdiv_rn(w, 2.0f)), y0 = __fsub_rn(cy, __fdiv_rn(h, 2.0f));
      const float x1 = __fadd_rn(cx, __fdiv_rn(w, 2.0f)), y1 = __fadd_rn(cy, __fdiv_rn(h, 2.0f));
      const float o = iou(j, x0, y0, x1, y1, __fmul_rn(__fsub_rn(x1, x0), __fsub_rn(y1, y0)));
      if (o > best) { best = o; arg = p; }             // p ascends within a thread: the first maximum stays
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, off);
      const int ab = __shfl_xor_sync(0xffffffffu, arg, off);
      if (ob > best || (ob == best && ab < arg)) { best = ob; arg = ab; }
    }
    if (lane == 0) { s_red_ov[warp] = best; s_red_idx[warp] = arg; }
    __syncthreads();
    if (tid == 0) {
      for (int w2 = 1; w2 < MB_THREADS / 32; ++w2)
        if (s_red_ov[w2] > best || (s_red_ov[w2] == best && s_red_idx[w2] < arg)) { best = s_red_ov[w2]; arg = s_red_idx[w2]; }
      s_bp_ov[j] = best;
      s_bp_idx[j] = arg;
    }
    __syncthreads();
  }
  // ---- every truth keeps its best prior (box_utils.py:98-103; a later truth wins a shared prior)
  if (tid == 0) {
    for (int j = 0; j < n; ++j) bo[s_bp_idx[j]] = 2.0f;
    for (int j = 0; j < n; ++j) bi[s_bp_idx[j]] = j;
  }
  __syncthreads();
  // ---- labels and encoded offsets
  for (int p = tid; p < P; p += MB_THREADS) {
    const int j = n > 0 ? bi[p] : 0;
    const int64_t o = (int64_t)img * P + p;
    if (n == 0) {
      conf_t[o] = 0;
      loc_t[4 * o] = loc_t[4 * o + 1] = loc_t[4 * o + 2] = loc_t[4 * o + 3] = 0.0f;
      continue;
    }
    conf_t[o] = bo[p] < threshold ? 0 : labels[(int64_t)img * max_obj + j] + 1;
    const float cx = priors[4 * p], cy = priors[4 * p + 1], w = priors[4 * p + 2], h = priors[4 * p + 3];
    const float gx = __fdiv_rn(__fsub_rn(__fdiv_rn(__fadd_rn(s_t[j][0], s_t[j][2]), 2.0f), cx), __fmul_rn(var0, w));
    const float gy = __fdiv_rn(__fsub_rn(__fdiv_rn(__fadd_rn(s_t[j][1], s_t[j][3]), 2.0f), cy), __fmul_rn(var0, h));
    const float gw = __fdiv_rn(logf(__fdiv_rn(__fsub_rn(s_t[j][2], s_t[j][0]), w)), var1);
    const float gh = __fdiv_rn(logf(__fdiv_rn(__fsub_rn(s_t[j][3], s_t[j][1]), h)), var1);
    loc_t[4 * o] = gx; loc_t[4 * o + 1] = gy; loc_t[4 * o + 2] = gw; loc_t[4 * o + 3] = gh;
  }
}

}  // namespace frost

using namespace frost;

extern "C" int frost_multibox_match(const float* truths, const int64_t* labels, const int* num_objs, int batch, int max_obj,
                                    const float* priors, int num_priors, float threshold, float var0, float var1, float* loc_t,
                                    int64_t* conf_t, float* scratch_overlap, int* scratch_index, void* stream) {
  FROST_REQUIRE(truths && labels && num_objs && priors && loc_t && conf_t && scratch_overlap && scratch_index,
                "frost_multibox_match: null pointer");
  FROST_REQUIRE(batch > 0 && num_priors > 0 && max_obj > 0 && max_obj <= MB_MAX_OBJ,
                "frost_multibox_match: batch, num_priors > 0 and 0 < max_obj <= %d", MB_MAX_OBJ);
  multibox_match_kernel<<<batch, MB_THREADS, 0, (cudaStream_t)stream>>>(truths, labels, num_objs, max_obj, priors, num_priors, threshold,
                                                                        var0, var1, loc_t, conf_t, scratch_overlap, scratch_index);
  FROST_LAUNCH_CHECK("multibox_match");
  return FROST_OK;
}
