// dw_dilated.cu - depthwise k x k convolution with dilation > 1 (forward with integer statistics, dgrad, wgrad): the
// convolutions of SSDLite-MobileNetV2's last two stages (Object_Detection/ssd_qmv2.py:40-52, 137-138: dilation 2, padding =
// dilation), MobileNetV3(dilated=True) (Classification/models/imagenet/mobilenetv3.py:129-131) and ESPNetV2's EESP branches.
// Same operand conventions as dw_conv.cu - uint8 NHWC activations with row pitch ldx, int8 weights [k*k][C] tap-major,
//   I = sum_taps (x - zp_a) * (w - zp_w)   (taps outside the image contribute 0),  any padding 0 .. dilation * (k - 1)
// (the SSD extras' depthwise 3x3 with padding 0, ssd_qmv2.py:188-203, takes this path too)
// - but plain gather kernels: one thread per (pixel, 4-channel group), no register transposes, no shared-memory tiles
// (shared memory only combines the statistics of a CTA).
// These layers sit on the 19x19 .. 10x10 planes of the detection / segmentation backbones; the tuned stride-1/2 kernels of
// dw_conv.cu remain the path for dilation 1.
#include <algorithm>
#include "common.cuh"

namespace frost {

__device__ __forceinline__ int dd_sext(unsigned w, int i) { return (int)(signed char)((w >> (8 * i)) & 0xffu); }
__device__ __forceinline__ int dd_zext(unsigned w, int i) { return (int)((w >> (8 * i)) & 0xffu); }

// CTA = 32 channel groups (lanes: 128 contiguous bytes per tap) x 8 pixel slots (warps); a thread keeps its channel group,
// its statistics stay in registers, the 8 warps meet in shared memory and one warp flushes: 8x fewer integer atomics than a
// flush per thread, which is what bounds this kernel once enough threads are in flight to hide the gather latency
__global__ void __launch_bounds__(256) dw_dil_fwd_kernel(const uint8_t* xq, const int32_t* x_zp_p, const int8_t* wq,
                                                         const int32_t* w_zp_p, int N, int H, int W, int C, int ldx, int k, int S, int D,
                                                         int pad, int Ho, int Wo, int32_t* acc_out, FrostChanStats* stats) {
  __shared__ long long s_sum[8][32][4];
  __shared__ unsigned long long s_sq[8][32][4];
  __shared__ int s_mn[8][32][4], s_mx[8][32][4];
  pdl_enter();
  const int lane = threadIdx.x & 31, slot = threadIdx.x >> 5;
  const int CG = C >> 2;
  const int cg = blockIdx.y * 32 + lane;
  const bool active = cg < CG;
  const int zp_a = *x_zp_p, zp_w = *w_zp_p;
  const int64_t M = (int64_t)N * Ho * Wo;
  long long st_sum[4] = {0, 0, 0, 0};
  unsigned long long st_sq[4] = {0, 0, 0, 0};
  int st_mn[4] = {INT_MAX, INT_MAX, INT_MAX, INT_MAX}, st_mx[4] = {INT_MIN, INT_MIN, INT_MIN, INT_MIN};
  if (active) {
    for (int64_t p = blockIdx.x * 8 + slot; p < M; p += (int64_t)gridDim.x * 8) {
      const int ow = (int)(p % Wo), oh = (int)((p / Wo) % Ho), n = (int)(p / ((int64_t)Wo * Ho));
      int acc[4] = {0, 0, 0, 0};
      for (int r = 0; r < k; ++r) {
        const int ih = oh * S - pad + r * D;
        if ((unsigned)ih >= (unsigned)H) continue;
        for (int s = 0; s < k; ++s) {
          const int iw = ow * S - pad + s * D;
          if ((unsigned)iw >= (unsigned)W) continue;
          const unsigned xw = __ldcg(reinterpret_cast<const unsigned*>(xq + (((int64_t)n * H + ih) * W + iw) * ldx + cg * 4));
          const unsigned ww = __ldcg(reinterpret_cast<const unsigned*>(wq + (int64_t)(r * k + s) * C + cg * 4));
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) acc[ch] += (dd_zext(xw, ch) - zp_a) * (dd_sext(ww, ch) - zp_w);
        }
      }
      *reinterpret_cast<int4*>(acc_out + p * C + cg * 4) = make_int4(acc[0], acc[1], acc[2], acc[3]);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const int I = acc[ch];
        st_sum[ch] += I;
        st_sq[ch] += (unsigned long long)((long long)I * (long long)I);
        st_mn[ch] = min(st_mn[ch], I);
        st_mx[ch] = max(st_mx[ch], I);
      }
    }
  }
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    s_sum[slot][lane][ch] = st_sum[ch]; s_sq[slot][lane][ch] = st_sq[ch];
    s_mn[slot][lane][ch] = st_mn[ch]; s_mx[slot][lane][ch] = st_mx[ch];
  }
  __syncthreads();
  if (slot == 0 && active) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      long long sum = 0;
      unsigned long long sq = 0;
      int mn = INT_MAX, mx = INT_MIN;
      for (int w = 0; w < 8; ++w) {
        sum += s_sum[w][lane][ch]; sq += s_sq[w][lane][ch];
        mn = min(mn, s_mn[w][lane][ch]); mx = max(mx, s_mx[w][lane][ch]);
      }
      chan_stats_flush(stats + cg * 4 + ch, sum, sq, mn, mx);
    }
  }
}

// dx[n][ih][iw][c] (+)= s_w * sum_{r,s} dz[n][oh][ow][c] * (w[r][s][c] - zp_w),  oh*S - pad + r*D == ih, ow*S - pad + s*D == iw
__global__ void __launch_bounds__(256) dw_dil_dgrad_kernel(const float* dz, const int8_t* wq, const float* w_scale_p,
                                                           const int32_t* w_zp_p, int N, int H, int W, int C, int k, int S, int D,
                                                           int pad, int Ho, int Wo, float* dx, int accumulate) {
  pdl_enter();
  const int CG = C >> 2;
  const int64_t total = (int64_t)N * H * W * CG;
  const float sw = *w_scale_p;
  const int zp_w = *w_zp_p;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % CG);
    const int64_t p = i / CG;
    const int iw = (int)(p % W), ih = (int)((p / W) % H), n = (int)(p / ((int64_t)W * H));
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < k; ++r) {
      const int nh = ih + pad - r * D;
      if (nh < 0 || nh % S != 0 || nh / S >= Ho) continue;
      for (int s = 0; s < k; ++s) {
        const int nw = iw + pad - s * D;
        if (nw < 0 || nw % S != 0 || nw / S >= Wo) continue;
        const float4 g = __ldcg(reinterpret_cast<const float4*>(dz + (((int64_t)n * Ho + nh / S) * Wo + nw / S) * C + cg * 4));
        const unsigned ww = __ldcg(reinterpret_cast<const unsigned*>(wq + (int64_t)(r * k + s) * C + cg * 4));
        a[0] = fmaf(g.x, (float)(dd_sext(ww, 0) - zp_w), a[0]);
        a[1] = fmaf(g.y, (float)(dd_sext(ww, 1) - zp_w), a[1]);
        a[2] = fmaf(g.z, (float)(dd_sext(ww, 2) - zp_w), a[2]);
        a[3] = fmaf(g.w, (float)(dd_sext(ww, 3) - zp_w), a[3]);
      }
    }
    float4* o = reinterpret_cast<float4*>(dx + p * C + cg * 4);
    float4 v = make_float4(a[0] * sw, a[1] * sw, a[2] * sw, a[3] * sw);
    if (accumulate) {
      const float4 old = *o;
      v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
    }
    *o = v;
  }
}

// dwq[r*k+s][c] += s_a * sum_{n,oh,ow} dz[n][oh][ow][c] * (x[n][oh*S-pad+r*D][ow*S-pad+s*D][c] - zp_a)
// thread <-> fixed (tap, channel group), output pixels strided; one atomicAdd per thread and channel at the end
__global__ void __launch_bounds__(256) dw_dil_wgrad_kernel(const float* dz, const uint8_t* xq, const float* x_scale_p,
                                                           const int32_t* x_zp_p, int N, int H, int W, int C, int ldx, int k, int S, int D,
                                                           int pad, int Ho, int Wo, int64_t n_threads, float* dwq) {
  pdl_enter();
  const int64_t tid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (tid >= n_threads) return;
  const int CG = C >> 2, KK = k * k;
  const int cg = (int)(tid % CG), tap = (int)((tid / CG) % KK);
  const int r = tap / k, s = tap % k;
  const int zp_a = *x_zp_p;
  const int64_t M = (int64_t)N * Ho * Wo, pstride = n_threads / ((int64_t)CG * KK);
  float a[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t p = tid / ((int64_t)CG * KK); p < M; p += pstride) {
    const int ow = (int)(p % Wo), oh = (int)((p / Wo) % Ho), n = (int)(p / ((int64_t)Wo * Ho));
    const int ih = oh * S - pad + r * D, iw = ow * S - pad + s * D;
    if ((unsigned)ih >= (unsigned)H || (unsigned)iw >= (unsigned)W) continue;
    const float4 g = __ldcg(reinterpret_cast<const float4*>(dz + p * C + cg * 4));
    const unsigned xw = __ldcg(reinterpret_cast<const unsigned*>(xq + (((int64_t)n * H + ih) * W + iw) * ldx + cg * 4));
    a[0] = fmaf(g.x, (float)(dd_zext(xw, 0) - zp_a), a[0]);
    a[1] = fmaf(g.y, (float)(dd_zext(xw, 1) - zp_a), a[1]);
    a[2] = fmaf(g.z, (float)(dd_zext(xw, 2) - zp_a), a[2]);
    a[3] = fmaf(g.w, (float)(dd_zext(xw, 3) - zp_a), a[3]);
  }
  const float sa = *x_scale_p;
  float* o = dwq + (int64_t)tap * C + cg * 4;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch)
    if (a[ch] != 0.f) atomicAdd(o + ch, a[ch] * sa);
}

static bool dil_shape_ok(int C, int k, int stride, int dilation, int pad) {
  return C > 0 && C % 4 == 0 && (k == 3 || k == 5) && (stride == 1 || stride == 2) && dilation >= 1 && dilation <= 16 && pad >= 0 &&
         pad <= dilation * (k - 1);
}
static void dil_out(int H, int W, int k, int S, int D, int pad, int* Ho, int* Wo) {
  const int span = D * (k - 1) + 1;
  *Ho = (H + 2 * pad - span) / S + 1;
  *Wo = (W + 2 * pad - span) / S + 1;
}

}  // namespace frost

using namespace frost;

extern "C" int frost_dw_conv_forward_dilated(const uint8_t* xq, int ldx, const int32_t* x_zp, const int8_t* wq, const int32_t* w_zp,
                                             int N, int H, int W, int C, int k, int stride, int dilation, int pad, int32_t* acc,
                                             FrostChanStats* stats, void* stream) {
  FROST_REQUIRE(xq && x_zp && wq && w_zp && acc && stats, "frost_dw_conv_forward_dilated: null pointer");
  FROST_REQUIRE(ldx >= C && ldx % 4 == 0, "frost_dw_conv_forward_dilated: ldx=%d must be >= C and a multiple of 4", ldx);
  FROST_REQUIRE(N > 0 && H > 0 && W > 0 && dil_shape_ok(C, k, stride, dilation, pad),
                "frost_dw_conv_forward_dilated: bad shape (C%%4==0, k in {3,5}, stride in {1,2}, dilation 1..16, 0 <= pad <= dilation*(k-1))");
  int Ho, Wo;
  dil_out(H, W, k, stride, dilation, pad, &Ho, &Wo);
  FROST_REQUIRE(Ho > 0 && Wo > 0, "frost_dw_conv_forward_dilated: empty output");
  const int64_t M = (int64_t)N * Ho * Wo;
  const int gy = (C / 4 + 31) / 32;
  // ~8 CTAs of 256 threads per SM: the loop is a chain of dependent gathers (latency-bound)
  const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(M, 8), (int64_t)kNumSMs * 8 / gy + 1));
  dw_dil_fwd_kernel<<<dim3(gx, gy), 256, 0, (cudaStream_t)stream>>>(xq, x_zp, wq, w_zp, N, H, W, C, ldx, k, stride, dilation, pad, Ho, Wo,
                                                                    acc, stats);
  FROST_LAUNCH_CHECK("dw_dil_fwd");
  return FROST_OK;
}

extern "C" int frost_dw_dgrad_dilated(const float* dz, const int8_t* wq, const float* w_scale, const int32_t* w_zp, int N, int H, int W,
                                      int C, int k, int stride, int dilation, int pad, float* dx, int accumulate, void* stream) {
  FROST_REQUIRE(dz && wq && w_scale && w_zp && dx, "frost_dw_dgrad_dilated: null pointer");
  FROST_REQUIRE(N > 0 && H > 0 && W > 0 && dil_shape_ok(C, k, stride, dilation, pad), "frost_dw_dgrad_dilated: bad shape");
  FROST_REQUIRE(((reinterpret_cast<uintptr_t>(dz) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0, "frost_dw_dgrad_dilated: 16-byte alignment");
  int Ho, Wo;
  dil_out(H, W, k, stride, dilation, pad, &Ho, &Wo);
  const int64_t total = (int64_t)N * H * W * (C / 4);
  dw_dil_dgrad_kernel<<<grid_for(total, 256, kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(dz, wq, w_scale, w_zp, N, H, W, C, k, stride,
                                                                                          dilation, pad, Ho, Wo, dx, accumulate);
  FROST_LAUNCH_CHECK("dw_dil_dgrad");
  return FROST_OK;
}

extern "C" int frost_dw_wgrad_dilated(const float* dz, const uint8_t* xq, int ldx, const float* x_scale, const int32_t* x_zp, int N, int H,
                                      int W, int C, int k, int stride, int dilation, int pad, float* dwq, void* stream) {
  FROST_REQUIRE(dz && xq && x_scale && x_zp && dwq, "frost_dw_wgrad_dilated: null pointer");
  FROST_REQUIRE(ldx >= C && ldx % 4 == 0, "frost_dw_wgrad_dilated: ldx=%d must be >= C and a multiple of 4", ldx);
  FROST_REQUIRE(N > 0 && H > 0 && W > 0 && dil_shape_ok(C, k, stride, dilation, pad), "frost_dw_wgrad_dilated: bad shape");
  FROST_REQUIRE((reinterpret_cast<uintptr_t>(dz) & 15) == 0, "frost_dw_wgrad_dilated: dz must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int Ho, Wo;
  dil_out(H, W, k, stride, dilation, pad, &Ho, &Wo);
  if (cudaMemsetAsync(dwq, 0, sizeof(float) * (size_t)k * k * C, st) != cudaSuccess) {
    set_error("frost_dw_wgrad_dilated: memset failed");
    return FROST_ECUDA;
  }
  const int64_t lanes = (int64_t)(C / 4) * k * k, M = (int64_t)N * Ho * Wo;
  const int64_t per = std::max<int64_t>(1, std::min<int64_t>(M, (int64_t)kNumSMs * 4096 / lanes + 1));
  const int64_t n_threads = per * lanes;
  dw_dil_wgrad_kernel<<<(unsigned)ceil_div(n_threads, 256), 256, 0, st>>>(dz, xq, x_scale, x_zp, N, H, W, C, ldx, k, stride, dilation, pad,
                                                                          Ho, Wo, n_threads, dwq);
  FROST_LAUNCH_CHECK("dw_dil_wgrad");
  return FROST_OK;
}
