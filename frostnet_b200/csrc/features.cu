// features.cu - what the feature-backbone variant (frostnet_features.py:342-352) needs on top of the
// classifier path: it has no QuantStub, so its stem convolves the raw fp32 NCHW image with the
// fake-quantised stem weights (the reference's nniqat.ConvBnReLU2d on an unquantised input), and it returns
// four dequantised NCHW feature maps whose gradients flow back into the NHWC gradient chain.
#include "common.cuh"

namespace frost {

constexpr int FS_MAXC = 32;
constexpr int FS_THREADS = 256;

__global__ void stats_reset_f32_kernel(FrostChanStats* s, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    s[i].sum = 0;      // double 0.0
    s[i].sq_lo = 0;    // double 0.0
    s[i].sq_hi = 0;
    s[i].min = __float_as_int(INFINITY);
    s[i].max = __float_as_int(-INFINITY);
  }
}

// z[m][co] = sum_{r,s,ci} x[n][ci][ih][iw] * (q_w[co][r][s][ci] - zp_w);  one thread per output pixel.
__global__ void __launch_bounds__(FS_THREADS) stem_conv_fwd_f32_kernel(const float* x, const int8_t* wq,
                                                                      const int32_t* w_zp_p, int N, int H, int W,
                                                                      int cin, int cout, int k, int stride, int pad, int Ho,
                                                                      int Wo, float* z, FrostChanStats* stats) {
  extern __shared__ float fs_mem[];
  const int KK = k * k * cin;
  float* s_w = fs_mem;                          // [KK][FS_MAXC]
  float* s_tile = fs_mem + KK * FS_MAXC;        // [FS_THREADS][FS_MAXC+1]
  const float zp_w = (float)*w_zp_p;
  for (int i = threadIdx.x; i < KK * FS_MAXC; i += blockDim.x) {
    const int t = i / FS_MAXC, co = i % FS_MAXC;
    s_w[i] = (co < cout) ? (float)wq[(int64_t)co * KK + t] - zp_w : 0.0f;
  }
  __syncthreads();
  const int64_t total = (int64_t)N * Ho * Wo;
  const int64_t p0 = (int64_t)blockIdx.x * FS_THREADS;
  const int64_t p = p0 + threadIdx.x;
  float acc[FS_MAXC];
#pragma unroll
  for (int c = 0; c < FS_MAXC; ++c) acc[c] = 0.0f;
  if (p < total) {
    const int ow = (int)(p % Wo);
    const int64_t t1 = p / Wo;
    const int oh = (int)(t1 % Ho);
    const int n = (int)(t1 / Ho);
    for (int r = 0; r < k; ++r) {
      const int ih = oh * stride - pad + r;
      if (ih < 0 || ih >= H) continue;
      for (int s = 0; s < k; ++s) {
        const int iw = ow * stride - pad + s;
        if (iw < 0 || iw >= W) continue;
        for (int ci = 0; ci < cin; ++ci) {
          const float xa = ld_cg(x + (((int64_t)n * cin + ci) * H + ih) * W + iw);
          const float* wrow = s_w + ((r * k + s) * cin + ci) * FS_MAXC;
#pragma unroll
          for (int c = 0; c < FS_MAXC; ++c) acc[c] = fmaf(xa, wrow[c], acc[c]);
        }
      }
    }
  }
#pragma unroll
  for (int c = 0; c < FS_MAXC; ++c) s_tile[threadIdx.x * (FS_MAXC + 1) + c] = acc[c];
  __syncthreads();
  const int npix = (int)min((int64_t)FS_THREADS, total - p0);
  const int c = threadIdx.x % FS_MAXC, g = threadIdx.x / FS_MAXC;
  constexpr int G = FS_THREADS / FS_MAXC;
  if (c < cout) {
    double s = 0.0, sq = 0.0;
    float mn = INFINITY, mx = -INFINITY;
    for (int px = g; px < npix; px += G) {
      const float v = s_tile[px * (FS_MAXC + 1) + c];
      z[(p0 + px) * cout + c] = v;
      s += (double)v;
      sq += (double)v * (double)v;
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
    if (mn <= mx) {
      FrostChanStats* st = stats + c;
      atomicAdd(reinterpret_cast<double*>(&st->sum), s);
      atomicAdd(reinterpret_cast<double*>(&st->sq_lo), sq);
      atomic_min_float(reinterpret_cast<float*>(&st->min), mn);
      atomic_max_float(reinterpret_cast<float*>(&st->max), mx);
    }
  }
}

// dwq[co][r][s][ci] += sum_pix dz[pix][co] * x[n][ci][ih][iw]
constexpr int FSW_PIX = 64;
__global__ void __launch_bounds__(256) stem_wgrad_f32_kernel(const float* dz, const float* x, int N, int H,
                                                            int W, int cin, int cout, int k, int stride, int pad, int Ho, int Wo,
                                                            int64_t pix_per_block, float* dwq) {
  extern __shared__ float fs_mem[];
  const int KK = k * k * cin;
  const int KP = KK | 1;
  float* s_patch = fs_mem;                     // [FSW_PIX][KP]
  float* s_dz = fs_mem + FSW_PIX * KP;         // [FSW_PIX][32]
  const int co = threadIdx.x & 31, j = threadIdx.x >> 5;
  const int64_t total = (int64_t)N * Ho * Wo;
  const int64_t p_begin = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p_end = min(total, p_begin + pix_per_block);
  const int fp = threadIdx.x >> 2, ff = threadIdx.x & 3;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int64_t pb = p_begin; pb < p_end; pb += FSW_PIX) {
    const int np = (int)min((int64_t)FSW_PIX, p_end - pb);
    {
      const int64_t p = pb + fp;
      const bool pv = fp < np;
      int ow = 0, oh = 0, n = 0;
      if (pv) {
        ow = (int)(p % Wo);
        const int64_t t1 = p / Wo;
        oh = (int)(t1 % Ho);
        n = (int)(t1 / Ho);
      }
      const int ih0 = oh * stride - pad, iw0 = ow * stride - pad;
      for (int t = ff; t < KK; t += 4) {
        const int ci = t % cin, rs = t / cin, r = rs / k, sx = rs - r * k;
        const int ih = ih0 + r, iw = iw0 + sx;
        float v = 0.0f;
        if (pv && (unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W)
          v = ld_cg(x + (((int64_t)n * cin + ci) * H + ih) * W + iw);
        s_patch[fp * KP + t] = v;
      }
    }
    for (int i = threadIdx.x; i < FSW_PIX * 32; i += blockDim.x) {
      const int pl = i >> 5, c = i & 31;
      s_dz[i] = (pl < np && c < cout) ? ld_cg(dz + (pb + pl) * cout + c) : 0.0f;
    }
    __syncthreads();
#pragma unroll 4
    for (int pl = 0; pl < FSW_PIX; ++pl) {
      const float d = s_dz[pl * 32 + co];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int t = j + 8 * i;
        if (t < KK) acc[i] = fmaf(d, s_patch[pl * KP + t], acc[i]);
      }
    }
    __syncthreads();
  }
  if (co < cout) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int t = j + 8 * i;
      if (t < KK) atomicAdd(dwq + (int64_t)co * KK + t, acc[i]);
    }
  }
}

// NHWC uint8 indices -> NCHW fp32 values, 32x32 (pixel x channel) tiles through shared memory
__global__ void __launch_bounds__(256) dequant_to_nchw_kernel(const uint8_t* q, int ldq, const float* scale_p,
                                                             const int32_t* zp_p, int HW, int C, float* y) {
  __shared__ float tile[32][33];
  const float s = *scale_p, zp = (float)*zp_p;
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows per pass
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    tile[r][tx] = (p < HW && c < C) ? fq_dequant((float)q[((int64_t)n * HW + p) * ldq + c], zp, s) : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    if (c < C && p < HW) y[((int64_t)n * C + c) * HW + p] = tile[tx][r];
  }
}

__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const float* g, int HW, int C, float* out,
                                                          int accumulate) {
  __shared__ float tile[32][33];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int r = ty; r < 32; r += 8) {
    const int c = c0 + r, p = p0 + tx;
    tile[r][tx] = (c < C && p < HW) ? g[((int64_t)n * C + c) * HW + p] : 0.0f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int p = p0 + r, c = c0 + tx;
    if (p < HW && c < C) {
      float* o = out + ((int64_t)n * HW + p) * C + c;
      *o = accumulate ? *o + tile[tx][r] : tile[tx][r];
    }
  }
}

}  // namespace frost

using namespace frost;

extern "C" int frost_stats_reset_f32(FrostChanStats* stats, int64_t n, void* stream) {
  FROST_REQUIRE(stats && n >= 0, "frost_stats_reset_f32: bad args");
  if (n == 0) return FROST_OK;
  stats_reset_f32_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(stats, n);
  FROST_LAUNCH_CHECK("stats_reset_f32");
  return FROST_OK;
}

extern "C" int frost_stem_conv_forward_f32(const float* x_nchw, const int8_t* wq, const int32_t* w_zp, int N, int H, int W,
                                           int cin, int cout, int k, int stride, int pad, float* z, FrostChanStats* stats,
                                           void* stream) {
  FROST_REQUIRE(x_nchw && wq && w_zp && z && stats, "frost_stem_conv_forward_f32: null pointer");
  FROST_REQUIRE(cout > 0 && cout <= FS_MAXC && cin > 0 && cin <= 8 && k > 0 && k <= 7,
                "frost_stem_conv_forward_f32: cout<=32, cin<=8, k<=7");
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int64_t total = (int64_t)N * Ho * Wo;
  const size_t smem = sizeof(float) * ((size_t)k * k * cin * FS_MAXC + (size_t)FS_THREADS * (FS_MAXC + 1));
  if (smem > 48 * 1024 && first_use_on_device(reinterpret_cast<const void*>(&stem_conv_fwd_f32_kernel)))
    cudaFuncSetAttribute(stem_conv_fwd_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  stem_conv_fwd_f32_kernel<<<(unsigned)ceil_div(total, FS_THREADS), FS_THREADS, smem, (cudaStream_t)stream>>>(
      x_nchw, wq, w_zp, N, H, W, cin, cout, k, stride, pad, Ho, Wo, z, stats);
  FROST_LAUNCH_CHECK("stem_conv_fwd_f32");
  return FROST_OK;
}

extern "C" int frost_stem_wgrad_f32(const float* dz, const float* x_nchw, int N, int H, int W, int cin, int cout, int k,
                                    int stride, int pad, float* dwq, void* stream) {
  FROST_REQUIRE(dz && x_nchw && dwq, "frost_stem_wgrad_f32: null pointer");
  FROST_REQUIRE(cout > 0 && cout <= 32 && cin > 0 && k > 0 && k * k * cin <= 32, "frost_stem_wgrad_f32: cout<=32, k*k*cin<=32");
  cudaStream_t st = (cudaStream_t)stream;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int KK = k * k * cin;
  if (cudaMemsetAsync(dwq, 0, sizeof(float) * (size_t)KK * cout, st) != cudaSuccess) {
    set_error("frost_stem_wgrad_f32: memset failed");
    return FROST_ECUDA;
  }
  const int64_t total = (int64_t)N * Ho * Wo;
  const int64_t nblk = std::min<int64_t>(ceil_div(total, FSW_PIX), (int64_t)kNumSMs * 6);
  const int64_t ppb = ceil_div(ceil_div(total, nblk), FSW_PIX) * FSW_PIX;
  const size_t smem = sizeof(float) * (FSW_PIX * (KK | 1) + FSW_PIX * 32);
  stem_wgrad_f32_kernel<<<(unsigned)ceil_div(total, ppb), 256, smem, st>>>(dz, x_nchw, N, H, W, cin, cout, k, stride, pad, Ho, Wo,
                                                                          ppb, dwq);
  FROST_LAUNCH_CHECK("stem_wgrad_f32");
  return FROST_OK;
}

extern "C" int frost_dequant_to_nchw(const uint8_t* q, int ldq, const float* scale, const int32_t* zp, int N, int H, int W, int C,
                                     float* y_nchw, void* stream) {
  FROST_REQUIRE(q && scale && zp && y_nchw && N > 0 && H > 0 && W > 0 && C > 0 && ldq >= C, "frost_dequant_to_nchw: bad args");
  dim3 grid((unsigned)ceil_div(H * W, 32), (unsigned)ceil_div(C, 32), (unsigned)N);
  dequant_to_nchw_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(q, ldq, scale, zp, H * W, C, y_nchw);
  FROST_LAUNCH_CHECK("dequant_to_nchw");
  return FROST_OK;
}

extern "C" int frost_nchw_to_nhwc(const float* g_nchw, int N, int C, int H, int W, float* g_nhwc, int accumulate, void* stream) {
  FROST_REQUIRE(g_nchw && g_nhwc && N > 0 && H > 0 && W > 0 && C > 0, "frost_nchw_to_nhwc: bad args");
  dim3 grid((unsigned)ceil_div(H * W, 32), (unsigned)ceil_div(C, 32), (unsigned)N);
  nchw_to_nhwc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(g_nchw, H * W, C, g_nhwc, accumulate);
  FROST_LAUNCH_CHECK("nchw_to_nhwc");
  return FROST_OK;
}
