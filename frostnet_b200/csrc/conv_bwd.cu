// conv_bwd.cu - dgrad / wgrad of the quantised convolutions and the classifier GEMMs
// (aten::convolution_backward / addmm in the reference's autograd graph; SURVEY.md 8a' steps 4-5).
// Operands: dz fp32 (gradient wrt the real-valued conv output), weights as int8 indices,
// activations as uint8 indices; the dequantisation scale is applied once in the epilogue.
// v1: shared-memory tiled SIMT fp32 kernels.
#include "common.cuh"

namespace frost {


constexpr int GB_T = 64;    // output tile edge
constexpr int GB_RC = 16;   // reduction chunk
constexpr int GB_LD = GB_T + 4;

// ---------------------------------------------------------------- pointwise dgrad
// dx[m][k] (+)= s_w * sum_co dz[m][co] * (wq[co][k] - zp_w)
__global__ void __launch_bounds__(256) pw_dgrad_kernel(const float* dz, const int8_t* wq,
                                                      const float* w_scale_p, const int32_t* w_zp_p,
                                                      int64_t M, int K, int cout, float* dx, int accumulate) {
  __shared__ __align__(16) float a_s[GB_RC][GB_LD];  // dz^T : [co][m]
  __shared__ __align__(16) float b_s[GB_RC][GB_LD];  // w    : [co][k]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * GB_T;
  const int k0 = blockIdx.y * GB_T;
  const float zp_w = (float)*w_zp_p;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int c0 = 0; c0 < cout; c0 += GB_RC) {
    {  // dz tile: 64 rows x 16 co, one float4 per thread
      const int r = tid >> 2, c4 = tid & 3;
      const int64_t m = m0 + r;
      const int co = c0 + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < M && co < cout) v = ld_cg(reinterpret_cast<const float4*>(dz + m * cout + co));  // cout % 4 == 0
      a_s[c4 * 4 + 0][r] = v.x;
      a_s[c4 * 4 + 1][r] = v.y;
      a_s[c4 * 4 + 2][r] = v.z;
      a_s[c4 * 4 + 3][r] = v.w;
    }
    {  // weight tile: 16 co x 64 k int8, 4 bytes per thread
      const int r = tid >> 4, k4 = tid & 15;
      const int co = c0 + r, k = k0 + k4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (co < cout && k < K) {
        const unsigned pk = ld_cg(reinterpret_cast<const unsigned*>(wq + (int64_t)co * K + k));
        v.x = (float)(int8_t)(pk & 0xff) - zp_w;
        v.y = (float)(int8_t)((pk >> 8) & 0xff) - zp_w;
        v.z = (float)(int8_t)((pk >> 16) & 0xff) - zp_w;
        v.w = (float)(int8_t)((pk >> 24) & 0xff) - zp_w;
      }
      *reinterpret_cast<float4*>(&b_s[r][k4 * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GB_RC; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&a_s[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&b_s[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float s_w = *w_scale_p;
  const int k = k0 + tx * 4;
  if (k < K) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int64_t m = m0 + ty * 4 + i;
      if (m < M) {
        float4* p = reinterpret_cast<float4*>(dx + m * K + k);
        float4 r = make_float4(acc[i][0] * s_w, acc[i][1] * s_w, acc[i][2] * s_w, acc[i][3] * s_w);
        if (accumulate) {
          const float4 o = *p;
          r.x += o.x; r.y += o.y; r.z += o.z; r.w += o.w;
        }
        *p = r;
      }
    }
  }
}

// ---------------------------------------------------------------- pointwise wgrad (split over M)
// dwq[co][k] += s_a * sum_m dz[m][co] * (x[m][k] - zp_a);  X = uint8 indices or fp32 (classifier)
template <typename XT>
__global__ void __launch_bounds__(256) pw_wgrad_kernel(const float* dz, const XT* x,
                                                      const float* x_scale_p, const int32_t* x_zp_p,
                                                      int64_t M, int K, int cout, int64_t rows_per_split,
                                                      float* dwq) {
  __shared__ __align__(16) float a_s[GB_RC][GB_LD];  // dz : [m][co]
  __shared__ __align__(16) float b_s[GB_RC][GB_LD];  // x  : [m][k]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int co0 = blockIdx.x * GB_T, k0 = blockIdx.y * GB_T;
  const int64_t m_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t m_end = min(M, m_begin + rows_per_split);
  const float zp_a = x_zp_p ? (float)*x_zp_p : 0.0f;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int64_t mc = m_begin; mc < m_end; mc += GB_RC) {
    const int r = tid >> 4, c4 = tid & 15;
    const int64_t m = mc + r;
    {
      const int co = co0 + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < m_end && co < cout) v = ld_cg(reinterpret_cast<const float4*>(dz + m * cout + co));
      *reinterpret_cast<float4*>(&a_s[r][c4 * 4]) = v;
    }
    {
      const int k = k0 + c4 * 4;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < m_end && k < K) {
        if constexpr (sizeof(XT) == 1) {
          const unsigned pk = ld_cg(reinterpret_cast<const unsigned*>(x + m * K + k));
          v.x = (float)(pk & 0xff) - zp_a;
          v.y = (float)((pk >> 8) & 0xff) - zp_a;
          v.z = (float)((pk >> 16) & 0xff) - zp_a;
          v.w = (float)((pk >> 24) & 0xff) - zp_a;
        } else {
          v = ld_cg(reinterpret_cast<const float4*>(x + m * K + k));
        }
      }
      *reinterpret_cast<float4*>(&b_s[r][c4 * 4]) = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GB_RC; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&a_s[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&b_s[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  const float s_a = x_scale_p ? *x_scale_p : 1.0f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= cout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K) atomicAdd(dwq + (int64_t)co * K + k, acc[i][j] * s_a);
    }
  }
}

// ---------------------------------------------------------------- stem wgrad
// dwq[co][r][s][ci] += s_a * sum_pix dz[pix][co] * (x[patch(pix)][r][s][ci] - zp_a)
// Batches of 64 pixels: the (pixel, tap) patch values and the dz rows are staged in shared memory (pitch 32,
// taps >= k*k*cin stay zero).  Compute role: thread = (4 couts, 4 taps, pixel subset of 4): two 16-byte
// shared loads feed 16 FMAs (the previous one-cout thread needed 5 loads per 4 FMAs and was LDS-bound).
// The four pixel subsets are summed through shared memory, then one scalar atomic per (co, tap) per CTA;
// the grid is one resident wave (frost::tunable).
constexpr int STEMW_PIX = 64;
__global__ void __launch_bounds__(256) stem_wgrad_kernel(const float* dz, const uint8_t* xq,
                                                        const float* x_scale_p, const int32_t* x_zp_p,
                                                        int N, int H, int W, int cin, int cout, int k, int stride, int pad,
                                                        int Ho, int Wo, int64_t pix_per_block, float* dwq) {
  constexpr int PP = 36;                       // patch pitch: 16-byte aligned rows, fill stores spread over banks
  __shared__ __align__(16) float s_buf[STEMW_PIX * PP + STEMW_PIX * 32];
  pdl_enter();
  const int KK = k * k * cin;
  float* s_patch = s_buf;                      // [STEMW_PIX][PP]
  float* s_dz = s_buf + STEMW_PIX * PP;        // [STEMW_PIX][32]
  const int co4 = threadIdx.x & 7, tg = (threadIdx.x >> 3) & 7, psub = threadIdx.x >> 6;
  const float zp_a = (float)*x_zp_p;
  const int64_t total = (int64_t)N * Ho * Wo;
  const int64_t p_begin = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p_end = min(total, p_begin + pix_per_block);
  // fill role: 4 threads per pixel, each takes kernel positions (r, s) = ff, ff+4, ... with all input channels
  const int fp = threadIdx.x >> 2, ff = threadIdx.x & 3;
  for (int i = threadIdx.x; i < STEMW_PIX * PP; i += blockDim.x) s_patch[i] = 0.0f;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  __syncthreads();
  for (int64_t pb = p_begin; pb < p_end; pb += STEMW_PIX) {
    const int np = (int)min((int64_t)STEMW_PIX, p_end - pb);
    {
      const bool pv = fp < np;
      int ow = 0, oh = 0, n = 0;
      if (pv) {
        const unsigned p = (unsigned)(pb + fp);          // total < 2^31 (checked on the host)
        ow = (int)(p % (unsigned)Wo);
        const unsigned t1 = p / (unsigned)Wo;
        oh = (int)(t1 % (unsigned)Ho);
        n = (int)(t1 / (unsigned)Ho);
      }
      const int ih0 = oh * stride - pad, iw0 = ow * stride - pad;
      const uint8_t* img = xq + (int64_t)n * H * W * cin;
      for (int rs = ff; rs < k * k; rs += 4) {
        const int r = rs / k, sx = rs - r * k;
        const int ih = ih0 + r, iw = iw0 + sx;
        const bool ok = pv && (unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W;
        const uint8_t* px = img + ((int64_t)ih * W + iw) * cin;
        for (int ci = 0; ci < cin; ++ci) s_patch[fp * PP + rs * cin + ci] = ok ? (float)ld_cg(px + ci) - zp_a : 0.0f;
      }
    }
    for (int i = threadIdx.x; i < STEMW_PIX * 32; i += blockDim.x) {
      const int pl = i >> 5, c = i & 31;
      s_dz[i] = (pl < np && c < cout) ? ld_cg(dz + (pb + pl) * cout + c) : 0.0f;
    }
    __syncthreads();
#pragma unroll 4
    for (int pl = psub; pl < STEMW_PIX; pl += 4) {
      const float4 d = *reinterpret_cast<const float4*>(s_dz + pl * 32 + 4 * co4);
      const float4 x = *reinterpret_cast<const float4*>(s_patch + pl * PP + 4 * tg);
      const float dv[4] = {d.x, d.y, d.z, d.w}, xv[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xv[i], dv[j], acc[i][j]);
    }
    __syncthreads();
  }
  // combine the 4 pixel subsets: s_buf reused as [psub][tap 32][co 32]
  float* s_red = s_buf;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(s_red + (psub * 32 + 4 * tg + i) * 32 + 4 * co4) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  __syncthreads();
  const float s_a = *x_scale_p;
  for (int i = threadIdx.x; i < 32 * 32; i += blockDim.x) {
    const int t = i >> 5, co = i & 31;
    if (t < KK && co < cout) {
      const float v = s_red[i] + s_red[1024 + i] + s_red[2048 + i] + s_red[3072 + i];
      atomicAdd(dwq + (int64_t)co * KK + t, v * s_a);
    }
  }
}

// ---------------------------------------------------------------- classifier forward
// out[n][co] = sum_k x[n][k] * (wq[co][k]-zp_w)*s_w + bias[co]
__global__ void __launch_bounds__(256) linear_fwd_kernel(const float* x, const int8_t* wq,
                                                        const float* w_scale_p, const int32_t* w_zp_p,
                                                        const float* bias, int N, int K, int cout,
                                                        float* out) {
  __shared__ __align__(16) float a_s[GB_RC][GB_LD];  // x^T : [k][n]
  __shared__ __align__(16) float b_s[GB_RC][GB_LD];  // w^T : [k][co]
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int n0 = blockIdx.x * GB_T, co0 = blockIdx.y * GB_T;
  const float zp_w = (float)*w_zp_p, s_w = *w_scale_p;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  for (int k0 = 0; k0 < K; k0 += GB_RC) {
    const int r = tid >> 2, k4 = tid & 3;
    const int k = k0 + k4 * 4;
    {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n0 + r < N && k < K) v = ld_cg(reinterpret_cast<const float4*>(x + (int64_t)(n0 + r) * K + k));
      a_s[k4 * 4 + 0][r] = v.x; a_s[k4 * 4 + 1][r] = v.y; a_s[k4 * 4 + 2][r] = v.z; a_s[k4 * 4 + 3][r] = v.w;
    }
    {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (co0 + r < cout && k < K) {
        const unsigned pk = ld_cg(reinterpret_cast<const unsigned*>(wq + (int64_t)(co0 + r) * K + k));
        v.x = ((float)(int8_t)(pk & 0xff) - zp_w) * s_w;
        v.y = ((float)(int8_t)((pk >> 8) & 0xff) - zp_w) * s_w;
        v.z = ((float)(int8_t)((pk >> 16) & 0xff) - zp_w) * s_w;
        v.w = ((float)(int8_t)((pk >> 24) & 0xff) - zp_w) * s_w;
      }
      b_s[k4 * 4 + 0][r] = v.x; b_s[k4 * 4 + 1][r] = v.y; b_s[k4 * 4 + 2][r] = v.z; b_s[k4 * 4 + 3][r] = v.w;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GB_RC; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&a_s[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&b_s[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = n0 + ty * 4 + i;
    if (n >= N) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int co = co0 + tx * 4 + j;
      if (co < cout) out[(int64_t)n * cout + co] = acc[i][j] + (bias ? bias[co] : 0.0f);
    }
  }
}

__global__ void __launch_bounds__(256) colsum_kernel(const float* d, int N, int C, float* out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float s = 0.0f;
  for (int n = 0; n < N; ++n) s += d[(int64_t)n * C + c];
  out[c] = s;
}

static int launch_pw_wgrad_u8(const float* dz, const uint8_t* x, const float* xs, const int32_t* xzp, int64_t M, int K,
                              int cout, float* dwq, cudaStream_t st) {
  const int ct = (int)ceil_div(cout, GB_T), kt = (int)ceil_div(K, GB_T);
  int64_t splits = std::max<int64_t>(1, std::min<int64_t>(ceil_div(M, GB_RC * 8), ceil_div((int64_t)kNumSMs * 4, (int64_t)ct * kt)));
  int64_t rows = ceil_div(ceil_div(M, splits), GB_RC) * GB_RC;
  splits = ceil_div(M, rows);
  pw_wgrad_kernel<uint8_t><<<dim3(ct, kt, (unsigned)splits), 256, 0, st>>>(dz, x, xs, xzp, M, K, cout, rows, dwq);
  return 0;
}

}  // namespace frost

using namespace frost;

extern "C" int frost_pw_dgrad(const float* dz, const int8_t* wq, const float* w_scale, const int32_t* w_zp, int64_t M,
                              int K, int cout, float* dx, int accumulate, void* stream) {
  FROST_REQUIRE(dz && wq && w_scale && w_zp && dx, "frost_pw_dgrad: null pointer");
  FROST_REQUIRE(M > 0 && K > 0 && cout > 0 && K % 4 == 0 && cout % 4 == 0, "frost_pw_dgrad: K and cout must be multiples of 4");
  pw_dgrad_kernel<<<dim3((unsigned)ceil_div(M, GB_T), (unsigned)ceil_div(K, GB_T)), 256, 0, (cudaStream_t)stream>>>(
      dz, wq, w_scale, w_zp, M, K, cout, dx, accumulate);
  FROST_LAUNCH_CHECK("pw_dgrad");
  return FROST_OK;
}

extern "C" int frost_pw_wgrad(const float* dz, const uint8_t* xq, const float* x_scale, const int32_t* x_zp, int64_t M,
                              int K, int cout, float* dwq, void* stream) {
  FROST_REQUIRE(dz && xq && x_scale && x_zp && dwq, "frost_pw_wgrad: null pointer");
  FROST_REQUIRE(M > 0 && K > 0 && cout > 0 && K % 4 == 0 && cout % 4 == 0, "frost_pw_wgrad: K and cout must be multiples of 4");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(dwq, 0, sizeof(float) * (size_t)K * cout, st) != cudaSuccess) {
    set_error("frost_pw_wgrad: memset failed");
    return FROST_ECUDA;
  }
  launch_pw_wgrad_u8(dz, xq, x_scale, x_zp, M, K, cout, dwq, st);
  FROST_LAUNCH_CHECK("pw_wgrad");
  return FROST_OK;
}

extern "C" int frost_stem_wgrad(const float* dz, const uint8_t* xq, const float* x_scale, const int32_t* x_zp, int N,
                                int H, int W, int cin, int cout, int k, int stride, int pad, float* dwq, void* stream) {
  FROST_REQUIRE(dz && xq && x_scale && x_zp && dwq, "frost_stem_wgrad: null pointer");
  FROST_REQUIRE(cout > 0 && cout <= 32 && cin > 0 && k > 0 && k * k * cin <= 32, "frost_stem_wgrad: cout<=32, k*k*cin<=32");
  cudaStream_t st = (cudaStream_t)stream;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int KK = k * k * cin;
  if (cudaMemsetAsync(dwq, 0, sizeof(float) * (size_t)KK * cout, st) != cudaSuccess) {
    set_error("frost_stem_wgrad: memset failed");
    return FROST_ECUDA;
  }
  const int64_t total = (int64_t)N * Ho * Wo;
  const int64_t nblk = std::min<int64_t>(ceil_div(total, STEMW_PIX), (int64_t)kNumSMs * tunable(FROST_TUNE_STEM_WGRAD_CTAS_PER_SM));
  const int64_t ppb = ceil_div(ceil_div(total, nblk), STEMW_PIX) * STEMW_PIX;
  FROST_REQUIRE(total < (int64_t)1 << 31, "frost_stem_wgrad: more than 2^31 output pixels");
  launch_pdl(stem_wgrad_kernel, dim3((unsigned)ceil_div(total, ppb)), dim3(256), 0, st, dz, xq, x_scale, x_zp, N, H, W, cin, cout, k,
             stride, pad, Ho, Wo, ppb, dwq);
  FROST_LAUNCH_CHECK("stem_wgrad");
  return FROST_OK;
}

extern "C" int frost_linear_forward(const float* x, const int8_t* wq, const float* w_scale, const int32_t* w_zp,
                                    const float* bias, int N, int K, int cout, float* out, void* stream) {
  FROST_REQUIRE(x && wq && w_scale && w_zp && out, "frost_linear_forward: null pointer");
  FROST_REQUIRE(N > 0 && K > 0 && cout > 0 && K % 4 == 0, "frost_linear_forward: K must be a multiple of 4");
  linear_fwd_kernel<<<dim3((unsigned)ceil_div(N, GB_T), (unsigned)ceil_div(cout, GB_T)), 256, 0, (cudaStream_t)stream>>>(
      x, wq, w_scale, w_zp, bias, N, K, cout, out);
  FROST_LAUNCH_CHECK("linear_fwd");
  return FROST_OK;
}

extern "C" int frost_linear_backward(const float* dout, const float* x, const int8_t* wq, const float* w_scale,
                                     const int32_t* w_zp, int N, int K, int cout, float* dx, float* dwq, float* dbias,
                                     void* stream) {
  FROST_REQUIRE(dout && x && wq && w_scale && w_zp && dx && dwq, "frost_linear_backward: null pointer");
  FROST_REQUIRE(N > 0 && K % 4 == 0 && cout % 4 == 0, "frost_linear_backward: K and cout must be multiples of 4");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = frost_pw_dgrad(dout, wq, w_scale, w_zp, N, K, cout, dx, 0, stream);
  if (rc) return rc;
  if (cudaMemsetAsync(dwq, 0, sizeof(float) * (size_t)K * cout, st) != cudaSuccess) {
    set_error("frost_linear_backward: memset failed");
    return FROST_ECUDA;
  }
  const int ct = (int)ceil_div(cout, GB_T), kt = (int)ceil_div(K, GB_T);
  int64_t splits = std::max<int64_t>(1, std::min<int64_t>(ceil_div(N, GB_RC), ceil_div((int64_t)kNumSMs * 2, (int64_t)ct * kt)));
  int64_t rows = ceil_div(ceil_div(N, splits), GB_RC) * GB_RC;
  splits = ceil_div(N, rows);
  pw_wgrad_kernel<float><<<dim3(ct, kt, (unsigned)splits), 256, 0, st>>>(dout, x, nullptr, nullptr, N, K, cout, rows, dwq);
  FROST_LAUNCH_CHECK("linear_wgrad");
  if (dbias) {
    colsum_kernel<<<(unsigned)ceil_div(cout, 256), 256, 0, st>>>(dout, N, cout, dbias);
    FROST_LAUNCH_CHECK("colsum");
  }
  return FROST_OK;
}
