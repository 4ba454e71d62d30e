// quant.cu - per-tensor fake-quant (observer + qparams + quantise), QuantStub, weight prep/backward.
// Restates torch.fused_moving_avg_obs_fake_quant (torch/ao/quantization/fake_quantize.py:423-438)
// and the weight side of nniqat.ConvBn2d (torch/ao/nn/intrinsic/qat/modules/conv_fused.py:131-146).
#include <cuda_bf16.h>
#include "common.cuh"

namespace frost {

// ---------------------------------------------------------------- min/max reduction (no atomics)
// Each block writes its partial (min,max) to partial[2*blockIdx.x ..]; a 1-block finalize reduces.
constexpr int kMinMaxThreads = 256;
constexpr int kMinMaxMaxBlocks = FROST_FQ_SCRATCH_FLOATS / 2;  // 1024

__global__ void __launch_bounds__(kMinMaxThreads) minmax_partial_kernel(const float* x, int64_t n,
                                                                        float* partial) {
  float mn = INFINITY, mx = -INFINITY;
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const bool vec_ok = ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
  const int64_t n4 = vec_ok ? (n >> 2) : 0;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (int64_t i = tid; i < n4; i += stride) {
    const float4 v = ld_cg(x4 + i);
    mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
    mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
  }
  for (int64_t i = (n4 << 2) + tid; i < n; i += stride) {
    const float v = ld_cg(x + i);
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  block_minmax(mn, mx);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = mn;
    partial[2 * blockIdx.x + 1] = mx;
  }
}

__global__ void __launch_bounds__(256) fq_apply_kernel(const float* x, int64_t n, const float* scale_p,
                                                       const int32_t* zp_p, int qmin, int qmax, float* y,
                                                       uint8_t* mask, int32_t* q) {
  const float s = *scale_p, zp = (float)*zp_p;
  const float inv = __fdiv_rn(1.0f, s);
  const float lo = (float)qmin, hi = (float)qmax;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float idx = fq_index(ld_cg(x + i), inv, zp);
    const float qc = fminf(fmaxf(idx, lo), hi);
    if (y) y[i] = fq_dequant(qc, zp, s);
    if (mask) mask[i] = (idx >= lo && idx <= hi) ? 1 : 0;
    if (q) q[i] = (int32_t)qc;
  }
}

__global__ void __launch_bounds__(256) fq_backward_kernel(const float* dy, const uint8_t* mask,
                                                          int64_t n, float* dx) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    dx[i] = mask[i] ? dy[i] : 0.0f;
}

// QuantStub apply: NCHW fp32 -> NHWC uint8.  One thread per (n,h,w) pixel.
__global__ void __launch_bounds__(256) input_quant_apply_kernel(const float* x, int N, int C, int HW,
                                                                const float* scale_p, const int32_t* zp_p,
                                                                uint8_t* q) {
  const float s = *scale_p, zp = (float)*zp_p;
  const float inv = __fdiv_rn(1.0f, s);
  const int64_t total = (int64_t)N * HW;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t n = i / HW, p = i - n * HW;
    const float* src = x + (n * C) * HW + p;
    uint8_t* dst = q + i * C;
    for (int c = 0; c < C; ++c) {
      const float idx = fq_index(ld_cg(src + (int64_t)c * HW), inv, zp);
      dst[c] = (uint8_t)fminf(fmaxf(idx, 0.0f), 255.0f);
    }
  }
}

// RGB images, HW % 4 == 0: one thread per 4 consecutive pixels - three 16-byte plane loads, three 4-byte stores
// (the per-pixel kernel writes single bytes: 1.6 TB/s on the 154 MB image batch).
__global__ void __launch_bounds__(256) input_quant_apply_rgb4_kernel(const float* x, int N, int HW, const float* scale_p,
                                                                     const int32_t* zp_p, uint8_t* q) {
  const float s = *scale_p, zp = (float)*zp_p;
  const float inv = __fdiv_rn(1.0f, s);
  const int HW4 = HW >> 2;
  const int64_t total = (int64_t)N * HW4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t n = i / HW4, p4 = i - n * HW4;
    const float4* src = reinterpret_cast<const float4*>(x + (n * 3) * HW) + p4;
    const float4 r = ld_cg(src), g = ld_cg(src + HW4), b = ld_cg(src + 2 * HW4);
    const float v[12] = {r.x, g.x, b.x, r.y, g.y, b.y, r.z, g.z, b.z, r.w, g.w, b.w};
    unsigned w[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      unsigned pk = 0u;
#pragma unroll
      for (int e = 0; e < 4; ++e)
        pk |= (unsigned)fminf(fmaxf(fq_index(v[4 * k + e], inv, zp), 0.0f), 255.0f) << (8 * e);
      w[k] = pk;
    }
    unsigned* dst = reinterpret_cast<unsigned*>(q + (n * HW + p4 * 4) * 3);
    dst[0] = w[0]; dst[1] = w[1]; dst[2] = w[2];
  }
}

// QuantStub fused with the stem's im2col: one row of k*k*C quantize indices per OUTPUT pixel of the dense kxk stem conv, in
// (kh, kw, c) order, padded to `ld` bytes - the activation operand of the fused 1x1 tensor-core kernels (the stem becomes a
// K = ld GEMM).  Taps that fall outside the image hold the zero point (the reference pads the DEQUANTISED tensor with 0.0).
__global__ void __launch_bounds__(256) input_quant_im2col_kernel(const float* x, int N, int C, int H, int W, int k, int stride,
                                                                 int pad, int Ho, int Wo, const float* scale_p, const int32_t* zp_p,
                                                                 uint8_t* q, int ld) {
  const float s = *scale_p, zp = (float)*zp_p;
  const float inv = __fdiv_rn(1.0f, s);
  const unsigned zpb = (unsigned)*zp_p & 0xffu;
  const int64_t total = (int64_t)N * Ho * Wo;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const int HW = H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += step) {
    const int ow = (int)(i % Wo);
    const int64_t t1 = i / Wo;
    const int oh = (int)(t1 % Ho);
    const int64_t n = t1 / Ho;
    const int ih0 = oh * stride - pad, iw0 = ow * stride - pad;
    unsigned words[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};       // ld <= 32
    int b = 0;
    for (int r = 0; r < k; ++r) {
      const int ih = ih0 + r;
      for (int sx = 0; sx < k; ++sx) {
        const int iw = iw0 + sx;
        const bool ok = (unsigned)ih < (unsigned)H && (unsigned)iw < (unsigned)W;
        for (int c = 0; c < C; ++c, ++b) {
          unsigned byte = zpb;
          if (ok) {
            const float v = ld_cg(x + (n * C + c) * HW + (int64_t)ih * W + iw);
            byte = (unsigned)fminf(fmaxf(fq_index(v, inv, zp), 0.0f), 255.0f);
          }
          words[b >> 2] |= byte << (8 * (b & 3));
        }
      }
    }
    uint4* dst = reinterpret_cast<uint4*>(q + i * ld);
    dst[0] = make_uint4(words[0], words[1], words[2], words[3]);
    if (ld > 16) dst[1] = make_uint4(words[4], words[5], words[6], words[7]);
  }
}

// ---------------------------------------------------------------- weights
__device__ __forceinline__ int64_t wq_index(const FrostWeightDesc& d, int c, int ci, int y, int x) {
  switch (d.layout) {
    case 1: return (int64_t)(y * d.kw + x) * d.cout + c;
    case 2: return (((int64_t)c * d.kh + y) * d.kw + x) * d.cin_g + ci;
    default: return (((int64_t)c * d.cin_g + ci) * d.kh + y) * d.kw + x;
  }
}

constexpr int kWeightThreads = 256;

__device__ __forceinline__ float weight_sf(const FrostWeightDesc& d, int c, float* rstd_out) {
  // scale_factor = gamma / sqrt(running_var + eps)      (conv_fused.py:138-139)
  float sf = 1.0f, rstd = 1.0f;
  if (d.bn_weight) {
    const float std_run = __fsqrt_rn(__fadd_rn(d.bn_var[c], d.bn_eps));
    sf = __fdiv_rn(d.bn_weight[c], std_run);
    rstd = __fdiv_rn(1.0f, std_run);
  }
  if (rstd_out) *rstd_out = rstd;
  return sf;
}

// pass 1 (one CTA per FROST_WEIGHT_CHUNK elements): min/max of Ws = W * scale_factor
__global__ void __launch_bounds__(kWeightThreads) weight_minmax_kernel(const FrostWeightDesc* descs,
                                                                      const FrostOptChunk* chunks,
                                                                      float* scratch) {
  const FrostOptChunk ck = chunks[blockIdx.x];
  const FrostWeightDesc d = descs[ck.tensor];
  if (!d.observe) return;
  const int K = d.cin_g * d.kh * d.kw;
  const int64_t total = (int64_t)d.cout * K;
  const int64_t e0 = (int64_t)ck.chunk * FROST_WEIGHT_CHUNK;
  const int64_t e1 = min(total, e0 + FROST_WEIGHT_CHUNK);
  float mn = INFINITY, mx = -INFINITY;
  for (int64_t e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
    const int c = (int)(e / K);
    const float ws = __fmul_rn(d.weight[e], weight_sf(d, c, nullptr));
    mn = fminf(mn, ws);
    mx = fmaxf(mx, ws);
  }
  block_minmax(mn, mx);
  if (threadIdx.x == 0) {
    atomic_min_float(scratch + 2 * ck.tensor, mn);
    atomic_max_float(scratch + 2 * ck.tensor + 1, mx);
  }
}

// pass 2 (one CTA per layer): scale_factor arrays, observer EMA + qparams, reset of the scratch slot
__global__ void __launch_bounds__(kWeightThreads) weight_finalize_kernel(const FrostWeightDesc* descs,
                                                                        float* scratch) {
  const FrostWeightDesc d = descs[blockIdx.x];
  for (int c = threadIdx.x; c < d.cout; c += blockDim.x) {
    float rstd;
    d.sf[c] = weight_sf(d, c, &rstd);
    d.rstd_run[c] = rstd;
    d.wsum[c] = 0;
  }
  if (threadIdx.x == 0) {
    if (d.observe) observer_update(d.wfq, scratch[2 * blockIdx.x], scratch[2 * blockIdx.x + 1], -128, 127, true, d.averaging_const);
    scratch[2 * blockIdx.x] = INFINITY;
    scratch[2 * blockIdx.x + 1] = -INFINITY;
  }
}

// pass 3 (one CTA per chunk): int8 indices in the kernel layout, STE mask, per-cout index sums
__global__ void __launch_bounds__(kWeightThreads) weight_quant_kernel(const FrostWeightDesc* descs,
                                                                     const FrostOptChunk* chunks) {
  const FrostOptChunk ck = chunks[blockIdx.x];
  const FrostWeightDesc d = descs[ck.tensor];
  const int K = d.cin_g * d.kh * d.kw;
  const int64_t total = (int64_t)d.cout * K;
  const int64_t e0 = (int64_t)ck.chunk * FROST_WEIGHT_CHUNK;
  const int64_t e1 = min(total, e0 + FROST_WEIGHT_CHUNK);
  const float s = *d.wfq.scale, zp = (float)*d.wfq.zero_point;
  const float inv = __fdiv_rn(1.0f, s);
  const int khw = d.kh * d.kw;
  const int lane = threadIdx.x & 31;
  // warp-uniform trip count; 32-bit index arithmetic (a weight tensor has < 2^31 elements: checked on the host side of
  // the descriptor table by construction, cout*K <= 1728*1728).  The per-cout index sum is reduced in the warp first:
  // consecutive elements share their output channel, one atomic per warp instead of one per element.
  for (int64_t base = e0 + (threadIdx.x - lane); base < e1; base += blockDim.x) {
    const int64_t e = base + lane;
    const bool live = e < e1;
    int c = -1, qi = 0;
    if (live) {
      const int e32 = (int)e;
      c = e32 / K;
      const int r = e32 - c * K;
      const int ci = r / khw, yx = r - ci * khw;
      const int y = yx / d.kw, x = yx - y * d.kw;
      const float ws = __fmul_rn(d.weight[e], d.sf[c]);
      const float idx = fq_index(ws, inv, zp);
      const float qc = fminf(fmaxf(idx, -128.0f), 127.0f);
      d.wq[wq_index(d, c, ci, y, x)] = (int8_t)qc;
      if (d.wq_mma) {   // the tensor-core operand bytes, zero point folded for one-signed weights.  1x1: column r == ci;
                        // dense kxk (the stem as an im2col GEMM): column (y, x, ci), the order of frost_input_quant_im2col
        const int zpi = (int)zp;
        const unsigned flip = zpi == 0 ? 0u : (zpi == -128 ? 0x80u : 0x7fu);
        const int col = d.layout == 2 ? (y * d.kw + x) * d.cin_g + ci : r;
        d.wq_mma[(int64_t)c * d.ldw + col] = (int8_t)(((unsigned)(int)qc ^ flip) & 0xffu);
        // the GEMM runs over ldw columns: the padding columns count as weights equal to the zero point (value 0)
        if (d.layout == 2 && r == 0 && zpi != 0) atomicAdd(d.wsum + c, (d.ldw - K) * zpi);
      }
      if (d.wt_bf16) d.wt_bf16[(int64_t)r * d.cout + c] = __bfloat16_as_ushort(__float2bfloat16_rn(qc - zp));  // 1x1: r == ci; exact
      d.wmask[e] = (idx >= -128.0f && idx <= 127.0f) ? 1 : 0;
      qi = (int)qc;
    }
    const int c0 = __shfl_sync(0xffffffffu, c, 0);
    if (__all_sync(0xffffffffu, c == c0 || !live)) {
      const int tot = __reduce_add_sync(0xffffffffu, qi);
      if (lane == 0 && c0 >= 0) atomicAdd(d.wsum + c0, tot);
    } else if (live) {
      atomicAdd(d.wsum + c, qi);
    }
  }
}

// dW = dWq*mask*sf ; dgamma = dgamma_bn + (dsf_bn + sum_k dWq*mask*W) * rstd_run   (SURVEY 8a' 5-6)
// one CTA (8 warps) per 8 output channels of one layer: chunks[i] = {layer, first channel / 8}
__global__ void __launch_bounds__(kWeightThreads) weight_backward_kernel(const FrostWeightDesc* descs,
                                                                        const FrostOptChunk* chunks) {
  const FrostOptChunk ck = chunks[blockIdx.x];
  const FrostWeightDesc d = descs[ck.tensor];
  const int K = d.cin_g * d.kh * d.kw;
  const int khw = d.kh * d.kw;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = ck.chunk * (kWeightThreads / 32) + warp;
  if (c >= d.cout) return;
  const float sf = d.sf[c];
  float acc = 0.0f;
  // 4 independent (mask, dWq, W) load triples in flight per lane: a dependent mask -> gradient load chain per
  // element made this kernel latency-bound (54 serial round trips on the K = 1728 rows)
  for (int r0 = lane; r0 < K; r0 += 128) {
    uint8_t mk[4];
    float dw[4], w[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + 32 * u;
      mk[u] = 0; dw[u] = 0.0f; w[u] = 0.0f;
      if (r < K) {
        const int64_t e = (int64_t)c * K + r;
        const int ci = r / khw, yx = r - ci * khw;
        const int y = yx / d.kw, x = yx - y * d.kw;
        mk[u] = ld_cg(d.wmask + e);
        // the im2col stem's weight gradient comes from the tensor-core wgrad as [cout][ldw]
        dw[u] = ld_cg(d.dwq + ((d.wq_mma && d.layout == 2) ? (int64_t)c * d.ldw + (y * d.kw + x) * d.cin_g + ci : wq_index(d, c, ci, y, x)));
        w[u] = ld_cg(d.weight + e);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = r0 + 32 * u;
      if (r < K) {
        const float dws = mk[u] ? dw[u] : 0.0f;
        d.dweight[(int64_t)c * K + r] = dws * sf;
        acc = fmaf(dws, w[u], acc);
      }
    }
  }
  acc = warp_sum(acc);
  if (lane == 0 && d.dgamma) d.dgamma[c] = d.dgamma_bn[c] + (d.dsf_bn[c] + acc) * d.rstd_run[c];
}

__global__ void stats_reset_kernel(FrostChanStats* s, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    s[i].sum = 0;
    s[i].sq_lo = 0;
    s[i].sq_hi = 0;
    s[i].min = INT_MAX;
    s[i].max = INT_MIN;
  }
}

static int run_observer(const float* x, int64_t n, FrostFQ fq, int qmin, int qmax, int symmetric, float c,
                        float* cur_minmax, float* scratch, cudaStream_t st, bool observe) {
  int nblk = 1;
  if (observe || cur_minmax) {
    nblk = grid_for(n, kMinMaxThreads * 8, kMinMaxMaxBlocks);
    minmax_partial_kernel<<<nblk, kMinMaxThreads, 0, st>>>(x, n, scratch);
    FROST_LAUNCH_CHECK("minmax_partial");
  }
  if (observe || cur_minmax) {
    fq_finalize_kernel<<<1, 1024, 0, st>>>(scratch, nblk, fq, qmin, qmax, symmetric, c, observe ? 1 : 0, cur_minmax);
    FROST_LAUNCH_CHECK("fq_finalize");
  }
  return FROST_OK;
}

}  // namespace frost

using namespace frost;

extern "C" int frost_stats_reset(FrostChanStats* stats, int64_t n, void* stream) {
  FROST_REQUIRE(stats && n >= 0, "frost_stats_reset: bad args");
  if (n == 0) return FROST_OK;
  stats_reset_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(stats, n);
  FROST_LAUNCH_CHECK("stats_reset");
  return FROST_OK;
}

extern "C" int frost_fq_forward(const float* x, int64_t n, FrostFQ fq, int qmin, int qmax, int symmetric,
                                int observe, float averaging_const, float* y, uint8_t* mask, int32_t* q,
                                float* scratch, void* stream) {
  FROST_REQUIRE(x && n > 0 && fq.scale && fq.zero_point && fq.min_val && fq.max_val && scratch,
                "frost_fq_forward: null pointer or empty tensor (n=%lld)", (long long)n);
  FROST_REQUIRE(qmin < qmax, "frost_fq_forward: qmin >= qmax");
  cudaStream_t st = (cudaStream_t)stream;
  if (observe) {
    int rc = run_observer(x, n, fq, qmin, qmax, symmetric, averaging_const, nullptr, scratch, st, true);
    if (rc) return rc;
  }
  if (y || mask || q) {
    fq_apply_kernel<<<grid_for(n, 256 * 4), 256, 0, st>>>(x, n, fq.scale, fq.zero_point, qmin, qmax, y, mask, q);
    FROST_LAUNCH_CHECK("fq_apply");
  }
  return FROST_OK;
}

extern "C" int frost_fq_backward(const float* dy, const uint8_t* mask, int64_t n, float* dx, void* stream) {
  FROST_REQUIRE(dy && mask && dx && n > 0, "frost_fq_backward: bad args");
  fq_backward_kernel<<<grid_for(n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(dy, mask, n, dx);
  FROST_LAUNCH_CHECK("fq_backward");
  return FROST_OK;
}

extern "C" int frost_input_quant(const float* x_nchw, int N, int C, int H, int W, FrostFQ fq, int observe,
                                 float averaging_const, uint8_t* q_nhwc, float* cur_minmax, float* scratch,
                                 void* stream) {
  FROST_REQUIRE(x_nchw && q_nhwc && scratch && N > 0 && C > 0 && H > 0 && W > 0, "frost_input_quant: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)N * C * H * W;
  if (observe || cur_minmax) {
    int rc = run_observer(x_nchw, n, fq, 0, 255, 0, averaging_const, cur_minmax, scratch, st, observe != 0);
    if (rc) return rc;
  }
  if (C == 3 && (H * W) % 4 == 0 && (reinterpret_cast<uintptr_t>(x_nchw) & 15) == 0 && (reinterpret_cast<uintptr_t>(q_nhwc) & 3) == 0)
    input_quant_apply_rgb4_kernel<<<grid_for((int64_t)N * H * W / 4, 256, kNumSMs * 8), 256, 0, st>>>(x_nchw, N, H * W, fq.scale,
                                                                                                   fq.zero_point, q_nhwc);
  else
    input_quant_apply_kernel<<<grid_for((int64_t)N * H * W, 256), 256, 0, st>>>(x_nchw, N, C, H * W, fq.scale,
                                                                                fq.zero_point, q_nhwc);
  FROST_LAUNCH_CHECK("input_quant_apply");
  return FROST_OK;
}

extern "C" int frost_input_quant_im2col(const float* x_nchw, int N, int C, int H, int W, int k, int stride, int pad, FrostFQ fq,
                                        int observe, float averaging_const, uint8_t* q, int ld, float* cur_minmax, float* scratch,
                                        void* stream) {
  FROST_REQUIRE(x_nchw && q && scratch && N > 0 && C > 0 && H > 0 && W > 0 && k > 0 && stride > 0 && pad >= 0,
                "frost_input_quant_im2col: bad args");
  FROST_REQUIRE(k * k * C <= ld && (ld == 16 || ld == 32) && (reinterpret_cast<uintptr_t>(q) & 15) == 0,
                "frost_input_quant_im2col: k*k*C=%d must fit the row pitch ld=%d (16 or 32), q 16-byte aligned", k * k * C, ld);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n = (int64_t)N * C * H * W;
  if (observe || cur_minmax) {
    int rc = run_observer(x_nchw, n, fq, 0, 255, 0, averaging_const, cur_minmax, scratch, st, observe != 0);
    if (rc) return rc;
  }
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  input_quant_im2col_kernel<<<grid_for((int64_t)N * Ho * Wo, 256, kNumSMs * 8), 256, 0, st>>>(x_nchw, N, C, H, W, k, stride, pad, Ho, Wo,
                                                                                           fq.scale, fq.zero_point, q, ld);
  FROST_LAUNCH_CHECK("input_quant_im2col");
  return FROST_OK;
}

extern "C" int frost_weight_prep_multi(const FrostWeightDesc* descs, int n, const FrostOptChunk* chunks, int n_chunks,
                                       float* scratch, void* stream) {
  FROST_REQUIRE(descs && n > 0 && chunks && n_chunks > 0 && scratch, "frost_weight_prep_multi: bad args");
  cudaStream_t st = (cudaStream_t)stream;
  weight_minmax_kernel<<<n_chunks, kWeightThreads, 0, st>>>(descs, chunks, scratch);
  FROST_LAUNCH_CHECK("weight_minmax");
  weight_finalize_kernel<<<n, kWeightThreads, 0, st>>>(descs, scratch);
  FROST_LAUNCH_CHECK("weight_finalize");
  weight_quant_kernel<<<n_chunks, kWeightThreads, 0, st>>>(descs, chunks);
  FROST_LAUNCH_CHECK("weight_quant");
  return FROST_OK;
}

extern "C" int frost_weight_backward_multi(const FrostWeightDesc* descs, int n, const FrostOptChunk* chunks, int n_chunks,
                                           void* stream) {
  FROST_REQUIRE(descs && n > 0 && chunks && n_chunks > 0, "frost_weight_backward_multi: bad args");
  weight_backward_kernel<<<n_chunks, kWeightThreads, 0, (cudaStream_t)stream>>>(descs, chunks);
  FROST_LAUNCH_CHECK("weight_backward");
  return FROST_OK;
}
