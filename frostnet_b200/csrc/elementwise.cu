// elementwise.cu - FloatFunctional.cat / .add with their own observers (frostnet.py:129,142;
// torch/ao/nn/quantized/modules/functional_modules.py:50-52,80-82), AdaptiveAvgPool+Dropout of the
// head (frostnet.py:295-299), and small helpers.  Everything streams uint8 indices; the fp32 value
// of an element is (q - zp) * scale, recomputed in registers.
#include "common.cuh"

namespace frost {

struct QSrc {
  float scale, zp;
};
__device__ __forceinline__ QSrc load_src(const FrostQTensor& t) { return QSrc{*t.scale, (float)*t.zp}; }
__device__ __forceinline__ float deq(unsigned byte, const QSrc& s) { return fq_dequant((float)byte, s.zp, s.scale); }

// ---------------------------------------------------------------- cat
// The observer of the cat output sees min/max over both sources, which are known analytically
// (each source carries the dequantised min/max of its tensor): no reduction pass.
__global__ void cat_finalize_kernel(FrostQTensor a, FrostQTensor b, FrostFQ fq, int observe, float c,
                                    float* cur_minmax_out) {
  pdl_enter();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float mn = fminf(a.cur_minmax[0], b.cur_minmax[0]);
  const float mx = fmaxf(a.cur_minmax[1], b.cur_minmax[1]);
  if (observe) observer_update(fq, mn, mx, 0, 255, false, c);
  const float s = *fq.scale, zp = (float)*fq.zero_point;
  const float inv = __fdiv_rn(1.0f, s);
  const float qa = fminf(fmaxf(fq_index(mn, inv, zp), 0.0f), 255.0f);
  const float qb = fminf(fmaxf(fq_index(mx, inv, zp), 0.0f), 255.0f);
  cur_minmax_out[0] = fq_dequant(qa, zp, s);
  cur_minmax_out[1] = fq_dequant(qb, zp, s);
}

// one thread-iteration = 8 output bytes (C1, C2 multiples of 8)
__global__ void __launch_bounds__(256) cat_requant_kernel(FrostQTensor a, FrostQTensor b, int64_t M, const float* out_scale,
                                                         const int32_t* out_zp, uint8_t* q_out, int ld_out) {
  pdl_enter();
  const QSrc sa = load_src(a), sb = load_src(b);
  const float so = *out_scale, zo = (float)*out_zp;
  const float inv = __fdiv_rn(1.0f, so);
  const int C1 = a.C, C2 = b.C, C = C1 + C2;
  const int G = C >> 3, G1 = C1 >> 3;
  const int64_t total = M * G;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t m = i / G;
    const int g = (int)(i - m * G);
    uint2 in;
    QSrc s;
    if (g < G1) { in = ld_cg(reinterpret_cast<const uint2*>(a.q + m * a.ld) + g); s = sa; }
    else { in = ld_cg(reinterpret_cast<const uint2*>(b.q + m * b.ld) + (g - G1)); s = sb; }
    const unsigned w[2] = {in.x, in.y};
    unsigned o[2] = {0u, 0u};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float x = deq((w[e >> 2] >> (8 * (e & 3))) & 0xff, s);
      const float qc = fminf(fmaxf(fq_index(x, inv, zo), 0.0f), 255.0f);
      o[e >> 2] |= ((unsigned)qc) << (8 * (e & 3));
    }
    *reinterpret_cast<uint2*>(q_out + m * ld_out + (int64_t)g * 8) = make_uint2(o[0], o[1]);
  }
}

// one thread-iteration = 4 fp32 gradient elements
__global__ void __launch_bounds__(256) cat_backward_kernel(const float* dcat, FrostQTensor a, FrostQTensor b,
                                                          int64_t M, const float* out_scale, const int32_t* out_zp,
                                                          float* da, float* db, int accumulate_b) {
  pdl_enter();
  const QSrc sa = load_src(a), sb = load_src(b);
  const float so = *out_scale, zo = (float)*out_zp;
  const float inv = __fdiv_rn(1.0f, so);
  const int C1 = a.C, C2 = b.C, C = C1 + C2;
  const int G = C >> 2, G1 = C1 >> 2;
  const int64_t total = M * G;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int64_t m = i / G;
    const int g = (int)(i - m * G);
    const float4 d = ld_cg(reinterpret_cast<const float4*>(dcat + m * C) + g);
    const bool from_a = g < G1;
    const unsigned in = from_a ? ld_cg(reinterpret_cast<const unsigned*>(a.q + m * a.ld) + g)
                               : ld_cg(reinterpret_cast<const unsigned*>(b.q + m * b.ld) + (g - G1));
    const QSrc s = from_a ? sa : sb;
    const float dv[4] = {d.x, d.y, d.z, d.w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float idx = fq_index(deq((in >> (8 * e)) & 0xff, s), inv, zo);
      o[e] = (idx >= 0.0f && idx <= 255.0f) ? dv[e] : 0.0f;
    }
    if (from_a) {
      reinterpret_cast<float4*>(da + m * C1)[g] = make_float4(o[0], o[1], o[2], o[3]);
    } else {
      float4* p = reinterpret_cast<float4*>(db + m * C2) + (g - G1);
      if (accumulate_b) {
        const float4 old = ld_cg(p);
        *p = make_float4(old.x + o[0], old.y + o[1], old.z + o[2], old.w + o[3]);
      } else {
        *p = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
}

// ---------------------------------------------------------------- add
// word i of the logical [M][C] tensor -> byte offset in a tensor whose rows are ld bytes apart
__device__ __forceinline__ int64_t pitched(int64_t i, int C4, int ld) {
  if (ld == C4 * 4) return i * 4;
  const int64_t row = i / C4;
  return row * ld + (i - row * C4) * 4;
}

__global__ void __launch_bounds__(256) add_minmax_kernel(FrostQTensor a, FrostQTensor b, int64_t n4, float* partial) {
  pdl_enter();
  const QSrc sa = load_src(a), sb = load_src(b);
  float mn = INFINITY, mx = -INFINITY;
  const int C4 = a.C >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const unsigned wa = ld_cg(reinterpret_cast<const unsigned*>(a.q + pitched(i, C4, a.ld)));
    const unsigned wb = ld_cg(reinterpret_cast<const unsigned*>(b.q + pitched(i, C4, b.ld)));
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v = __fadd_rn(deq((wa >> (8 * e)) & 0xff, sa), deq((wb >> (8 * e)) & 0xff, sb));
      mn = fminf(mn, v);
      mx = fmaxf(mx, v);
    }
  }
  block_minmax(mn, mx);
  if (threadIdx.x == 0) {
    partial[2 * blockIdx.x] = mn;
    partial[2 * blockIdx.x + 1] = mx;
  }
}

__global__ void __launch_bounds__(256) add_requant_kernel(FrostQTensor a, FrostQTensor b, int64_t n4, const float* out_scale,
                                                         const int32_t* out_zp, uint8_t* q_out, int ld_out) {
  pdl_enter();
  const QSrc sa = load_src(a), sb = load_src(b);
  const float so = *out_scale, zo = (float)*out_zp;
  const float inv = __fdiv_rn(1.0f, so);
  const int C4 = a.C >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const unsigned wa = ld_cg(reinterpret_cast<const unsigned*>(a.q + pitched(i, C4, a.ld)));
    const unsigned wb = ld_cg(reinterpret_cast<const unsigned*>(b.q + pitched(i, C4, b.ld)));
    unsigned o = 0u;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v = __fadd_rn(deq((wa >> (8 * e)) & 0xff, sa), deq((wb >> (8 * e)) & 0xff, sb));
      const float qc = fminf(fmaxf(fq_index(v, inv, zo), 0.0f), 255.0f);
      o |= ((unsigned)qc) << (8 * e);
    }
    *reinterpret_cast<unsigned*>(q_out + pitched(i, C4, ld_out)) = o;
  }
}

__global__ void __launch_bounds__(256) add_backward_kernel(const float* dout, FrostQTensor a, FrostQTensor b,
                                                          int64_t n4, const float* out_scale, const int32_t* out_zp,
                                                          float* dsum, float* da, int accumulate_a) {
  pdl_enter();
  const QSrc sa = load_src(a), sb = load_src(b);
  const float so = *out_scale, zo = (float)*out_zp;
  const float inv = __fdiv_rn(1.0f, so);
  const int C4 = a.C >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const unsigned wa = ld_cg(reinterpret_cast<const unsigned*>(a.q + pitched(i, C4, a.ld)));
    const unsigned wb = ld_cg(reinterpret_cast<const unsigned*>(b.q + pitched(i, C4, b.ld)));
    const float4 d = ld_cg(reinterpret_cast<const float4*>(dout) + i);
    const float dv[4] = {d.x, d.y, d.z, d.w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float v = __fadd_rn(deq((wa >> (8 * e)) & 0xff, sa), deq((wb >> (8 * e)) & 0xff, sb));
      const float idx = fq_index(v, inv, zo);
      o[e] = (idx >= 0.0f && idx <= 255.0f) ? dv[e] : 0.0f;
    }
    const float4 r = make_float4(o[0], o[1], o[2], o[3]);
    reinterpret_cast<float4*>(dsum)[i] = r;
    float4* p = reinterpret_cast<float4*>(da) + i;
    if (accumulate_a) {
      const float4 old = ld_cg(p);
      *p = make_float4(old.x + r.x, old.y + r.y, old.z + r.z, old.w + r.w);
    } else {
      *p = r;
    }
  }
}

__global__ void __launch_bounds__(256) axpy_kernel(const float* x, float* y, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) y[i] += x[i];
}

// ---------------------------------------------------------------- head: avg-pool + dropout
// thread -> (n, 4 channels); integer sum over HW is exact.
__global__ void __launch_bounds__(256) pool_dropout_fwd_kernel(const uint8_t* q, const float* scale_p,
                                                              const int32_t* zp_p, int N, int HW, int C,
                                                              const float* keep, float keep_scale,
                                                              float* pooled) {
  const float s = *scale_p;
  const int zp = *zp_p;
  const int CG = C >> 2;
  const int64_t total = (int64_t)N * CG;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(i / CG), cg = (int)(i % CG);
    int sum[4] = {0, 0, 0, 0};
    const uint8_t* base = q + ((int64_t)n * HW) * C + cg * 4;
    for (int p = 0; p < HW; ++p) {
      const unsigned w = ld_cg(reinterpret_cast<const unsigned*>(base + (int64_t)p * C));
#pragma unroll
      for (int e = 0; e < 4; ++e) sum[e] += (int)((w >> (8 * e)) & 0xff);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v = __fdiv_rn(__fmul_rn((float)(sum[e] - HW * zp), s), (float)HW);
      const int64_t o = (int64_t)n * C + cg * 4 + e;
      if (keep) v = __fmul_rn(v, __fmul_rn(keep[o], keep_scale));
      pooled[o] = v;
    }
  }
}

__global__ void __launch_bounds__(256) pool_dropout_bwd_kernel(const float* dpooled, int N, int HW, int C,
                                                              const float* keep, float keep_scale,
                                                              float* dy) {
  const int CG = C >> 2;
  const int64_t total = (int64_t)N * HW * CG;
  const float inv_hw = 1.0f / (float)HW;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cg = (int)(i % CG);
    const int64_t n = i / ((int64_t)HW * CG);
    float4 d = ld_cg(reinterpret_cast<const float4*>(dpooled + n * C) + cg);
    if (keep) {
      const float4 k = ld_cg(reinterpret_cast<const float4*>(keep + n * C) + cg);
      d.x *= k.x * keep_scale; d.y *= k.y * keep_scale; d.z *= k.z * keep_scale; d.w *= k.w * keep_scale;
    }
    reinterpret_cast<float4*>(dy)[i] = make_float4(d.x * inv_hw, d.y * inv_hw, d.z * inv_hw, d.w * inv_hw);
  }
}

}  // namespace frost

using namespace frost;

static int check_qt(const FrostQTensor& t) { return t.q && t.scale && t.zp && t.C > 0 && t.ld >= t.C && t.ld % 4 == 0; }

extern "C" int frost_cat_forward(FrostQTensor a, FrostQTensor b, int64_t M, FrostFQ fq, int observe,
                                 float averaging_const, uint8_t* q_out, int ld_out, float* cur_minmax_out, void* stream) {
  FROST_REQUIRE(check_qt(a) && check_qt(b) && a.cur_minmax && b.cur_minmax && q_out && cur_minmax_out && fq.scale &&
                    fq.zero_point && fq.min_val && fq.max_val,
                "frost_cat_forward: null pointer");
  FROST_REQUIRE(M > 0 && a.C % 8 == 0 && b.C % 8 == 0, "frost_cat_forward: channel counts must be multiples of 8");
  FROST_REQUIRE(ld_out >= a.C + b.C && ld_out % 8 == 0 && a.ld % 8 == 0 && b.ld % 8 == 0, "frost_cat_forward: row pitches must be multiples of 8");
  cudaStream_t st = (cudaStream_t)stream;
  launch_pdl(cat_finalize_kernel, dim3(1), dim3(32), 0, st, a, b, fq, observe, averaging_const, cur_minmax_out);
  FROST_LAUNCH_CHECK("cat_finalize");
  const int64_t total = M * ((a.C + b.C) / 8);
  launch_pdl(cat_requant_kernel, dim3(grid_for(total, 256 * 2)), dim3(256), 0, st, a, b, M, fq.scale, fq.zero_point, q_out, ld_out);
  FROST_LAUNCH_CHECK("cat_requant");
  return FROST_OK;
}

extern "C" int frost_cat_backward(const float* dcat, FrostQTensor a, FrostQTensor b, int64_t M, const float* out_scale,
                                  const int32_t* out_zp, float* da, float* db, int accumulate_b, void* stream) {
  FROST_REQUIRE(dcat && check_qt(a) && check_qt(b) && out_scale && out_zp && da && db, "frost_cat_backward: null pointer");
  FROST_REQUIRE(M > 0 && a.C % 4 == 0 && b.C % 4 == 0, "frost_cat_backward: channel counts must be multiples of 4");
  const int64_t total = M * ((a.C + b.C) / 4);
  launch_pdl(cat_backward_kernel, dim3(grid_for(total, 256 * 2)), dim3(256), 0, (cudaStream_t)stream, dcat, a, b, M, out_scale, out_zp,
             da, db, accumulate_b);
  FROST_LAUNCH_CHECK("cat_backward");
  return FROST_OK;
}

extern "C" int frost_add_forward(FrostQTensor a, FrostQTensor b, int64_t n, FrostFQ fq, int observe,
                                 float averaging_const, uint8_t* q_out, int ld_out, float* cur_minmax_out, float* scratch,
                                 void* stream) {
  FROST_REQUIRE(check_qt(a) && check_qt(b) && q_out && cur_minmax_out && scratch && fq.scale && fq.zero_point &&
                    fq.min_val && fq.max_val,
                "frost_add_forward: null pointer");
  FROST_REQUIRE(n > 0 && n % 4 == 0, "frost_add_forward: n must be a positive multiple of 4");
  FROST_REQUIRE(a.C == b.C && a.C % 4 == 0 && n % a.C == 0 && ld_out >= a.C && ld_out % 4 == 0,
                "frost_add_forward: both operands are [n/C][C] with C %% 4 == 0; ld_out >= C");
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t n4 = n / 4;
  const int nblk = grid_for(n4, 256 * 4, FROST_FQ_SCRATCH_FLOATS / 2);
  launch_pdl(add_minmax_kernel, dim3(nblk), dim3(256), 0, st, a, b, n4, scratch);
  FROST_LAUNCH_CHECK("add_minmax");
  launch_pdl(fq_finalize_kernel, dim3(1), dim3(1024), 0, st, scratch, nblk, fq, 0, 255, 0, averaging_const, observe ? 1 : 0, cur_minmax_out);
  FROST_LAUNCH_CHECK("add_finalize");
  launch_pdl(add_requant_kernel, dim3(grid_for(n4, 256 * 4)), dim3(256), 0, st, a, b, n4, fq.scale, fq.zero_point, q_out, ld_out);
  FROST_LAUNCH_CHECK("add_requant");
  return FROST_OK;
}

extern "C" int frost_add_backward(const float* dout, FrostQTensor a, FrostQTensor b, int64_t n, const float* out_scale,
                                  const int32_t* out_zp, float* dsum, float* da, int accumulate_a, void* stream) {
  FROST_REQUIRE(dout && check_qt(a) && check_qt(b) && out_scale && out_zp && dsum && da, "frost_add_backward: null pointer");
  FROST_REQUIRE(n > 0 && n % 4 == 0 && a.C == b.C && a.C % 4 == 0 && n % a.C == 0, "frost_add_backward: n must be a positive multiple of C, C %% 4 == 0");
  launch_pdl(add_backward_kernel, dim3(grid_for(n / 4, 256 * 4)), dim3(256), 0, (cudaStream_t)stream, dout, a, b, n / 4, out_scale, out_zp,
             dsum, da, accumulate_a);
  FROST_LAUNCH_CHECK("add_backward");
  return FROST_OK;
}

extern "C" int frost_axpy(const float* x, float* y, int64_t n, void* stream) {
  FROST_REQUIRE(x && y && n > 0, "frost_axpy: bad args");
  axpy_kernel<<<grid_for(n, 256 * 4), 256, 0, (cudaStream_t)stream>>>(x, y, n);
  FROST_LAUNCH_CHECK("axpy");
  return FROST_OK;
}

extern "C" int frost_pool_dropout_forward(const uint8_t* q, const float* scale, const int32_t* zp, int N, int HW, int C,
                                          const float* keep, float keep_scale, float* pooled, void* stream) {
  FROST_REQUIRE(q && scale && zp && pooled && N > 0 && HW > 0 && C > 0 && C % 4 == 0, "frost_pool_dropout_forward: bad args");
  pool_dropout_fwd_kernel<<<grid_for((int64_t)N * C / 4, 256), 256, 0, (cudaStream_t)stream>>>(q, scale, zp, N, HW, C,
                                                                                             keep, keep_scale, pooled);
  FROST_LAUNCH_CHECK("pool_dropout_fwd");
  return FROST_OK;
}

extern "C" int frost_pool_dropout_backward(const float* dpooled, int N, int HW, int C, const float* keep,
                                           float keep_scale, float* dy, void* stream) {
  FROST_REQUIRE(dpooled && dy && N > 0 && HW > 0 && C > 0 && C % 4 == 0, "frost_pool_dropout_backward: bad args");
  pool_dropout_bwd_kernel<<<grid_for((int64_t)N * HW * C / 4, 256 * 2), 256, 0, (cudaStream_t)stream>>>(
      dpooled, N, HW, C, keep, keep_scale, dy);
  FROST_LAUNCH_CHECK("pool_dropout_bwd");
  return FROST_OK;
}
