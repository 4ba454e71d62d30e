// hswish.cu - the quantization-aware hard-swish of the reference's MobileNetV3 blocks
// (Classification/models/imagenet/mobilenetv3.py:43-56):
//     a  = x + 3                         FloatFunctional.add_scalar  (not observed)
//     ra = FQ_A(relu6(a))                nn.ReLU6 + its activation_post_process
//     mb = FQ_B(x * ra)                  FloatFunctional.mul         (observed)
//     y  = mb * (1/6)                    FloatFunctional.mul_scalar  (not observed)
// The input sits on a uint8 grid, so it takes at most 256 distinct values and EVERYTHING above is a function of the
// input index: both observers' min / max are extrema over the indices that are present, and the element-wise work
// collapses to table lookups.
//   pass 1  index + presence : q = clamp(rint(x/s)+zp), 256 presence flags                 (reads 4 B, writes 1 B / element)
//   tables  (one CTA, thread = index): value, relu6, observer A -> qparams, FQ_A, product, observer B -> qparams, FQ_B,
//           output value / index, and the per-index factors of the backward
//   pass 2  y = OUT[q]                                                                       (reads 1 B, writes 4 B / element)
//   backward dx = gm*RA[q] + [inner mask] gm*V[q],  gm = [FQ_B in range] dy*(1/6)           (reads 5 B, writes 4 B / element)
// instead of 4 fake-quant kernels, 2 min/max reductions and 4 element-wise ops over fp32 tensors.
#include "common.cuh"

namespace frost {

constexpr int HS_N = 256;
// workspace (HS_N-float planes): 0 presence (as uint32), 1 V, 2 RA, 3 inner mask, 4 outer mask, 5 OUT (fp32), 6 OUT index
constexpr int HS_PLANES = 7;

__global__ void __launch_bounds__(256) hswish_index_kernel(const float* x, int64_t n, int vec, const float* in_scale, const int32_t* in_zp,
                                                          uint8_t* q_in, unsigned* presence) {
  __shared__ unsigned s_p[HS_N];
  s_p[threadIdx.x] = 0u;
  __syncthreads();
  const float inv = __fdiv_rn(1.0f, *in_scale), zp = (float)*in_zp;
  auto index_of = [&](float v) { return (unsigned)fminf(fmaxf(fq_index(v, inv, zp), 0.0f), 255.0f); };
  // 4 elements per thread and iteration (16-byte load, 4-byte store); the host guarantees the alignment when vec != 0
  const int64_t n4 = vec ? n / 4 : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = ld_cg(reinterpret_cast<const float4*>(x) + i);
    const unsigned q0 = index_of(v.x), q1 = index_of(v.y), q2 = index_of(v.z), q3 = index_of(v.w);
    reinterpret_cast<unsigned*>(q_in)[i] = q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
    s_p[q0] = 1u; s_p[q1] = 1u; s_p[q2] = 1u; s_p[q3] = 1u;       // benign race: every writer stores 1
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned q = index_of(ld_cg(x + i));
    q_in[i] = (uint8_t)q;
    s_p[q] = 1u;
  }
  __syncthreads();
  if (s_p[threadIdx.x]) presence[threadIdx.x] = 1u;
}

// sigmoid != 0: the reference's _Hsigmoid (mobilenetv3.py:59-69)  y = mul_scalar(FQ_relu6(relu6(x + 3)), 1/6) - no product,
// no second fake-quant; the backward tables are set so that hswish_bwd_kernel computes  dx = [inner mask] dy/6.
__global__ void __launch_bounds__(HS_N) hswish_tables_kernel(const float* in_scale, const int32_t* in_zp, FrostFQ fq_a, int observe_a,
                                                           FrostFQ fq_b, int observe_b, float avg_c, float* ws, float* out_scale,
                                                           int sigmoid) {
  const int i = threadIdx.x;
  const bool present = reinterpret_cast<const unsigned*>(ws)[i] != 0u;
  const float v = fq_dequant((float)i, (float)*in_zp, *in_scale);
  const float a = __fadd_rn(v, 3.0f);
  const float r = fminf(fmaxf(a, 0.0f), 6.0f);
  float mn = present ? r : INFINITY, mx = present ? r : -INFINITY;
  block_minmax(mn, mx);
  if (i == 0 && observe_a) observer_update(fq_a, mn, mx, 0, 255, false, avg_c);
  __syncthreads();
  const float s_a = *fq_a.scale, zp_a = (float)*fq_a.zero_point;
  const float idx_a = fq_index(r, __fdiv_rn(1.0f, s_a), zp_a);
  const bool pass_a = idx_a >= 0.0f && idx_a <= 255.0f;
  const float ra = fq_dequant(fminf(fmaxf(idx_a, 0.0f), 255.0f), zp_a, s_a);
  const float c6s = (float)(1.0 / 6.0);
  if (sigmoid) {
    ws[1 * HS_N + i] = 1.0f;                                              // "V": d(out)/d(r) path only
    ws[2 * HS_N + i] = 0.0f;                                              // no direct x factor
    ws[3 * HS_N + i] = (pass_a && a > 0.0f && a < 6.0f) ? 1.0f : 0.0f;
    ws[4 * HS_N + i] = 1.0f;
    ws[5 * HS_N + i] = __fmul_rn(ra, c6s);
    ws[6 * HS_N + i] = fminf(fmaxf(idx_a, 0.0f), 255.0f);
    if (i == 0 && out_scale) *out_scale = __fmul_rn(s_a, c6s);            // grid of the result: (q - zp_a) * (s_a / 6)
    return;
  }
  const float m = __fmul_rn(v, ra);
  mn = present ? m : INFINITY;
  mx = present ? m : -INFINITY;
  block_minmax(mn, mx);
  if (i == 0 && observe_b) observer_update(fq_b, mn, mx, 0, 255, false, avg_c);
  __syncthreads();
  const float s_b = *fq_b.scale, zp_b = (float)*fq_b.zero_point;
  const float idx_b = fq_index(m, __fdiv_rn(1.0f, s_b), zp_b);
  const bool pass_b = idx_b >= 0.0f && idx_b <= 255.0f;
  const float qb = fminf(fmaxf(idx_b, 0.0f), 255.0f);
  const float c6 = (float)(1.0 / 6.0);
  ws[1 * HS_N + i] = v;
  ws[2 * HS_N + i] = ra;
  ws[3 * HS_N + i] = (pass_a && a > 0.0f && a < 6.0f) ? 1.0f : 0.0f;     // FQ_A's STE x hardtanh'(a)
  ws[4 * HS_N + i] = pass_b ? 1.0f : 0.0f;
  ws[5 * HS_N + i] = __fmul_rn(fq_dequant(qb, zp_b, s_b), c6);
  ws[6 * HS_N + i] = qb;
  if (i == 0 && out_scale) *out_scale = __fmul_rn(s_b, c6);             // grid of the result: (q - zp_b) * (s_b / 6)
}

__global__ void __launch_bounds__(256) hswish_apply_kernel(const uint8_t* q_in, int64_t n, int vec, const float* ws, float* y, uint8_t* y_q) {
  __shared__ float s_out[HS_N];
  __shared__ uint8_t s_q[HS_N];
  s_out[threadIdx.x] = ws[5 * HS_N + threadIdx.x];
  s_q[threadIdx.x] = (uint8_t)ws[6 * HS_N + threadIdx.x];
  __syncthreads();
  const int64_t n4 = vec ? n / 4 : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned w = ld_cg(reinterpret_cast<const unsigned*>(q_in) + i);
    const unsigned q0 = w & 255u, q1 = (w >> 8) & 255u, q2 = (w >> 16) & 255u, q3 = w >> 24;
    if (y) reinterpret_cast<float4*>(y)[i] = make_float4(s_out[q0], s_out[q1], s_out[q2], s_out[q3]);
    if (y_q) reinterpret_cast<unsigned*>(y_q)[i] = (unsigned)s_q[q0] | ((unsigned)s_q[q1] << 8) | ((unsigned)s_q[q2] << 16) | ((unsigned)s_q[q3] << 24);
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned q = q_in[i];
    if (y) y[i] = s_out[q];
    if (y_q) y_q[i] = s_q[q];
  }
}

__global__ void __launch_bounds__(256) hswish_bwd_kernel(const float* dy, const uint8_t* q_in, int64_t n, int vec, const float* ws, float* dx) {
  __shared__ float s_v[HS_N], s_ra[HS_N], s_mi[HS_N], s_mo[HS_N];
  s_v[threadIdx.x] = ws[1 * HS_N + threadIdx.x];
  s_ra[threadIdx.x] = ws[2 * HS_N + threadIdx.x];
  s_mi[threadIdx.x] = ws[3 * HS_N + threadIdx.x];
  s_mo[threadIdx.x] = ws[4 * HS_N + threadIdx.x];
  __syncthreads();
  const float c6 = (float)(1.0 / 6.0);
  // autograd of the reference, op by op: mul_scalar, FQ_B (STE), mul (both operands), FQ_A (STE), hardtanh, add_scalar
  auto grad = [&](float g, unsigned q) {
    const float gm = s_mo[q] != 0.0f ? __fmul_rn(g, c6) : 0.0f;
    const float g1 = __fmul_rn(gm, s_ra[q]);
    const float g2 = s_mi[q] != 0.0f ? __fmul_rn(gm, s_v[q]) : 0.0f;
    return __fadd_rn(g1, g2);
  };
  const int64_t n4 = vec ? n / 4 : 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const unsigned w = ld_cg(reinterpret_cast<const unsigned*>(q_in) + i);
    const float4 g = ld_cg(reinterpret_cast<const float4*>(dy) + i);
    reinterpret_cast<float4*>(dx)[i] = make_float4(grad(g.x, w & 255u), grad(g.y, (w >> 8) & 255u), grad(g.z, (w >> 16) & 255u), grad(g.w, w >> 24));
  }
  for (int64_t i = n4 * 4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = grad(ld_cg(dy + i), q_in[i]);
}

}  // namespace frost

using namespace frost;

extern "C" int frost_hswish_workspace_floats(void) { return HS_PLANES * HS_N; }

static int hs_forward(const float* x, int64_t n, const float* in_scale, const int32_t* in_zp, FrostFQ fq_relu6, int observe_relu6,
                      FrostFQ fq_mul, int observe_mul, float averaging_const, uint8_t* q_in, float* y, uint8_t* y_q, float* workspace,
                      float* out_scale, int sigmoid, void* stream);

extern "C" int frost_hswish_forward(const float* x, int64_t n, const float* in_scale, const int32_t* in_zp, FrostFQ fq_relu6,
                                    int observe_relu6, FrostFQ fq_mul, int observe_mul, float averaging_const, uint8_t* q_in,
                                    float* y, uint8_t* y_q, float* workspace, float* out_scale, void* stream) {
  FROST_REQUIRE(fq_mul.scale && fq_mul.zero_point && fq_mul.min_val && fq_mul.max_val, "frost_hswish_forward: null fake-quant state");
  return hs_forward(x, n, in_scale, in_zp, fq_relu6, observe_relu6, fq_mul, observe_mul, averaging_const, q_in, y, y_q, workspace,
                    out_scale, 0, stream);
}

extern "C" int frost_hsigmoid_forward(const float* x, int64_t n, const float* in_scale, const int32_t* in_zp, FrostFQ fq_relu6,
                                      int observe_relu6, float averaging_const, uint8_t* q_in, float* y, uint8_t* y_q,
                                      float* workspace, float* out_scale, void* stream) {
  return hs_forward(x, n, in_scale, in_zp, fq_relu6, observe_relu6, fq_relu6, 0, averaging_const, q_in, y, y_q, workspace, out_scale, 1,
                    stream);
}

static int hs_forward(const float* x, int64_t n, const float* in_scale, const int32_t* in_zp, FrostFQ fq_relu6, int observe_relu6,
                      FrostFQ fq_mul, int observe_mul, float averaging_const, uint8_t* q_in, float* y, uint8_t* y_q, float* workspace,
                      float* out_scale, int sigmoid, void* stream) {
  FROST_REQUIRE(x && n > 0 && in_scale && in_zp && q_in && workspace && (y || y_q), "frost_hswish_forward: bad args");
  FROST_REQUIRE(fq_relu6.scale && fq_relu6.zero_point && fq_relu6.min_val && fq_relu6.max_val, "frost_hswish_forward: null fake-quant state");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(workspace, 0, sizeof(float) * HS_N, st) != cudaSuccess) {
    set_error("frost_hswish_forward: memset failed");
    return FROST_ECUDA;
  }
  const unsigned blocks = grid_for(n, 256 * 8, kNumSMs * 8);
  auto al = [](const void* p, uintptr_t a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; };
  const int vec = al(x, 16) && al(q_in, 4) && al(y, 16) && al(y_q, 4) ? 1 : 0;
  hswish_index_kernel<<<blocks, 256, 0, st>>>(x, n, vec, in_scale, in_zp, q_in, reinterpret_cast<unsigned*>(workspace));
  FROST_LAUNCH_CHECK("hswish_index");
  hswish_tables_kernel<<<1, HS_N, 0, st>>>(in_scale, in_zp, fq_relu6, observe_relu6, fq_mul, observe_mul, averaging_const, workspace,
                                          out_scale, sigmoid);
  FROST_LAUNCH_CHECK("hswish_tables");
  hswish_apply_kernel<<<blocks, 256, 0, st>>>(q_in, n, vec, workspace, y, y_q);
  FROST_LAUNCH_CHECK("hswish_apply");
  return FROST_OK;
}

extern "C" int frost_hswish_backward(const float* dy, const uint8_t* q_in, int64_t n, const float* workspace, float* dx, void* stream) {
  FROST_REQUIRE(dy && q_in && workspace && dx && n > 0, "frost_hswish_backward: bad args");
  const int vec = ((reinterpret_cast<uintptr_t>(dy) | reinterpret_cast<uintptr_t>(dx)) & 15) == 0 && (reinterpret_cast<uintptr_t>(q_in) & 3) == 0;
  hswish_bwd_kernel<<<grid_for(n, 256 * 8, kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(dy, q_in, n, vec, workspace, dx);
  FROST_LAUNCH_CHECK("hswish_bwd");
  return FROST_OK;
}
