// bn.cu - BatchNorm(+ReLU)+activation fake-quant of one fused conv, forward and backward.
// Restates torch/ao/nn/intrinsic/qat/modules/conv_fused.py:156-167 (+ :708-710 ReLU) and the
// activation_post_process hook (fake_quantize.py:423-438) on top of the integer accumulators I:
//   conv_orig = conv/scale_factor = I * m_c,  m_c = s_a*s_w/scale_factor_c
//   v = bn(conv_orig) = A_c*I + B_c          (training: batch statistics from exact integer sums)
//   y = FQ_a(relu(v))
#include <cuda_bf16.h>
#include "bn_math.cuh"

namespace frost {

void dw_launch_shape(int C, int max_cgb, int* cg_per_block, int* nchunks, int* threads);

// ---------------------------------------------------------------- finalize (1 CTA)
__global__ void __launch_bounds__(1024) bn_finalize_kernel(FrostBnFinalizeArgs a) {
  __shared__ float s_mn[32], s_mx[32];
  pdl_enter();
  const double M = (double)a.count;
  const double sa_sw = (double)(*a.x_scale) * (double)(*a.w_scale);
  // momentum < 0 <-> nn.BatchNorm2d(momentum=None): cumulative moving average, factor 1/num_batches_tracked
  // (the counter is incremented by thread 0 after the barrier below, so every thread reads the old value here)
  const double mom = a.momentum >= 0.0f ? (double)a.momentum
                                        : 1.0 / (double)((a.num_batches_tracked ? *a.num_batches_tracked : 0) + 1);
  float gmn = INFINITY, gmx = -INFINITY;
  for (int c = threadIdx.x; c < a.C; c += blockDim.x) {
    const BnChannel r = bn_channel_finalize(a.stats[c], a.stats_format, M, a.count > 1, sa_sw, a.sf[c], a.gamma[c], a.beta[c],
                                            a.running_mean[c], a.running_var[c], a.eps, mom, a.training, a.relu);
    if (a.training) {
      a.running_mean[c] = r.new_running_mean;
      a.running_var[c] = r.new_running_var;
    }
    a.A[c] = r.A;
    a.B[c] = r.B;
    a.mean_I[c] = r.mean_I;
    a.kfac[c] = r.kfac;
    gmn = fminf(gmn, r.v_lo);
    gmx = fmaxf(gmx, r.v_hi);
  }
  gmn = warp_min(gmn);
  gmx = warp_max(gmx);
  if ((threadIdx.x & 31) == 0) { s_mn[threadIdx.x >> 5] = gmn; s_mx[threadIdx.x >> 5] = gmx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { gmn = fminf(gmn, s_mn[w]); gmx = fmaxf(gmx, s_mx[w]); }
    if (a.training && a.num_batches_tracked) *a.num_batches_tracked += 1;
    if (a.observe) observer_update(a.afq, gmn, gmx, 0, 255, false, a.averaging_const);
    const float s = *a.afq.scale, zp = (float)*a.afq.zero_point;
    const float inv = __fdiv_rn(1.0f, s);
    const float qa = fminf(fmaxf(fq_index(gmn, inv, zp), 0.0f), 255.0f);
    const float qb = fminf(fmaxf(fq_index(gmx, inv, zp), 0.0f), 255.0f);
    a.cur_minmax[0] = fq_dequant(qa, zp, s);
    a.cur_minmax[1] = fq_dequant(qb, zp, s);
  }
}

// ---------------------------------------------------------------- apply: I -> uint8 index
// 4 consecutive channels per thread-iteration (one 16-byte load, one 4-byte store; the per-channel affine
// is read from shared memory as float4 - consecutive lanes hit consecutive banks), two iterations in flight.
__device__ __forceinline__ unsigned bnq4(const int4 v, int fmt, const float4 A, const float4 B, int relu, float inv, float zp) {
  const unsigned q0 = bnq1(acc_val(v.x, fmt), A.x, B.x, relu, inv, zp), q1 = bnq1(acc_val(v.y, fmt), A.y, B.y, relu, inv, zp);
  const unsigned q2 = bnq1(acc_val(v.z, fmt), A.z, B.z, relu, inv, zp), q3 = bnq1(acc_val(v.w, fmt), A.w, B.w, relu, inv, zp);
  return q0 | (q1 << 8) | (q2 << 16) | (q3 << 24);
}

template <int U>   // 16-byte loads in flight per thread
__global__ void __launch_bounds__(256) bnq_apply_kernel(const int32_t* acc, int fmt, int64_t n4, int C,
                                                       const float* A, const float* B,
                                                       int relu, const float* scale_p,
                                                       const int32_t* zp_p, uint8_t* q, int ldq) {
  extern __shared__ __align__(16) float s_ab[];  // A[C], B[C]
  pdl_enter();
  for (int c = threadIdx.x; c < C; c += blockDim.x) { s_ab[c] = A[c]; s_ab[C + c] = B[c]; }
  __syncthreads();
  const float s = *scale_p, zp = (float)*zp_p;
  const float inv = __fdiv_rn(1.0f, s);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int C4 = C >> 2;
  // (row, 4-channel group) of this thread's element, advanced without divisions in the loop; the output rows are
  // ldq bytes apart (ldq == C: dense)
  int cg = (int)(i0 % C4);
  int64_t row = i0 / C4;
  const int cstep = (int)(stride % C4);
  const int64_t rstep = stride / C4;
  const int ld4 = ldq >> 2;
  const int4* in = reinterpret_cast<const int4*>(acc);
  unsigned* out = reinterpret_cast<unsigned*>(q);
  const float4* sA = reinterpret_cast<const float4*>(s_ab);
  const float4* sB = reinterpret_cast<const float4*>(s_ab + C);
  for (int64_t i = i0; i < n4; i += U * stride) {
    int4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (i + u * stride < n4) v[u] = ld_cg(in + i + u * stride);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (i + u * stride < n4) out[row * ld4 + cg] = bnq4(v[u], fmt, sA[cg], sB[cg], relu, inv, zp);
      cg += cstep;
      row += rstep;
      if (cg >= C4) { cg -= C4; ++row; }
    }
  }
}

// ---------------------------------------------------------------- backward
// Per-channel S1 = sum dv, S2 = sum dv*(I - mean_I).  Thread -> fixed 4-channel group, strided rows.
// A CTA owns a narrow slice of channels (cg_per_block <= 16 groups: 256-byte row segments) and many rows, so
// that few CTAs contribute to any one channel: the partials are combined through shared memory (no atomics)
// and each CTA issues ONE fp64 atomic per (channel, sum) - same-sector L2 atomics serialise (~12 ns each), and
// with all-channel CTAs that tail was longer than the streaming pass for most layers.
template <int U, int MINB>   // U rows in flight per thread (2*U independent 16-byte loads), MINB resident CTAs per SM
__global__ void __launch_bounds__(256, MINB) bn_bwd_reduce_kernel(FrostBnBackwardArgs a, int cg_per_block) {
  extern __shared__ double s_part[];  // [rows_per_block][cg_per_block][4 channels][2 sums], then [cg_per_block][3] float4 (A, B, mean)
  pdl_enter();
  const int C = a.C;
  const int cg_local = threadIdx.x % cg_per_block;
  const int cg = blockIdx.y * cg_per_block + cg_local;
  const int rows_per_block = blockDim.x / cg_per_block;
  const int row_local = threadIdx.x / cg_per_block;
  // this thread's fp64 accumulators live in its own shared-memory slot [row_local][cg_local*8 + 2*ch + {0,1}]
  // (column t of the CTA's slice is the global sum index blockIdx.y*cg_per_block*8 + t; sums is [C][2])
  double* mine = s_part + ((size_t)row_local * cg_per_block + cg_local) * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) mine[j] = 0.0;
  {
    const float inv = __fdiv_rn(1.0f, *a.out_scale), zp = (float)*a.out_zp;
    // per-channel-group coefficients live in shared memory (12 registers less: 4 CTAs per SM instead of 3)
    float4* s_coef = reinterpret_cast<float4*>(s_part + (size_t)blockDim.x * 8) + cg_local * 3;
    if (row_local == 0) {
      s_coef[0] = ld_cg(reinterpret_cast<const float4*>(a.A) + cg);
      s_coef[1] = ld_cg(reinterpret_cast<const float4*>(a.B) + cg);
      s_coef[2] = ld_cg(reinterpret_cast<const float4*>(a.mean_I) + cg);
    }
    __syncthreads();
    const int64_t stride = (int64_t)gridDim.x * rows_per_block;
    int batches = 0;
    float p1[4] = {0.f, 0.f, 0.f, 0.f}, p2[4] = {0.f, 0.f, 0.f, 0.f};
    for (int64_t m = (int64_t)blockIdx.x * rows_per_block + row_local; m < a.M; m += U * stride) {
      float4 dy[U];
      int4 I[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int64_t mm = m + u * stride;
        dy[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        I[u] = make_int4(0, 0, 0, 0);
        if (mm < a.M) {
          dy[u] = ld_cg(reinterpret_cast<const float4*>(a.dy + mm * C) + cg);
          I[u] = ld_cg(reinterpret_cast<const int4*>(a.acc + mm * C) + cg);
        }
      }
      const float4 A = s_coef[0], B = s_coef[1], mu = s_coef[2];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const float i0 = acc_val(I[u].x, a.acc_format), i1 = acc_val(I[u].y, a.acc_format);
        const float i2 = acc_val(I[u].z, a.acc_format), i3 = acc_val(I[u].w, a.acc_format);
        const float d0 = bn_dv(dy[u].x, i0, A.x, B.x, a.relu, inv, zp);
        const float d1 = bn_dv(dy[u].y, i1, A.y, B.y, a.relu, inv, zp);
        const float d2 = bn_dv(dy[u].z, i2, A.z, B.z, a.relu, inv, zp);
        const float d3 = bn_dv(dy[u].w, i3, A.w, B.w, a.relu, inv, zp);
        p1[0] += d0; p1[1] += d1; p1[2] += d2; p1[3] += d3;
        // centred: no cancellation between sum dv*I and mean*sum dv, so fp32 partial sums suffice
        p2[0] = fmaf(d0, i0 - mu.x, p2[0]);
        p2[1] = fmaf(d1, i1 - mu.y, p2[1]);
        p2[2] = fmaf(d2, i2 - mu.z, p2[2]);
        p2[3] = fmaf(d3, i3 - mu.w, p2[3]);
      }
      if (++batches == 64 / U) {  // flush the fp32 partials (<= 64 terms) into the fp64 accumulators
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) { mine[2 * ch] += (double)p1[ch]; mine[2 * ch + 1] += (double)p2[ch]; p1[ch] = 0.f; p2[ch] = 0.f; }
        batches = 0;
      }
    }
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) { mine[2 * ch] += (double)p1[ch]; mine[2 * ch + 1] += (double)p2[ch]; }
  }
  __syncthreads();
  const int cols = cg_per_block * 8;
  for (int t = threadIdx.x; t < cols; t += blockDim.x) {
    double tot = 0.0;
    for (int r = 0; r < rows_per_block; ++r) tot += s_part[(size_t)r * cols + t];
    atomicAdd(a.sums + (size_t)blockIdx.y * cols + t, tot);
  }
}

// Per-channel coefficients of  dz = c1*(dv - a0 - a1*(I - mean_I))  plus the BN parameter grads.
template <int U>   // elements (2 x 16-byte loads each) in flight per thread
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(FrostBnBackwardArgs a, int64_t n4) {
  extern __shared__ __align__(16) float s_c[];  // A,B,mean_I,c1,a0,a1 : 6*C
  pdl_enter();
  const int C = a.C;
  {
    // per-channel coefficients of dz = c1*(dv - a0 - a1*(I - mean_I)) from the reduced sums; every block
    // recomputes them (C <= a few thousand, a handful of double flops each); block 0 also emits the BN grads
    const double M = (double)a.M;
    const double sa_sw = (double)(*a.x_scale) * (double)(*a.w_scale);
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      const float A = a.A[c], mean_I = a.mean_I[c];
      const BnBwdChannel r = bn_bwd_channel(a.sums[2 * c], a.sums[2 * c + 1], M, sa_sw, A, a.kfac[c], mean_I, a.gamma[c], a.sf[c],
                                            a.eps, a.frozen ? 0 : 1);
      s_c[c] = A;
      s_c[C + c] = a.B[c];
      s_c[2 * C + c] = mean_I;
      s_c[3 * C + c] = r.c1;
      s_c[4 * C + c] = r.a0;
      s_c[5 * C + c] = r.a1;
      if (blockIdx.x == 0) {
        a.dgamma_bn[c] = r.dgamma_bn;
        a.dbeta[c] = r.dbeta;
        a.dsf_bn[c] = r.dsf_bn;
      }
    }
  }
  __syncthreads();
  const float inv = __fdiv_rn(1.0f, *a.out_scale), zp = (float)*a.out_zp;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int C4 = C >> 2;
  int cg = (int)(i0 % C4);              // C % 4 == 0: a float4 never straddles a row
  const int cstep = (int)(stride % C4);
  const float4* s4 = reinterpret_cast<const float4*>(s_c);
  for (int64_t ib = i0; ib < n4; ib += U * stride) {
    float4 dyv[U];
    int4 Iv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (ib + u * stride < n4) {
        dyv[u] = ld_cg(reinterpret_cast<const float4*>(a.dy) + ib + u * stride);
        Iv[u] = ld_cg(reinterpret_cast<const int4*>(a.acc) + ib + u * stride);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = ib + u * stride;
      const float4 cA = s4[cg], cB = s4[C4 + cg], cM = s4[2 * C4 + cg], c1 = s4[3 * C4 + cg], a0 = s4[4 * C4 + cg], a1 = s4[5 * C4 + cg];
      cg += cstep;
      if (cg >= C4) cg -= C4;
      if (i >= n4) continue;
      const float4 dy = dyv[u];
      const int4 I4 = Iv[u];
      float o[4];
      {
        const float i0 = acc_val(I4.x, a.acc_format), i1 = acc_val(I4.y, a.acc_format);
        const float i2 = acc_val(I4.z, a.acc_format), i3 = acc_val(I4.w, a.acc_format);
        const float dv0 = bn_dv(dy.x, i0, cA.x, cB.x, a.relu, inv, zp), dv1 = bn_dv(dy.y, i1, cA.y, cB.y, a.relu, inv, zp);
        const float dv2 = bn_dv(dy.z, i2, cA.z, cB.z, a.relu, inv, zp), dv3 = bn_dv(dy.w, i3, cA.w, cB.w, a.relu, inv, zp);
        o[0] = c1.x * (dv0 - a0.x - a1.x * (i0 - cM.x));
        o[1] = c1.y * (dv1 - a0.y - a1.y * (i1 - cM.y));
        o[2] = c1.z * (dv2 - a0.z - a1.z * (i2 - cM.z));
        o[3] = c1.w * (dv3 - a0.w - a1.w * (i3 - cM.w));
      }
      if (a.dz_format == 0) {
        reinterpret_cast<float4*>(a.dz)[i] = make_float4(o[0], o[1], o[2], o[3]);
      } else {
        // bf16 hi + lo planes: hi = bf16(dz), lo = bf16(dz - hi)  (dz - hi is exact in fp32)
        __nv_bfloat16 h[4], l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          h[e] = __float2bfloat16_rn(o[e]);
          l[e] = __float2bfloat16_rn(o[e] - __bfloat162float(h[e]));
        }
        uint2 hv, lv;
        hv.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
        hv.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
        lv.x = (uint32_t)__bfloat16_as_ushort(l[0]) | ((uint32_t)__bfloat16_as_ushort(l[1]) << 16);
        lv.y = (uint32_t)__bfloat16_as_ushort(l[2]) | ((uint32_t)__bfloat16_as_ushort(l[3]) << 16);
        reinterpret_cast<uint2*>(a.dz)[i] = hv;
        reinterpret_cast<uint2*>(a.dz_lo)[i] = lv;
      }
    }
  }
}

}  // namespace frost

using namespace frost;

extern "C" int frost_bn_finalize(const FrostBnFinalizeArgs* a, void* stream) {
  FROST_REQUIRE(a && a->stats && a->x_scale && a->w_scale && a->sf && a->gamma && a->beta && a->running_mean &&
                    a->running_var && a->A && a->B && a->mean_I && a->kfac && a->cur_minmax && a->afq.scale &&
                    a->afq.zero_point && a->afq.min_val && a->afq.max_val,
                "frost_bn_finalize: null pointer");
  FROST_REQUIRE(a->C > 0 && a->count > 0, "frost_bn_finalize: empty tensor");
  launch_pdl(bn_finalize_kernel, dim3(1), dim3(a->C >= 512 ? 1024 : (a->C >= 128 ? 256 : 64)), 0, (cudaStream_t)stream, *a);
  FROST_LAUNCH_CHECK("bn_finalize");
  return FROST_OK;
}

extern "C" int frost_bnq_apply(const int32_t* acc, int acc_format, int64_t M, int C, const float* A, const float* B, int relu,
                               const float* out_scale, const int32_t* out_zp, uint8_t* q, int ldq, void* stream) {
  FROST_REQUIRE(acc && A && B && out_scale && out_zp && q, "frost_bnq_apply: null pointer");
  FROST_REQUIRE(M > 0 && C > 0 && C % 4 == 0, "frost_bnq_apply: C=%d must be a positive multiple of 4", C);
  FROST_REQUIRE(ldq >= C && ldq % 4 == 0, "frost_bnq_apply: ldq=%d must be >= C and a multiple of 4", ldq);
  const int64_t n4 = M * C / 4;
  const size_t smem = 2 * C * sizeof(float);
  // one resident wave, U loads in flight per thread: bytes in flight per SM is what the HBM pipe needs
  const int u = tunable(FROST_TUNE_BNQ_UNROLL);
#define LAUNCH_BNQ(U)                                                                                                      \
  do {                                                                                                                     \
    static int per_sm = 0;                                                                                                 \
    if (!per_sm && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bnq_apply_kernel<U>, 256, 16 * 1024) != cudaSuccess) \
      per_sm = 4;                                                                                                          \
    launch_pdl(bnq_apply_kernel<U>, dim3(grid_for(n4, 256 * U, kNumSMs * std::max(per_sm, 1))), dim3(256), smem,         \
               (cudaStream_t)stream, acc, acc_format, n4, C, A, B, relu, out_scale, out_zp, q, ldq);                            \
  } while (0)
  if (u >= 8) LAUNCH_BNQ(8);
  else if (u >= 4) LAUNCH_BNQ(4);
  else LAUNCH_BNQ(2);
#undef LAUNCH_BNQ
  FROST_LAUNCH_CHECK("bnq_apply");
  return FROST_OK;
}

static int bn_backward_check(const FrostBnBackwardArgs* a) {
  FROST_REQUIRE(a && a->dy && a->acc && a->A && a->B && a->mean_I && a->kfac && a->gamma && a->sf && a->x_scale &&
                    a->w_scale && a->out_scale && a->out_zp && a->sums && a->coef && a->dz && a->dgamma_bn &&
                    a->dbeta && a->dsf_bn,
                "frost_bn_backward: null pointer");
  FROST_REQUIRE(a->M > 0 && a->C > 0 && a->C % 4 == 0, "frost_bn_backward: C=%d must be a positive multiple of 4", a->C);
  FROST_REQUIRE(a->dz_format == 0 || (a->dz_format == 1 && a->dz_lo), "frost_bn_backward: dz_format 1 needs dz_lo");
  FROST_REQUIRE(6 * (size_t)a->C * sizeof(float) <= 96 * 1024, "frost_bn_backward: C=%d too large for the coefficient tile", a->C);
  return FROST_OK;
}

extern "C" int frost_bn_backward_reduce(const FrostBnBackwardArgs* a, void* stream) {
  int rc = bn_backward_check(a);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(a->sums, 0, sizeof(double) * 2 * a->C, st) != cudaSuccess) {
    set_error("frost_bn_backward: memset failed");
    return FROST_ECUDA;
  }
  int cgb, chunks, threads;
  dw_launch_shape(a->C, std::min(64, tunable(FROST_TUNE_BN_RED_MAX_CGB)), &cgb, &chunks, &threads);
  const int rows_per_block = threads / cgb;
  const int64_t wave = std::max<int64_t>(1, (int64_t)kNumSMs * tunable(FROST_TUNE_BN_RED_CTAS_PER_SM) / chunks);
  int gx = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(a->M, (int64_t)rows_per_block * 8), wave));
  const size_t smem = sizeof(double) * threads * 8 + sizeof(float4) * 3 * cgb;
  if (tunable(FROST_TUNE_BN_RED_UNROLL) >= 8)
    launch_pdl(bn_bwd_reduce_kernel<8, 2>, dim3(gx, chunks), dim3(threads), smem, st, *a, cgb);
  else if (tunable(FROST_TUNE_BN_RED_CTAS_PER_SM) >= 4)
    launch_pdl(bn_bwd_reduce_kernel<4, 4>, dim3(gx, chunks), dim3(threads), smem, st, *a, cgb);
  else
    launch_pdl(bn_bwd_reduce_kernel<4, 3>, dim3(gx, chunks), dim3(threads), smem, st, *a, cgb);
  FROST_LAUNCH_CHECK("bn_bwd_reduce");
  return FROST_OK;
}

extern "C" int frost_bn_backward_apply(const FrostBnBackwardArgs* a, void* stream) {
  int rc = bn_backward_check(a);
  if (rc) return rc;
  const int64_t n4 = a->M * a->C / 4;
  const size_t smem = 6 * (size_t)a->C * sizeof(float);
  const int u = tunable(FROST_TUNE_BN_APPLY_UNROLL);
#define LAUNCH_APPLY(U)                                                                                                    \
  do {                                                                                                                     \
    static int per_sm = 0;                                                                                                 \
    if (first_use_on_device(reinterpret_cast<const void*>(&bn_bwd_apply_kernel<U>)))                                       \
      cudaFuncSetAttribute(bn_bwd_apply_kernel<U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);                \
    if (!per_sm) {                                                                                                         \
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, bn_bwd_apply_kernel<U>, 256, 24 * 1024) != cudaSuccess)   \
        per_sm = 3;                                                                                                        \
    }                                                                                                                      \
    launch_pdl(bn_bwd_apply_kernel<U>, dim3(grid_for(n4, 256 * U, kNumSMs * std::max(per_sm, 1))), dim3(256), smem,      \
               (cudaStream_t)stream, *a, n4);                                                                              \
  } while (0)
  if (u >= 4) LAUNCH_APPLY(4);
  else if (u >= 2) LAUNCH_APPLY(2);
  else LAUNCH_APPLY(1);
#undef LAUNCH_APPLY
  FROST_LAUNCH_CHECK("bn_bwd_apply");
  return FROST_OK;
}

extern "C" int frost_bn_backward(const FrostBnBackwardArgs* a, void* stream) {
  int rc = frost_bn_backward_reduce(a, stream);
  if (rc) return rc;
  return frost_bn_backward_apply(a, stream);
}
