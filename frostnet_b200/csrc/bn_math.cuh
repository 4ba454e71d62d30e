// bn_math.cuh - the per-channel / per-element arithmetic of BatchNorm(+ReLU)+activation fake-quant on top of
// the integer accumulators I, shared by the stand-alone passes (bn.cu) and the fused tensor-core kernels
// (pw_fused.cu) so that both produce bit-identical results.
//   conv_orig = conv/scale_factor = I * m_c,  m_c = s_a*s_w/scale_factor_c
//   v = bn(conv_orig) = A_c*I + B_c          (training: batch statistics from exact integer sums)
//   y = FQ_a(relu(v))
// Restates torch/ao/nn/intrinsic/qat/modules/conv_fused.py:156-167 (+ :708-710 ReLU) and the
// activation_post_process hook (fake_quantize.py:423-438).
#pragma once
#include "common.cuh"

namespace frost {

// raw conv output of one element as a float: int32 accumulator (format 0) or fp32 bits (format 1)
__device__ __forceinline__ float acc_val(int raw, int fmt) { return fmt ? __int_as_float(raw) : (float)raw; }
__device__ __forceinline__ float bn_affine(float I, float A, float B) { return fmaf(I, A, B); }

// ---------------------------------------------------------------- per-channel finalize
struct BnChannel {
  float A, B;        // v = A*I + B
  float mean_I;      // batch mean of I (eval mode: running_mean / m_c, so that xhat = (I - mean_I) * kfac either way)
  float kfac;        // m_c * invstd
  float v_lo, v_hi;  // extrema of relu(v) over the channel (v is monotone in I)
  float new_running_mean, new_running_var;   // training only
};

__device__ __forceinline__ BnChannel bn_channel_finalize(const FrostChanStats& st, int stats_format, double M, int count_gt1,
                                                         double sa_sw, float sf, float gamma, float beta, float run_mean,
                                                         float run_var, float eps, double mom, int training, int relu) {
  BnChannel r;
  const double sum = stats_format ? __longlong_as_double(st.sum) : (double)st.sum;
  const double sq = stats_format ? __longlong_as_double((long long)st.sq_lo)
                                 : (double)st.sq_hi * 4294967296.0 + (double)st.sq_lo;
  const double mean_I = sum / M;
  double var_I = sq / M - mean_I * mean_I;
  if (var_I < 0.0) var_I = 0.0;
  const double m_c = sa_sw / (double)sf;
  double mean_u, invstd;
  r.new_running_mean = run_mean;
  r.new_running_var = run_var;
  if (training) {
    mean_u = m_c * mean_I;
    const double var_u = m_c * m_c * var_I;
    invstd = 1.0 / sqrt(var_u + (double)eps);
    const double unbiased = count_gt1 ? var_u * (M / (M - 1.0)) : var_u;
    r.new_running_mean = (float)((1.0 - mom) * (double)run_mean + mom * mean_u);
    r.new_running_var = (float)((1.0 - mom) * (double)run_var + mom * unbiased);
    r.mean_I = (float)mean_I;
  } else {
    mean_u = (double)run_mean;
    invstd = 1.0 / sqrt((double)run_var + (double)eps);
    r.mean_I = (float)(mean_u / m_c);
  }
  const double g = (double)gamma;
  r.A = (float)(m_c * invstd * g);
  r.B = (float)((double)beta - mean_u * invstd * g);
  r.kfac = (float)(m_c * invstd);
  // v is monotone in I for fixed (A,B): the channel extrema of v sit at the integer extrema.
  float v0 = bn_affine(acc_val(st.min, stats_format), r.A, r.B), v1 = bn_affine(acc_val(st.max, stats_format), r.A, r.B);
  if (relu) { v0 = fmaxf(v0, 0.0f); v1 = fmaxf(v1, 0.0f); }
  r.v_lo = fminf(v0, v1);
  r.v_hi = fmaxf(v0, v1);
  return r;
}

// MovingAverageMinMaxObserver step as a pure function (common.cuh::observer_update writes the state in place)
__device__ __forceinline__ void observer_ema(float& rmin, float& rmax, float cur_min, float cur_max, float c) {
  if (isinf(rmin) || isinf(rmax)) {
    rmin = cur_min;
    rmax = cur_max;
  } else {
    rmin = __fadd_rn(rmin, __fmul_rn(c, __fsub_rn(cur_min, rmin)));
    rmax = __fadd_rn(rmax, __fmul_rn(c, __fsub_rn(cur_max, rmax)));
  }
}

// ---------------------------------------------------------------- per-element forward / backward
// round-to-nearest-even + saturate to [0, 255] in one conversion (NaN -> 0)
__device__ __forceinline__ unsigned cvt_rni_sat_u8(float t) {
  unsigned q;
  asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(q) : "f"(t));
  return q;
}

// q = clamp(rint(relu(A*I + B) * inv) + zp, 0, 255)   - the arithmetic of ATen's fake-quantize kernel, step by step.
// zp == 0 (every post-ReLU tensor: its observed minimum is 0) needs neither the add nor the explicit ReLU / clamp:
// negative values saturate to 0 exactly like relu() followed by the clamp, and rint-then-clamp == saturating rint.
__device__ __forceinline__ unsigned bnq1(float I, float A, float B, int relu, float inv, float zp) {
  float r = bn_affine(I, A, B);
  if (zp == 0.0f) return cvt_rni_sat_u8(__fmul_rn(r, inv));
  if (relu) r = fmaxf(r, 0.0f);
  return (unsigned)fminf(fmaxf(fq_index(r, inv, zp), 0.0f), 255.0f);
}

// dv = dy * [0 <= idx <= 255] * [v > 0 if relu]
__device__ __forceinline__ float bn_dv(float dy, float I, float A, float B, int relu, float inv, float zp) {
  const float v = bn_affine(I, A, B);
  const float r = relu ? fmaxf(v, 0.0f) : v;
  const float idx = fq_index(r, inv, zp);
  const bool pass = (idx >= 0.0f) && (idx <= 255.0f) && (!relu || v > 0.0f);
  return pass ? dy : 0.0f;
}

// The STE / ReLU mask as an INTEGER INTERVAL of the accumulator.  Every step of
//   I -> v = fma(I, A, B) -> relu -> * inv -> rint -> + zp
// is monotone in I (non-decreasing for A >= 0, non-increasing for A < 0; fp rounding preserves monotonicity), so
//   pass(I) = [0 <= idx <= 255] && [v > 0 if relu]        (exactly bn_dv's predicate)
// holds on one interval [lo, lo + width) of integers.  The two ends are found by bisection on the exact predicate over
// |I| <= 2^28 (accumulators are bounded by 255*128*K), once per channel; the per-element test is then one subtract and
// one unsigned compare instead of ~10 float instructions.
struct MaskInterval {
  int lo;
  unsigned width;     // pass(I) <=> (unsigned)(I - lo) < width ; 0: never
};
__device__ __forceinline__ bool mask_passes(const MaskInterval& m, int I) { return (unsigned)(I - m.lo) < m.width; }

// Smallest I in [a, b] for which the monotone (false ... false true ... true) predicate holds, b if it never does.  `est`
// is a float guess of the answer: when the predicate confirms a small bracket around it, the bisection runs inside that
// bracket (~8 evaluations instead of ~29 over the whole 2^29 range - the search sits on the critical path of every small
// launch); otherwise the whole range is searched.  Either way the result is the exact first-true point.
template <typename Pred>
__device__ __forceinline__ int first_true_near(Pred pred, float est, int a, int b) {
  if (est >= (float)a && est <= (float)b) {            // false for NaN / inf / out of range: whole-range search
    const int c = (int)rintf(est);
    const int w = 8 + (abs(c) >> 17);                  // float error of the guess grows with |I|
    const int lo = c - w, hi = c + w;
    if (lo > a && !pred(lo)) a = lo + 1;
    if (hi < b && pred(hi)) b = hi;
  }
  while (a < b) {
    const int mid = a + ((b - a) >> 1);
    if (pred(mid)) b = mid; else a = mid + 1;
  }
  return a;
}

__device__ inline MaskInterval bn_mask_interval(float A, float B, int relu, float inv, float zp) {
  constexpr int LIM = 1 << 28;
  const bool inc = A >= 0.0f;
  // low side of the index range (and v > 0) / high side of the index range
  auto ok_low = [&](int I) {
    const float v = bn_affine((float)I, A, B);
    const float r = relu ? fmaxf(v, 0.0f) : v;
    return fq_index(r, inv, zp) >= 0.0f && (!relu || v > 0.0f);
  };
  auto ok_high = [&](int I) {
    const float v = bn_affine((float)I, A, B);
    const float r = relu ? fmaxf(v, 0.0f) : v;
    return fq_index(r, inv, zp) <= 255.0f;
  };
  // where v crosses the two thresholds, in units of I (guesses only)
  float t_low = __fdividef(-zp - 0.5f, inv);
  if (relu) t_low = fmaxf(t_low, 0.0f);
  const float t_high = __fdividef(255.5f - zp, inv);
  const float rA = __fdividef(1.0f, A);
  const float e_low = (t_low - B) * rA, e_high = (t_high - B) * rA;
  // left end: first I where the condition that fails for very small I holds (A >= 0: ok_low, else ok_high)
  const int first = inc ? first_true_near(ok_low, e_low, -LIM, LIM + 1) : first_true_near(ok_high, e_high, -LIM, LIM + 1);
  // right end: first I >= first where the other condition fails
  const int last = inc ? first_true_near([&](int I) { return !ok_high(I); }, e_high, first, LIM + 1)
                       : first_true_near([&](int I) { return !ok_low(I); }, e_low, first, LIM + 1);
  MaskInterval m;
  m.lo = first;
  m.width = (unsigned)(last - first);
  return m;
}

// Per-channel coefficients of  dz = c1*(dv - a0 - a1*(I - mean_I))  and the BN parameter gradients, from the reduced
// sums S1 = sum dv, S2 = sum dv*(I - mean_I).  training == 0 (frozen BatchNorm: eval-mode statistics are constants):
// dz = c1*dv, no mean / variance terms.
struct BnBwdChannel {
  float c1, a0, a1;
  float dgamma_bn, dbeta, dsf_bn;
};
__device__ __forceinline__ BnBwdChannel bn_bwd_channel(double S1, double S2, double M, double sa_sw, float A, float kfac,
                                                       float mean_I, float gamma, float sf, float eps, int training) {
  BnBwdChannel r;
  const double k = (double)kfac;
  const double T = k * S2;                                // sum dv*xhat (S2 is already centred)
  r.c1 = (float)((double)A / sa_sw);                      // gamma*invstd/sf
  r.dgamma_bn = (float)T;
  r.dbeta = (float)S1;
  const double g = (double)gamma, sfd = (double)sf;
  const double invstd = k * sfd / sa_sw;                  // k = m_c*invstd, m_c = sa_sw/sf
  if (training) {
    r.a0 = (float)(S1 / M);
    r.a1 = (float)(k * T / M);
    r.dsf_bn = (float)(-g * T * (double)eps * invstd * invstd / sfd);
  } else {
    r.a0 = 0.0f;
    r.a1 = 0.0f;
    // u = z/sf with constant statistics: dL/dsf = -(gamma*invstd/sf) * sum dv*u,  sum dv*u = T/invstd + mean_u*S1,
    // mean_u = mean_I*m_c (bn_channel_finalize stores running_mean/m_c as mean_I in eval mode)
    const double mean_u = (double)mean_I * sa_sw / sfd;
    r.dsf_bn = (float)(-g * invstd / sfd * (T / invstd + mean_u * S1));
  }
  return r;
}

}  // namespace frost
