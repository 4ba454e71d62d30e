// optim.cu - GradBoost optimizers (optimizer.py:121-206 QSGD, :264-359 QRMSprop, :411-512 QAdam,
// :564-667 QAdamW) as one multi-tensor kernel: sensitivity EMA (exp_min/exp_max with the
// reference's in-place bias-correction division), sign-aligned coin-tossed clipped |Laplace| boost,
// weight decay and the base update - one pass over (p, g, state), one CTA per 2048-element chunk.
// The reference draws its noise on the host (numpy, optimizer.py:178) and copies it to the GPU for
// every tensor; here it is a counter-based Philox4x32-10 stream, or injected arrays for parity tests.
#include "common.cuh"

namespace frost {

struct Philox {
  static constexpr uint32_t kM0 = 0xD2511F53u, kM1 = 0xCD9E8D57u, kW0 = 0x9E3779B9u, kW1 = 0xBB67AE85u;
  __device__ static uint4 gen(uint4 ctr, uint2 key) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
      const uint32_t hi0 = __umulhi(kM0, ctr.x), lo0 = kM0 * ctr.x;
      const uint32_t hi1 = __umulhi(kM1, ctr.z), lo1 = kM1 * ctr.z;
      ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
      key.x += kW0;
      key.y += kW1;
    }
    return ctr;
  }
};

struct HyperDev {
  FrostOptHyper h;
};

__global__ void __launch_bounds__(256) gradboost_kernel(const FrostOptTensor* tensors,
                                                       const FrostOptChunk* chunks, HyperDev hd) {
  const FrostOptHyper& h = hd.h;
  const FrostOptChunk ck = chunks[blockIdx.x];
  const FrostOptTensor t = tensors[ck.tensor];
  const int64_t base = (int64_t)ck.chunk * FROST_OPT_CHUNK;
  const float beta_s = (h.kind == FROST_OPT_QADAM || h.kind == FROST_OPT_QADAMW) ? h.beta1 : h.beta;
  const float bc1 = (float)(1.0 - pow((double)beta_s, (double)t.step));
  const float one_m_beta = (float)(1.0 - (double)beta_s);
  const float noise_scale = (float)pow(1.0 - (double)h.noise_decay, (double)t.restart_step);
  const float lr = t.lr, wd = t.weight_decay;
  // Adam-family scalars
  const double bc1_adam = 1.0 - pow((double)h.beta1, (double)t.step);
  const double bc2_adam = 1.0 - pow((double)h.beta2, (double)t.step);
  const float step_size = (float)((double)lr / bc1_adam);
  const float sqrt_bc2 = (float)sqrt(bc2_adam);
  const float one_m_b1 = (float)(1.0 - (double)h.beta1), one_m_b2 = (float)(1.0 - (double)h.beta2);
  const float one_m_alpha = (float)(1.0 - (double)h.alpha);

  for (int64_t i = base + threadIdx.x; i < min(t.n, base + FROST_OPT_CHUNK); i += blockDim.x) {
    float p = t.p[i];
    float g = t.g[i];
    if (h.grad_scale != 1.0f) g *= h.grad_scale;
    if (h.kind == FROST_OPT_QADAMW) p = __fmul_rn(p, (float)(1.0 - (double)lr * (double)wd));   // optimizer.py:580
    if (h.kind == FROST_OPT_QADAM && wd != 0.0f) g = fmaf(wd, p, g);                            // :466-467
    // sensitivity statistics (optimizer.py:165-168)
    const float ag = fabsf(g);
    float emin = t.exp_min[i], emax = t.exp_max[i];
    const float new_min = fminf(emin, ag), new_max = fmaxf(emax, ag);
    emin = __fdiv_rn(fmaf(one_m_beta, new_min, __fmul_rn(emin, beta_s)), bc1);
    emax = __fdiv_rn(fmaf(one_m_beta, new_max, __fmul_rn(emax, beta_s)), bc1);
    t.exp_min[i] = emin;
    t.exp_max[i] = emax;
    if (!h.is_warmup) {  // boost (optimizer.py:170-189)
      float noise, coin = 1.0f;
      if (t.noise) {
        noise = t.noise[i];
        if (h.toss_coin) coin = t.coin[i];
      } else {
        const uint4 r = Philox::gen(make_uint4((uint32_t)i, (uint32_t)(i >> 32), (uint32_t)ck.tensor, (uint32_t)t.step),
                                    make_uint2((uint32_t)h.seed, (uint32_t)(h.seed >> 32)));
        const float u = ((float)r.x + 0.5f) * 2.3283064365386963e-10f;  // (0,1]
        noise = -logf(u);                                               // |Laplace(0,1)| == Exp(1)
        coin = (float)(r.y >> 31);
      }
      const float sens = __fmul_rn(__fsub_rn(emax, emin), noise_scale);
      noise = __fmul_rn(noise, sens);
      if (h.toss_coin) {
        t.coin_toss[i] = coin;
        noise = __fmul_rn(noise, coin);
      }
      const float sgn = (g > 0.0f) ? 1.0f : ((g < 0.0f) ? -1.0f : 0.0f);
      noise = __fmul_rn(noise, sgn);
      if (h.clip_by > 0.0f) noise = fminf(fmaxf(noise, -h.clip_by), h.clip_by);
      g = __fadd_rn(g, noise);
    }
    if (h.kind == FROST_OPT_QSGD) {
      if (wd != 0.0f) g = fmaf(wd, p, g);
      float d = g;
      if (h.momentum != 0.0f) {
        float buf;
        if (t.first_momentum) buf = g;
        else buf = fmaf((float)(1.0 - (double)h.dampening), g, __fmul_rn(t.buf0[i], h.momentum));
        t.buf0[i] = buf;
        d = h.nesterov ? fmaf(h.momentum, buf, g) : buf;
      }
      t.g[i] = g;
      t.p[i] = fmaf(-lr, d, p);
    } else if (h.kind == FROST_OPT_QRMS) {
      t.g[i] = g;                                            // reference: grad.add(wd, p) is out of place here
      const float g2 = (wd != 0.0f) ? fmaf(wd, p, g) : g;
      float sq = __fadd_rn(__fmul_rn(t.buf0[i], h.alpha), __fmul_rn(__fmul_rn(one_m_alpha, g2), g2));  // addcmul: value*t1*t2
      t.buf0[i] = sq;
      float avg;
      if (h.centered) {
        float ga = fmaf(one_m_alpha, g2, __fmul_rn(t.buf2[i], h.alpha));
        t.buf2[i] = ga;
        avg = __fadd_rn(__fsqrt_rn(fmaf(-ga, ga, sq)), h.eps);
      } else {
        avg = __fadd_rn(__fsqrt_rn(sq), h.eps);
      }
      if (h.momentum > 0.0f) {
        const float buf = __fadd_rn(__fmul_rn(t.buf1[i], h.momentum), __fdiv_rn(g2, avg));
        t.buf1[i] = buf;
        t.p[i] = fmaf(-lr, buf, p);
      } else {
        t.p[i] = fmaf(-lr, __fdiv_rn(g2, avg), p);
      }
    } else {  // QAdam / QAdamW (optimizer.py:497-510)
      t.g[i] = g;
      const float m = fmaf(one_m_b1, g, __fmul_rn(t.buf0[i], h.beta1));
      const float v = __fadd_rn(__fmul_rn(t.buf1[i], h.beta2), __fmul_rn(__fmul_rn(one_m_b2, g), g));
      t.buf0[i] = m;
      t.buf1[i] = v;
      float vv = v;
      if (h.amsgrad) {
        vv = fmaxf(t.buf2[i], v);
        t.buf2[i] = vv;
      }
      const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv), sqrt_bc2), h.eps);
      t.p[i] = fmaf(-step_size, __fdiv_rn(m, denom), p);
    }
  }
}

}  // namespace frost

using namespace frost;

extern "C" int frost_gradboost_multi(const FrostOptTensor* tensors, int n, const FrostOptChunk* chunks, int n_chunks,
                                     const FrostOptHyper* hyper, void* stream) {
  FROST_REQUIRE(tensors && chunks && hyper && n > 0 && n_chunks > 0, "frost_gradboost_multi: bad args");
  FROST_REQUIRE(hyper->kind >= FROST_OPT_QSGD && hyper->kind <= FROST_OPT_QADAMW, "frost_gradboost_multi: unknown kind %d",
                hyper->kind);
  HyperDev hd;
  hd.h = *hyper;
  gradboost_kernel<<<n_chunks, 256, 0, (cudaStream_t)stream>>>(tensors, chunks, hd);
  FROST_LAUNCH_CHECK("gradboost");
  return FROST_OK;
}
