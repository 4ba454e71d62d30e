/* frost_b200.h - C ABI of the B200-native FrostNet QAT hot path (libfrost_b200.so).
 *
 * The reference (clovaai/frostnet) has NO FFI: it is pure Python wiring of torch.ao.quantization
 * eager-mode QAT.  The boundary this library replaces is therefore the set of ATen ops that
 * the reference dispatches to for one QAT training step.  Every entry
 * point cites the reference call site / torch module it stands in for.  All pointers are raw
 * DEVICE pointers borrowed for the duration of the call (PyTorch owns every tensor); sizes are
 * plain integers; `stream` is a cudaStream_t passed as void*.  No torch types cross the boundary.
 * Every function returns 0 on success or a negative FROST_E* code; frost_last_error() gives text.
 *
 * Data layout in HBM (the B200-first part - see DESIGN.md):
 *   - activations: NHWC uint8 quantize INDICES q (value = (q - zero_point) * scale), 1 B/element;
 *   - raw conv accumulators: NHWC int32  I = sum (q_a - zp_a) * (q_w - zp_w)  (exact integers);
 *   - weights: int8 indices in a kernel-friendly layout (+ STE mask in the PyTorch layout);
 *   - gradients: NHWC fp32.
 */
#ifndef FROST_B200_H_
#define FROST_B200_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FROST_OK 0
#define FROST_EINVAL (-1) /* bad argument (shape, alignment, null pointer) */
#define FROST_ECUDA (-2)  /* a CUDA runtime call or launch failed           */
#define FROST_ENOSUP (-3) /* configuration not supported by this build      */

/* ABI version; bumped on any signature change (2: pw_fused entry points, FrostWeightDesc.wq_mma, FrostBnBackwardArgs.frozen). */
int frost_abi_version(void);
/* Text of the last error on the calling thread ("" if none). */
const char* frost_last_error(void);
/* Number of kernels this library has launched since load (all threads); bench.py's gpu_launches. */
int64_t frost_launch_count(void);

/* Launch-shape knobs for measurement (tools/microbench_ops.py sweeps them on the device; results never depend
 * on them).  value 0 restores the built-in default.  Not part of the reference-facing surface. */
enum {
  FROST_TUNE_DW_FWD_CTAS_PER_SM = 0,   /* depthwise forward: resident CTAs per SM (grid = 148 * value)      */
  FROST_TUNE_DW_WGRAD_CTAS_PER_SM = 1, /* depthwise wgrad                                                    */
  FROST_TUNE_BN_RED_CTAS_PER_SM = 2,   /* bn_backward_reduce                                                 */
  FROST_TUNE_BN_RED_MAX_CGB = 3,       /* bn_backward_reduce: max 4-channel groups per CTA (channel chunking) */
  FROST_TUNE_STEM_FWD_CTAS_PER_SM = 4, /* stem forward (persistent over pixel tiles)                         */
  FROST_TUNE_STEM_WGRAD_CTAS_PER_SM = 5,
  FROST_TUNE_DW_DGRAD_CTAS_PER_SM = 6,
  FROST_TUNE_PDL = 7,                  /* 1: programmatic dependent launch between the per-layer kernels, 2: off */
  FROST_TUNE_BN_RED_UNROLL = 8,        /* bn_backward_reduce: rows in flight per thread (4 or 8)               */
  FROST_TUNE_BN_APPLY_UNROLL = 9,      /* bn_backward_apply: elements in flight per thread (1, 2 or 4)          */
  FROST_TUNE_BNQ_UNROLL = 10,          /* bnq_apply: 16-byte loads in flight per thread (2, 4 or 8)             */
  FROST_TUNE_DW_FWD_TILED = 11,        /* depthwise forward smem tiles: 1 wide stride-1 planes only, 2 never, 3 always */
  FROST_TUNE_DW_DGRAD_TILED = 12,      /* depthwise dgrad, stride 1: 1 shared-memory tiles filled by TMA, 2 the gather kernel  */
  FROST_TUNE_COUNT = 13
};
int frost_set_tunable(int which, int value);
int frost_get_tunable(int which);
/* Measurement aid, only in a library built with -DFROST_TRACE (python -m frostnet_b200.build --force --trace; FROST_ENOSUP
 * otherwise): while `device_stamps` (>= 32 int64 on the device) is set, CTA (0,0) of every frost_pw_fused_* launch writes
 * SM-clock stamps of its timeline there (tools/trace_fused.py explains the indices).  NULL switches it off. */
int frost_debug_set_trace(long long* device_stamps);

/* One FusedMovingAvgObsFakeQuantize instance (torch/ao/quantization/fake_quantize.py:423-438):
 * the four state buffers of the module, on the device.  Kernels read AND update them in place
 * (observer EMA, then qparams), exactly the side effects of torch.fused_moving_avg_obs_fake_quant. */
typedef struct {
  float* min_val;      /* activation_post_process.min_val  (scalar, +inf initially) */
  float* max_val;      /* activation_post_process.max_val  (scalar, -inf initially) */
  float* scale;        /* scale[1]       */
  int32_t* zero_point; /* zero_point[1]  */
} FrostFQ;

/* Per-channel integer statistics of one conv output (filled by the conv kernels with integer
 * atomics: exact and order-independent, hence deterministic).  32 bytes per channel. */
typedef struct {
  long long sum;            /* sum I                                  */
  unsigned long long sq_lo; /* sum (I*I) & 0xffffffff                 */
  unsigned long long sq_hi; /* sum (I*I) >> 32                        */
  int32_t min;              /* min I  (INT32_MAX when empty)          */
  int32_t max;              /* max I  (INT32_MIN when empty)          */
} FrostChanStats;

/* Reset `n` channel-stat records (sum=0, min=INT_MAX, max=INT_MIN). */
int frost_stats_reset(FrostChanStats* stats, int64_t n, void* stream);
/* Same record reused for a conv whose raw output is fp32 (the feature backbone's stem, which has no QuantStub
 * in front of it - frostnet_features.py:342-343): sum / sq_lo hold doubles (sum z, sum z^2), min / max hold
 * float bits.  Reset: sums 0, min +inf, max -inf. */
int frost_stats_reset_f32(FrostChanStats* stats, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Generic per-tensor fake-quant on fp32 data (logits of classifier.2, unit tests).
 * Replaces torch.fused_moving_avg_obs_fake_quant as called from
 * torch/ao/quantization/fake_quantize.py:423-438 (reference call sites: frostnet.py:305-306 via
 * prepare_qat, Classification/train.py:168-173).
 *   observe!=0 : aminmax(x) -> EMA (first call copies; else m += 0.01*(cur-m)) -> qparams
 *   y    = (clamp(rint(x*(1/scale)) + zp, qmin, qmax) - zp) * scale
 *   mask = 1 iff qmin <= rint(x*(1/scale)) + zp <= qmax      (STE mask; may be NULL)
 *   q    = clamped index as int32 (may be NULL)
 * scratch: FROST_FQ_SCRATCH_FLOATS floats of device scratch. */
#define FROST_FQ_SCRATCH_FLOATS 2048
int frost_fq_forward(const float* x, int64_t n, FrostFQ fq, int qmin, int qmax, int symmetric,
                     int observe, float averaging_const, float* y, uint8_t* mask, int32_t* q,
                     float* scratch, void* stream);
/* dx = dy * mask   (backward of the cachemask fake-quant kernel). */
int frost_fq_backward(const float* dy, const uint8_t* mask, int64_t n, float* dx, void* stream);

/* QuantStub (frostnet.py:304,320): observe + quantise the fp32 NCHW input image to NHWC uint8.
 * cur_minmax[2] receives the dequantised min/max of the produced tensor. */
int frost_input_quant(const float* x_nchw, int N, int C, int H, int W, FrostFQ fq, int observe,
                      float averaging_const, uint8_t* q_nhwc, float* cur_minmax, float* scratch,
                      void* stream);

/* QuantStub fused with the im2col of the dense kxk stem conv (frostnet.py:277,320): q[N*Ho*Wo][ld] holds, per OUTPUT pixel, the
 * k*k*C quantize indices of its receptive field in (kh, kw, c) order (taps outside the image = zero point), so that the
 * stem runs on the fused 1x1 tensor-core kernels as a K = ld GEMM (FrostWeightDesc.wq_mma of a layout-2 conv uses the
 * same column order).  k*k*C <= ld, ld in {16, 32}. */
int frost_input_quant_im2col(const float* x_nchw, int N, int C, int H, int W, int k, int stride, int pad, FrostFQ fq,
                             int observe, float averaging_const, uint8_t* q, int ld, float* cur_minmax, float* scratch,
                             void* stream);

/* ---------------------------------------------------------------------------------------------
 * Weight side of nniqat.ConvBn(ReLU)2d / nnqat.Conv2d
 * (torch/ao/nn/intrinsic/qat/modules/conv_fused.py:131-146; frostnet.py:14-28,46-60 after
 * fuse_model): scale_factor = gamma/sqrt(running_var+eps); Ws = W*scale_factor;
 * observer+qparams (qint8, per_tensor_symmetric); q_w = clamp(rint(Ws/s_w)+zp_w).
 * `descs` lives in DEVICE memory (n of them). */
typedef struct {
  const float* weight;      /* [cout][cin_g][kh][kw] fp32, PyTorch layout                     */
  const float* bn_weight;   /* gamma[cout], or NULL (classifier: scale_factor = 1)            */
  const float* bn_var;      /* running_var[cout] (read BEFORE this step's update)             */
  float bn_eps;
  int32_t cout, cin_g, kh, kw;
  int32_t layout;           /* 0: wq[cout][cin_g*kh*kw] (1x1: [cout][cin]);
                               1: depthwise  wq[kh*kw][cout];
                               2: dense kxk  wq[cout][kh][kw][cin_g]                           */
  int32_t observe;
  float averaging_const;
  FrostFQ wfq;
  int8_t* wq;               /* out: int8 indices in `layout`                                   */
  int8_t* wq_mma;           /* out (layout 0 and 2; may be NULL): the same indices as the tensor-core A operand of the
                               fused kernels, rows `ldw` bytes apart (ldw % 16 == 0, padding stays zero).  zp_w == 0: the
                               s8 indices; zp_w == -128 / 127 (one-signed weights): q_w - zp_w resp. zp_w - q_w
                               as u8 (byte ^ 0x80 / ^ 0x7f).  Layout 2 (the stem as an im2col GEMM): columns in (kh, kw, cin)
                               order; wsum then counts the ldw - K padding columns as zero-point weights, and the backward
                               reads dwq as [cout][ldw]                                          */
  int32_t ldw;
  uint16_t* wt_bf16;        /* out (layout 0, 1x1 only; may be NULL): (q_w - zp_w) as bf16, transposed [cin][cout],
                               the B operand of the tensor-core dgrad                          */
  uint8_t* wmask;           /* out: STE mask, PyTorch layout                                   */
  float* sf;               /* out: scale_factor[cout]                                         */
  float* rstd_run;          /* out: 1/sqrt(running_var+eps)[cout] as seen by this forward       */
  int32_t* wsum;            /* out: per-cout sum_k q_w (raw index sum)                          */
  /* backward (frost_weight_backward_multi): */
  const float* dwq;         /* in : d/d(dequantised weight) in `layout`, fp32                  */
  const float* dgamma_bn;   /* in : [cout] sum dv*xhat   (NULL for classifier)                 */
  const float* dsf_bn;      /* in : [cout] d/d(scale_factor) through conv/scale_factor         */
  float* dweight;           /* out: grad of weight, PyTorch layout                             */
  float* dgamma;            /* out: grad of bn.weight [cout] (NULL for classifier)             */
} FrostWeightDesc;
/* Work list entry shared by the multi-tensor kernels: {tensor/layer index, chunk index within it}. */
typedef struct { int32_t tensor; int32_t chunk; } FrostOptChunk;
/* prep: `chunks` (device) lists every FROST_WEIGHT_CHUNK-element slice of every layer; `scratch` is
 * 2*n floats initialised to (+inf,-inf) pairs once by the caller (the kernels re-arm it).
 * backward: `chunks` lists every group of FROST_WEIGHT_BWD_CHANNELS output channels of every layer. */
#define FROST_WEIGHT_CHUNK 8192
#define FROST_WEIGHT_BWD_CHANNELS 8
int frost_weight_prep_multi(const FrostWeightDesc* descs, int n, const FrostOptChunk* chunks,
                            int n_chunks, float* scratch, void* stream);
int frost_weight_backward_multi(const FrostWeightDesc* descs, int n, const FrostOptChunk* chunks,
                                int n_chunks, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Convolutions on quantize indices (the F.conv2d at conv_fused.py:155, restated exactly in
 * integers: conv = s_a*s_w*I).  All write I (int32 NHWC) and accumulate FrostChanStats. */
/* 1x1 pointwise: xq[M][K] u8, wq[cout][K] s8 -> acc[M][cout]. K%8==0, cout%4==0, 16-byte aligned.
 * tcgen05.mma kind::i8 with TMEM accumulators (pw_conv_tc.cu); wsum[co] = sum_k q_w[co][k]. */
int frost_pw_conv_forward(const uint8_t* xq, const int32_t* x_zp, const int8_t* wq,
                          const int32_t* w_zp, const int32_t* wsum, int64_t M, int K, int cout,
                          int32_t* acc, FrostChanStats* stats, void* stream);
/* Same contract on CUDA cores (dp4a); kept as the cross-check of the tensor-core kernel. */
int frost_pw_conv_forward_simt(const uint8_t* xq, const int32_t* x_zp, const int8_t* wq,
                               const int32_t* w_zp, const int32_t* wsum, int64_t M, int K, int cout,
                               int32_t* acc, FrostChanStats* stats, void* stream);
/* depthwise kxk (k in {3,5}), stride in {1,2}, pad=(k-1)/2: xq[N][H][W][C] -> acc[N][Ho][Wo][C]. C%4==0.
 * ldx: bytes between consecutive input pixels (>= C; == C: dense) */
int frost_dw_conv_forward(const uint8_t* xq, int ldx, const int32_t* x_zp, const int8_t* wq,
                          const int32_t* w_zp, int N, int H, int W, int C, int k, int stride,
                          int32_t* acc, FrostChanStats* stats, void* stream);
/* dense kxk stem (frostnet.py:277: 3->32, 3x3, s2, p1): xq[N][H][W][cin] u8, wq[cout][k][k][cin]. */
int frost_stem_conv_forward(const uint8_t* xq, const int32_t* x_zp, const int8_t* wq,
                            const int32_t* w_zp, int N, int H, int W, int cin, int cout, int k,
                            int stride, int pad, int32_t* acc, FrostChanStats* stats, void* stream);

/* ---------------------------------------------------------------------------------------------
 * BN + ReLU + activation fake-quant of one fused conv (conv_fused.py:156-167, :708-710 and the
 * activation_post_process hook).  frost_bn_finalize turns the integer statistics into the
 * per-channel affine  v = A_c*I + B_c  (= bn(conv/scale_factor)), updates running stats
 * (momentum 0.1, unbiased var), derives the post-ReLU global min/max from the per-channel
 * integer extrema (v is monotone in I), runs the observer EMA + qparams, and saves what the
 * backward needs.  frost_bnq_apply then writes q = clamp(rint(relu(v)/s)+zp). */
typedef struct {
  const FrostChanStats* stats;
  int32_t C;
  int64_t count;               /* N*Ho*Wo                                                     */
  const float* x_scale;        /* scale of the conv input activation                          */
  const float* w_scale;        /* scale of the weight fake-quant                              */
  const float* sf;             /* scale_factor[C] from weight prep                            */
  const float* gamma;          /* bn.weight                                                   */
  const float* beta;           /* bn.bias                                                     */
  float* running_mean;
  float* running_var;
  int64_t* num_batches_tracked;
  float momentum, eps;
  int32_t stats_format;        /* 0: integer statistics of int32 accumulators; 1: fp32 raw conv output (see above) */
  int32_t training;            /* 1: batch statistics + running-stat update; 0: running stats */
  int32_t relu;
  int32_t observe;
  float averaging_const;
  FrostFQ afq;                 /* activation_post_process of the fused conv                   */
  float* A;                    /* out [C] */
  float* B;                    /* out [C] */
  float* mean_I;               /* out [C] batch mean of I (for backward)                      */
  float* kfac;                 /* out [C] k_c = m_c*invstd: xhat = (I-mean_I)*k_c             */
  float* cur_minmax;           /* out [2] dequantised min/max of the produced tensor          */
} FrostBnFinalizeArgs;
int frost_bn_finalize(const FrostBnFinalizeArgs* a, void* stream);
/* acc_format: 0 = int32 accumulators, 1 = fp32 raw conv output (bit pattern in the same buffer). */
int frost_bnq_apply(const int32_t* acc, int acc_format, int64_t M, int C, const float* A, const float* B, int relu,
                    const float* out_scale, const int32_t* out_zp, uint8_t* q, int ldq, void* stream);

/* Backward of the same chain (SURVEY.md 8a'): dy = grad wrt the fake-quantised output.
 * reduce: S1_c = sum dv, S2_c = sum dv*(I-mean_I)  with dv = dy*[0<=idx<=255]*[v>0]   (double[2*C], zeroed by callee)
 * apply : dz = grad wrt the real-valued conv output (what dgrad/wgrad consume), and
 *         dgamma_bn = sum dv*xhat, dbeta = S1, dsf_bn (see FrostWeightDesc). */
typedef struct {
  const float* dy;             /* [M][C] fp32                                                 */
  const int32_t* acc;          /* [M][C] int32 I saved by the forward (fp32 bits if acc_format 1) */
  int32_t acc_format;
  int64_t M;
  int32_t C;
  int32_t relu;
  const float* A;
  const float* B;
  const float* mean_I;
  const float* kfac;
  const float* gamma;
  const float* sf;
  const float* x_scale;        /* scale of the conv input activation                          */
  const float* w_scale;        /* scale of the weight fake-quant                              */
  const float* out_scale;
  const int32_t* out_zp;
  float eps;
  double* sums;                /* scratch [2*C]                                               */
  float* coef;                 /* scratch [3*C]                                               */
  float* dz;                   /* out [M][C] fp32 (dz_format 0)  |  bf16 hi plane [M][C] (dz_format 1)  */
  void* dz_lo;                 /* out bf16 lo plane [M][C] (dz_format 1): dz == hi + lo to 16 mantissa bits */
  int32_t dz_format;           /* 0: fp32 (depthwise / stem consumers); 1: bf16 hi+lo planes (tensor-core dgrad/wgrad) */
  float* dgamma_bn;            /* out [C]                                                     */
  float* dbeta;                /* out [C]                                                     */
  float* dsf_bn;               /* out [C]                                                     */
  int32_t frozen;              /* 1: the forward ran this BatchNorm in eval mode (running statistics are constants:
                                  nn.BatchNorm2d.eval() / frostnet_features.py:354-359 _freeze_stages): dz = c1*dv */
} FrostBnBackwardArgs;
int frost_bn_backward(const FrostBnBackwardArgs* a, void* stream);          /* reduce, then apply */
int frost_bn_backward_reduce(const FrostBnBackwardArgs* a, void* stream);   /* pass 1: the two per-channel sums */
int frost_bn_backward_apply(const FrostBnBackwardArgs* a, void* stream);    /* pass 2: dz + BN parameter grads  */

/* ---------------------------------------------------------------------------------------------
 * Fused 1x1 ConvBn(ReLU)2d (pw_fused.cu): the conv_fused.py:131-167,708-710 chain + activation fake-quant of one
 * squeeze / expand / reduce / last_layer conv (frostnet.py:98-119,293) as ONE launch, and the first two stages of its
 * autograd.  The int32 accumulator is never stored: every pass recomputes it on the tensor cores from the uint8 rows
 * (tcgen05.mma kind::i8; operands staged by TMA: ldx, ldw, ldq must be multiples of 16 bytes).
 *   forward     : stats (sum I, sum I^2, min, max per channel) -> grid barrier -> BN finalize + observer + qparams
 *                 (same arithmetic and side effects as frost_bn_finalize) -> q = clamp(rint(relu(A*I+B)/s)+zp)
 *   bwd_reduce  : S1 = sum dv, S2 = sum dv*(I-mean_I)            (== frost_bn_backward_reduce)
 *   bwd_apply   : dz as bf16 hi/lo planes + BN parameter grads    (== frost_bn_backward_apply, dz_format 1)
 * All CTAs of the forward launch must be co-resident (the callee sizes the grid to <= #SMs). */
typedef struct {
  const uint8_t* x;            /* [M][ldx] uint8 indices of the conv input (K valid bytes per row)          */
  int64_t M;
  int32_t K, ldx;              /* K % 8 == 0, ldx % 16 == 0                                                 */
  const int8_t* w_mma;         /* FrostWeightDesc.wq_mma [cout][ldw]                                        */
  int32_t ldw, cout;           /* cout % 4 == 0 (backward: cout % 8 == 0)                                   */
  const int32_t* x_zp;         /* zero point of x                                                           */
  const int32_t* w_zp;         /* zero point of the weight fake-quant (0, -128 or 127)                      */
  const int32_t* wsum;         /* FrostWeightDesc.wsum [cout]: sum_k q_w                                    */
} FrostPwOperands;
typedef struct {
  FrostPwOperands op;
  FrostBnFinalizeArgs bn;      /* stats: this layer's records, reset; C == cout, count == M                 */
  uint32_t* grid_barrier;      /* one zeroed uint32 in device memory per launch                             */
  uint8_t* q;                  /* out [M][ldq]                                                              */
  int32_t ldq;
} FrostPwFusedFwdArgs;
typedef struct {
  FrostPwOperands op;
  FrostBnBackwardArgs bn;      /* acc is ignored (recomputed); apply: dz / dz_lo planes with dz_format 1    */
} FrostPwFusedBwdArgs;
/* The whole backward of an expand conv (small K, wide cout) in one kernel (pw_chain.cu): BatchNorm-backward apply + STE /
 * ReLU mask + dgrad + wgrad; dz stays in shared memory as the tensor-core operand.  Run frost_pw_fused_bwd_reduce first.
 *   dx[M][K]   (+)= s_w * sum_co dz[m][co] * (q_w[co][k] - zp_w)
 *   dwq[cout][K] =  s_a * sum_m  dz[m][co] * (q_a[m][k]  - zp_a)          (zeroed by the callee)
 * frost_pw_chain_supported(K, cout) != 0 tells which layers qualify (K <= 64, cout tiles wider than 64 channels). */
typedef struct {
  FrostPwOperands op;
  FrostBnBackwardArgs bn;      /* acc, dz, dz_lo are ignored                                                 */
  const void* wt_bf16;         /* FrostWeightDesc.wt_bf16 [K][cout]                                          */
  float* dx;                   /* [M][K] fp32                                                                */
  int32_t accumulate;          /* dx += instead of dx =                                                      */
  float* dwq;                  /* [cout][K] fp32                                                             */
} FrostPwChainArgs;
int frost_pw_chain_supported(int K, int cout);
int frost_pw_chain_backward(const FrostPwChainArgs* a, void* stream);
int frost_pw_fused_forward(const FrostPwFusedFwdArgs* a, void* stream);
int frost_pw_fused_bwd_reduce(const FrostPwFusedBwdArgs* a, void* stream);
int frost_pw_fused_bwd_apply(const FrostPwFusedBwdArgs* a, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Quantization-aware hard-swish: the reference's _Hswish module
 * (Classification/models/imagenet/mobilenetv3.py:43-56) after prepare_qat,
 *     y = mul_scalar( FQ_mul( x * FQ_relu6( relu6( add_scalar(x, 3) ) ) ), 1/6 )
 * with both observers (EMA + qparams, updated in place), on a tensor x that sits on the uint8 grid (in_scale, in_zp).
 * The input takes <= 256 values, so the op is an index pass + one 256-entry table pass (hswish.cu).
 *   q_in      out [n] uint8: the input indices (kept for the backward)
 *   y / y_q   out [n] fp32 values and / or uint8 indices of the result (either may be NULL); the result's grid is
 *             (*out_scale = scale_mul/6, zero point of fq_mul)
 *   workspace frost_hswish_workspace_floats() floats, written by forward, read by backward of the same step
 * backward: dx = grad through mul_scalar, FQ_mul (STE), mul (both operands), FQ_relu6 (STE), hardtanh, add_scalar. */
int frost_hswish_workspace_floats(void);
int frost_hswish_forward(const float* x, int64_t n, const float* in_scale, const int32_t* in_zp, FrostFQ fq_relu6,
                         int observe_relu6, FrostFQ fq_mul, int observe_mul, float averaging_const, uint8_t* q_in,
                         float* y, uint8_t* y_q, float* workspace, float* out_scale, void* stream);
int frost_hswish_backward(const float* dy, const uint8_t* q_in, int64_t n, const float* workspace, float* dx, void* stream);
/* The reference's _Hsigmoid (mobilenetv3.py:59-69): y = mul_scalar(FQ_relu6(relu6(add_scalar(x, 3))), 1/6); same two passes, one
 * observer; result grid (scale_relu6/6, zero point of fq_relu6).  Backward: frost_hswish_backward with this call's workspace. */
int frost_hsigmoid_forward(const float* x, int64_t n, const float* in_scale, const int32_t* in_zp, FrostFQ fq_relu6,
                           int observe_relu6, float averaging_const, uint8_t* q_in, float* y, uint8_t* y_q, float* workspace,
                           float* out_scale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Quantization-aware squeeze-and-excite: the reference's SEModule
 * (Classification/models/imagenet/mobilenetv3.py:85-102) after fuse_model() + prepare_qat,
 *     gate = Hsigmoid( FQ( Linear_q( FQ( relu( Linear_q( avg_pool(x) ) ) ) ) ) ) ;  y = FQ_mul( x * gate.expand_as(x) )
 * is a composition (frostnet_b200/se.py): frost_pool_dropout_forward (keep = NULL), frost_fq_forward on the weights
 * (symmetric int8) and activations, frost_linear_forward / _backward, frost_hsigmoid_forward - plus the three element-wise
 * pieces below.  fp32; a row is one (image, channel) plane of hw values. */
/* y = max(x, 0) ; mask = [x > 0] (may be NULL).  Backward: frost_fq_backward(dy, mask). */
int frost_relu_forward(const float* x, int64_t n, float* y, uint8_t* mask, void* stream);
/* y[r][i] = x[r][i] * gate[r]        (torch.mul(x, gate.expand_as(x)): one fp32 rounding, bit-identical) */
int frost_bcast_mul_forward(const float* x, const float* gate, int64_t rows, int hw, float* y, void* stream);
/* dx[r][i] = dy[r][i] * gate[r] ; dgate[r] = sum_i dy[r][i] * x[r][i] */
int frost_bcast_mul_backward(const float* dy, const float* x, const float* gate, int64_t rows, int hw, float* dx,
                             float* dgate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * FloatFunctional.cat / .add (frostnet.py:129,142;
 * torch/ao/nn/quantized/modules/functional_modules.py:50-52,80-82): op + own observer + FQ. */
typedef struct {
  const uint8_t* q;
  const float* scale;
  const int32_t* zp;
  const float* cur_minmax;     /* dequantised min/max of this tensor */
  int32_t C;
  int32_t ld;                  /* row pitch in bytes (>= C; == C: dense NHWC).  The engine pads rows to multiples of 16
                                  bytes so that TMA can address them; padding bytes are don't-care */
} FrostQTensor;
/* out[M][C1+C2] = FQ(cat([a, b], channel)) */
int frost_cat_forward(FrostQTensor a, FrostQTensor b, int64_t M, FrostFQ fq, int observe,
                      float averaging_const, uint8_t* q_out, int ld_out, float* cur_minmax_out, void* stream);
/* da[M][C1] = dcat[:, :C1]*mask ; db[M][C2] (+)= dcat[:, C1:]*mask */
int frost_cat_backward(const float* dcat, FrostQTensor a, FrostQTensor b, int64_t M,
                       const float* out_scale, const int32_t* out_zp, float* da, float* db,
                       int accumulate_b, void* stream);
/* out[n] = FQ(a + b)  (two passes: min/max, then quantise). scratch: FROST_FQ_SCRATCH_FLOATS. */
int frost_add_forward(FrostQTensor a, FrostQTensor b, int64_t n, FrostFQ fq, int observe,
                      float averaging_const, uint8_t* q_out, int ld_out, float* cur_minmax_out, float* scratch,
                      void* stream);
/* dsum = dout*mask ; da (+)= dsum */
int frost_add_backward(const float* dout, FrostQTensor a, FrostQTensor b, int64_t n,
                       const float* out_scale, const int32_t* out_zp, float* dsum, float* da,
                       int accumulate_a, void* stream);
/* y (+)= x, fp32 */
int frost_axpy(const float* x, float* y, int64_t n, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Head: AdaptiveAvgPool2d(1) -> Dropout -> nnqat.Conv2d(1280,nclass,1)  (frostnet.py:295-299,328) */
/* pooled[N][C] = mean_hw((q-zp)*s) * keep[N][C] * keep_scale   (keep may be NULL) */
int frost_pool_dropout_forward(const uint8_t* q, const float* scale, const int32_t* zp, int N,
                               int HW, int C, const float* keep, float keep_scale, float* pooled,
                               void* stream);
/* dy[N][HW][C] = dpooled[N][C]*keep*keep_scale/HW */
int frost_pool_dropout_backward(const float* dpooled, int N, int HW, int C, const float* keep,
                                float keep_scale, float* dy, void* stream);
/* out[N][cout] = x[N][K] . ((wq - zp_w)*s_w)^T + bias   (fp32 SIMT; wq[cout][K]) */
int frost_linear_forward(const float* x, const int8_t* wq, const float* w_scale,
                         const int32_t* w_zp, const float* bias, int N, int K, int cout, float* out,
                         void* stream);
/* dx[N][K] = dout . Wq ; dwq[cout][K] = dout^T . x ; dbias[cout] = sum_n dout */
int frost_linear_backward(const float* dout, const float* x, const int8_t* wq,
                          const float* w_scale, const int32_t* w_zp, int N, int K, int cout,
                          float* dx, float* dwq, float* dbias, void* stream);

/* ---------------------------------------------------------------------------------------------
 * SSD MultiBox target matching: replaces the per-image Python loop of
 * Object_Detection/layers/modules/multibox_loss.py:66-74 over layers/box_utils.py:71-113 `match` (+ `encode`, :115-138) - the
 * whole batch in one launch, on the device the predictions already live on.
 *   truths [batch][max_obj][4] point-form boxes, labels [batch][max_obj], num_objs[batch] valid entries per image (<= 128)
 *   priors [num_priors][4] centre-size form;  threshold, var0, var1 = overlap threshold and cfg['variance']
 *   loc_t [batch][num_priors][4] encoded offsets, conf_t [batch][num_priors] class index + 1 (0 = background)
 *   scratch_overlap / scratch_index: [batch][num_priors] floats / ints of device scratch
 * conf_t is the reference's bit for bit (same fp32 operations, first-maximum ties); loc_t up to logf's last ulp. */
int frost_multibox_match(const float* truths, const int64_t* labels, const int* num_objs, int batch, int max_obj,
                         const float* priors, int num_priors, float threshold, float var0, float var1, float* loc_t,
                         int64_t* conf_t, float* scratch_overlap, int* scratch_index, void* stream);

/* ---------------------------------------------------------------------------------------------
 * dgrad / wgrad of the convolutions (aten::convolution_backward in the reference). */
/* dx[M][K] (+)= s_w * sum_co dz[M][co]*(wq[co][K]-zp_w)      (CUDA-core fp32 version) */
int frost_pw_dgrad(const float* dz, const int8_t* wq, const float* w_scale, const int32_t* w_zp,
                   int64_t M, int K, int cout, float* dx, int accumulate, void* stream);
/* Same result on tensor cores: tcgen05.mma kind::f16, bf16 operands, fp32 accumulation in TMEM.
 * dz comes as two bf16 planes (hi, lo; written by frost_bn_backward with dz_format 1), the integer weights
 * are exact in bf16 (wt_bf16 [K][cout] from weight prep); 2 MMAs per k-step. */
int frost_pw_dgrad_tc(const void* dz_hi, const void* dz_lo, const void* wt_bf16, const float* w_scale,
                      int64_t M, int K, int cout, float* dx, int accumulate, void* stream);
/* Tensor-core wgrad: reduction over the rows m with MN-major bf16 operands (dz hi/lo planes), fp32 atomics. */
int frost_pw_wgrad_tc(const void* dz_hi, const void* dz_lo, const uint8_t* xq, int ldx, const float* x_scale,
                      const int32_t* x_zp, int64_t M, int K, int cout, float* dwq, void* stream);
/* dwq[cout][K] = s_a * sum_m dz[m][co]*(xq[m][K]-zp_a)     (dwq zeroed by callee; CUDA-core fp32 version) */
int frost_pw_wgrad(const float* dz, const uint8_t* xq, const float* x_scale, const int32_t* x_zp,
                   int64_t M, int K, int cout, float* dwq, void* stream);
int frost_dw_dgrad(const float* dz, const int8_t* wq, const float* w_scale, const int32_t* w_zp,
                   int N, int H, int W, int C, int k, int stride, float* dx, int accumulate,
                   void* stream);
/* dwq[k*k][C] */
int frost_dw_wgrad(const float* dz, const uint8_t* xq, int ldx, const float* x_scale, const int32_t* x_zp,
                   int N, int H, int W, int C, int k, int stride, float* dwq, void* stream);
/* The same three with dilation >= 1 and an explicit padding 0 .. dilation*(k-1) (the reference's layers use dilation*(k-1)/2, the
 * SSD extras' depthwise 3x3 padding 0; Object_Detection/ssd_qmv2.py:40-52,137-138,
 * Classification/models/imagenet/mobilenetv3.py:129-131 with dilated=True): plain gather kernels (dw_dilated.cu), same operand
 * conventions - acc / stats as frost_dw_conv_forward, dx (+)= as frost_dw_dgrad, dwq[k*k][C] as frost_dw_wgrad. */
int frost_dw_conv_forward_dilated(const uint8_t* xq, int ldx, const int32_t* x_zp, const int8_t* wq, const int32_t* w_zp,
                                  int N, int H, int W, int C, int k, int stride, int dilation, int pad, int32_t* acc,
                                  FrostChanStats* stats, void* stream);
int frost_dw_dgrad_dilated(const float* dz, const int8_t* wq, const float* w_scale, const int32_t* w_zp, int N, int H, int W,
                           int C, int k, int stride, int dilation, int pad, float* dx, int accumulate, void* stream);
int frost_dw_wgrad_dilated(const float* dz, const uint8_t* xq, int ldx, const float* x_scale, const int32_t* x_zp, int N, int H,
                           int W, int C, int k, int stride, int dilation, int pad, float* dwq, void* stream);

/* Feature backbone (frostnet_features.py:342-352): the stem convolves the raw fp32 NCHW image with the
 * fake-quantised weights: z[m][co] = sum x * (q_w - zp_w)   (conv == s_w * z), fp32 statistics. */
int frost_stem_conv_forward_f32(const float* x_nchw, const int8_t* wq, const int32_t* w_zp, int N, int H, int W,
                                int cin, int cout, int k, int stride, int pad, float* z, FrostChanStats* stats,
                                void* stream);
int frost_stem_wgrad_f32(const float* dz, const float* x_nchw, int N, int H, int W, int cin, int cout, int k,
                         int stride, int pad, float* dwq, void* stream);
/* feature taps: y_nchw = (q - zp) * s ; and the way back: g_nhwc (+)= transpose(g_nchw) */
int frost_dequant_to_nchw(const uint8_t* q, int ldq, const float* scale, const int32_t* zp, int N, int H, int W, int C,
                          float* y_nchw, void* stream);
int frost_nchw_to_nhwc(const float* g_nchw, int N, int C, int H, int W, float* g_nhwc, int accumulate, void* stream);
/* dwq[cout][k][k][cin] (the stem has no dgrad: its input is the image) */
int frost_stem_wgrad(const float* dz, const uint8_t* xq, const float* x_scale, const int32_t* x_zp,
                     int N, int H, int W, int cin, int cout, int k, int stride, int pad, float* dwq,
                     void* stream);

/* ---------------------------------------------------------------------------------------------
 * GradBoost optimizers (optimizer.py:121-206 QSGD, :264-359 QRMSprop, :411-512 QAdam,
 * :564-667 QAdamW) as ONE multi-tensor kernel.  The |Laplace(0,1)| noise and the coin toss that
 * the reference draws on the host with numpy (optimizer.py:178-185) come from a counter-based
 * Philox4x32-10 stream keyed by (seed, tensor index, element index, step) - or from explicit
 * arrays (noise/coin != NULL) so tests can inject the reference's own draws. */
#define FROST_OPT_QSGD 0
#define FROST_OPT_QRMS 1
#define FROST_OPT_QADAM 2
#define FROST_OPT_QADAMW 3
typedef struct {
  float* p;
  float* g;            /* mutated in place like the reference (grad.add_(noise), wd)           */
  float* exp_min;
  float* exp_max;
  float* coin_toss;    /* state['coin_toss'] (NULL if !toss_coin)                               */
  float* buf0;         /* QSGD momentum_buffer | QRMS square_avg | QAdam exp_avg               */
  float* buf1;         /* QRMS momentum_buffer | QAdam exp_avg_sq                              */
  float* buf2;         /* QRMS grad_avg (centered) | QAdam max_exp_avg_sq (amsgrad)            */
  const float* noise;  /* optional injected |Laplace| draws                                    */
  const float* coin;   /* optional injected coin draws                                         */
  int64_t n;
  float lr, weight_decay;
  int32_t step;        /* state['step'] AFTER increment                                        */
  int32_t restart_step;/* state['restart_step'] AFTER increment                                */
  int32_t first_momentum; /* QSGD: 1 if momentum_buffer is being created this step             */
  int32_t pad_;
} FrostOptTensor;
typedef struct {
  int32_t kind;
  int32_t is_warmup;
  int32_t toss_coin, nesterov, centered, amsgrad;
  float momentum, dampening, beta, beta1, beta2, eps, alpha, clip_by, noise_decay;
  float grad_scale;    /* multiplied into g first (1/world_size after an all-reduce sum)       */
  uint64_t seed;
} FrostOptHyper;
/* `tensors` (n of them) and `chunks` live in DEVICE memory; `hyper` is a host struct.
 * chunks[i] = {tensor index, chunk index within the tensor}; a chunk is FROST_OPT_CHUNK
 * consecutive elements, one CTA each, so no CTA is launched without work. */
#define FROST_OPT_CHUNK 2048
int frost_gradboost_multi(const FrostOptTensor* tensors, int n, const FrostOptChunk* chunks,
                          int n_chunks, const FrostOptHyper* hyper, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FROST_B200_H_ */
