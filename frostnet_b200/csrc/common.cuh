// common.cuh - shared device/host helpers for libfrost_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <limits.h>
#include <math.h>
#include "../../include/frost_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libfrost_b200 is written for sm_100a (B200) only"
#endif

namespace frost {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);
int tunable(int which);  // FROST_TUNE_* launch-shape knob (api.cu)
// true the first time `key` (a kernel's address) is seen on the CURRENT device: per-kernel function attributes
// (cudaFuncSetAttribute) are per device, and one process may drive several GPUs (nn.DataParallel threads).
bool first_use_on_device(const void* key);

#define FROST_REQUIRE(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      frost::set_error(__VA_ARGS__);        \
      return FROST_EINVAL;                  \
    }                                       \
  } while (0)

#define FROST_LAUNCH_CHECK(name)                                                     \
  do {                                                                               \
    cudaError_t e__ = cudaGetLastError();                                            \
    if (e__ != cudaSuccess) {                                                        \
      frost::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));      \
      return FROST_ECUDA;                                                            \
    }                                                                                \
    frost::count_launch();                                                           \
  } while (0)

constexpr int kNumSMs = 148;  // B200

// ---------------------------------------------------------------- global loads
// Every load of data that another kernel of the step produced goes to L2 (ld.global.cg), and no pointer carries
// the restrict qualifier: the NON-COHERENT path (ld.global.nc - the ldg intrinsic, or the compiler's promotion
// of loads through const restrict pointers) is exempt from the visibility guarantee of griddepcontrol.wait.  With
// programmatic dependent launch a CTA can be resident on an SM while an older kernel's CTA on the same SM still
// pulls lines into L1; when the torch caching allocator hands the same address to the next layer's tensor, a
// .nc load after the wait hit those stale lines (measured: flaky index mismatches, gone with this rule).
// The streaming operands have no L1 reuse to lose; the depthwise windows are served by L2.
template <typename T>
__device__ __forceinline__ T ld_cg(const T* p) { return __ldcg(p); }

// ---------------------------------------------------------------- programmatic dependent launch (PDL)
// A QAT step is ~590 dependent launches of 5-600 us kernels.  Kernels launched through launch_pdl() may become
// resident while their predecessor in the stream is still draining (the predecessor's CTAs call
// griddepcontrol.launch_dependents first thing), and block in griddepcontrol.wait until the predecessor has
// completed and its writes are visible.  Rule for every kernel launched this way: EVERY CTA executes pdl_enter()
// (or pdl_trigger() ... pdl_wait()) before it touches global memory and before it exits - completion of a grid
// that skipped the wait would release its own dependents too early.
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_trigger(); pdl_wait(); }

// early = false: ordinary stream-ordered launch (for kernels that still load through L1, e.g. 8-byte cp.async.ca)
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_if(bool early, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                        Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = (early && tunable(FROST_TUNE_PDL) == 1) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  return launch_pdl_if(true, kernel, grid, block, smem, st, args...);
}

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline int grid_for(int64_t work_items, int per_block, int max_blocks = kNumSMs * 16) {
  int64_t b = ceil_div(work_items, per_block);
  if (b < 1) b = 1;
  if (b > max_blocks) b = max_blocks;
  return (int)b;
}

// ---------------------------------------------------------------- device helpers
__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// float atomic min / max via the ordered-int trick (works for any mix of signs; target must be
// initialised to +inf / -inf).
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  if (v >= 0.0f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// One CTA's contribution to a channel's integer statistics.  Hundreds of CTAs hit the same 32-byte record at
// the end of a kernel and same-sector atomics serialise in L2 (measured ~12 ns each), so the min/max atomics
// are skipped when a (possibly stale, hence only ever too large / too small) read shows they cannot win.
__device__ __forceinline__ void chan_stats_flush(FrostChanStats* g, long long sum, unsigned long long sq, int mn, int mx) {
  if (mn > mx) return;  // nothing accumulated
  atomicAdd(reinterpret_cast<unsigned long long*>(&g->sum), (unsigned long long)sum);
  if (sq >> 32) {
    atomicAdd(&g->sq_lo, sq & 0xffffffffull);
    atomicAdd(&g->sq_hi, sq >> 32);
  } else {
    atomicAdd(&g->sq_lo, sq);
  }
  if (mn < __ldcg(&g->min)) atomicMin(&g->min, mn);
  if (mx > __ldcg(&g->max)) atomicMax(&g->max, mx);
}

// ---------------------------------------------------------------- quantisation parameters
constexpr float kSmallScaleThreshold = 6.1e-5f;

// ATen quant_utils.h::ChooseQuantizationParams(min,max,qmin,qmax,preserve_sparsity=symmetric),
// the function torch.fused_moving_avg_obs_fake_quant uses for per-tensor qparams.  `mn`/`mx` are
// C floats, intermediates double - restated operation by operation (oracle: choose_qparams).
__device__ __forceinline__ void choose_qparams(float mn, float mx, int qmin, int qmax, bool symmetric,
                                               float* scale_out, int* zp_out) {
  const bool sym_case = (mn < 0.0f) && (mx > 0.0f) && symmetric;
  if (sym_case) {
    const int sym_qmin = -((qmax - qmin) / 2 + 1);
    const int sym_qmax = (qmax - qmin) / 2;
    const float a = fabsf(__fdiv_rn(mn, (float)sym_qmin));
    const float b = fabsf(__fdiv_rn(mx, (float)sym_qmax));
    const double max_scale = (double)fmaxf(a, b);
    mn = (float)(max_scale * (double)sym_qmin);
    mx = (float)(max_scale * (double)sym_qmax);
  }
  mn = fminf(mn, 0.0f);
  mx = fmaxf(mx, 0.0f);
  double scale = ((double)mx - (double)mn) / (double)(qmax - qmin);
  const float fs = (float)scale;
  if (fs == 0.0f || isinf(__fdiv_rn(1.0f, fs))) scale = 0.1;
  if (scale < (double)kSmallScaleThreshold) {
    const float org_scale = (float)scale;
    scale = (double)kSmallScaleThreshold;
    if (mn == 0.0f) {
      mx = __fmul_rn(kSmallScaleThreshold, (float)(qmax - qmin));
    } else if (mx == 0.0f) {
      mn = -__fmul_rn(kSmallScaleThreshold, (float)(qmax - qmin));
    } else {
      const float amp = __fdiv_rn(kSmallScaleThreshold, org_scale);
      mn = __fmul_rn(mn, amp);
      mx = __fmul_rn(mx, amp);
    }
  }
  const double zp_from_min = (double)qmin - (double)mn / scale;
  const double zp_from_max = (double)qmax - (double)mx / scale;
  const double err_min = fabs((double)qmin) - fabs((double)mn / scale);
  const double err_max = fabs((double)qmax) - fabs((double)mx / scale);
  double init_zp = (err_min < err_max) ? zp_from_min : zp_from_max;
  if (sym_case) init_zp = (double)(qmin + qmax) / 2.0;
  int zp;
  if (init_zp < (double)qmin) zp = qmin;
  else if (init_zp > (double)qmax) zp = qmax;
  else zp = (int)rint(init_zp);  // nearbyint, ties-to-even
  *scale_out = (float)scale;
  *zp_out = zp;
}

// MovingAverageMinMaxObserver step + qparams on one FQ state (single thread).
// ATen fused_obs_fake_quant.cpp::calculate_moving_average (== observer.py:668-683).
__device__ __forceinline__ void observer_update(const FrostFQ& fq, float cur_min, float cur_max, int qmin,
                                                int qmax, bool symmetric, float c) {
  float rmin = *fq.min_val, rmax = *fq.max_val;
  if (isinf(rmin) || isinf(rmax)) {
    rmin = cur_min;
    rmax = cur_max;
  } else {
    rmin = __fadd_rn(rmin, __fmul_rn(c, __fsub_rn(cur_min, rmin)));
    rmax = __fadd_rn(rmax, __fmul_rn(c, __fsub_rn(cur_max, rmax)));
  }
  *fq.min_val = rmin;
  *fq.max_val = rmax;
  float s;
  int zp;
  choose_qparams(rmin, rmax, qmin, qmax, symmetric, &s, &zp);
  *fq.scale = s;
  *fq.zero_point = zp;
}

// Unclamped quantize index  rint(x * inv_scale) + zp  (ATen fake_quantize cachemask kernel).
__device__ __forceinline__ float fq_index(float x, float inv_scale, float zp_f) {
  return __fadd_rn(rintf(__fmul_rn(x, inv_scale)), zp_f);
}
__device__ __forceinline__ float fq_dequant(float idx_clamped, float zp_f, float scale) {
  return __fmul_rn(__fsub_rn(idx_clamped, zp_f), scale);
}

// ---------------------------------------------------------------- block min/max + FQ finalize
__device__ __forceinline__ void block_minmax(float& mn, float& mx) {
  __shared__ float s_mn[32], s_mx[32];
  mn = warp_min(mn);
  mx = warp_max(mx);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) { s_mn[w] = mn; s_mx[w] = mx; }
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  if (w == 0) {
    mn = lane < nw ? s_mn[lane] : INFINITY;
    mx = lane < nw ? s_mx[lane] : -INFINITY;
    mn = warp_min(mn);
    mx = warp_max(mx);
  }
  __syncthreads();
}


// Reduce the partials, run observer + qparams; optionally emit the dequantised min/max of the
// tensor that the apply kernel will produce (quantisation is monotone).
static __global__ void __launch_bounds__(1024) fq_finalize_kernel(const float* partial, int nparts, FrostFQ fq,
                                                           int qmin, int qmax, int symmetric, float c,
                                                           int observe, float* cur_minmax) {
  pdl_enter();
  float mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < nparts; i += blockDim.x) {
    mn = fminf(mn, partial[2 * i]);
    mx = fmaxf(mx, partial[2 * i + 1]);
  }
  block_minmax(mn, mx);
  if (threadIdx.x == 0) {
    if (observe) observer_update(fq, mn, mx, qmin, qmax, symmetric != 0, c);
    if (cur_minmax) {
      const float s = *fq.scale, zp = (float)*fq.zero_point;
      const float inv = __fdiv_rn(1.0f, s);
      const float qa = fminf(fmaxf(fq_index(mn, inv, zp), (float)qmin), (float)qmax);
      const float qb = fminf(fmaxf(fq_index(mx, inv, zp), (float)qmin), (float)qmax);
      cur_minmax[0] = fq_dequant(qa, zp, s);
      cur_minmax[1] = fq_dequant(qb, zp, s);
    }
  }
}

}  // namespace frost
