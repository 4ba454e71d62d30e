// conv_fwd.cu - forward convolutions on quantize indices (exact integer restatement of the
// F.conv2d at torch/ao/nn/intrinsic/qat/modules/conv_fused.py:155 for frostnet.py:14-28,46-60).
//   I[m][co] = sum_k (q_a[m][k] - zp_a) * (q_w[co][k] - zp_w)          (int32, exact)
// so that conv == s_a*s_w*I.  Every kernel also accumulates the per-channel integer statistics
// (sum, sum of squares, min, max) that training-mode BatchNorm and the activation observer need.
#include "common.cuh"

namespace frost {

__device__ __forceinline__ int dp4a_us(unsigned a, int b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp4a_uu(unsigned a, unsigned b, int c) {
  int d;
  asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}

// Block-level per-channel statistic slots in shared memory, flushed with one set of global
// integer atomics per channel per CTA.
struct SmemStat {
  long long sum;
  unsigned long long sq;
  int mn, mx;
};
__device__ __forceinline__ void smem_stat_init(SmemStat* s, int n) {
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s[i].sum = 0;
    s[i].sq = 0;
    s[i].mn = INT_MAX;
    s[i].mx = INT_MIN;
  }
}
__device__ __forceinline__ void smem_stat_add(SmemStat* s, long long sum, unsigned long long sq, int mn, int mx) {
  atomicAdd(reinterpret_cast<unsigned long long*>(&s->sum), (unsigned long long)sum);
  atomicAdd(&s->sq, sq);
  atomicMin(&s->mn, mn);
  atomicMax(&s->mx, mx);
}
__device__ __forceinline__ void global_stat_flush(FrostChanStats* g, const SmemStat& s) {
  chan_stats_flush(g, s.sum, s.sq, s.mn, s.mx);
}

// ================================================================= 1x1 pointwise (dp4a, SIMT)
// CTA tile 128 rows x BN couts, K consumed in 64-byte chunks staged in shared memory with a
// register prefetch of the next chunk.  256 threads; thread (tx,ty) owns rows ty+16*i (i<8) and
// couts tx+16*j (j<BN/16).  smem rows padded to 80 B so 16-byte LDS are conflict-free.
constexpr int PW_BM = 128;
constexpr int PW_KC = 64;
constexpr int PW_LD = 80;

template <int BN>
__global__ void __launch_bounds__(256) pw_conv_fwd_kernel(const uint8_t* xq, const int32_t* x_zp_p,
                                                         const int8_t* wq, const int32_t* w_zp_p,
                                                         const int32_t* wsum, int64_t M, int K, int cout,
                                                         int32_t* acc_out, FrostChanStats* stats) {
  constexpr int TN = BN / 16;
  constexpr int A_LOADS = PW_BM * PW_KC / 8 / 256;  // uint2 per thread = 4
  constexpr int B_LOADS = (BN * PW_KC / 8 + 255) / 256;
  __shared__ __align__(16) uint8_t As[PW_BM * PW_LD];
  __shared__ __align__(16) uint8_t Bs[BN * PW_LD];
  __shared__ SmemStat s_stat[BN];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.x * PW_BM;
  const int n0 = blockIdx.y * BN;
  const int zp_a = *x_zp_p, zp_w = *w_zp_p;
  const bool need_rowsum = (zp_w != 0);

  smem_stat_init(s_stat, BN);

  int acc[8][TN];
  int rowsum[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    rowsum[i] = 0;
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0;
  }

  uint2 pa[A_LOADS], pb[B_LOADS];
  auto load_chunk = [&](int k0) {
#pragma unroll
    for (int l = 0; l < A_LOADS; ++l) {
      const int idx = tid + l * 256;  // 0..1023 : row = idx/8, col8 = idx%8
      const int r = idx >> 3, c8 = idx & 7;
      const int64_t m = m0 + r;
      const int k = k0 + c8 * 8;
      pa[l] = make_uint2(0u, 0u);
      if (m < M && k < K) pa[l] = ld_cg(reinterpret_cast<const uint2*>(xq + m * K + k));
    }
#pragma unroll
    for (int l = 0; l < B_LOADS; ++l) {
      const int idx = tid + l * 256;
      const int r = idx >> 3, c8 = idx & 7;
      const int k = k0 + c8 * 8;
      pb[l] = make_uint2(0u, 0u);
      if (r < BN && n0 + r < cout && k < K) pb[l] = ld_cg(reinterpret_cast<const uint2*>(wq + (int64_t)(n0 + r) * K + k));
    }
  };
  auto store_chunk = [&]() {
#pragma unroll
    for (int l = 0; l < A_LOADS; ++l) {
      const int idx = tid + l * 256;
      *reinterpret_cast<uint2*>(As + (idx >> 3) * PW_LD + (idx & 7) * 8) = pa[l];
    }
#pragma unroll
    for (int l = 0; l < B_LOADS; ++l) {
      const int idx = tid + l * 256;
      if ((idx >> 3) < BN) *reinterpret_cast<uint2*>(Bs + (idx >> 3) * PW_LD + (idx & 7) * 8) = pb[l];
    }
  };

  load_chunk(0);
  for (int k0 = 0; k0 < K; k0 += PW_KC) {
    __syncthreads();  // previous chunk fully consumed
    store_chunk();
    __syncthreads();
    if (k0 + PW_KC < K) load_chunk(k0 + PW_KC);
#pragma unroll
    for (int kk = 0; kk < PW_KC / 16; ++kk) {
      uint4 b[TN];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = *reinterpret_cast<const uint4*>(Bs + (tx + 16 * j) * PW_LD + kk * 16);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint4 a = *reinterpret_cast<const uint4*>(As + (ty + 16 * i) * PW_LD + kk * 16);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          int v = acc[i][j];
          v = dp4a_us(a.x, (int)b[j].x, v);
          v = dp4a_us(a.y, (int)b[j].y, v);
          v = dp4a_us(a.z, (int)b[j].z, v);
          v = dp4a_us(a.w, (int)b[j].w, v);
          acc[i][j] = v;
        }
        if (need_rowsum) {
          int r = rowsum[i];
          r = dp4a_uu(a.x, 0x01010101u, r);
          r = dp4a_uu(a.y, 0x01010101u, r);
          r = dp4a_uu(a.z, 0x01010101u, r);
          r = dp4a_uu(a.w, 0x01010101u, r);
          rowsum[i] = r;
        }
      }
    }
  }

  // epilogue: zero-point corrections, store, statistics
  const int kzz = K * zp_a * zp_w;
#pragma unroll
  for (int j = 0; j < TN; ++j) {
    const int co = n0 + tx + 16 * j;
    const bool cvalid = co < cout;
    const int corr_c = cvalid ? zp_a * ld_cg(wsum + co) : 0;
    long long s = 0;
    unsigned long long sq = 0;
    int mn = INT_MAX, mx = INT_MIN;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int64_t m = m0 + ty + 16 * i;
      if (cvalid && m < M) {
        const int I = acc[i][j] - corr_c - zp_w * rowsum[i] + kzz;
        acc_out[m * cout + co] = I;
        s += I;
        sq += (unsigned long long)((long long)I * (long long)I);
        mn = min(mn, I);
        mx = max(mx, I);
      }
    }
    // combine the two ty's that share a warp, then one smem update per (warp, cout)
    s += __shfl_xor_sync(0xffffffffu, s, 16);
    sq += __shfl_xor_sync(0xffffffffu, sq, 16);
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, 16));
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, 16));
    if ((tid & 16) == 0 && cvalid && mn <= mx) smem_stat_add(&s_stat[tx + 16 * j], s, sq, mn, mx);
  }
  __syncthreads();
  if (tid < BN && n0 + tid < cout) global_stat_flush(stats + n0 + tid, s_stat[tid]);
}

// ================================================================= dense kxk stem (SIMT int)
// One thread per output pixel, all couts (<=32) in registers; weights broadcast from smem.
constexpr int STEM_MAXC = 32;
constexpr int STEM_THREADS = 256;

__global__ void __launch_bounds__(STEM_THREADS) stem_conv_fwd_kernel(
    const uint8_t* xq, const int32_t* x_zp_p, const int8_t* wq,
    const int32_t* w_zp_p, int N, int H, int W, int cin, int cout, int k, int stride, int pad, int Ho,
    int Wo, int32_t* acc_out, FrostChanStats* stats) {
  extern __shared__ int s_mem[];
  pdl_enter();
  const int KK = k * k * cin;
  int* s_w = s_mem;                            // [KK][STEM_MAXC]
  int* s_tile = s_mem + KK * STEM_MAXC;        // [STEM_THREADS][STEM_MAXC+1]
  __shared__ SmemStat s_stat[STEM_MAXC];
  const int zp_a = *x_zp_p, zp_w = *w_zp_p;
  for (int i = threadIdx.x; i < KK * STEM_MAXC; i += blockDim.x) {
    const int t = i / STEM_MAXC, co = i % STEM_MAXC;
    s_w[i] = (co < cout) ? (int)wq[(int64_t)co * KK + t] - zp_w : 0;
  }
  smem_stat_init(s_stat, STEM_MAXC);
  __syncthreads();

  const int64_t total = (int64_t)N * Ho * Wo;
  const int64_t tiles = (total + STEM_THREADS - 1) / STEM_THREADS;
  // statistics role: thread -> (pixel group g, channel c), fixed for the CTA's lifetime; the CTA is persistent
  // over pixel tiles so that only gridDim.x (one resident wave) sets of atomics reach each channel's record
  const int sc = threadIdx.x % STEM_MAXC;
  const int sg = threadIdx.x / STEM_MAXC;      // 0..7
  constexpr int G = STEM_THREADS / STEM_MAXC;  // 8
  long long st_s = 0;
  unsigned long long st_sq = 0;
  int st_mn = INT_MAX, st_mx = INT_MIN;
  for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const int64_t p0 = tile * STEM_THREADS;
    const int64_t p = p0 + threadIdx.x;
    int acc[STEM_MAXC];
#pragma unroll
    for (int c = 0; c < STEM_MAXC; ++c) acc[c] = 0;
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
          const uint8_t* px = xq + (((int64_t)n * H + ih) * W + iw) * cin;
          for (int ci = 0; ci < cin; ++ci) {
            const int xa = (int)ld_cg(px + ci) - zp_a;
            const int* wrow = s_w + ((r * k + s) * cin + ci) * STEM_MAXC;
#pragma unroll
            for (int c = 0; c < STEM_MAXC; ++c) acc[c] += xa * wrow[c];
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < STEM_MAXC; ++c) s_tile[threadIdx.x * (STEM_MAXC + 1) + c] = acc[c];
    __syncthreads();
    // coalesced store + per-channel statistics
    const int npix = (int)min((int64_t)STEM_THREADS, total - p0);
    if (sc < cout) {
      int s32 = 0;   // |I| <= 27*255*128 < 2^20 and <= 32 pixels per thread and tile: no overflow
      for (int px = sg; px < npix; px += G) {
        const int I = s_tile[px * (STEM_MAXC + 1) + sc];
        acc_out[(p0 + px) * cout + sc] = I;
        s32 += I;
        st_sq += (unsigned long long)((long long)I * (long long)I);
        st_mn = min(st_mn, I);
        st_mx = max(st_mx, I);
      }
      st_s += s32;
    }
    __syncthreads();
  }
  if (sc < cout && st_mn <= st_mx) smem_stat_add(&s_stat[sc], st_s, st_sq, st_mn, st_mx);
  __syncthreads();
  if (threadIdx.x < cout) {
    const SmemStat t = s_stat[threadIdx.x];
    chan_stats_flush(stats + threadIdx.x, t.sum, t.sq, t.mn, t.mx);
  }
}

}  // namespace frost

using namespace frost;

extern "C" int frost_pw_conv_forward_simt(const uint8_t* xq, const int32_t* x_zp, const int8_t* wq, const int32_t* w_zp,
                                     const int32_t* wsum, int64_t M, int K, int cout, int32_t* acc,
                                     FrostChanStats* stats, void* stream) {
  FROST_REQUIRE(xq && x_zp && wq && w_zp && wsum && acc && stats, "frost_pw_conv_forward_simt: null pointer");
  FROST_REQUIRE(M > 0 && K > 0 && cout > 0, "frost_pw_conv_forward_simt: empty problem M=%lld K=%d cout=%d", (long long)M, K, cout);
  FROST_REQUIRE(K % 8 == 0, "frost_pw_conv_forward_simt: K=%d must be a multiple of 8", K);
  FROST_REQUIRE((reinterpret_cast<uintptr_t>(xq) & 7) == 0 && (reinterpret_cast<uintptr_t>(wq) & 7) == 0,
                "frost_pw_conv_forward_simt: operands must be 8-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned gx = (unsigned)ceil_div(M, PW_BM);
  if (cout <= 32 || (cout % 64 != 0 && cout % 64 <= 32 && cout < 128)) {
    dim3 grid(gx, (unsigned)ceil_div(cout, 32));
    pw_conv_fwd_kernel<32><<<grid, 256, 0, st>>>(xq, x_zp, wq, w_zp, wsum, M, K, cout, acc, stats);
  } else {
    dim3 grid(gx, (unsigned)ceil_div(cout, 64));
    pw_conv_fwd_kernel<64><<<grid, 256, 0, st>>>(xq, x_zp, wq, w_zp, wsum, M, K, cout, acc, stats);
  }
  FROST_LAUNCH_CHECK("pw_conv_fwd");
  return FROST_OK;
}

extern "C" int frost_stem_conv_forward(const uint8_t* xq, const int32_t* x_zp, const int8_t* wq, const int32_t* w_zp,
                                       int N, int H, int W, int cin, int cout, int k, int stride, int pad,
                                       int32_t* acc, FrostChanStats* stats, void* stream) {
  FROST_REQUIRE(xq && x_zp && wq && w_zp && acc && stats, "frost_stem_conv_forward: null pointer");
  FROST_REQUIRE(cout > 0 && cout <= STEM_MAXC, "frost_stem_conv_forward: cout=%d must be in 1..%d", cout, STEM_MAXC);
  FROST_REQUIRE(cin > 0 && cin <= 8 && k > 0 && k <= 7, "frost_stem_conv_forward: cin<=8, k<=7");
  cudaStream_t st = (cudaStream_t)stream;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  const int64_t total = (int64_t)N * Ho * Wo;
  const size_t smem = sizeof(int) * ((size_t)k * k * cin * STEM_MAXC + (size_t)STEM_THREADS * (STEM_MAXC + 1));
  if (smem > 48 * 1024 && first_use_on_device(reinterpret_cast<const void*>(&stem_conv_fwd_kernel)))
    cudaFuncSetAttribute(stem_conv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int64_t wave = (int64_t)kNumSMs * tunable(FROST_TUNE_STEM_FWD_CTAS_PER_SM);
  launch_pdl(stem_conv_fwd_kernel, dim3((unsigned)std::min<int64_t>(ceil_div(total, STEM_THREADS), wave)), dim3(STEM_THREADS), smem, st,
             xq, x_zp, wq, w_zp, N, H, W, cin, cout, k, stride, pad, Ho, Wo, acc, stats);
  FROST_LAUNCH_CHECK("stem_conv_fwd");
  return FROST_OK;
}
