// pw_bwd_tc.cu - dgrad of the 1x1 convolutions on 5th-gen tensor cores (aten::convolution_backward's
// input gradient for the squeeze / expand / reduce / last_layer convs of frostnet.py:98-119,293).
//   dx[m][k] (+)= s_w * sum_co dz[m][co] * (q_w[co][k] - zp_w)
// dz is fp32 (it is not on a quantisation grid), the weights are small integers.  tcgen05.mma kind::f16
// with bf16 operands and fp32 accumulation in TMEM: the integer weights are exact in bf16, dz is split into
// bf16 hi + bf16 lo (16 mantissa bits, relative error 2^-17) and both halves are accumulated, i.e. 2 MMAs
// per k-step.  Same warp-specialised persistent structure as pw_conv_tc.cu; the producer warps convert
// (fp32 -> hi/lo bf16, int8 -> bf16) in registers and store straight into the SWIZZLE_128B operand tiles.
#include <cuda_bf16.h>
#include "tc_common.cuh"

namespace frost {

using namespace tc;

constexpr int DG_BM = 128;
constexpr int DG_KE = 64;        // reduction elements per stage (64 bf16 = one 128-byte swizzle span)
constexpr int DG_THREADS = 288;
constexpr int DG_SCR = 32 * 36;

template <int BN>
__host__ __device__ constexpr int dg_stages() { return BN >= 256 ? 3 : 4; }
template <int BN>
constexpr size_t dg_smem_bytes() {
  return 1024 + (size_t)dg_stages<BN>() * (2 * DG_BM * 128 + BN * 128) + 4 * DG_SCR * 4 + (2 * dg_stages<BN>() + 4) * 8 + 16;
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ void split_bf16(float x, float& hi, float& lo) {
  hi = __bfloat162float(__float2bfloat16_rn(x));
  lo = x - hi;
}

template <int BN>
__global__ void __launch_bounds__(DG_THREADS, 1) pw_dgrad_tc_kernel(const float* __restrict__ dz, const int8_t* __restrict__ wq_t,
                                                                   const float* __restrict__ w_scale_p,
                                                                   const int32_t* __restrict__ w_zp_p, int64_t M, int K,
                                                                   int cout, float* __restrict__ dx, int accumulate) {
  constexpr int STAGES = dg_stages<BN>();
  constexpr int A_BYTES = DG_BM * 128, B_BYTES = BN * 128, STAGE = 2 * A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* scratch = reinterpret_cast<float*>(smem + STAGES * STAGE);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(scratch + 4 * DG_SCR);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.y * BN;                    // first input channel (k) of this CTA's column tile
  const int n_valid = min(BN, K - n0);
  const int n_eff = (n_valid + 15) & ~15;
  const int num_kb = (cout + DG_KE - 1) / DG_KE;     // reduction blocks over cout
  const int64_t m_tiles = (M + DG_BM - 1) / DG_BM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 128); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc<2 * BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 4 && warp < 8) {
    // ================================================================= producer: load + convert + swizzled store
    const int tp = threadIdx.x - 128;
    const int c16 = tp & 7, r0 = tp >> 3;
    const float zp_w = (float)*w_zp_p;
    uint32_t it = 0;
    for (int64_t mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      const int64_t m0 = mt * DG_BM;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        uint8_t* a_hi = smem + s * STAGE;
        uint8_t* a_lo = a_hi + A_BYTES;
        uint8_t* b_s = a_lo + A_BYTES;
        const int co = kb * DG_KE + c16 * 8;         // first of this thread's 8 reduction elements
        // issue all global loads of the stage before touching shared memory (memory-level parallelism)
        float4 f[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t m = m0 + r0 + 16 * i;
          f[i][0] = f[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m < M && co < cout) {                  // cout % 8 == 0: the 8-element chunk is in or out
            const float4* p = reinterpret_cast<const float4*>(dz + m * cout + co);
            f[i][0] = __ldg(p);
            f[i][1] = __ldg(p + 1);
          }
        }
        // weights: all loads first (BN*8/128 <= 16 per thread), then convert + store
        constexpr int WL = BN * 8 / 128;
        uint2 wpk[WL];
#pragma unroll
        for (int q = 0; q < WL; ++q) {
          const int idx = tp + 128 * q;
          const int r = idx >> 3, c = idx & 7;
          const int cc = kb * DG_KE + c * 8;
          wpk[q] = make_uint2(0u, 0u);
          if (idx < n_eff * 8 && r < n_valid && cc < cout)
            wpk[q] = __ldg(reinterpret_cast<const uint2*>(wq_t + (int64_t)(n0 + r) * cout + cc));
        }
        mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 16 * i;
          const float x[8] = {f[i][0].x, f[i][0].y, f[i][0].z, f[i][0].w, f[i][1].x, f[i][1].y, f[i][1].z, f[i][1].w};
          float h[8], l[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) split_bf16(x[e], h[e], l[e]);
          const uint32_t off = sw128_offset(r, c16);
          *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
          *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(pack_bf16(l[0], l[1]), pack_bf16(l[2], l[3]), pack_bf16(l[4], l[5]), pack_bf16(l[6], l[7]));
        }
#pragma unroll
        for (int q = 0; q < WL; ++q) {
          const int idx = tp + 128 * q;
          if (idx < n_eff * 8) {
            const int r = idx >> 3, c = idx & 7;
            const bool v = (r < n_valid) && (kb * DG_KE + c * 8 < cout);
            float w[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const unsigned word = e < 4 ? wpk[q].x : wpk[q].y;
              w[e] = v ? (float)(int)(int8_t)((word >> (8 * (e & 3))) & 0xff) - zp_w : 0.0f;
            }
            *reinterpret_cast<uint4*>(b_s + sw128_offset(r, c)) =
                make_uint4(pack_bf16(w[0], w[1]), pack_bf16(w[2], w[3]), pack_bf16(w[4], w[5]), pack_bf16(w[6], w[7]));
          }
        }
        fence_proxy_async();
        mbar_arrive(&full_bar[s]);
      }
    }
  } else if (warp == 8) {
    // ================================================================= MMA issuer
    const uint32_t idesc = umma_idesc(1 /*F32*/, 1 /*BF16*/, 1 /*BF16*/, DG_BM, n_eff);
    uint32_t it = 0, tile_i = 0;
    for (int64_t mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++tile_i) {
      const uint32_t acc = tile_i & 1;
      mbar_wait(&tempty_bar[acc], ((tile_i >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&full_bar[s], (it / STAGES) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_s = smem_u32(smem + s * STAGE);
          const uint64_t ahi = umma_desc_sw128(a_s), alo = umma_desc_sw128(a_s + A_BYTES), bd = umma_desc_sw128(a_s + 2 * A_BYTES);
          const int nk = min(DG_KE / 16, (cout - kb * DG_KE + 15) / 16);
          for (int k4 = 0; k4 < nk; ++k4) {
            umma_f16(d_tmem, ahi + (uint64_t)(2 * k4), bd + (uint64_t)(2 * k4), idesc, (kb | k4) != 0 ? 1u : 0u);
            umma_f16(d_tmem, alo + (uint64_t)(2 * k4), bd + (uint64_t)(2 * k4), idesc, 1u);
          }
          umma_commit(&empty_bar[s]);
          if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================================================= epilogue
    float* my = scratch + warp * DG_SCR;
    const float s_w = *w_scale_p;
    uint32_t tile_i = 0;
    for (int64_t mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++tile_i) {
      const uint32_t acc = tile_i & 1;
      const int64_t m0 = mt * DG_BM + warp * 32;
      mbar_wait(&tfull_bar[acc], (tile_i >> 1) & 1);
      tc_fence_after();
      for (int chunk = 0; chunk * 32 < n_valid; ++chunk) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + acc * BN + chunk * 32 + ((uint32_t)(warp * 32) << 16), v);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          *reinterpret_cast<float4*>(my + lane * 36 + 4 * jj) =
              make_float4(__uint_as_float(v[4 * jj]) * s_w, __uint_as_float(v[4 * jj + 1]) * s_w,
                          __uint_as_float(v[4 * jj + 2]) * s_w, __uint_as_float(v[4 * jj + 3]) * s_w);
        __syncwarp();
        const int col4 = (lane & 7) * 4;
        const int n = n0 + chunk * 32 + col4;
        if (n < K) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = (lane >> 3) + 4 * i;
            const int64_t m = m0 + row;
            if (m < M) {
              float4 val = *reinterpret_cast<const float4*>(my + row * 36 + col4);
              float4* dst = reinterpret_cast<float4*>(dx + m * K + n);
              if (accumulate) {
                const float4 o = *dst;
                val.x += o.x; val.y += o.y; val.z += o.z; val.w += o.w;
              }
              *dst = val;
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<2 * BN>(tmem_base);
  }
}

template <int BN>
static int launch_dgrad_tc(const float* dz, const int8_t* wq_t, const float* w_scale, const int32_t* w_zp, int64_t M, int K,
                           int cout, float* dx, int accumulate, cudaStream_t st) {
  constexpr size_t smem = dg_smem_bytes<BN>();
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(pw_dgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("pw_dgrad_tc: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return FROST_ECUDA;
    }
    attr_done = true;
  }
  const int n_tiles = (K + BN - 1) / BN;
  const int64_t m_tiles = ceil_div(M, DG_BM);
  const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(m_tiles, kNumSMs / n_tiles));
  pw_dgrad_tc_kernel<BN><<<dim3(gx, n_tiles), DG_THREADS, smem, st>>>(dz, wq_t, w_scale, w_zp, M, K, cout, dx, accumulate);
  return FROST_OK;
}


// ================================================================= wgrad
//   dwq[co][k] += s_a * sum_m dz[m][co] * (q_a[m][k] - zp_a)
// GEMM with the reduction over the rows m: D[co][k] = sum_m A[co][m] * B[k][m].  Both operands are stored
// exactly as they sit in HBM (row m = 128-byte lines of co resp. k), which is the canonical MN-MAJOR
// SWIZZLE_128B layout: 8 (m) x 64 (co|k) bf16 atoms of 1024 B, the next 8 rows 1024 B further (SBO), the next
// 64 columns 8192 B further (LBO).  The transposes the reference's wgrad needs are done by the descriptors.
constexpr int WG_ROWS = 64;       // reduction rows (m) per stage
constexpr int WG_STAGES = 3;
constexpr int WG_BLK = 64 * 128;  // one 64-column block of a stage: 64 rows x 128 B
constexpr int WG_THREADS = 288;
constexpr int WG_BN = 256;
constexpr size_t wg_smem_bytes() {
  return 1024 + (size_t)WG_STAGES * (2 * 2 * WG_BLK + 4 * WG_BLK) + 4 * DG_SCR * 4 + (2 * WG_STAGES + 2) * 8 + 16;
}

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(WG_BLK >> 4) << 16;            // LBO: next 64-element block along M/N
  d |= (uint64_t)(1024 >> 4) << 32;              // SBO: next 8 rows along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(WG_THREADS, 1) pw_wgrad_tc_kernel(const float* __restrict__ dz, const uint8_t* __restrict__ xq,
                                                                   const float* __restrict__ x_scale_p,
                                                                   const int32_t* __restrict__ x_zp_p, int64_t M, int K,
                                                                   int cout, int64_t rows_per_split, float* __restrict__ dwq) {
  constexpr int A_BYTES = 2 * WG_BLK;                   // 128 co x 64 rows, one of (hi, lo)
  constexpr int STAGE = 2 * A_BYTES + 4 * WG_BLK;       // hi, lo, up to 256 k
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* scratch = reinterpret_cast<float*>(smem + WG_STAGES * STAGE);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(scratch + 4 * DG_SCR);
  uint64_t* empty_bar = full_bar + WG_STAGES;
  uint64_t* tfull_bar = empty_bar + WG_STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.x * 128;
  const int k0 = blockIdx.y * WG_BN;
  const int n_valid = min(WG_BN, K - k0);
  const int n_eff = (n_valid + 15) & ~15;
  const int64_t m_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t m_end = min(M, m_begin + rows_per_split);
  const int num_kb = (int)((m_end - m_begin + WG_ROWS - 1) / WG_ROWS);

  if (threadIdx.x == 0) {
    for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full_bar[s], 128); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull_bar[0], 1);
    mbar_fence_init();
  }
  if (warp == 8) tmem_alloc<WG_BN>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (num_kb > 0) {
    if (warp >= 4 && warp < 8) {
      // ================================================================= producer
      const int tp = threadIdx.x - 128;
      const int cidx = tp & 15, r0 = tp >> 4;          // 16 chunks of 8 co per row, 8 rows per pass
      const float zp_a = (float)*x_zp_p;
      const int cpr = n_eff >> 3;                      // 8-channel chunks per row of x
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % WG_STAGES;
        uint8_t* a_hi = smem + s * STAGE;
        uint8_t* a_lo = a_hi + A_BYTES;
        uint8_t* b_s = a_lo + A_BYTES;
        const int64_t mb = m_begin + (int64_t)kb * WG_ROWS;
        const int co = co0 + cidx * 8;
        float4 f[8][2];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t m = mb + r0 + 8 * i;
          f[i][0] = f[i][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (m < m_end && co < cout) {
            const float4* p = reinterpret_cast<const float4*>(dz + m * cout + co);
            f[i][0] = __ldg(p);
            f[i][1] = __ldg(p + 1);
          }
        }
        // activations: all loads first (64 rows x 256/8 chunks / 128 threads <= 16 per thread)
        constexpr int XL = WG_ROWS * (WG_BN / 8) / 128;
        uint2 xpk[XL];
#pragma unroll
        for (int q = 0; q < XL; ++q) {
          const int idx = tp + 128 * q;
          const int r = idx / cpr, kc = idx - r * cpr;
          xpk[q] = make_uint2(0u, 0u);
          if (idx < WG_ROWS * cpr && mb + r < m_end && k0 + kc * 8 < K)
            xpk[q] = __ldg(reinterpret_cast<const uint2*>(xq + (mb + r) * K + k0 + kc * 8));
        }
        mbar_wait(&empty_bar[s], ((kb / WG_STAGES) & 1) ^ 1);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 8 * i;
          const float x[8] = {f[i][0].x, f[i][0].y, f[i][0].z, f[i][0].w, f[i][1].x, f[i][1].y, f[i][1].z, f[i][1].w};
          float h[8], l[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) split_bf16(x[e], h[e], l[e]);
          const uint32_t off = (uint32_t)(cidx >> 3) * WG_BLK + sw128_offset(r, cidx & 7);
          *reinterpret_cast<uint4*>(a_hi + off) = make_uint4(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]), pack_bf16(h[4], h[5]), pack_bf16(h[6], h[7]));
          *reinterpret_cast<uint4*>(a_lo + off) = make_uint4(pack_bf16(l[0], l[1]), pack_bf16(l[2], l[3]), pack_bf16(l[4], l[5]), pack_bf16(l[6], l[7]));
        }
#pragma unroll
        for (int q = 0; q < XL; ++q) {
          const int idx = tp + 128 * q;
          if (idx < WG_ROWS * cpr) {
            const int r = idx / cpr, kc = idx - r * cpr;
            const bool v = (mb + r < m_end) && (k0 + kc * 8 < K);
            float w[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const unsigned word = e < 4 ? xpk[q].x : xpk[q].y;
              w[e] = v ? (float)((word >> (8 * (e & 3))) & 0xff) - zp_a : 0.0f;
            }
            *reinterpret_cast<uint4*>(b_s + (uint32_t)(kc >> 3) * WG_BLK + sw128_offset(r, kc & 7)) =
                make_uint4(pack_bf16(w[0], w[1]), pack_bf16(w[2], w[3]), pack_bf16(w[4], w[5]), pack_bf16(w[6], w[7]));
          }
        }
        fence_proxy_async();
        mbar_arrive(&full_bar[s]);
      }
    } else if (warp == 8) {
      // ================================================================= MMA issuer (A and B MN-major)
      const uint32_t idesc = umma_idesc(1, 1, 1, 128, n_eff) | (1u << 15) | (1u << 16);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % WG_STAGES;
        mbar_wait(&full_bar[s], (kb / WG_STAGES) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_s = smem_u32(smem + s * STAGE);
          const uint64_t ahi = umma_desc_mn_sw128(a_s), alo = umma_desc_mn_sw128(a_s + A_BYTES), bd = umma_desc_mn_sw128(a_s + 2 * A_BYTES);
          const int64_t rows = min((int64_t)WG_ROWS, m_end - (m_begin + (int64_t)kb * WG_ROWS));
          const int nk = (int)((rows + 15) / 16);
          for (int j = 0; j < nk; ++j) {             // 16 rows = two 8-row atoms = 2048 B
            const uint64_t adv = (uint64_t)((2048 * j) >> 4);
            umma_f16(tmem_base, ahi + adv, bd + adv, idesc, (kb | j) != 0 ? 1u : 0u);
            umma_f16(tmem_base, alo + adv, bd + adv, idesc, 1u);
          }
          umma_commit(&empty_bar[s]);
          if (kb == num_kb - 1) umma_commit(&tfull_bar[0]);
        }
        __syncwarp();
      }
    } else {
      // ================================================================= epilogue: fp32 atomics into dwq[co][k]
      float* my = scratch + warp * DG_SCR;
      const float s_a = *x_scale_p;
      mbar_wait(&tfull_bar[0], 0);
      tc_fence_after();
      for (int chunk = 0; chunk * 32 < n_valid; ++chunk) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + chunk * 32 + ((uint32_t)(warp * 32) << 16), v);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          *reinterpret_cast<float4*>(my + lane * 36 + 4 * jj) =
              make_float4(__uint_as_float(v[4 * jj]) * s_a, __uint_as_float(v[4 * jj + 1]) * s_a,
                          __uint_as_float(v[4 * jj + 2]) * s_a, __uint_as_float(v[4 * jj + 3]) * s_a);
        __syncwarp();
        const int col4 = (lane & 7) * 4;
        const int k = k0 + chunk * 32 + col4;
        if (k < K) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = (lane >> 3) + 4 * i;
            const int co = co0 + warp * 32 + row;
            if (co < cout) {
              const float4 val = *reinterpret_cast<const float4*>(my + row * 36 + col4);
              atomicAdd(reinterpret_cast<float4*>(dwq + (int64_t)co * K + k), val);
            }
          }
        }
        __syncwarp();
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<WG_BN>(tmem_base);
  }
}

}  // namespace frost

using namespace frost;

extern "C" int frost_pw_dgrad_tc(const float* dz, const int8_t* wq_t, const float* w_scale, const int32_t* w_zp, int64_t M,
                                 int K, int cout, float* dx, int accumulate, void* stream) {
  FROST_REQUIRE(dz && wq_t && w_scale && w_zp && dx, "frost_pw_dgrad_tc: null pointer");
  FROST_REQUIRE(M > 0 && K > 0 && cout > 0 && K % 4 == 0 && cout % 8 == 0,
                "frost_pw_dgrad_tc: K=%d must be a multiple of 4 and cout=%d of 8", K, cout);
  FROST_REQUIRE((reinterpret_cast<uintptr_t>(dz) & 15) == 0 && (reinterpret_cast<uintptr_t>(dx) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(wq_t) & 7) == 0,
                "frost_pw_dgrad_tc: operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  int rc;
  if (K <= 64) rc = launch_dgrad_tc<64>(dz, wq_t, w_scale, w_zp, M, K, cout, dx, accumulate, st);
  else if (K <= 128) rc = launch_dgrad_tc<128>(dz, wq_t, w_scale, w_zp, M, K, cout, dx, accumulate, st);
  else rc = launch_dgrad_tc<256>(dz, wq_t, w_scale, w_zp, M, K, cout, dx, accumulate, st);
  if (rc) return rc;
  FROST_LAUNCH_CHECK("pw_dgrad_tc");
  return FROST_OK;
}

extern "C" int frost_pw_wgrad_tc(const float* dz, const uint8_t* xq, const float* x_scale, const int32_t* x_zp, int64_t M,
                                 int K, int cout, float* dwq, void* stream) {
  FROST_REQUIRE(dz && xq && x_scale && x_zp && dwq, "frost_pw_wgrad_tc: null pointer");
  FROST_REQUIRE(M > 0 && K > 0 && cout > 0 && K % 8 == 0 && cout % 8 == 0, "frost_pw_wgrad_tc: K and cout must be multiples of 8");
  FROST_REQUIRE((reinterpret_cast<uintptr_t>(dz) & 15) == 0 && (reinterpret_cast<uintptr_t>(dwq) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(xq) & 7) == 0,
                "frost_pw_wgrad_tc: operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(dwq, 0, sizeof(float) * (size_t)K * cout, st) != cudaSuccess) {
    set_error("frost_pw_wgrad_tc: memset failed");
    return FROST_ECUDA;
  }
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(pw_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wg_smem_bytes());
    if (e != cudaSuccess) {
      set_error("pw_wgrad_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return FROST_ECUDA;
    }
    attr_done = true;
  }
  const int ct = (int)ceil_div(cout, 128), kt = (int)ceil_div(K, WG_BN);
  int64_t splits = std::max<int64_t>(1, std::min<int64_t>(ceil_div(M, WG_ROWS * 4), (int64_t)kNumSMs / ((int64_t)ct * kt)));
  int64_t rows = ceil_div(ceil_div(M, splits), WG_ROWS) * WG_ROWS;
  splits = ceil_div(M, rows);
  pw_wgrad_tc_kernel<<<dim3(ct, kt, (unsigned)splits), WG_THREADS, wg_smem_bytes(), st>>>(dz, xq, x_scale, x_zp, M, K, cout, rows, dwq);
  FROST_LAUNCH_CHECK("pw_wgrad_tc");
  return FROST_OK;
}
