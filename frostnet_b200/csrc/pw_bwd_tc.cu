// pw_bwd_tc.cu - dgrad and wgrad of the 1x1 convolutions on 5th-gen tensor cores
// (aten::convolution_backward for the squeeze / expand / reduce / last_layer convs of frostnet.py:98-119,293).
//   dgrad: dx[m][k]    (+)= s_w * sum_co dz[m][co] * (q_w[co][k] - zp_w)
//   wgrad: dwq[co][k]   +=  s_a * sum_m  dz[m][co] * (q_a[m][k]  - zp_a)
// dz is fp32-valued (it is not on a quantisation grid); frost_bn_backward hands it over as two bf16 planes
// hi + lo (16 mantissa bits, relative error 2^-17).  The integer operands are exact in bf16.  tcgen05.mma
// kind::f16 with fp32 accumulation in TMEM, 2 MMAs (hi, lo) per k-step.
// Because dz already sits in HBM in operand format, the producers are pure cp.async (LDGSTS) into the
// SWIZZLE_128B tiles - the same bytes serve as the K-major A operand of dgrad and, untouched, as the MN-major
// A operand of wgrad (the transposes of the reference's wgrad are done by the UMMA descriptors).
#include <cuda_bf16.h>
#include <cuda.h>
#include <algorithm>
#include "pw_tma.cuh"

namespace frost {

using namespace tc;

constexpr int BW_SCR = 32 * 36;

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// ================================================================= dgrad
constexpr int DG_BM = 128;
constexpr int DG_KE = 64;          // reduction elements (cout) per stage: 64 bf16 = one 128-byte swizzle span
constexpr int DG_THREADS = 416;    // warps 0-7 epilogue (2 groups), 8-11 producer, 12 MMA

template <int BN>
__host__ __device__ constexpr int dg_stages() { return BN >= 256 ? 2 : (BN >= 128 ? 3 : 4); }
template <int BN>
constexpr size_t dg_smem_bytes() {
  return 1024 + (size_t)dg_stages<BN>() * (2 * DG_BM * 128 + BN * 128) + 8 * BW_SCR * 4 + (2 * dg_stages<BN>() + 4) * 8 + 16;
}

template <int BN>
__global__ void __launch_bounds__(DG_THREADS, 1) pw_dgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_hi,
                                                                   const __grid_constant__ CUtensorMap tm_lo,
                                                                   const __grid_constant__ CUtensorMap tm_wt,
                                                                   const uint16_t* dz_hi, const uint16_t* dz_lo, const uint16_t* wt,
                                                                   int use_tma, const float* w_scale_p, int64_t M, int K,
                                                                   int cout, float* dx, int accumulate) {
  constexpr int STAGES = dg_stages<BN>();
  constexpr int BW_LAG = STAGES > 2 ? 2 : 1;   // cp.async groups in flight per producer thread (< STAGES)
  constexpr int A_BYTES = DG_BM * 128, B_BYTES = BN * 128, STAGE = 2 * A_BYTES + B_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* scratch = reinterpret_cast<float*>(smem + STAGES * STAGE);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(scratch + 8 * BW_SCR);
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

  // barrier init and the TMEM allocation overlap the previous kernel's tail; dependents are released after the
  // allocation is complete (see pw_conv_tc.cu)
  if (threadIdx.x == 0) {
    // full: the TMA bytes of the stage, or one arrival per cp.async producer warp
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], use_tma ? 1 : 4); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }
    mbar_fence_init();
  }
  if (warp == 12) tmem_alloc<2 * BN>(tmem_slot);
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= 8 && warp < 12 && use_tma) {
    // ================================================================= producer: one thread, three TMA boxes per stage
    // (dz hi / lo: 64 channels x 128 rows; W^T: 64 channels x BN rows - all K-major SWIZZLE_128B exactly as they sit in HBM;
    // rows / channels past the tensors arrive as zeros).  STAGES stages are in flight; the cp.async producers this replaces
    // kept two groups per thread.
    if (warp == 8 && lane == 0) {
      tma_prefetch_desc(&tm_hi);
      tma_prefetch_desc(&tm_lo);
      tma_prefetch_desc(&tm_wt);
      uint32_t it = 0;
      for (int64_t mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
        const int m0 = (int)(mt * DG_BM);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          const uint32_t a_hi = smem_u32(smem + s * STAGE);
          const uint32_t a_lo = a_hi + A_BYTES, b_s = a_lo + A_BYTES;
          mbar_expect_tx(&full_bar[s], STAGE);
          tma_load_2d(&tm_hi, &full_bar[s], a_hi, kb * DG_KE, m0);
          tma_load_2d(&tm_lo, &full_bar[s], a_lo, kb * DG_KE, m0);
          tma_load_2d(&tm_wt, &full_bar[s], b_s, kb * DG_KE, n0);
        }
      }
    }
  } else if (warp >= 8 && warp < 12) {
    // ================================================================= producer (cp.async, 4 warps): the long layers - many
    // row tiles per CTA - stream faster this way than through one TMA thread (measured: 214 vs 291 us on 3.2 M x 16 -> 96)
    const int tp = threadIdx.x - 256;
    const int c16 = tp & 7, r0 = tp >> 3;
    uint32_t it = 0;
    for (int64_t mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      const int64_t m0 = mt * DG_BM;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
        const uint32_t a_hi = smem_u32(smem + s * STAGE);
        const uint32_t a_lo = a_hi + A_BYTES, b_s = a_lo + A_BYTES;
        const int co = kb * DG_KE + c16 * 8;           // 8 bf16 = 16 bytes; cout % 8 == 0
        // one 64-bit base per stage, 32-bit row steps (the producers' instruction count is the stage latency)
        const int64_t abase = (m0 + r0) * cout + co;
        const int rows_left = (int)min((int64_t)DG_BM, M - m0);
        const bool co_ok = co < cout;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 16 * i;
          const bool v = (r < rows_left) && co_ok;
          const int64_t off = v ? abase + i * (16 * cout) : 0;
          const uint32_t d = sw128_offset(r, c16);
          cp_async_zfill<16>(a_hi + d, dz_hi + off, v);
          cp_async_zfill<16>(a_lo + d, dz_lo + off, v);
        }
        for (int idx = tp; idx < n_eff * 8; idx += 128) {
          const int r = idx >> 3, c = idx & 7;
          const int cc = kb * DG_KE + c * 8;
          const bool v = (r < n_valid) && (cc < cout);
          cp_async_zfill<16>(b_s + sw128_offset(r, c), wt + (v ? (n0 + r) * cout + cc : 0), v);   // K*cout < 2^31
        }
        cp_async_commit();
        if (it >= (uint32_t)BW_LAG) {
          cp_async_wait<BW_LAG>();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) mbar_arrive(&full_bar[(it - BW_LAG) % STAGES]);
        }
      }
    }
    cp_async_wait<0>();
    fence_proxy_async();
    __syncwarp();
    if (lane == 0)
      for (uint32_t j = (it > (uint32_t)BW_LAG ? it - BW_LAG : 0u); j < it; ++j) mbar_arrive(&full_bar[j % STAGES]);
  } else if (warp == 12) {
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
    // ================================================================= epilogue: both groups on every tile (even / odd 32-column
    // chunks): the 14x14 / 7x7 layers give a CTA one or two tiles, and a group per TMEM buffer left half the warps idle there
    const int grp = warp >> 2, wq4 = warp & 3;
    float* my = scratch + warp * BW_SCR;
    const float s_w = *w_scale_p;
    uint32_t tile_i = 0;
    for (int64_t mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++tile_i) {
      const uint32_t acc = tile_i & 1;
      const int64_t m0 = mt * DG_BM + wq4 * 32;
      mbar_wait(&tfull_bar[acc], (tile_i >> 1) & 1);
      tc_fence_after();
      for (int chunk = grp; chunk * 32 < n_valid; chunk += 2) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + acc * BN + chunk * 32 + ((uint32_t)(wq4 * 32) << 16), v);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj)
          *reinterpret_cast<float4*>(my + lane * 36 + 4 * jj) =
              make_float4(__uint_as_float(v[4 * jj]) * s_w, __uint_as_float(v[4 * jj + 1]) * s_w,
                          __uint_as_float(v[4 * jj + 2]) * s_w, __uint_as_float(v[4 * jj + 3]) * s_w);
        __syncwarp();
        const int col4 = (lane & 7) * 4;
        const int n = n0 + chunk * 32 + col4;
        if (n < K) {
          float* obase = dx + (m0 + (lane >> 3)) * K + n;                 // one 64-bit base per chunk, 32-bit row steps
          const int rows_left = (int)min((int64_t)32, M - m0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = (lane >> 3) + 4 * i;
            if (row < rows_left) {
              float4 val = *reinterpret_cast<const float4*>(my + row * 36 + col4);
              float4* dst = reinterpret_cast<float4*>(obase + i * (4 * K));
              if (accumulate) {
                const float4 o = ld_cg(dst);
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
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc<2 * BN>(tmem_base);
  }
}

template <int BN>
static int launch_dgrad_tc(const uint16_t* dz_hi, const uint16_t* dz_lo, const uint16_t* wt, const float* w_scale, int64_t M,
                           int K, int cout, float* dx, int accumulate, cudaStream_t st) {
  constexpr size_t smem = dg_smem_bytes<BN>();
  if (first_use_on_device(reinterpret_cast<const void*>(&pw_dgrad_tc_kernel<BN>))) {
    cudaError_t e = cudaFuncSetAttribute(pw_dgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("pw_dgrad_tc: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      (void)cudaGetLastError();
      return FROST_ECUDA;
    }
  }
  const int n_tiles = (K + BN - 1) / BN;
  const int64_t m_tiles = ceil_div(M, DG_BM);
  const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(m_tiles, kNumSMs / n_tiles));
  CUtensorMap tm_hi, tm_lo, tm_wt;
  if (!make_map_b16(&tm_hi, dz_hi, (uint64_t)cout, (uint64_t)M, (uint64_t)cout * 2, DG_KE, DG_BM) ||
      !make_map_b16(&tm_lo, dz_lo, (uint64_t)cout, (uint64_t)M, (uint64_t)cout * 2, DG_KE, DG_BM) ||
      !make_map_b16(&tm_wt, wt, (uint64_t)cout, (uint64_t)K, (uint64_t)cout * 2, DG_KE, BN)) {
    set_error("pw_dgrad_tc: cuTensorMapEncodeTiled failed (driver too old, or operands not 16-byte aligned)");
    return FROST_ENOSUP;
  }
  // short layers (a few row tiles per CTA: the 14x14 / 7x7 stages) are latency-bound: TMA keeps STAGES stages in flight
  const int use_tma = ceil_div(m_tiles, gx) <= 8 ? 1 : 0;
  launch_pdl(pw_dgrad_tc_kernel<BN>, dim3(gx, n_tiles), dim3(DG_THREADS), smem, st, tm_hi, tm_lo, tm_wt, dz_hi, dz_lo, wt, use_tma, w_scale, M, K,
             cout, dx, accumulate);
  return FROST_OK;
}

// ================================================================= wgrad
// GEMM with the reduction over the rows m: D[co][k] = sum_m A[co][m] * B[k][m].  Both operands are stored
// exactly as they sit in HBM (row m = 128-byte lines of co resp. k), which is the canonical MN-MAJOR
// SWIZZLE_128B layout: 8 (m) x 64 (co|k) bf16 atoms of 1024 B, the next 8 rows 1024 B further (SBO), the next
// 64 columns one block further (LBO).
constexpr int WG_ROWS = 64;       // reduction rows (m) per stage
constexpr int WG_BLK = 64 * 128;  // one 64-column block of a stage: 64 rows x 128 B
constexpr int WG_PROD_WARPS = 8;   // the producers convert x (u8 -> bf16) in registers: 4 warps were the per-stage critical path
constexpr int WG_THREADS = (4 + WG_PROD_WARPS + 1) * 32;   // warps 0-3 epilogue, 4-11 producer, 12 MMA

template <int BN>
__host__ __device__ constexpr int wg_stages() { return BN >= 256 ? 3 : 4; }
template <int BN>
constexpr size_t wg_smem_bytes() {
  return 1024 + (size_t)wg_stages<BN>() * (4 * WG_BLK + (BN / 64) * WG_BLK) + 4 * BW_SCR * 4 + (2 * wg_stages<BN>() + 2) * 8 + 16;
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

template <int BN>
__global__ void __launch_bounds__(WG_THREADS, 1) pw_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tm_hi,
                                                                   const __grid_constant__ CUtensorMap tm_lo,
                                                                   const uint8_t* xq,
                                                                   const float* x_scale_p,
                                                                   const int32_t* x_zp_p, int64_t M, int K, int ldx,
                                                                   int cout, int64_t rows_per_split, float* dwq) {
  constexpr int STAGES = wg_stages<BN>();
  constexpr int A_BYTES = 2 * WG_BLK;                       // 128 co x 64 rows, one of (hi, lo)
  constexpr int STAGE = 2 * A_BYTES + (BN / 64) * WG_BLK;   // hi, lo, BN k-columns
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* scratch = reinterpret_cast<float*>(smem + STAGES * STAGE);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(scratch + 4 * BW_SCR);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int co0 = blockIdx.x * 128;
  const int k0 = blockIdx.y * BN;
  const int n_valid = min(BN, K - k0);
  const int n_eff = (n_valid + 15) & ~15;
  const int64_t m_begin = (int64_t)blockIdx.z * rows_per_split;
  const int64_t m_end = min(M, m_begin + rows_per_split);
  const int num_kb = (int)((m_end - m_begin + WG_ROWS - 1) / WG_ROWS);

  if (threadIdx.x == 0) {
    // a stage is full when every producer WARP has stored its converted x chunks (one arrival per warp: 256 single-thread
    // arrivals on one mbarrier took ~0.5 us per stage) and the four dz boxes have landed
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], WG_PROD_WARPS + 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull_bar[0], 1);
    mbar_fence_init();
  }
  if (warp == 4 + WG_PROD_WARPS) tmem_alloc<BN>(tmem_slot);
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  const uint32_t tmem_base = *tmem_slot;

  if (num_kb > 0) {
    if (warp >= 4 && warp < 4 + WG_PROD_WARPS) {
      // ================================================================= producer
      constexpr int PT = WG_PROD_WARPS * 32;           // producer threads
      const int tp = threadIdx.x - 128;
      // byte -> float without I2F: 0x4B0000bb is the float 2^23 + bb; subtracting 2^23 + zp_a is exact
      const float zp_magic = 8388608.0f + (float)*x_zp_p;
      const int cpr = n_eff >> 3;                      // 8-channel chunks per row of x
      constexpr int XL = (WG_ROWS * (BN / 8) + PT - 1) / PT;   // x chunks per thread (<= 8)
      // stage-invariant coordinates of this thread's x chunks (an integer division per chunk and stage otherwise)
      int xoff[XL];          // element offset of the chunk inside a 64-row stage of x, -1: none
      uint32_t xdst[XL];     // its byte offset inside the B tile
      short xrow[XL];
#pragma unroll
      for (int q = 0; q < XL; ++q) {
        const int idx = tp + PT * q;
        const int r = idx / cpr, kc = idx - r * cpr;
        const bool has = idx < WG_ROWS * cpr && k0 + kc * 8 < K;
        xoff[q] = has ? r * ldx + k0 + kc * 8 : -1;      // x rows are ldx bytes apart (ldx == K: dense)
        xrow[q] = (short)r;
        xdst[q] = (uint32_t)(kc >> 3) * WG_BLK + sw128_offset(r, kc & 7);
        if (idx >= WG_ROWS * cpr) xrow[q] = -1;
      }
      // x of stage kb + 2 is requested while stage kb is converted (three register buffers, the loop unrolled by three): with
      // the loads issued in the iteration that consumes them every 64-row stage paid one DRAM round trip (~0.9 us), with one
      // stage of distance still most of it - an iteration is shorter than the round trip
      auto load_x = [&](uint2 (&dst)[XL], int kb) {
        const int64_t mb = m_begin + (int64_t)kb * WG_ROWS;
        const uint8_t* xstage = xq + mb * ldx;
        const int rows_left = (int)min((int64_t)WG_ROWS, m_end - mb);
#pragma unroll
        for (int q = 0; q < XL; ++q) {
          dst[q] = make_uint2(0u, 0u);
          if (kb < num_kb && xoff[q] >= 0 && xrow[q] < rows_left) dst[q] = ld_cg(reinterpret_cast<const uint2*>(xstage + xoff[q]));
        }
      };
      auto stage_body = [&](const uint2 (&cur)[XL], uint2 (&fut)[XL], int kb) {
        const int s = kb % STAGES;
        uint8_t* stage = smem + s * STAGE;
        const uint32_t a_hi = smem_u32(stage), a_lo = a_hi + A_BYTES;
        uint8_t* b_s = stage + 2 * A_BYTES;
        const int64_t mb = m_begin + (int64_t)kb * WG_ROWS;
        const int rows_left = (int)min((int64_t)WG_ROWS, m_end - mb);
        load_x(fut, kb + 2);
        mbar_wait(&empty_bar[s], ((kb / STAGES) & 1) ^ 1);
        // dz hi / lo: four TMA boxes (64 channels x 64 rows each) straight into the MN-major SWIZZLE_128B tiles; rows past the
        // end of the tensor and channels past cout arrive as zeros.  Up to STAGES stages of dz are in flight (the cp.async
        // producers this replaces kept two), and no load-queue entries are spent on them.
        if (tp == 0) {
          mbar_expect_tx(&full_bar[s], 4 * WG_BLK);
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
            tma_load_2d(&tm_hi, &full_bar[s], a_hi + blk * WG_BLK, co0 + 64 * blk, (int)mb);
            tma_load_2d(&tm_lo, &full_bar[s], a_lo + blk * WG_BLK, co0 + 64 * blk, (int)mb);
          }
        }
#pragma unroll
        for (int q = 0; q < XL; ++q) {
          if (xrow[q] >= 0) {
            const bool v = xoff[q] >= 0 && xrow[q] < rows_left;
            float w[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const unsigned word = e < 4 ? cur[q].x : cur[q].y;
              w[e] = v ? __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7650u + (e & 3))) - zp_magic : 0.0f;
            }
            *reinterpret_cast<uint4*>(b_s + xdst[q]) =
                make_uint4(pack_bf16(w[0], w[1]), pack_bf16(w[2], w[3]), pack_bf16(w[4], w[5]), pack_bf16(w[6], w[7]));
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_bar[s]);
      };
      uint2 xa[XL], xb[XL], xc[XL];
      load_x(xa, 0);
      load_x(xb, 1);
      for (int kb = 0; kb < num_kb; kb += 3) {
        stage_body(xa, xc, kb);
        if (kb + 1 < num_kb) stage_body(xb, xa, kb + 1);
        if (kb + 2 < num_kb) stage_body(xc, xb, kb + 2);
      }
    } else if (warp == 4 + WG_PROD_WARPS) {
      // ================================================================= MMA issuer (A and B MN-major)
      const uint32_t idesc = umma_idesc(1, 1, 1, 128, n_eff) | (1u << 15) | (1u << 16);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        mbar_wait(&full_bar[s], (kb / STAGES) & 1);
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
      float* my = scratch + warp * BW_SCR;
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
  if (warp == 4 + WG_PROD_WARPS) {
    tc_fence_after();
    tmem_dealloc<BN>(tmem_base);
  }
}

template <int BN>
static int launch_wgrad_tc(const uint16_t* dz_hi, const uint16_t* dz_lo, const uint8_t* xq, int ldx, const float* x_scale,
                           const int32_t* x_zp, int64_t M, int K, int cout, float* dwq, cudaStream_t st) {
  constexpr size_t smem = wg_smem_bytes<BN>();
  if (first_use_on_device(reinterpret_cast<const void*>(&pw_wgrad_tc_kernel<BN>))) {
    cudaError_t e = cudaFuncSetAttribute(pw_wgrad_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("pw_wgrad_tc: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return FROST_ECUDA;
    }
  }
  const int ct = (int)ceil_div(cout, 128), kt = (int)ceil_div(K, BN);
  int64_t splits = std::max<int64_t>(1, std::min<int64_t>(ceil_div(M, WG_ROWS * 4), (int64_t)kNumSMs / ((int64_t)ct * kt)));
  int64_t rows = ceil_div(ceil_div(M, splits), WG_ROWS) * WG_ROWS;
  splits = ceil_div(M, rows);
  CUtensorMap tm_hi, tm_lo;
  if (!make_map_b16(&tm_hi, dz_hi, (uint64_t)cout, (uint64_t)M, (uint64_t)cout * 2, 64, WG_ROWS) ||
      !make_map_b16(&tm_lo, dz_lo, (uint64_t)cout, (uint64_t)M, (uint64_t)cout * 2, 64, WG_ROWS)) {
    set_error("pw_wgrad_tc: cuTensorMapEncodeTiled failed (driver too old, or dz planes not 16-byte aligned)");
    return FROST_ENOSUP;
  }
  launch_pdl(pw_wgrad_tc_kernel<BN>, dim3(ct, kt, (unsigned)splits), dim3(WG_THREADS), smem, st, tm_hi, tm_lo, xq, x_scale, x_zp, M, K, ldx, cout, rows, dwq);
  return FROST_OK;
}

}  // namespace frost

using namespace frost;

extern "C" int frost_pw_dgrad_tc(const void* dz_hi, const void* dz_lo, const void* wt_bf16, const float* w_scale, int64_t M,
                                 int K, int cout, float* dx, int accumulate, void* stream) {
  FROST_REQUIRE(dz_hi && dz_lo && wt_bf16 && w_scale && dx, "frost_pw_dgrad_tc: null pointer");
  FROST_REQUIRE(M > 0 && K > 0 && cout > 0 && K % 4 == 0 && cout % 8 == 0,
                "frost_pw_dgrad_tc: K=%d must be a multiple of 4 and cout=%d of 8", K, cout);
  FROST_REQUIRE((int64_t)K * cout < ((int64_t)1 << 31), "frost_pw_dgrad_tc: weight tensor K*cout must be < 2^31 elements");
  FROST_REQUIRE((reinterpret_cast<uintptr_t>(dz_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(dz_lo) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(dx) & 15) == 0 && (reinterpret_cast<uintptr_t>(wt_bf16) & 15) == 0,
                "frost_pw_dgrad_tc: operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const uint16_t* hi = static_cast<const uint16_t*>(dz_hi);
  const uint16_t* lo = static_cast<const uint16_t*>(dz_lo);
  const uint16_t* wt = static_cast<const uint16_t*>(wt_bf16);
  int rc;
  if (K <= 64) rc = launch_dgrad_tc<64>(hi, lo, wt, w_scale, M, K, cout, dx, accumulate, st);
  else if (K <= 128) rc = launch_dgrad_tc<128>(hi, lo, wt, w_scale, M, K, cout, dx, accumulate, st);
  else rc = launch_dgrad_tc<256>(hi, lo, wt, w_scale, M, K, cout, dx, accumulate, st);
  if (rc) return rc;
  FROST_LAUNCH_CHECK("pw_dgrad_tc");
  return FROST_OK;
}

extern "C" int frost_pw_wgrad_tc(const void* dz_hi, const void* dz_lo, const uint8_t* xq, int ldx, const float* x_scale,
                                 const int32_t* x_zp, int64_t M, int K, int cout, float* dwq, void* stream) {
  FROST_REQUIRE(dz_hi && dz_lo && xq && x_scale && x_zp && dwq, "frost_pw_wgrad_tc: null pointer");
  FROST_REQUIRE(ldx >= K && ldx % 8 == 0, "frost_pw_wgrad_tc: ldx=%d must be >= K and a multiple of 8", ldx);
  FROST_REQUIRE(M > 0 && K > 0 && cout > 0 && K % 8 == 0 && cout % 8 == 0, "frost_pw_wgrad_tc: K and cout must be multiples of 8");
  FROST_REQUIRE((reinterpret_cast<uintptr_t>(dz_hi) & 15) == 0 && (reinterpret_cast<uintptr_t>(dz_lo) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(dwq) & 15) == 0 && (reinterpret_cast<uintptr_t>(xq) & 7) == 0,
                "frost_pw_wgrad_tc: operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(dwq, 0, sizeof(float) * (size_t)K * cout, st) != cudaSuccess) {
    set_error("frost_pw_wgrad_tc: memset failed");
    return FROST_ECUDA;
  }
  const uint16_t* hi = static_cast<const uint16_t*>(dz_hi);
  const uint16_t* lo = static_cast<const uint16_t*>(dz_lo);
  int rc;
  if (K <= 64) rc = launch_wgrad_tc<64>(hi, lo, xq, ldx, x_scale, x_zp, M, K, cout, dwq, st);
  else if (K <= 128) rc = launch_wgrad_tc<128>(hi, lo, xq, ldx, x_scale, x_zp, M, K, cout, dwq, st);
  else rc = launch_wgrad_tc<256>(hi, lo, xq, ldx, x_scale, x_zp, M, K, cout, dwq, st);
  if (rc) return rc;
  FROST_LAUNCH_CHECK("pw_wgrad_tc");
  return FROST_OK;
}
