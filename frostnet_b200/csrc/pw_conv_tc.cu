// pw_conv_tc.cu - the 1x1 (pointwise) convolution of the Frost bottleneck on 5th-gen tensor cores.
//   I[m][co] = sum_k (q_a[m][k] - zp_a) * (q_w[co][k] - zp_w)      (int32, exact)
// == the F.conv2d at torch/ao/nn/intrinsic/qat/modules/conv_fused.py:155 for the squeeze / expand /
// reduce / last_layer convs of frostnet.py:98-119,293, restated on quantize indices.
//
// tcgen05.mma kind::i8 (u8 x s8 -> s32), accumulators in TMEM (double buffered), operands staged in
// shared memory in the canonical K-major SWIZZLE_128B layout.  Warp-specialised persistent CTA:
//   warps 0-15 epilogue : tcgen05.ld -> zero-point correction -> smem transpose -> coalesced int32 stores
//                         + per-channel integer statistics (sum, sum^2, min, max) kept in smem for the
//                         CTA's lifetime, flushed once with integer atomics.  Four groups of 4 warps: group g
//                         drains accumulator buffer g&1 and the 32-column chunks of parity g>>1.  The epilogue
//                         (~26 instructions per output with the statistics) is latency-bound, so it gets 16 of
//                         the CTA's 21 warps (8 warps: 475 us on the 16->96 112x112 layer)
//   warps 16-19 producer: cp.async (LDGSTS, zero-filled) global -> swizzled smem, 3-4-stage mbarrier ring.
//                         (NHWC rows are K bytes apart and K % 16 != 0 for half the layers (24, 40, 56, 72,
//                         104, ...), which a TMA tensor map cannot describe; 8/16-byte cp.async can.)
//   warp 20    MMA      : one elected lane issues tcgen05.mma, tcgen05.commit releases stages / signals TMEM
// Weight zero-points: per_tensor_symmetric gives zp_w = 0 (s8 operand).  One-signed weight tensors make the
// reference fall back to affine with zp_w = -128 or 127; then q_w - zp_w (or its negation) fits
// u8: the producer rewrites the bytes (xor 0x80 / 0x7f) and the MMA runs u8 x u8 with a sign in the epilogue.
#include "tc_common.cuh"

namespace frost {

using namespace tc;

struct TcStat {
  long long sum;
  unsigned long long sq;
  int mn, mx;
};

constexpr int TC_BM = 128;
constexpr int TC_BK = 128;  // bytes of K per stage (= one swizzle span)
template <int BN>
__host__ __device__ constexpr int tc_stages() { return BN >= 256 ? 3 : 4; }
constexpr int TC_LAG = 2;   // cp.async groups in flight per producer thread
constexpr int TC_EPI_WARPS = 16;
constexpr int TC_THREADS = (TC_EPI_WARPS + 5) * 32;   // warps 0-15 epilogue (4 groups), 16-19 producer, 20 MMA
constexpr int TC_SCR = 32 * 36;  // ints of transpose scratch per epilogue warp

template <int BN>
constexpr size_t tc_smem_bytes() {
  return 1024 + (size_t)tc_stages<BN>() * (TC_BM * TC_BK + BN * TC_BK) + TC_EPI_WARPS * TC_SCR * 4 + BN * 4 + BN * sizeof(TcStat) +
         (2 * tc_stages<BN>() + 4) * 8 + 16;
}

template <int BN, bool VEC16>
__global__ void __launch_bounds__(TC_THREADS, 1) pw_conv_fwd_tc_kernel(
    const uint8_t* xq, const int32_t* x_zp_p, const int8_t* wq,
    const int32_t* w_zp_p, const int32_t* wsum, int64_t M, int K, int cout,
    int32_t* acc_out, FrostChanStats* stats) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  constexpr int TC_STAGES = tc_stages<BN>();
  constexpr int A_BYTES = TC_BM * TC_BK, B_BYTES = BN * TC_BK, STAGE = A_BYTES + B_BYTES;
  int* scratch = reinterpret_cast<int*>(smem + TC_STAGES * STAGE);
  int* s_corr = scratch + TC_EPI_WARPS * TC_SCR;
  TcStat* s_stat = reinterpret_cast<TcStat*>(s_corr + BN);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_stat + BN);
  uint64_t* empty_bar = full_bar + TC_STAGES;
  uint64_t* tfull_bar = empty_bar + TC_STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // ---- one-time setup.  Everything up to pdl_wait() touches only this CTA's shared memory / TMEM and overlaps
  // the tail of the previous kernel in the stream (common.cuh: programmatic dependent launch).  Dependents are
  // released only after this CTA's TMEM allocation is complete: a dependent CTA that became resident on this SM and
  // allocated first would block our tcgen05.alloc while waiting for our grid - a deadlock.
  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full_bar[s], 128); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], TC_EPI_WARPS / 2); }
    mbar_fence_init();
  }
  if (warp == TC_EPI_WARPS + 4) tmem_alloc<2 * BN>(tmem_slot);
  pdl_wait();
  const int zp_a = *x_zp_p, zp_w = *w_zp_p;
  if (zp_w != 0 && zp_w != -128 && zp_w != 127) __trap();  // not reachable with ChooseQuantizationParams
  const uint32_t wxor = zp_w == 0 ? 0u : (zp_w == -128 ? 0x80808080u : 0x7f7f7f7fu);
  const int wsign = (zp_w == 127) ? -1 : 1;
  const int n0 = blockIdx.y * BN;
  const int n_valid = min(BN, cout - n0);
  const int n_eff = (n_valid + 15) & ~15;
  const int num_kb = (K + TC_BK - 1) / TC_BK;
  const int64_t m_tiles = (M + TC_BM - 1) / TC_BM;
  for (int j = threadIdx.x; j < BN; j += blockDim.x) {
    int corr = 0;
    if (j < n_valid) {
      const int ws = wsum[n0 + j];
      const int ws_eff = zp_w == 0 ? ws : (zp_w == -128 ? ws + 128 * K : 127 * K - ws);
      corr = zp_a * ws_eff;
    }
    s_corr[j] = corr;
    s_stat[j].sum = 0; s_stat[j].sq = 0; s_stat[j].mn = INT_MAX; s_stat[j].mx = INT_MIN;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  const uint32_t tmem_base = *tmem_slot;

  if (warp >= TC_EPI_WARPS && warp < TC_EPI_WARPS + 4) {
    // ================================================================= producer
    const int tp = threadIdx.x - TC_EPI_WARPS * 32;
    const int c16 = tp & 7, r0 = tp >> 3;  // 16 rows per pass
    uint32_t it = 0;
    for (int64_t mt = blockIdx.x; mt < m_tiles; mt += gridDim.x) {
      const int64_t m0 = mt * TC_BM;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % TC_STAGES;
        mbar_wait(&empty_bar[s], ((it / TC_STAGES) & 1) ^ 1);
        const uint32_t a_s = smem_u32(smem + s * STAGE);
        const uint32_t b_s = a_s + A_BYTES;
        const int kbyte = kb * TC_BK + c16 * 16;
        const uint8_t* abase = xq + (m0 + r0) * K + kbyte;      // one 64-bit base per stage, 32-bit row steps
        const int rows_left = (int)min((int64_t)TC_BM, M - m0);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = r0 + 16 * i;
          const bool rv = r < rows_left;
          const uint8_t* src = abase + i * (16 * K);             // only dereferenced when rv
          const uint32_t dst = a_s + sw128_offset(r, c16);
          if constexpr (VEC16) {
            const bool v = rv && kbyte < K;
            cp_async_zfill<16>(dst, v ? src : xq, v);
          } else {
            const bool v0 = rv && kbyte < K, v1 = rv && kbyte + 8 < K;
            cp_async_zfill<8>(dst, v0 ? src : xq, v0);
            cp_async_zfill<8>(dst + 8, v1 ? src + 8 : xq, v1);
          }
        }
        if (wxor == 0u) {
          for (int idx = tp; idx < n_eff * 8; idx += 128) {
            const int r = idx >> 3, c = idx & 7;
            const int kk = kb * TC_BK + c * 16;
            const bool rv = r < n_valid;
            const int8_t* src = wq + (n0 + (rv ? r : 0)) * K + kk;                   // cout*K < 2^31
            const uint32_t dst = b_s + sw128_offset(r, c);
            if constexpr (VEC16) {
              const bool v = rv && kk < K;
              cp_async_zfill<16>(dst, v ? (const void*)src : (const void*)wq, v);
            } else {
              const bool v0 = rv && kk < K, v1 = rv && kk + 8 < K;
              cp_async_zfill<8>(dst, v0 ? (const void*)src : (const void*)wq, v0);
              cp_async_zfill<8>(dst + 8, v1 ? (const void*)(src + 8) : (const void*)wq, v1);
            }
          }
        } else {
          for (int idx = tp; idx < n_eff * 8; idx += 128) {
            const int r = idx >> 3, c = idx & 7;
            const int kk = kb * TC_BK + c * 16;
            const bool rv = r < n_valid;
            const int8_t* src = wq + (n0 + (rv ? r : 0)) * K + kk;                   // cout*K < 2^31
            uint2 lo = make_uint2(0u, 0u), hi = make_uint2(0u, 0u);
            if (rv && kk < K) { lo = ld_cg(reinterpret_cast<const uint2*>(src)); lo.x ^= wxor; lo.y ^= wxor; }
            if (rv && kk + 8 < K) { hi = ld_cg(reinterpret_cast<const uint2*>(src + 8)); hi.x ^= wxor; hi.y ^= wxor; }
            *reinterpret_cast<uint4*>(smem + s * STAGE + A_BYTES + sw128_offset(r, c)) = make_uint4(lo.x, lo.y, hi.x, hi.y);
          }
        }
        cp_async_commit();
        if (it >= (uint32_t)TC_LAG) {
          cp_async_wait<TC_LAG>();
          fence_proxy_async();
          mbar_arrive(&full_bar[(it - TC_LAG) % TC_STAGES]);
        }
      }
    }
    // drain
    cp_async_wait<0>();
    fence_proxy_async();
    for (uint32_t j = (it > (uint32_t)TC_LAG ? it - TC_LAG : 0u); j < it; ++j) mbar_arrive(&full_bar[j % TC_STAGES]);
  } else if (warp == TC_EPI_WARPS + 4) {
    // ================================================================= MMA issuer
    const uint32_t idesc = umma_idesc(2 /*S32*/, 0 /*A: u8*/, zp_w == 0 ? 1 : 0 /*B: s8 | u8*/, TC_BM, n_eff);
    uint32_t it = 0, tile_i = 0;
    for (int64_t mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++tile_i) {
      const uint32_t acc = tile_i & 1;
      mbar_wait(&tempty_bar[acc], ((tile_i >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const int s = it % TC_STAGES;
        mbar_wait(&full_bar[s], (it / TC_STAGES) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_s = smem_u32(smem + s * STAGE);
          const uint64_t adesc = umma_desc_sw128(a_s), bdesc = umma_desc_sw128(a_s + A_BYTES);
          const int nk = min(TC_BK / 32, (K - kb * TC_BK + 31) / 32);
          for (int k4 = 0; k4 < nk; ++k4)
            umma_i8(d_tmem, adesc + (uint64_t)(2 * k4), bdesc + (uint64_t)(2 * k4), idesc, (kb | k4) != 0 ? 1u : 0u);
          umma_commit(&empty_bar[s]);
          if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // ================================================================= epilogue
    // four groups of 4 warps; group g drains accumulator buffer g&1 (tiles g&1, (g&1)+2, ...) and the chunks of
    // parity g>>1; warp w of a group owns TMEM lanes 32*(w%4).. (the hardware restricts a warp to the lane quarter
    // warp_id % 4)
    const int grp = (warp >> 2) & 1, cpar = warp >> 3, wq4 = warp & 3;
    int* my = scratch + warp * TC_SCR;
    uint32_t tile_i = 0;
    for (int64_t mt = blockIdx.x; mt < m_tiles; mt += gridDim.x, ++tile_i) {
      const uint32_t acc = tile_i & 1;
      if ((int)acc != grp) continue;
      const int64_t m0 = mt * TC_BM + wq4 * 32;
      mbar_wait(&tfull_bar[acc], (tile_i >> 1) & 1);
      tc_fence_after();
      for (int chunk = cpar; chunk * 32 < n_valid; chunk += 2) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + acc * BN + chunk * 32 + ((uint32_t)(wq4 * 32) << 16), v);
#pragma unroll
        for (int jj = 0; jj < 8; ++jj) {
          int4 o;
          o.x = wsign * ((int)v[4 * jj + 0] - s_corr[chunk * 32 + 4 * jj + 0]);
          o.y = wsign * ((int)v[4 * jj + 1] - s_corr[chunk * 32 + 4 * jj + 1]);
          o.z = wsign * ((int)v[4 * jj + 2] - s_corr[chunk * 32 + 4 * jj + 2]);
          o.w = wsign * ((int)v[4 * jj + 3] - s_corr[chunk * 32 + 4 * jj + 3]);
          *reinterpret_cast<int4*>(my + lane * 36 + 4 * jj) = o;
        }
        __syncwarp();
        const int col4 = (lane & 7) * 4;
        const int n = n0 + chunk * 32 + col4;
        long long s4[4] = {0, 0, 0, 0};
        unsigned long long q4[4] = {0, 0, 0, 0};
        int mn4[4] = {INT_MAX, INT_MAX, INT_MAX, INT_MAX}, mx4[4] = {INT_MIN, INT_MIN, INT_MIN, INT_MIN};
        if (n < cout) {
          int32_t* obase = acc_out + (m0 + (lane >> 3)) * cout + n;      // one 64-bit base per chunk, 32-bit row steps
          const int rows_left = (int)min((int64_t)32, M - m0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = (lane >> 3) + 4 * i;
            if (row < rows_left) {
              const int4 val = *reinterpret_cast<const int4*>(my + row * 36 + col4);
              *reinterpret_cast<int4*>(obase + i * (4 * cout)) = val;
              const int e[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                s4[c] += e[c];
                q4[c] += (unsigned long long)((long long)e[c] * (long long)e[c]);
                mn4[c] = min(mn4[c], e[c]);
                mx4[c] = max(mx4[c], e[c]);
              }
            }
          }
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
#pragma unroll
          for (int o = 8; o <= 16; o <<= 1) {
            s4[c] += __shfl_xor_sync(0xffffffffu, s4[c], o);
            q4[c] += __shfl_xor_sync(0xffffffffu, q4[c], o);
            mn4[c] = min(mn4[c], __shfl_xor_sync(0xffffffffu, mn4[c], o));
            mx4[c] = max(mx4[c], __shfl_xor_sync(0xffffffffu, mx4[c], o));
          }
          if (lane < 8 && n < cout && mn4[c] <= mx4[c]) {
            TcStat* st = &s_stat[chunk * 32 + col4 + c];
            atomicAdd(reinterpret_cast<unsigned long long*>(&st->sum), (unsigned long long)s4[c]);
            atomicAdd(&st->sq, q4[c]);
            atomicMin(&st->mn, mn4[c]);
            atomicMax(&st->mx, mx4[c]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  // ---- teardown: flush the CTA's statistics, release TMEM
  tc_fence_before();
  __syncthreads();
  for (int j = threadIdx.x; j < n_valid; j += blockDim.x) {
    const TcStat s = s_stat[j];
    chan_stats_flush(stats + n0 + j, s.sum, s.sq, s.mn, s.mx);
  }
  if (warp == TC_EPI_WARPS + 4) {
    tc_fence_after();
    tmem_dealloc<2 * BN>(tmem_base);
  }
}

template <int BN, bool VEC16>
static int launch_tc(const uint8_t* xq, const int32_t* x_zp, const int8_t* wq, const int32_t* w_zp, const int32_t* wsum,
                     int64_t M, int K, int cout, int32_t* acc, FrostChanStats* stats, cudaStream_t st) {
  constexpr size_t smem = tc_smem_bytes<BN>();
  if (first_use_on_device(reinterpret_cast<const void*>(&pw_conv_fwd_tc_kernel<BN, VEC16>))) {
    cudaError_t e = cudaFuncSetAttribute(pw_conv_fwd_tc_kernel<BN, VEC16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("pw_conv_fwd_tc: cudaFuncSetAttribute(%zu) failed: %s", smem, cudaGetErrorString(e));
      return FROST_ECUDA;
    }
  }
  const int n_tiles = (cout + BN - 1) / BN;
  const int64_t m_tiles = ceil_div(M, TC_BM);
  const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(m_tiles, kNumSMs / n_tiles));
  // the 8-byte cp.async of the K % 16 != 0 variant allocates in L1 (cp.async.ca): ordinary launch (common.cuh)
  launch_pdl_if(VEC16, pw_conv_fwd_tc_kernel<BN, VEC16>, dim3(gx, n_tiles), dim3(TC_THREADS), smem, st, xq, x_zp, wq, w_zp, wsum, M, K, cout, acc, stats);
  return FROST_OK;
}

}  // namespace frost

using namespace frost;

extern "C" int frost_pw_conv_forward(const uint8_t* xq, const int32_t* x_zp, const int8_t* wq, const int32_t* w_zp,
                                     const int32_t* wsum, int64_t M, int K, int cout, int32_t* acc,
                                     FrostChanStats* stats, void* stream) {
  FROST_REQUIRE(xq && x_zp && wq && w_zp && wsum && acc && stats, "frost_pw_conv_forward: null pointer");
  FROST_REQUIRE(M > 0 && K > 0 && cout > 0, "frost_pw_conv_forward: empty problem M=%lld K=%d cout=%d", (long long)M, K, cout);
  FROST_REQUIRE((int64_t)K * cout < ((int64_t)1 << 31), "frost_pw_conv_forward: weight tensor K*cout must be < 2^31 elements");
  FROST_REQUIRE(K % 8 == 0 && cout % 4 == 0, "frost_pw_conv_forward: K=%d must be a multiple of 8 and cout=%d of 4", K, cout);
  FROST_REQUIRE((reinterpret_cast<uintptr_t>(xq) & 15) == 0 && (reinterpret_cast<uintptr_t>(wq) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(acc) & 15) == 0,
                "frost_pw_conv_forward: operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const bool v16 = (K % 16 == 0);
  int rc;
  if (cout <= 64) rc = v16 ? launch_tc<64, true>(xq, x_zp, wq, w_zp, wsum, M, K, cout, acc, stats, st)
                           : launch_tc<64, false>(xq, x_zp, wq, w_zp, wsum, M, K, cout, acc, stats, st);
  else if (cout <= 128) rc = v16 ? launch_tc<128, true>(xq, x_zp, wq, w_zp, wsum, M, K, cout, acc, stats, st)
                                 : launch_tc<128, false>(xq, x_zp, wq, w_zp, wsum, M, K, cout, acc, stats, st);
  else rc = v16 ? launch_tc<256, true>(xq, x_zp, wq, w_zp, wsum, M, K, cout, acc, stats, st)
                : launch_tc<256, false>(xq, x_zp, wq, w_zp, wsum, M, K, cout, acc, stats, st);
  if (rc) return rc;
  FROST_LAUNCH_CHECK("pw_conv_fwd_tc");
  return FROST_OK;
}
