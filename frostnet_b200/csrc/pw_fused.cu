// pw_fused.cu - the 1x1 (pointwise) ConvBn(ReLU)2d of the Frost bottleneck as ONE kernel per direction and pass:
// convolution on tcgen05 tensor cores + BatchNorm (training-mode batch statistics) + ReLU + activation fake-quant
// in the epilogue, and the matching backward passes - none of which ever stores the int32 accumulator I.
// Replaces, for the squeeze / expand / reduce / last_layer convs of frostnet.py:98-119,293, the chain
//   aten::convolution -> div -> native_batch_norm -> relu -> fused_moving_avg_obs_fake_quant   (conv_fused.py:131-167,708-710)
// and the first two stages of its autograd (SURVEY.md 8a' steps 1-3).
//
// Orientation: D[channel][pixel] = W[channel][k] * X[pixel][k]^T.  The weight tile is the UMMA A operand (M = 128 output
// channels -> the 128 TMEM lanes), the activation tile the B operand (N = 256 pixels -> TMEM columns).  A thread of the
// epilogue therefore owns ONE OUTPUT CHANNEL for the CTA's lifetime: the BatchNorm statistics, the per-channel affine and
// the backward coefficients live in its registers, the per-channel reductions (sum I, sum I^2, min, max; sum dv, sum dv*xhat)
// need no shuffles and no shared memory, and global accesses to [pixel][channel] gradients are coalesced across the warp.
//
//   forward   phase A: GEMM -> per-channel integer statistics (registers) -> integer atomics
//             grid barrier (all CTAs resident: grid <= 148, 1 CTA/SM) -> every CTA finalises BN / observer / qparams
//             phase B: the same GEMM again (the operands are 1/6..1x the output bytes; the tensor pipe is idle anyway)
//                      -> A_c*I+B_c -> ReLU -> quantise -> uint8 tile in shared memory -> TMA store
//             With frozen BatchNorm and the observer off (eval / late QAT) phase A and the barrier are skipped.
//   backward  reduce: GEMM -> dv = dy*mask -> S1, S2 per channel (registers) -> fp64 atomics
//             apply : GEMM -> dz = c1*(dv - a0 - a1*(I - mean)) -> bf16 hi/lo planes (operands of the tensor-core dgrad/wgrad)
//
// Operand staging: TMA (cp.async.bulk.tensor, SWIZZLE_128B, zero fill out of bounds) when the row pitches are multiples
// of 16 bytes - one elected thread, the mbarrier counts bytes; otherwise (dense NHWC rows with C % 16 != 0) the
// cp.async path of the first-generation kernel (pw_conv_tc.cu).  Weights stay resident in shared memory when K <= 384.
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <cstring>
#include "tc_common.cuh"
#include "bn_math.cuh"

namespace frost {

using namespace tc;

constexpr int PF_NPX = 256;          // pixels per tile (UMMA N)
constexpr int PF_CH = 128;           // channel rows per tile (UMMA M)
constexpr int PF_BK = 128;           // bytes of K per k-block (one swizzle span)
constexpr int PF_RES_KB = 3;         // weight k-blocks that stay resident
constexpr int PF_EPI_WARPS = 8;      // two per TMEM lane quarter: columns [0,128) and [128,256)
constexpr int PF_EPI_THREADS = PF_EPI_WARPS * 32;
constexpr int PF_PROD_WARPS = 4;
constexpr int PF_THREADS = (PF_EPI_WARPS + PF_PROD_WARPS + 1) * 32;   // warps 0-7 epilogue, 8-11 producer, 12 MMA
constexpr int PF_W_BYTES = PF_CH * PF_BK;       // 16 KB
constexpr int PF_X_BYTES = PF_NPX * PF_BK;      // 32 KB
constexpr int PF_STAGING = PF_NPX * PF_CH;      // 32 KB uint8 output tile
constexpr int PF_TAIL = 8192;                   // barriers + per-channel combine buffers

enum { PF_FWD = 0, PF_BWD_REDUCE = 1, PF_BWD_APPLY = 2 };

struct PwFusedParams {
  const uint8_t* x;
  int64_t M;
  int K, ldx;
  const int8_t* w;        // MMA-ready weight bytes [cout][ldw]
  int ldw, cout, bn, n_kb;
  const int32_t* x_zp;
  const int32_t* w_zp;
  const int32_t* wsum;
  int tma_in, tma_out, vec16;
  // forward
  FrostBnFinalizeArgs fin;
  unsigned* grid_bar;
  uint8_t* q;
  int ldq;
  // backward
  FrostBnBackwardArgs bwd;
};

struct PfCombine {       // per channel of the tile: the two column halves meet here
  unsigned long long sum, sq;
  int mn, mx;
  double s1, s2;
};

// ---------------------------------------------------------------- TMA / mbarrier helpers
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* tm, uint64_t* bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tm, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(src), "r"(c0), "r"(c1)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(PF_EPI_THREADS) : "memory"); }

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ================================================================= the kernel
template <int MODE>
__global__ void __launch_bounds__(PF_THREADS, 1) pw_fused_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                                const __grid_constant__ CUtensorMap tm_w,
                                                                const __grid_constant__ CUtensorMap tm_q,
                                                                const __grid_constant__ PwFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const bool resident = p.n_kb <= PF_RES_KB;
  const int S = resident ? 4 : 3;                                    // pipeline stages
  const int stage_bytes = resident ? PF_X_BYTES : PF_X_BYTES + PF_W_BYTES;
  uint8_t* wres = smem;                                              // [n_kb][128 rows][128 B] when resident
  uint8_t* stages = smem + (resident ? PF_RES_KB * PF_W_BYTES : 0);  // [S][x 32 KB (+ w 16 KB)]
  uint8_t* staging = stages + S * stage_bytes;                       // forward: [256 px][bn] uint8
  uint8_t* tail = staging + (MODE == PF_FWD ? PF_STAGING : 0);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);            // [4]
  uint64_t* empty_bar = full_bar + 4;                                // [4]
  uint64_t* tfull_bar = empty_bar + 4;                               // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                              // [2]
  uint64_t* wfull_bar = tempty_bar + 2;                              // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(wfull_bar + 1);
  float* s_red = reinterpret_cast<float*>(tmem_slot + 2);            // [2 * PF_EPI_WARPS] block min / max
  float* s_qp = s_red + 2 * PF_EPI_WARPS;                            // scale, zero point (as float), 1/scale, spare
  PfCombine* s_comb = reinterpret_cast<PfCombine*>(tail + 256);      // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ct = blockIdx.y;
  const int c_tile0 = ct * p.bn;
  const int n_valid = min(p.bn, p.cout - c_tile0);
  const int64_t n_ptiles = (p.M + PF_NPX - 1) / PF_NPX;
  const int prod_arrivals = p.tma_in ? 1 : PF_PROD_WARPS * 32;

  // ---- one-time setup (overlaps the tail of the previous kernel: programmatic dependent launch, common.cuh)
  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) { mbar_init(&full_bar[s], prod_arrivals); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], PF_EPI_WARPS); }
    mbar_init(wfull_bar, prod_arrivals);
    mbar_fence_init();
  }
  if (warp == PF_EPI_WARPS + PF_PROD_WARPS) tmem_alloc<512>(tmem_slot);
  if (threadIdx.x < PF_CH) {
    PfCombine& c = s_comb[threadIdx.x];
    c.sum = 0; c.sq = 0; c.mn = INT_MAX; c.mx = INT_MIN; c.s1 = 0.0; c.s2 = 0.0;
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  const uint32_t tmem_base = *tmem_slot;

  const int zp_a = *p.x_zp, zp_w = *p.w_zp;
  if (zp_w != 0 && zp_w != -128 && zp_w != 127) __trap();  // not reachable with ChooseQuantizationParams
  // forward: does this launch need batch statistics / the observer's min-max (phase A + grid barrier)?
  const bool need_stats = (MODE == PF_FWD) && (p.fin.training || p.fin.observe);
  const int n_phases = (MODE == PF_FWD && need_stats) ? 2 : 1;

  if (warp >= PF_EPI_WARPS && warp < PF_EPI_WARPS + PF_PROD_WARPS) {
    // ================================================================= producer
    const int tp = threadIdx.x - PF_EPI_WARPS * 32;
    if (p.tma_in) {
      if (tp == 0) {
        tma_prefetch_desc(&tm_x);
        tma_prefetch_desc(&tm_w);
        if (resident) {
          mbar_expect_tx(wfull_bar, (uint32_t)(p.n_kb * PF_W_BYTES));
          for (int kb = 0; kb < p.n_kb; ++kb) tma_load_2d(&tm_w, wfull_bar, smem_u32(wres + kb * PF_W_BYTES), kb * PF_BK, c_tile0);
        }
        uint32_t it = 0;
        for (int ph = 0; ph < n_phases; ++ph) {
          for (int64_t pt = blockIdx.x; pt < n_ptiles; pt += gridDim.x) {
            for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
              const int s = it % S;
              mbar_wait(&empty_bar[s], ((it / S) & 1) ^ 1);
              uint8_t* st = stages + s * stage_bytes;
              mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
              tma_load_2d(&tm_x, &full_bar[s], smem_u32(st), kb * PF_BK, (int)(pt * PF_NPX));
              if (!resident) tma_load_2d(&tm_w, &full_bar[s], smem_u32(st + PF_X_BYTES), kb * PF_BK, c_tile0);
            }
          }
        }
      }
    } else {
      // cp.async (LDGSTS) with zero fill into the same SWIZZLE_128B tiles
      const int c16 = tp & 7, r0 = tp >> 3;            // 16 rows per pass
      auto load_w = [&](uint8_t* dst_tile, int kb) {
        const uint32_t dst0 = smem_u32(dst_tile);
        const int kk = kb * PF_BK + c16 * 16;
#pragma unroll
        for (int i = 0; i < PF_CH / 16; ++i) {
          const int r = r0 + 16 * i;
          const bool v = (r < n_valid) && (kk < p.K);     // ldw is a multiple of 16 and its padding bytes are zero
          cp_async_zfill<16>(dst0 + sw128_offset(r, c16), v ? (const void*)(p.w + (int64_t)(c_tile0 + r) * p.ldw + kk) : (const void*)p.w, v);
        }
      };
      if (resident) {
        for (int kb = 0; kb < p.n_kb; ++kb) load_w(wres + kb * PF_W_BYTES, kb);
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async();
        mbar_arrive(wfull_bar);
      }
      constexpr int LAG = 2;
      uint32_t it = 0;
      for (int ph = 0; ph < n_phases; ++ph) {
        for (int64_t pt = blockIdx.x; pt < n_ptiles; pt += gridDim.x) {
          const int64_t m0 = pt * PF_NPX;
          const int rows_left = (int)min((int64_t)PF_NPX, p.M - m0);
          for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
            const int s = it % S;
            mbar_wait(&empty_bar[s], ((it / S) & 1) ^ 1);
            uint8_t* st = stages + s * stage_bytes;
            const uint32_t a_s = smem_u32(st);
            const int kbyte = kb * PF_BK + c16 * 16;
            const uint8_t* abase = p.x + (m0 + r0) * p.ldx + kbyte;
#pragma unroll
            for (int i = 0; i < PF_NPX / 16; ++i) {
              const int r = r0 + 16 * i;
              const bool rv = r < rows_left;
              const uint8_t* src = abase + (int64_t)i * 16 * p.ldx;
              const uint32_t dst = a_s + sw128_offset(r, c16);
              if (p.vec16) {
                const bool v = rv && kbyte < p.K;
                cp_async_zfill<16>(dst, v ? src : p.x, v);
              } else {
                const bool v0 = rv && kbyte < p.K, v1 = rv && kbyte + 8 < p.K;
                cp_async_zfill<8>(dst, v0 ? src : p.x, v0);
                cp_async_zfill<8>(dst + 8, v1 ? src + 8 : p.x, v1);
              }
            }
            if (!resident) load_w(st + PF_X_BYTES, kb);
            cp_async_commit();
            if (it >= (uint32_t)LAG) {
              cp_async_wait<LAG>();
              fence_proxy_async();
              mbar_arrive(&full_bar[(it - LAG) % S]);
            }
          }
        }
      }
      cp_async_wait<0>();
      fence_proxy_async();
      for (uint32_t j = (it > (uint32_t)LAG ? it - LAG : 0u); j < it; ++j) mbar_arrive(&full_bar[j % S]);
    }
  } else if (warp == PF_EPI_WARPS + PF_PROD_WARPS) {
    // ================================================================= MMA issuer
    // A = weights (s8, or u8 after the producer-side zero-point rewrite done by weight prep), B = activations (u8)
    const uint32_t idesc = umma_idesc(2 /*S32*/, zp_w == 0 ? 1 : 0, 0, PF_CH, PF_NPX);
    if (resident) {
      mbar_wait(wfull_bar, 0);
      tc_fence_after();
    }
    uint32_t it = 0, tile_i = 0;
    for (int ph = 0; ph < n_phases; ++ph) {
      for (int64_t pt = blockIdx.x; pt < n_ptiles; pt += gridDim.x, ++tile_i) {
        const uint32_t acc = tile_i & 1;
        mbar_wait(&tempty_bar[acc], ((tile_i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * PF_NPX;
        for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
          const int s = it % S;
          mbar_wait(&full_bar[s], (it / S) & 1);
          tc_fence_after();
          if (lane == 0) {
            uint8_t* st = stages + s * stage_bytes;
            const uint64_t bdesc = umma_desc_sw128(smem_u32(st));
            const uint64_t adesc = umma_desc_sw128(smem_u32(resident ? wres + kb * PF_W_BYTES : st + PF_X_BYTES));
            const int nk = min(PF_BK / 32, (p.K - kb * PF_BK + 31) / 32);
            for (int k4 = 0; k4 < nk; ++k4)
              umma_i8(d_tmem, adesc + (uint64_t)(2 * k4), bdesc + (uint64_t)(2 * k4), idesc, (kb | k4) != 0 ? 1u : 0u);
            umma_commit(&empty_bar[s]);
            if (kb == p.n_kb - 1) umma_commit(&tfull_bar[acc]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================================================================= epilogue: thread <-> output channel
    const int quarter = warp & 3, half = warp >> 2;
    const int c_local = quarter * 32 + lane;
    const bool active = c_local < n_valid;
    const bool warp_active = quarter * 32 < n_valid;
    const int c = c_tile0 + (active ? c_local : 0);
    const int tid = threadIdx.x;                         // 0..255 within the epilogue
    // I = wsign * (raw - corr): activation zero point folded through the (rewritten) weight row sum
    const int wsign = (zp_w == 127) ? -1 : 1;
    int corr_s;                                          // wsign * corr
    {
      const int ws = p.wsum[c];
      const int ws_eff = zp_w == 0 ? ws : (zp_w == -128 ? ws + 128 * p.K : 127 * p.K - ws);
      corr_s = wsign * zp_a * ws_eff;
    }
    uint32_t tile_i = 0;

    // per-thread channel state
    float cA = 0.f, cB = 0.f, cMean = 0.f;
    float inv = 1.f, zpf = 0.f;
    int relu = 0;

    if constexpr (MODE == PF_FWD) {
      relu = p.fin.relu;
      const double Mcount = (double)p.fin.count;
      const double sa_sw = (double)(*p.fin.x_scale) * (double)(*p.fin.w_scale);
      const double mom = p.fin.momentum >= 0.0f ? (double)p.fin.momentum
                                                : 1.0 / (double)((p.fin.num_batches_tracked ? *p.fin.num_batches_tracked : 0) + 1);
      // the observer state as it was BEFORE this layer ran: every CTA derives the same new state from it after the barrier
      float rmin = *p.fin.afq.min_val, rmax = *p.fin.afq.max_val;
      if (need_stats) {
        // ---------------- phase A: statistics
        long long sum = 0;
        unsigned long long sq = 0;
        int mn = INT_MAX, mx = INT_MIN;
        for (int64_t pt = blockIdx.x; pt < n_ptiles; pt += gridDim.x, ++tile_i) {
          const uint32_t acc = tile_i & 1;
          const int px_valid = (int)min((int64_t)PF_NPX, p.M - pt * PF_NPX);
          mbar_wait(&tfull_bar[acc], (tile_i >> 1) & 1);
          tc_fence_after();
          if (warp_active) {
#pragma unroll 1
            for (int ch = 0; ch < 4; ++ch) {
              const int col0 = half * 128 + ch * 32;
              if (col0 >= px_valid) break;
              uint32_t v[32];
              tmem_ld_32x32(tmem_base + acc * PF_NPX + col0 + ((uint32_t)(quarter * 32) << 16), v);
              const int nv = min(32, px_valid - col0);
              int s32 = 0;
              if (nv == 32) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const int I = wsign * (int)v[j] - corr_s;
                  s32 += I;
                  sq += (unsigned long long)((long long)I * (long long)I);
                  mn = min(mn, I);
                  mx = max(mx, I);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  if (j < nv) {
                    const int I = wsign * (int)v[j] - corr_s;
                    s32 += I;
                    sq += (unsigned long long)((long long)I * (long long)I);
                    mn = min(mn, I);
                    mx = max(mx, I);
                  }
                }
              }
              sum += s32;
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
        // the two column halves of a channel meet in shared memory, one set of integer atomics per channel and CTA
        if (active && mn <= mx) {
          PfCombine& cb = s_comb[c_local];
          atomicAdd(&cb.sum, (unsigned long long)sum);
          atomicAdd(&cb.sq, sq);
          atomicMin(&cb.mn, mn);
          atomicMax(&cb.mx, mx);
        }
        epi_bar_sync();
        if (active && half == 0) {
          const PfCombine cb = s_comb[c_local];
          chan_stats_flush(const_cast<FrostChanStats*>(p.fin.stats) + c, (long long)cb.sum, cb.sq, cb.mn, cb.mx);
        }
        // ---------------- grid barrier: every CTA of this launch is resident (host: grid <= #SMs, 1 CTA per SM)
        __threadfence();
        epi_bar_sync();
        if (tid == 0) {
          const unsigned target = gridDim.x * gridDim.y;
          __threadfence();                                   // cumulativity: the CTA's atomics above are ordered before the arrival
          atomicAdd(p.grid_bar, 1u);
          bool ok = false;
          for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
            if (ld_acquire_gpu(p.grid_bar) >= target) { ok = true; break; }
            __nanosleep(64);
          }
          if (!ok) __trap();
        }
        epi_bar_sync();
      }
      // ---------------- finalize: BN affine of my channel, observer over all channels, qparams (same result in every CTA)
      BnChannel mine;
      mine.A = mine.B = mine.mean_I = mine.kfac = 0.f;
      float gmn = INFINITY, gmx = -INFINITY;
      for (int cc = tid; cc < p.cout; cc += PF_EPI_THREADS) {
        FrostChanStats st;
        const FrostChanStats* g = p.fin.stats + cc;
        st.sum = __ldcg(&g->sum); st.sq_lo = __ldcg(&g->sq_lo); st.sq_hi = __ldcg(&g->sq_hi);
        st.min = __ldcg(&g->min); st.max = __ldcg(&g->max);
        if (!need_stats) { st.min = 0; st.max = 0; }
        const BnChannel r = bn_channel_finalize(st, 0, Mcount, p.fin.count > 1, sa_sw, p.fin.sf[cc], p.fin.gamma[cc], p.fin.beta[cc],
                                                p.fin.running_mean[cc], p.fin.running_var[cc], p.fin.eps, mom, p.fin.training, relu);
        gmn = fminf(gmn, r.v_lo);
        gmx = fmaxf(gmx, r.v_hi);
      }
      if (active) {
        FrostChanStats st;
        const FrostChanStats* g = p.fin.stats + c;
        st.sum = __ldcg(&g->sum); st.sq_lo = __ldcg(&g->sq_lo); st.sq_hi = __ldcg(&g->sq_hi);
        st.min = __ldcg(&g->min); st.max = __ldcg(&g->max);
        if (!need_stats) { st.min = 0; st.max = 0; }
        mine = bn_channel_finalize(st, 0, Mcount, p.fin.count > 1, sa_sw, p.fin.sf[c], p.fin.gamma[c], p.fin.beta[c],
                                   p.fin.running_mean[c], p.fin.running_var[c], p.fin.eps, mom, p.fin.training, relu);
      }
      gmn = warp_min(gmn);
      gmx = warp_max(gmx);
      if (lane == 0) { s_red[2 * warp] = gmn; s_red[2 * warp + 1] = gmx; }
      epi_bar_sync();          // also orders every thread's reads of the old running statistics before the writes below
      if (tid == 0) {
        for (int w = 0; w < PF_EPI_WARPS; ++w) { gmn = fminf(gmn, s_red[2 * w]); gmx = fmaxf(gmx, s_red[2 * w + 1]); }
        float s;
        int zp;
        if (p.fin.observe) {
          observer_ema(rmin, rmax, gmn, gmx, p.fin.averaging_const);
          choose_qparams(rmin, rmax, 0, 255, false, &s, &zp);
        } else {
          s = *p.fin.afq.scale;
          zp = *p.fin.afq.zero_point;
        }
        s_qp[0] = s;
        s_qp[1] = (float)zp;
        s_qp[2] = __fdiv_rn(1.0f, s);
        if (blockIdx.x == 0 && blockIdx.y == 0) {
          if (p.fin.training && p.fin.num_batches_tracked) *p.fin.num_batches_tracked += 1;
          if (p.fin.observe) {
            *p.fin.afq.min_val = rmin;
            *p.fin.afq.max_val = rmax;
            *p.fin.afq.scale = s;
            *p.fin.afq.zero_point = zp;
          }
          if (need_stats) {
            const float zf = (float)zp, iv = __fdiv_rn(1.0f, s);
            const float qa = fminf(fmaxf(fq_index(gmn, iv, zf), 0.0f), 255.0f);
            const float qb = fminf(fmaxf(fq_index(gmx, iv, zf), 0.0f), 255.0f);
            p.fin.cur_minmax[0] = fq_dequant(qa, zf, s);
            p.fin.cur_minmax[1] = fq_dequant(qb, zf, s);
          }
        }
      }
      epi_bar_sync();
      inv = s_qp[2];
      zpf = s_qp[1];
      cA = mine.A;
      cB = mine.B;
      if (blockIdx.x == 0 && active && half == 0) {       // one writer per channel
        p.fin.A[c] = mine.A;
        p.fin.B[c] = mine.B;
        p.fin.mean_I[c] = mine.mean_I;
        p.fin.kfac[c] = mine.kfac;
        if (p.fin.training) {
          p.fin.running_mean[c] = mine.new_running_mean;
          p.fin.running_var[c] = mine.new_running_var;
        }
      }
      // ---------------- phase B: quantise
      const int bn = p.bn;
      for (int64_t pt = blockIdx.x; pt < n_ptiles; pt += gridDim.x, ++tile_i) {
        const uint32_t acc = tile_i & 1;
        const int px_valid = (int)min((int64_t)PF_NPX, p.M - pt * PF_NPX);
        if (p.tma_out && tid == 0) tma_store_wait_read();      // the previous tile's store has finished reading `staging`
        epi_bar_sync();
        mbar_wait(&tfull_bar[acc], (tile_i >> 1) & 1);
        tc_fence_after();
        if (warp_active) {
#pragma unroll 1
          for (int ch = 0; ch < 4; ++ch) {
            const int col0 = half * 128 + ch * 32;
            if (col0 >= px_valid) break;
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + acc * PF_NPX + col0 + ((uint32_t)(quarter * 32) << 16), v);
            if (active) {
              uint8_t* dst = staging + col0 * bn + c_local;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const float I = (float)(wsign * (int)v[j] - corr_s);
                dst[j * bn] = (uint8_t)bnq1(I, cA, cB, relu, inv, zpf);
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        if (p.tma_out) {
          fence_proxy_async();
          epi_bar_sync();
          if (tid == 0) tma_store_2d(&tm_q, smem_u32(staging), c_tile0, (int)(pt * PF_NPX));
        } else {
          epi_bar_sync();
          // cooperative copy: 4-byte words, consecutive threads -> consecutive words of a row (coalesced)
          const int wpr = n_valid >> 2;                        // words per row (cout % 4 == 0)
          uint8_t* qbase = p.q + pt * PF_NPX * (int64_t)p.ldq + c_tile0;
          for (int i = tid; i < px_valid * wpr; i += PF_EPI_THREADS) {
            const int r = i / wpr, wd = i - r * wpr;
            *reinterpret_cast<uint32_t*>(qbase + (int64_t)r * p.ldq + 4 * wd) = *reinterpret_cast<const uint32_t*>(staging + r * bn + 4 * wd);
          }
        }
      }
      if (p.tma_out && tid == 0) tma_store_wait_all();
    } else {
      // ================================================================= backward passes
      const FrostBnBackwardArgs& b = p.bwd;
      relu = b.relu;
      inv = __fdiv_rn(1.0f, *b.out_scale);
      zpf = (float)*b.out_zp;
      cA = b.A[c];
      cB = b.B[c];
      cMean = b.mean_I[c];
      const int cout = p.cout;
      float c1 = 0.f, a0 = 0.f, a1 = 0.f;
      if constexpr (MODE == PF_BWD_APPLY) {
        const double sa_sw = (double)(*b.x_scale) * (double)(*b.w_scale);
        const BnBwdChannel r = bn_bwd_channel(__ldcg(b.sums + 2 * c), __ldcg(b.sums + 2 * c + 1), (double)b.M, sa_sw, cA, b.kfac[c], cMean,
                                              b.gamma[c], b.sf[c], b.eps, b.frozen ? 0 : 1);
        c1 = r.c1; a0 = r.a0; a1 = r.a1;
        if (blockIdx.x == 0 && active && half == 0) {
          b.dgamma_bn[c] = r.dgamma_bn;
          b.dbeta[c] = r.dbeta;
          b.dsf_bn[c] = r.dsf_bn;
        }
      }
      double S1 = 0.0, S2 = 0.0;
      uint16_t* dz_hi = reinterpret_cast<uint16_t*>(b.dz);
      uint16_t* dz_lo = reinterpret_cast<uint16_t*>(b.dz_lo);
      for (int64_t pt = blockIdx.x; pt < n_ptiles; pt += gridDim.x, ++tile_i) {
        const uint32_t acc = tile_i & 1;
        const int px_valid = (int)min((int64_t)PF_NPX, p.M - pt * PF_NPX);
        // dy of the first chunk is requested before the accumulator is waited for
        mbar_wait(&tfull_bar[acc], (tile_i >> 1) & 1);
        tc_fence_after();
        if (warp_active) {
#pragma unroll 1
          for (int ch = 0; ch < 4; ++ch) {
            const int col0 = half * 128 + ch * 32;
            if (col0 >= px_valid) break;
            const int nv = min(32, px_valid - col0);
            const int64_t e0 = (pt * PF_NPX + col0) * (int64_t)cout + c;      // element (pixel, channel) of column 0
            float dy[32];
            if (active) {
              const float* src = b.dy + e0;
#pragma unroll
              for (int j = 0; j < 32; ++j) dy[j] = (j < nv) ? ld_cg(src + j * cout) : 0.0f;
            }
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + acc * PF_NPX + col0 + ((uint32_t)(quarter * 32) << 16), v);
            if (active) {
              if constexpr (MODE == PF_BWD_REDUCE) {
                float p1 = 0.f, p2 = 0.f;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  const float I = (float)(wsign * (int)v[j] - corr_s);
                  const float dv = bn_dv(dy[j], I, cA, cB, relu, inv, zpf);
                  p1 += dv;
                  p2 = fmaf(dv, I - cMean, p2);
                }
                S1 += (double)p1;
                S2 += (double)p2;
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                  if (j < nv) {
                    const float I = (float)(wsign * (int)v[j] - corr_s);
                    const float dv = bn_dv(dy[j], I, cA, cB, relu, inv, zpf);
                    const float o = c1 * (dv - a0 - a1 * (I - cMean));
                    const __nv_bfloat16 h = __float2bfloat16_rn(o);
                    const __nv_bfloat16 l = __float2bfloat16_rn(o - __bfloat162float(h));
                    dz_hi[e0 + (int64_t)j * cout] = __bfloat16_as_ushort(h);
                    dz_lo[e0 + (int64_t)j * cout] = __bfloat16_as_ushort(l);
                  }
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      }
      if constexpr (MODE == PF_BWD_REDUCE) {
        if (active) {
          atomicAdd(&s_comb[c_local].s1, S1);
          atomicAdd(&s_comb[c_local].s2, S2);
        }
        epi_bar_sync();
        if (active && half == 0) {
          atomicAdd(b.sums + 2 * c, s_comb[c_local].s1);
          atomicAdd(b.sums + 2 * c + 1, s_comb[c_local].s2);
        }
      }
    }
  }

  // ---- teardown
  tc_fence_before();
  __syncthreads();
  if (warp == PF_EPI_WARPS + PF_PROD_WARPS) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ================================================================= host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
    (void)cudaGetLastError();
  }
  return fn;
}

// 2-D uint8 tensor [rows][pitch] with `cols` valid bytes per row; box = box_cols x box_rows
static bool make_map_u8(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch, uint32_t box_cols,
                        uint32_t box_rows, bool swizzle128) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {cols, rows};
  const cuuint64_t strides[1] = {pitch};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct PwOperands {
  const uint8_t* x; int64_t M; int K, ldx;
  const int8_t* w; int ldw, cout;
  const int32_t *x_zp, *w_zp, *wsum;
};

static int check_operands(const char* who, const PwOperands& o) {
  FROST_REQUIRE(o.x && o.w && o.x_zp && o.w_zp && o.wsum, "%s: null pointer", who);
  FROST_REQUIRE(o.M > 0 && o.M < ((int64_t)1 << 31) && o.K > 0 && o.cout > 0, "%s: empty problem or more than 2^31 pixels", who);
  FROST_REQUIRE(o.K % 8 == 0 && o.cout % 4 == 0, "%s: K=%d must be a multiple of 8 and cout=%d of 4", who, o.K, o.cout);
  FROST_REQUIRE(o.ldx >= o.K && o.ldx % 8 == 0, "%s: ldx=%d must be >= K and a multiple of 8", who, o.ldx);
  FROST_REQUIRE(o.ldw >= o.K && o.ldw % 16 == 0, "%s: ldw=%d must be >= K and a multiple of 16", who, o.ldw);
  FROST_REQUIRE((reinterpret_cast<uintptr_t>(o.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(o.w) & 15) == 0,
                "%s: operands must be 16-byte aligned", who);
  return FROST_OK;
}

template <int MODE>
static int launch_fused(const char* who, const PwOperands& o, PwFusedParams& p, cudaStream_t st) {
  p.x = o.x; p.M = o.M; p.K = o.K; p.ldx = o.ldx;
  p.w = o.w; p.ldw = o.ldw; p.cout = o.cout;
  p.x_zp = o.x_zp; p.w_zp = o.w_zp; p.wsum = o.wsum;
  p.n_kb = (o.K + PF_BK - 1) / PF_BK;
  const int n_ct = (o.cout + PF_CH - 1) / PF_CH;
  p.bn = (((o.cout + n_ct - 1) / n_ct) + 15) & ~15;          // balanced channel tiles, multiple of 16 (<= 128)
  const int n_ct_eff = (o.cout + p.bn - 1) / p.bn;
  const bool tma_allowed = tunable(FROST_TUNE_PW_TMA) == 1 && encode_fn() != nullptr;
  p.tma_in = (tma_allowed && o.ldx % 16 == 0) ? 1 : 0;
  p.vec16 = (o.ldx % 16 == 0 && o.K % 16 == 0) ? 1 : 0;
  p.tma_out = 0;
  CUtensorMap tm_x, tm_w, tm_q;
  memset(&tm_x, 0, sizeof(tm_x));
  memset(&tm_w, 0, sizeof(tm_w));
  memset(&tm_q, 0, sizeof(tm_q));
  if (p.tma_in) {
    if (!make_map_u8(&tm_x, o.x, (uint64_t)o.K, (uint64_t)o.M, (uint64_t)o.ldx, PF_BK, PF_NPX, true) ||
        !make_map_u8(&tm_w, o.w, (uint64_t)o.K, (uint64_t)o.cout, (uint64_t)o.ldw, PF_BK, PF_CH, true)) {
      set_error("%s: cuTensorMapEncodeTiled failed for the operands", who);
      return FROST_ECUDA;
    }
  }
  if (MODE == PF_FWD) {
    FROST_REQUIRE(p.q && p.ldq >= o.cout && p.ldq % 4 == 0 && (reinterpret_cast<uintptr_t>(p.q) & 15) == 0,
                  "%s: q must be 16-byte aligned with ldq >= cout, ldq %% 4 == 0", who);
    if (tma_allowed && p.ldq % 16 == 0) {
      if (!make_map_u8(&tm_q, p.q, (uint64_t)o.cout, (uint64_t)o.M, (uint64_t)p.ldq, (uint32_t)p.bn, PF_NPX, false)) {
        set_error("%s: cuTensorMapEncodeTiled failed for the output", who);
        return FROST_ECUDA;
      }
      p.tma_out = 1;
    }
  }
  const bool resident = p.n_kb <= PF_RES_KB;
  const size_t smem = 1024 + (resident ? PF_RES_KB * PF_W_BYTES + 4 * PF_X_BYTES : 3 * (PF_X_BYTES + PF_W_BYTES)) +
                      (MODE == PF_FWD ? PF_STAGING : 0) + PF_TAIL;
  if (first_use_on_device(reinterpret_cast<const void*>(&pw_fused_kernel<MODE>))) {
    cudaError_t e = cudaFuncSetAttribute(pw_fused_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute failed: %s", who, cudaGetErrorString(e));
      return FROST_ECUDA;
    }
  }
  const int64_t n_ptiles = ceil_div(o.M, PF_NPX);
  const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(n_ptiles, kNumSMs / n_ct_eff));
  // the 8-byte cp.async variant allocates in L1 (cp.async.ca): ordinary stream-ordered launch (common.cuh)
  const bool pdl_ok = p.tma_in || p.vec16;
  cudaError_t e = launch_pdl_if(pdl_ok, pw_fused_kernel<MODE>, dim3(gx, n_ct_eff), dim3(PF_THREADS), smem, st, tm_x, tm_w, tm_q, p);
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", who, cudaGetErrorString(e));
    return FROST_ECUDA;
  }
  return FROST_OK;
}

}  // namespace frost

using namespace frost;

static PwOperands operands_of(const FrostPwOperands& a) {
  PwOperands o;
  o.x = a.x; o.M = a.M; o.K = a.K; o.ldx = a.ldx;
  o.w = a.w_mma; o.ldw = a.ldw; o.cout = a.cout;
  o.x_zp = a.x_zp; o.w_zp = a.w_zp; o.wsum = a.wsum;
  return o;
}

extern "C" int frost_pw_fused_forward(const FrostPwFusedFwdArgs* a, void* stream) {
  FROST_REQUIRE(a, "frost_pw_fused_forward: null args");
  const PwOperands o = operands_of(a->op);
  int rc = check_operands("frost_pw_fused_forward", o);
  if (rc) return rc;
  const FrostBnFinalizeArgs& f = a->bn;
  FROST_REQUIRE(f.stats && f.x_scale && f.w_scale && f.sf && f.gamma && f.beta && f.running_mean && f.running_var && f.A && f.B &&
                    f.mean_I && f.kfac && f.cur_minmax && f.afq.scale && f.afq.zero_point && f.afq.min_val && f.afq.max_val &&
                    a->grid_barrier && a->q,
                "frost_pw_fused_forward: null pointer");
  FROST_REQUIRE(f.C == o.cout && f.count == o.M && f.stats_format == 0, "frost_pw_fused_forward: bn.C / bn.count must match the conv");
  PwFusedParams p;
  memset(&p, 0, sizeof(p));
  p.fin = f;
  p.grid_bar = a->grid_barrier;
  p.q = a->q;
  p.ldq = a->ldq;
  rc = launch_fused<PF_FWD>("frost_pw_fused_forward", o, p, (cudaStream_t)stream);
  if (rc) return rc;
  FROST_LAUNCH_CHECK("pw_fused_forward");
  return FROST_OK;
}

static int fused_backward(const FrostPwFusedBwdArgs* a, void* stream, bool apply) {
  const char* who = apply ? "frost_pw_fused_bwd_apply" : "frost_pw_fused_bwd_reduce";
  FROST_REQUIRE(a, "%s: null args", who);
  const PwOperands o = operands_of(a->op);
  int rc = check_operands(who, o);
  if (rc) return rc;
  const FrostBnBackwardArgs& b = a->bn;
  FROST_REQUIRE(b.dy && b.A && b.B && b.mean_I && b.kfac && b.gamma && b.sf && b.x_scale && b.w_scale && b.out_scale && b.out_zp &&
                    b.sums && b.dgamma_bn && b.dbeta && b.dsf_bn,
                "%s: null pointer", who);
  FROST_REQUIRE(b.C == o.cout && b.M == o.M && b.acc_format == 0, "%s: bn.C / bn.M must match the conv", who);
  FROST_REQUIRE(!apply || (b.dz && b.dz_lo && b.dz_format == 1), "%s: needs the bf16 hi/lo planes (dz_format 1)", who);
  cudaStream_t st = (cudaStream_t)stream;
  if (!apply && cudaMemsetAsync(b.sums, 0, sizeof(double) * 2 * b.C, st) != cudaSuccess) {
    set_error("%s: memset failed", who);
    return FROST_ECUDA;
  }
  PwFusedParams p;
  memset(&p, 0, sizeof(p));
  p.bwd = b;
  rc = apply ? launch_fused<PF_BWD_APPLY>(who, o, p, st) : launch_fused<PF_BWD_REDUCE>(who, o, p, st);
  if (rc) return rc;
  FROST_LAUNCH_CHECK(who);
  return FROST_OK;
}

extern "C" int frost_pw_fused_bwd_reduce(const FrostPwFusedBwdArgs* a, void* stream) { return fused_backward(a, stream, false); }
extern "C" int frost_pw_fused_bwd_apply(const FrostPwFusedBwdArgs* a, void* stream) { return fused_backward(a, stream, true); }
