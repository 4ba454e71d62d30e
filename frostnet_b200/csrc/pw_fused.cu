// pw_fused.cu - the 1x1 (pointwise) ConvBn(ReLU)2d of the Frost bottleneck as ONE kernel per direction and pass:
// convolution on tcgen05 tensor cores + BatchNorm (training-mode batch statistics) + ReLU + activation fake-quant
// in the epilogue, and the matching backward passes - none of which ever stores the int32 accumulator I.
// Replaces, for the squeeze / expand / reduce / last_layer convs of frostnet.py:98-119,293, the chain
//   aten::convolution -> div -> native_batch_norm -> relu -> fused_moving_avg_obs_fake_quant   (conv_fused.py:131-167,708-710)
// and the first two stages of its autograd (SURVEY.md 8a' steps 1-3).
//
// One pair of shared-memory tiles (activations: 256 pixels x 128 bytes of K, weights: 128 rows x 128 bytes of K, both
// staged by TMA in the K-major SWIZZLE_128B layout) feeds the tensor cores in TWO orientations, chosen per pass:
//
//   channel-per-thread ("cpt", the reduction passes):  D[channel][pixel] = W * X^T.  Weight tile = UMMA A (128 TMEM lanes =
//     output channels), activation tile = UMMA B (N = 256 pixels = TMEM columns).  An epilogue thread owns ONE output channel:
//     sum I, sum I^2, min, max (forward statistics) and sum dv, sum dv*(I-mean) (BatchNorm backward) accumulate in registers,
//     no shuffles, no shared memory; dy[pixel][channel] loads are coalesced across the warp.  When a tile has <= 64 (<= 32)
//     channels the weight rows are REPLICATED 2x (4x) down the 128 rows, so every lane quarter holds a copy of the result and
//     all warps work on different pixel columns.
//   pixel-per-thread ("ppt", the elementwise passes):  D[pixel][channel] = X * W^T (two M = 128 MMAs per 256-pixel tile).
//     A thread owns one pixel and 16 consecutive channels per step: the quantised bytes / bf16 gradient planes leave as 16-byte
//     vector stores, dy arrives as 16-byte vector loads, the per-channel coefficients are shared-memory broadcasts.
//
//   forward   phase A (cpt): GEMM -> per-channel integer statistics -> integer atomics
//             grid barrier (all CTAs resident: grid <= 148, 1 CTA/SM) -> every CTA finalises BN / observer / qparams
//             phase B (ppt): the same GEMM again (the operands are 1/6..1x the output bytes; the tensor pipe is idle anyway)
//                      -> A_c*I+B_c -> ReLU -> quantise -> uint8 rows
//             With frozen BatchNorm and the observer off (eval / late QAT) phase A and the barrier are skipped.
//   backward  reduce (cpt): GEMM -> dv = dy*mask -> S1, S2 per channel -> fp64 atomics
//             apply  (cpt): GEMM -> dz = c1*(dv - a0 - a1*(I - mean)) -> bf16 hi/lo planes (operands of the tensor-core dgrad/wgrad)
//             The STE / ReLU mask is an integer interval of I per channel (bn_math.cuh): one subtract + one compare per element.
//
// Warp roles: 16 epilogue warps (4 per TMEM lane quarter), 1 TMA producer (one elected thread; the mbarriers count bytes),
// 1 MMA issuer.  TMEM: 2 x 256 columns, double buffered (MMAs of tile i+1 overlap the epilogue of tile i).  Weights stay
// resident in shared memory when K <= 384.  Row pitches must be multiples of 16 bytes (TMA's stride rule; the engine pads).
#include <cuda.h>
#include <cuda_bf16.h>
#include <algorithm>
#include <cstring>
#include "pw_tma.cuh"
#include "bn_math.cuh"

namespace frost {

using namespace tc;

constexpr int PF_NPX = 256;          // pixels per tile
constexpr int PF_CH = 128;           // weight rows per tile
constexpr int PF_BK = 128;           // bytes of K per k-block (one swizzle span)
constexpr int PF_RES_KB = 3;         // weight k-blocks that stay resident
constexpr int PF_EPI_WARPS = 16;
constexpr int PF_EPI_THREADS = PF_EPI_WARPS * 32;
constexpr int PF_THREADS = (PF_EPI_WARPS + 2) * 32;   // warps 0-15 epilogue, 16 TMA producer, 17 MMA
constexpr int PF_W_BYTES = PF_CH * PF_BK;       // 16 KB
constexpr int PF_X_BYTES = PF_NPX * PF_BK;      // 32 KB
constexpr int PF_TAIL = 12288;                  // barriers + per-channel buffers
constexpr int PF_DY_SLOTS = 3;                  // backward: ring of gradient boxes (TMA)
constexpr int PF_DY_BYTES = 32768;              // one box: 64*R pixels x 128/R channels fp32 (R = weight-row replication)
constexpr int PF_BWD_STAGES = 2;                // activation stages of the backward modes (the ring takes the rest)

enum { PF_FWD = 0, PF_BWD_REDUCE = 1, PF_BWD_APPLY = 2 };

struct PwFusedParams {
  const uint8_t* x;
  int64_t M;
  int K, ldx;
  const int8_t* w;        // MMA-ready weight bytes [cout][ldw]
  int ldw, cout, bn, bnr, n_kb;
  const int32_t* x_zp;
  const int32_t* w_zp;
  const int32_t* wsum;
  // forward
  FrostBnFinalizeArgs fin;
  unsigned* grid_bar;
  uint8_t* q;
  int ldq;
  // backward
  FrostBnBackwardArgs bwd;
  long long* trace;       // measurement aid (frost_debug_set_trace): SM clock stamps of CTA (0,0), NULL = off
};

// stamp i of the launch timeline (one thread per call site); compiled in with -DFROST_TRACE only (python -m frostnet_b200.build --trace)
#ifdef FROST_TRACE
#define PF_STAMP(i) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) p.trace[i] = clock64(); } while (0)
#else
#define PF_STAMP(i) do { } while (0)
#endif

struct PfCombine {       // per channel of the tile: the partial results of the threads that share a channel meet here
  unsigned long long sum, sq;
  int mn, mx;
  double s1, s2;
};

// ================================================================= the kernel
template <int MODE>
__global__ void __launch_bounds__(PF_THREADS, 1) pw_fused_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                                const __grid_constant__ CUtensorMap tm_w,
                                                                const __grid_constant__ CUtensorMap tm_dy,
                                                                const __grid_constant__ PwFusedParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSETTING the __shared__ array (a cast through uintptr_t would turn every later access into a
  // generic LD/ST: ncu showed 10 % of the forward's instructions as generic loads of the per-channel coefficients)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const bool resident = p.n_kb <= PF_RES_KB;
  const int S = MODE != PF_FWD ? PF_BWD_STAGES : (resident ? 4 : 3);   // pipeline stages
  const int stage_bytes = resident ? PF_X_BYTES : PF_X_BYTES + PF_W_BYTES;
  uint8_t* wres = smem;                                              // [n_kb][128 rows][128 B] when resident
  uint8_t* stages = smem + (resident ? PF_RES_KB * PF_W_BYTES : 0);  // [S][x 32 KB (+ w 16 KB)]
  uint8_t* dy_ring = stages + S * stage_bytes;                       // backward: [PF_DY_SLOTS][32 KB] gradient boxes
  uint8_t* tail = dy_ring + (MODE != PF_FWD ? PF_DY_SLOTS * PF_DY_BYTES : 0);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);            // [4]
  uint64_t* empty_bar = full_bar + 4;                                // [4]
  uint64_t* tfull_bar = empty_bar + 4;                               // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                              // [2]
  uint64_t* wfull_bar = tempty_bar + 2;                              // [1]
  uint64_t* dyfull_bar = wfull_bar + 1;                              // [PF_DY_SLOTS]
  uint64_t* dyempty_bar = dyfull_bar + PF_DY_SLOTS;                  // [PF_DY_SLOTS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(dyempty_bar + PF_DY_SLOTS);
  float* s_red = reinterpret_cast<float*>(tmem_slot + 2);            // [2 * PF_EPI_WARPS] block min / max
  float* s_qp = s_red + 2 * PF_EPI_WARPS;                            // scale, zero point (as float), 1/scale, spare
  float4* s_cf = reinterpret_cast<float4*>(tail + 512);              // [128] per channel: A, B, corr (int bits), -
  PfCombine* s_comb = reinterpret_cast<PfCombine*>(tail + 512 + 2048 + 1024);   // [128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ct = blockIdx.y;
  const int c_tile0 = ct * p.bn;
  const int n_valid = min(p.bn, p.cout - c_tile0);
  const int64_t n_ptiles = (p.M + PF_NPX - 1) / PF_NPX;
  const int bnr = p.bnr;                       // replica stride of the weight rows: 32, 64 or 128
  const int R = PF_CH / bnr;

  // ---- one-time setup (overlaps the tail of the previous kernel: programmatic dependent launch, common.cuh)
  if (threadIdx.x == 0) {
    PF_STAMP(0);
    for (int s = 0; s < 4; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], PF_EPI_WARPS); }
    mbar_init(wfull_bar, 1);
    for (int i = 0; i < PF_DY_SLOTS; ++i) { mbar_init(&dyfull_bar[i], 1); mbar_init(&dyempty_bar[i], PF_EPI_WARPS); }
    mbar_fence_init();
  }
  if (warp == PF_EPI_WARPS + 1) tmem_alloc<512>(tmem_slot);
  if (threadIdx.x == 0) PF_STAMP(1);
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
  if (threadIdx.x == 0) PF_STAMP(2);

  // (the TMA producer needs neither zero point: it must not wait an L2 round trip for them before its first load)
  const bool is_producer = warp == PF_EPI_WARPS;
  const int zp_a = is_producer ? 0 : *p.x_zp, zp_w = is_producer ? 0 : *p.w_zp;
  if (zp_w != 0 && zp_w != -128 && zp_w != 127) __trap();  // not reachable with ChooseQuantizationParams
  // forward: does this launch need batch statistics / the observer's min-max (phase A + grid barrier)?
  const bool need_stats = (MODE == PF_FWD) && (p.fin.training || p.fin.observe);
  const int n_phases = (MODE == PF_FWD && need_stats) ? 2 : 1;
  // orientation of phase `ph`: forward = [cpt,] ppt ; backward reduce / apply = cpt
  auto phase_is_cpt = [&](int ph) { return MODE != PF_FWD || (need_stats && ph == 0); };

  if (warp == PF_EPI_WARPS) {
    // ================================================================= TMA producer (one elected thread)
    if (lane == 0) {
      tma_prefetch_desc(&tm_x);
      tma_prefetch_desc(&tm_w);
      if (resident) {
        mbar_expect_tx(wfull_bar, (uint32_t)(p.n_kb * PF_W_BYTES));
        for (int kb = 0; kb < p.n_kb; ++kb)
          for (int j = 0; j < R; ++j)          // the same bnr weight rows, replicated down the 128-row tile
            tma_load_2d(&tm_w, wfull_bar, smem_u32(wres + kb * PF_W_BYTES + j * bnr * PF_BK), kb * PF_BK, c_tile0);
      }
      uint32_t it = 0, dy_it = 0;
      for (int ph = 0; ph < n_phases; ++ph) {
        for (int64_t pt = blockIdx.x; pt < n_ptiles; pt += gridDim.x) {
          for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
            const int s = it % S;
            mbar_wait_parked(&empty_bar[s], ((it / S) & 1) ^ 1);
            uint8_t* st = stages + s * stage_bytes;
            mbar_expect_tx(&full_bar[s], (uint32_t)stage_bytes);
            tma_load_2d(&tm_x, &full_bar[s], smem_u32(st), kb * PF_BK, (int)(pt * PF_NPX));
            if (!resident)
              for (int j = 0; j < R; ++j)
                tma_load_2d(&tm_w, &full_bar[s], smem_u32(st + PF_X_BYTES + j * bnr * PF_BK), kb * PF_BK, c_tile0);
            if (it == 0) PF_STAMP(3);
          }
          if constexpr (MODE != PF_FWD) {
            // this tile's gradient rows, as boxes of 64*R pixels x 128/R channels: bulk copies keep HBM busy without tying up
            // load-queue entries and registers of the epilogue warps (the per-thread loads they replace were throttled there)
            const int box_px = 64 * R;
            for (int bj = 0; bj < PF_NPX / box_px; ++bj, ++dy_it) {
              const int slot = dy_it % PF_DY_SLOTS;
              mbar_wait_parked(&dyempty_bar[slot], ((dy_it / PF_DY_SLOTS) & 1) ^ 1);
              mbar_expect_tx(&dyfull_bar[slot], PF_DY_BYTES);
              tma_load_2d(&tm_dy, &dyfull_bar[slot], smem_u32(dy_ring + slot * PF_DY_BYTES), c_tile0, (int)(pt * PF_NPX + bj * box_px));
            }
          }
        }
      }
    }
  } else if (warp == PF_EPI_WARPS + 1) {
    // ================================================================= MMA issuer
    // weights: s8, or u8 after the zero-point rewrite done by weight prep (one-signed weights); activations: u8
    const int w_fmt = zp_w == 0 ? 1 : 0;
    const int n_mma = (n_valid + 15) & ~15;
    const uint32_t idesc_cpt = umma_idesc(2 /*S32*/, w_fmt, 0, PF_CH, PF_NPX);   // A = W (128 rows), B = X (256 pixels)
    const uint32_t idesc_ppt = umma_idesc(2 /*S32*/, 0, w_fmt, 128, n_mma);      // A = X (128 pixels), B = W (n_mma rows)
    if (resident) {
      mbar_wait_parked(wfull_bar, 0);
      tc_fence_after();
    }
    if (lane == 0) PF_STAMP(4);
    uint32_t it = 0, tile_i = 0;
    for (int ph = 0; ph < n_phases; ++ph) {
      const bool cpt = phase_is_cpt(ph);
      for (int64_t pt = blockIdx.x; pt < n_ptiles; pt += gridDim.x, ++tile_i) {
        const uint32_t acc = tile_i & 1;
        mbar_wait_parked(&tempty_bar[acc], ((tile_i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * PF_NPX;
        const bool second_half = p.M - pt * PF_NPX > 128;
        for (int kb = 0; kb < p.n_kb; ++kb, ++it) {
          const int s = it % S;
          mbar_wait_parked(&full_bar[s], (it / S) & 1);
          tc_fence_after();
          if (lane == 0) {
            if (it == 0) PF_STAMP(5);
            uint8_t* st = stages + s * stage_bytes;
            const uint64_t xdesc = umma_desc_sw128(smem_u32(st));
            const uint64_t wdesc = umma_desc_sw128(smem_u32(resident ? wres + kb * PF_W_BYTES : st + PF_X_BYTES));
            const int nk = min(PF_BK / 32, (p.K - kb * PF_BK + 31) / 32);
            if (cpt) {
              for (int k4 = 0; k4 < nk; ++k4)
                umma_i8(d_tmem, wdesc + (uint64_t)(2 * k4), xdesc + (uint64_t)(2 * k4), idesc_cpt, (kb | k4) != 0 ? 1u : 0u);
            } else {
              for (int h = 0; h < (second_half ? 2 : 1); ++h) {
                const uint64_t xh = xdesc + (uint64_t)((h * 128 * PF_BK) >> 4);       // pixels 128h .. 128h+127 of the tile
                for (int k4 = 0; k4 < nk; ++k4)
                  umma_i8(d_tmem + h * 128, xh + (uint64_t)(2 * k4), wdesc + (uint64_t)(2 * k4), idesc_ppt, (kb | k4) != 0 ? 1u : 0u);
              }
            }
            umma_commit(&empty_bar[s]);
            if (kb == p.n_kb - 1) umma_commit(&tfull_bar[acc]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ================================================================= epilogue (16 warps)
    const int quarter = warp & 3, grp = warp >> 2;        // TMEM lane quarter (hardware: warp_id % 4), group 0..3
    const int tid = threadIdx.x;                          // 0..511 within the epilogue
    const int wsign = (zp_w == 127) ? -1 : 1;             // I = wsign * raw - corr_s
    auto corr_of = [&](int c) {
      const int ws = p.wsum[c];
      const int ws_eff = zp_w == 0 ? ws : (zp_w == -128 ? ws + 128 * p.K : 127 * p.K - ws);
      return wsign * zp_a * ws_eff;
    };
    // ---- channel-per-thread mapping
    const int lane_g = quarter * 32 + lane;
    const int rep = lane_g / bnr;                         // which replica of the weight rows this lane holds
    const int c_local = lane_g - rep * bnr;
    const bool active = c_local < n_valid;
    const bool warp_active = (quarter * 32) % bnr < n_valid;
    const int c = c_tile0 + (active ? c_local : 0);
    const int cols_per = 64 / R;                          // pixel columns per (replica, group): 64, 32 or 16
    const int col_begin = (rep * 4 + grp) * cols_per;
    // ---- pixel-per-thread mapping
    const int ppt_h = grp & 1, ppt_par = grp >> 1;        // 128-pixel block, chunk parity
    const int ppt_pix = ppt_h * 128 + quarter * 32 + lane;

    uint32_t tile_i = 0;
    float inv = 1.f, zpf = 0.f;
    int relu = 0;

    if constexpr (MODE == PF_FWD) {
      relu = p.fin.relu;
      const int corr_s = corr_of(c);
      const double Mcount = (double)p.fin.count;
      const double sa_sw = (double)(*p.fin.x_scale) * (double)(*p.fin.w_scale);
      const double mom = p.fin.momentum >= 0.0f ? (double)p.fin.momentum
                                                : 1.0 / (double)((p.fin.num_batches_tracked ? *p.fin.num_batches_tracked : 0) + 1);
      // the observer state as it was BEFORE this layer ran: every CTA derives the same new state from it after the barrier
      float rmin = *p.fin.afq.min_val, rmax = *p.fin.afq.max_val;
      if (need_stats) {
        // ---------------- phase A (cpt): statistics
        long long sum = 0;
        unsigned long long sq = 0;
        int mn = INT_MAX, mx = INT_MIN;
        for (int64_t pt = blockIdx.x; pt < n_ptiles; pt += gridDim.x, ++tile_i) {
          const uint32_t acc = tile_i & 1;
          const int px_valid = (int)min((int64_t)PF_NPX, p.M - pt * PF_NPX);
          mbar_wait_parked(&tfull_bar[acc], (tile_i >> 1) & 1);
          tc_fence_after();
          if (tid == 0 && tile_i == 0) PF_STAMP(6);
          if (warp_active) {
#pragma unroll 1
            for (int col0 = col_begin; col0 < col_begin + cols_per; col0 += 16) {
              if (col0 >= px_valid) break;
              uint32_t v[16];
              tmem_ld_32x16(tmem_base + acc * PF_NPX + col0 + ((uint32_t)(quarter * 32) << 16), v);
              const int nv = min(16, px_valid - col0);
              int s32 = 0;
              if (nv == 16) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const int I = wsign * (int)v[j] - corr_s;
                  s32 += I;
                  sq += (unsigned long long)((long long)I * (long long)I);
                  mn = min(mn, I);
                  mx = max(mx, I);
                }
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
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
        if (tid == 0) PF_STAMP(7);
        // the threads that share a channel meet in shared memory; one set of integer atomics per channel and CTA
        if (active && mn <= mx) {
          PfCombine& cb = s_comb[c_local];
          atomicAdd(&cb.sum, (unsigned long long)sum);
          atomicAdd(&cb.sq, sq);
          atomicMin(&cb.mn, mn);
          atomicMax(&cb.mx, mx);
        }
        named_bar_sync<PF_EPI_THREADS>();
        if (tid < n_valid) {
          const PfCombine cb = s_comb[tid];
          chan_stats_flush(const_cast<FrostChanStats*>(p.fin.stats) + c_tile0 + tid, (long long)cb.sum, cb.sq, cb.mn, cb.mx);
        }
        // ---------------- grid barrier: every CTA of this launch is resident (host: grid <= #SMs, 1 CTA per SM)
        named_bar_sync<PF_EPI_THREADS>();
        if (tid == 0) {
          PF_STAMP(8);
          const unsigned target = gridDim.x * gridDim.y;
          __threadfence();                                   // cumulativity: the CTA's atomics (ordered by the barrier above) come before the arrival
          atomicAdd(p.grid_bar, 1u);
          bool ok = false;
          for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
            if (ld_acquire_gpu(p.grid_bar) >= target) { ok = true; break; }
            __nanosleep(64);
          }
          if (!ok) __trap();
          PF_STAMP(9);
        }
        named_bar_sync<PF_EPI_THREADS>();
      }
      // ---------------- finalize: BN affine of the tile's channels, observer over all channels, qparams (same in every CTA)
      // every channel once per CTA (the observer needs the extrema over ALL channels), in an order rotated so that thread
      // tid < n_valid meets channel c_tile0 + tid - the one whose coefficients this CTA's tiles use - in its first iteration
      float gmn = INFINITY, gmx = -INFINITY;
      BnChannel mine;
      mine.A = mine.B = mine.mean_I = mine.kfac = 0.f;
      mine.new_running_mean = mine.new_running_var = 0.f;
      for (int j = tid; j < p.cout; j += PF_EPI_THREADS) {
        int cc = c_tile0 + j;
        if (cc >= p.cout) cc -= p.cout;
        FrostChanStats st;
        const FrostChanStats* g = p.fin.stats + cc;
        st.sum = __ldcg(&g->sum); st.sq_lo = __ldcg(&g->sq_lo); st.sq_hi = __ldcg(&g->sq_hi);
        st.min = __ldcg(&g->min); st.max = __ldcg(&g->max);
        if (!need_stats) { st.min = 0; st.max = 0; }
        const BnChannel r = bn_channel_finalize(st, 0, Mcount, p.fin.count > 1, sa_sw, p.fin.sf[cc], p.fin.gamma[cc], p.fin.beta[cc],
                                                p.fin.running_mean[cc], p.fin.running_var[cc], p.fin.eps, mom, p.fin.training, relu);
        gmn = fminf(gmn, r.v_lo);
        gmx = fmaxf(gmx, r.v_hi);
        if (j < n_valid) {                                   // j == tid here: thread tid <-> channel c_tile0 + tid
          mine = r;
          s_cf[tid] = make_float4(r.A, r.B, __int_as_float(corr_of(cc)), 0.f);
        }
      }
      gmn = warp_min(gmn);
      gmx = warp_max(gmx);
      if (lane == 0) { s_red[2 * warp] = gmn; s_red[2 * warp + 1] = gmx; }
      named_bar_sync<PF_EPI_THREADS>();          // also orders every thread's reads of the old running statistics before the writes below
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
      named_bar_sync<PF_EPI_THREADS>();
      if (tid == 0) PF_STAMP(10);
      inv = s_qp[2];
      zpf = s_qp[1];
      if (blockIdx.x == 0 && tid < n_valid) {                // one writer per channel
        const int cc = c_tile0 + tid;
        p.fin.A[cc] = mine.A;
        p.fin.B[cc] = mine.B;
        p.fin.mean_I[cc] = mine.mean_I;
        p.fin.kfac[cc] = mine.kfac;
        if (p.fin.training) {
          p.fin.running_mean[cc] = mine.new_running_mean;
          p.fin.running_var[cc] = mine.new_running_var;
        }
      }
      // ---------------- phase B (ppt): quantise
      for (int64_t pt = blockIdx.x; pt < n_ptiles; pt += gridDim.x, ++tile_i) {
        const uint32_t acc = tile_i & 1;
        const int64_t pix = pt * PF_NPX + ppt_pix;
        const bool pvalid = pix < p.M;
        const bool wvalid = pt * PF_NPX + ppt_h * 128 + quarter * 32 < p.M;      // warp-uniform
        mbar_wait_parked(&tfull_bar[acc], (tile_i >> 1) & 1);
        tc_fence_after();
        if (tid == 0 && pt == blockIdx.x) PF_STAMP(11);
        if (wvalid) {
          uint8_t* qrow = p.q + pix * (int64_t)p.ldq + c_tile0;
#pragma unroll 1
          for (int ci = ppt_par; ci * 16 < n_valid; ci += 2) {
            uint32_t v[16];
            tmem_ld_32x16(tmem_base + acc * PF_NPX + ppt_h * 128 + ci * 16 + ((uint32_t)(quarter * 32) << 16), v);
            unsigned w4[4];
#pragma unroll
            for (int j4 = 0; j4 < 4; ++j4) {
              unsigned word = 0u;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float4 cf = s_cf[ci * 16 + j4 * 4 + e];
                const float I = (float)(wsign * (int)v[j4 * 4 + e] - __float_as_int(cf.z));
                word |= bnq1(I, cf.x, cf.y, relu, inv, zpf) << (8 * e);
              }
              w4[j4] = word;
            }
            if (pvalid) {
              const int nvc = min(16, n_valid - ci * 16);         // 4, 8, 12 or 16 valid channels in this chunk
              uint8_t* dst = qrow + ci * 16;
              if (nvc == 16) {
                *reinterpret_cast<uint4*>(dst) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
              } else {
#pragma unroll
                for (int j4 = 0; j4 < 3; ++j4)
                  if (j4 * 4 < nvc) *reinterpret_cast<unsigned*>(dst + 4 * j4) = w4[j4];
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      }
    } else {
      // ================================================================= backward passes (cpt): reduce / apply
      const FrostBnBackwardArgs& b = p.bwd;
      relu = b.relu;
      inv = __fdiv_rn(1.0f, *b.out_scale);
      zpf = (float)*b.out_zp;
      const int corr_s = corr_of(c);
      const float cA = b.A[c], cB = b.B[c], cMean = b.mean_I[c];
      const int cout = p.cout;
      // the STE / ReLU mask of this channel as an integer interval of the accumulator (bn_math.cuh)
      // everything the coefficients need is requested before the interval search: one L2 latency, not two
      double gS1 = 0.0, gS2 = 0.0, sa_sw = 1.0;
      float cK = 0.f, cG = 0.f, cSf = 1.f;
      if constexpr (MODE == PF_BWD_APPLY) {
        gS1 = __ldcg(b.sums + 2 * c);
        gS2 = __ldcg(b.sums + 2 * c + 1);
        sa_sw = (double)(*b.x_scale) * (double)(*b.w_scale);
        cK = b.kfac[c]; cG = b.gamma[c]; cSf = b.sf[c];
      }
      if (tid == 0) PF_STAMP(13);
      const MaskInterval mk = bn_mask_interval(cA, cB, relu, inv, zpf);
      if (tid == 0) PF_STAMP(14);
      float cP = 0.f, cQ = 0.f, cR = 0.f;
      if constexpr (MODE == PF_BWD_APPLY) {
        const BnBwdChannel r = bn_bwd_channel(gS1, gS2, (double)b.M, sa_sw, cA, cK, cMean, cG, cSf, b.eps, b.frozen ? 0 : 1);
        // dz = c1*(dv - a0 - a1*(I - mean)) = P*dv + R*I + Q
        cP = r.c1;
        cR = -r.c1 * r.a1;
        cQ = r.c1 * (r.a1 * cMean - r.a0);
        if (blockIdx.x == 0 && active && rep == 0 && grp == 0) {
          b.dgamma_bn[c] = r.dgamma_bn;
          b.dbeta[c] = r.dbeta;
          b.dsf_bn[c] = r.dsf_bn;
        }
      }
      uint16_t* dz_hi = reinterpret_cast<uint16_t*>(b.dz);
      uint16_t* dz_lo = reinterpret_cast<uint16_t*>(b.dz_lo);
      double S1 = 0.0, S2 = 0.0;
      // The gradient arrives through the shared-memory ring (TMA producer): per box of 64*R pixels every thread takes 16
      // pixels of its channel - the same 16 accumulator columns it reads from TMEM.
      const int box_px = 64 * R;
      const int n_box = PF_NPX / box_px;                         // boxes per tile: 4, 2 or 1
      const int px0 = (rep * 4 + grp) * 16;                      // my pixels inside a box
      const uint32_t dy_row_b = (uint32_t)bnr * 4;               // bytes between consecutive pixels inside a box
      const uint32_t dy_thread = smem_u32(dy_ring) + (uint32_t)px0 * dy_row_b + (uint32_t)c_local * 4;
      auto lds = [](uint32_t a) {
        float v;
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
        return v;
      };
      float p1 = 0.f, p2 = 0.f;
      auto proc16 = [&](uint32_t dy_addr, uint32_t taddr, int64_t e_col, int col, int px_valid) {
        uint32_t v[16];
        tmem_ld_32x16(taddr, v);
        if (!active) return;
        float d[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) d[j] = lds(dy_addr + (uint32_t)j * dy_row_b);
        if constexpr (MODE == PF_BWD_REDUCE) {
          // pixels past the end of the tensor: the box rows are zero-filled, dv = 0
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int I = wsign * (int)v[j] - corr_s;
            const float dv = mask_passes(mk, I) ? d[j] : 0.0f;
            p1 += dv;
            p2 = fmaf(dv, (float)I - cMean, p2);
          }
        } else {
          char* ph = reinterpret_cast<char*>(dz_hi + e_col);
          const int64_t lo_off = reinterpret_cast<char*>(dz_lo) - reinterpret_cast<char*>(dz_hi);   // one running pointer for both planes
          const int64_t row_h = (int64_t)cout * 2;
          auto one = [&](int j) {
            const int I = wsign * (int)v[j] - corr_s;
            const float base = fmaf(cR, (float)I, cQ);
            const float o = mask_passes(mk, I) ? fmaf(cP, d[j], base) : base;
            const __nv_bfloat16 h = __float2bfloat16_rn(o);
            const __nv_bfloat16 l = __float2bfloat16_rn(o - __bfloat162float(h));
            *reinterpret_cast<uint16_t*>(ph) = __bfloat16_as_ushort(h);
            *reinterpret_cast<uint16_t*>(ph + lo_off) = __bfloat16_as_ushort(l);
            ph += row_h;
          };
          if (col + 16 <= px_valid) {            // every tile but the last: no per-element predicates
#pragma unroll
            for (int j = 0; j < 16; ++j) one(j);
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (col + j < px_valid) one(j);
          }
        }
      };
      uint32_t dy_it = 0;
      for (int64_t pt = blockIdx.x; pt < n_ptiles; pt += gridDim.x, ++tile_i) {
        const uint32_t acc = tile_i & 1;
        const int px_valid = (int)min((int64_t)PF_NPX, p.M - pt * PF_NPX);
        const int64_t e_tile = pt * PF_NPX * (int64_t)cout + c;                // element (pixel 0 of the tile, my channel)
        const uint32_t t0 = tmem_base + acc * PF_NPX + ((uint32_t)(quarter * 32) << 16);
        if (tid == 0 && tile_i == 0) PF_STAMP(15);
        mbar_wait_parked(&tfull_bar[acc], (tile_i >> 1) & 1);
        tc_fence_after();
        if (tid == 0 && tile_i == 0) PF_STAMP(16);
        p1 = 0.f;
        p2 = 0.f;
#pragma unroll 1
        for (int bj = 0; bj < n_box; ++bj, ++dy_it) {
          const int slot = dy_it % PF_DY_SLOTS;
          // every warp, also one whose lanes hold no channel, takes part in the ring protocol (16 arrivals free a slot)
          mbar_wait_parked(&dyfull_bar[slot], (dy_it / PF_DY_SLOTS) & 1);
          const int col = bj * box_px + px0;                                   // first of my 16 pixels, tile-relative
          if (warp_active && col < px_valid)
            proc16(dy_thread + (uint32_t)slot * PF_DY_BYTES, t0 + col, e_tile + (int64_t)col * cout, col, px_valid);
          __syncwarp();
          if (lane == 0) mbar_arrive(&dyempty_bar[slot]);
        }
        S1 += (double)p1;                        // <= 64 terms per fp32 partial
        S2 += (double)p2;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      }
      if (tid == 0) PF_STAMP(17);
      if constexpr (MODE == PF_BWD_REDUCE) {
        if (active) {
          atomicAdd(&s_comb[c_local].s1, S1);
          atomicAdd(&s_comb[c_local].s2, S2);
        }
        named_bar_sync<PF_EPI_THREADS>();
        if (tid < n_valid) {
          atomicAdd(b.sums + 2 * (c_tile0 + tid), s_comb[tid].s1);
          atomicAdd(b.sums + 2 * (c_tile0 + tid) + 1, s_comb[tid].s2);
        }
      }
    }
  }

  // ---- teardown
  if (threadIdx.x == 0) PF_STAMP(12);
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) PF_STAMP(18);
  if (warp == PF_EPI_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ================================================================= host side
struct PwOperands {
  const uint8_t* x; int64_t M; int K, ldx;
  const int8_t* w; int ldw, cout;
  const int32_t *x_zp, *w_zp, *wsum;
};

static int check_operands(const char* who, const PwOperands& o) {
  FROST_REQUIRE(o.x && o.w && o.x_zp && o.w_zp && o.wsum, "%s: null pointer", who);
  FROST_REQUIRE(o.M > 0 && o.M < ((int64_t)1 << 31) && o.K > 0 && o.cout > 0, "%s: empty problem or more than 2^31 pixels", who);
  FROST_REQUIRE(o.K % 8 == 0 && o.cout % 4 == 0, "%s: K=%d must be a multiple of 8 and cout=%d of 4", who, o.K, o.cout);
  FROST_REQUIRE(o.ldx >= o.K && o.ldx % 16 == 0, "%s: ldx=%d must be >= K and a multiple of 16 (TMA stride rule)", who, o.ldx);
  FROST_REQUIRE(o.ldw >= o.K && o.ldw % 16 == 0, "%s: ldw=%d must be >= K and a multiple of 16", who, o.ldw);
  FROST_REQUIRE((reinterpret_cast<uintptr_t>(o.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(o.w) & 15) == 0,
                "%s: operands must be 16-byte aligned", who);
  return FROST_OK;
}

static long long* g_trace = nullptr;

template <int MODE>
static int launch_fused(const char* who, const PwOperands& o, PwFusedParams& p, cudaStream_t st) {
  p.trace = g_trace;
  if (!encode_fn()) {
    set_error("%s: cuTensorMapEncodeTiled is not available from this driver (the fused kernels stage operands by TMA)", who);
    return FROST_ENOSUP;
  }
  p.x = o.x; p.M = o.M; p.K = o.K; p.ldx = o.ldx;
  p.w = o.w; p.ldw = o.ldw; p.cout = o.cout;
  p.x_zp = o.x_zp; p.w_zp = o.w_zp; p.wsum = o.wsum;
  p.n_kb = (o.K + PF_BK - 1) / PF_BK;
  const int n_ct = (o.cout + PF_CH - 1) / PF_CH;
  p.bn = (((o.cout + n_ct - 1) / n_ct) + 15) & ~15;          // balanced channel tiles, multiple of 16 (<= 128)
  p.bnr = p.bn <= 32 ? 32 : (p.bn <= 64 ? 64 : 128);          // replica stride of the weight rows in the 128-row tile
  const int n_ct_eff = (o.cout + p.bn - 1) / p.bn;
  CUtensorMap tm_x, tm_w;
  if (!make_map_u8(&tm_x, o.x, (uint64_t)o.K, (uint64_t)o.M, (uint64_t)o.ldx, PF_BK, PF_NPX) ||
      !make_map_u8(&tm_w, o.w, (uint64_t)o.K, (uint64_t)o.cout, (uint64_t)o.ldw, PF_BK, (uint32_t)p.bnr)) {
    set_error("%s: cuTensorMapEncodeTiled failed", who);
    return FROST_ECUDA;
  }
  const bool resident = p.n_kb <= PF_RES_KB;
  CUtensorMap tm_dy = tm_x;                   // forward: unused
  size_t smem = 1024 + (resident ? PF_RES_KB * PF_W_BYTES + 4 * PF_X_BYTES : 3 * (PF_X_BYTES + PF_W_BYTES)) + PF_TAIL;
  if (MODE != PF_FWD) {
    // gradient boxes of 64*R pixels x 128/R channels (R = PF_CH / bnr), channel coordinate = the tile's first channel
    const int R = PF_CH / p.bnr;
    if (!make_map_f32(&tm_dy, p.bwd.dy, (uint64_t)o.cout, (uint64_t)o.M, (uint32_t)p.bnr, (uint32_t)(64 * R))) {
      set_error("%s: cuTensorMapEncodeTiled failed (dy)", who);
      return FROST_ECUDA;
    }
    smem = 1024 + (resident ? PF_RES_KB * PF_W_BYTES + PF_BWD_STAGES * PF_X_BYTES : PF_BWD_STAGES * (PF_X_BYTES + PF_W_BYTES)) +
           PF_DY_SLOTS * PF_DY_BYTES + PF_TAIL;
  }
  if (first_use_on_device(reinterpret_cast<const void*>(&pw_fused_kernel<MODE>))) {
    cudaError_t e = cudaFuncSetAttribute(pw_fused_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute failed: %s", who, cudaGetErrorString(e));
      return FROST_ECUDA;
    }
  }
  const int64_t n_ptiles = ceil_div(o.M, PF_NPX);
  const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(n_ptiles, kNumSMs / n_ct_eff));
  cudaError_t e = launch_pdl(pw_fused_kernel<MODE>, dim3(gx, n_ct_eff), dim3(PF_THREADS), smem, st, tm_x, tm_w, tm_dy, p);
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", who, cudaGetErrorString(e));
    return FROST_ECUDA;
  }
  return FROST_OK;
}

}  // namespace frost

using namespace frost;

extern "C" int frost_debug_set_trace(long long* device_stamps) {
#ifdef FROST_TRACE
  g_trace = device_stamps;
  return FROST_OK;
#else
  (void)device_stamps;
  set_error("frost_debug_set_trace: this library was built without -DFROST_TRACE (python -m frostnet_b200.build --force --trace)");
  return FROST_ENOSUP;
#endif
}

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
  FROST_REQUIRE(a->ldq >= o.cout && a->ldq % 16 == 0 && (reinterpret_cast<uintptr_t>(a->q) & 15) == 0,
                "frost_pw_fused_forward: q must be 16-byte aligned with ldq >= cout, ldq %% 16 == 0");
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
  FROST_REQUIRE(o.cout % 8 == 0, "%s: cout=%d must be a multiple of 8", who, o.cout);
  FROST_REQUIRE((reinterpret_cast<uintptr_t>(b.dy) & 15) == 0, "%s: dy must be 16-byte aligned", who);
  FROST_REQUIRE(!apply || (b.dz && b.dz_lo && b.dz_format == 1 && (reinterpret_cast<uintptr_t>(b.dz) & 15) == 0 &&
                           (reinterpret_cast<uintptr_t>(b.dz_lo) & 15) == 0),
                "%s: needs 16-byte aligned bf16 hi/lo planes (dz_format 1)", who);
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
