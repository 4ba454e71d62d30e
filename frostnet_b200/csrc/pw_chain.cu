// pw_chain.cu - the fused BACKWARD of a 1x1 ConvBn(ReLU)2d for the expand convolutions of the Frost bottleneck (small K,
// wide cout - where 60 % of the network's activation bytes live): BatchNorm-backward "apply", the fake-quant STE / ReLU
// mask, dgrad AND wgrad in ONE kernel.  dz (the gradient wrt the raw convolution output) never leaves the SM:
//
//   per 128-pixel tile, per 128-channel tile (channel-per-thread orientation, see pw_fused.cu):
//     MMA1 (kind::i8)   I[ch][px]  = W * X^T                     recomputed accumulator, TMEM
//     epilogue          dz = P*dv + R*I + Q  (dv = dy * mask(I); per-channel coefficients in registers, dy coalesced)
//                       -> bf16 hi/lo planes written STRAIGHT INTO SHARED MEMORY in the UMMA operand layout [ch][px]
//     MMA2 (kind::f16)  dx[px][k]  += dz[px][ch] * W'[ch][k]      A = the dz tile read MN-major, B = W'^T (bf16, exact ints)
//     MMA3 (kind::f16)  dW[ch][k]  += dz^T[ch][px] * X'[px][k]    A = the SAME dz tile read K-major, B = X' (u8 -> bf16 in smem)
//   dx leaves through a pixel-per-thread epilogue (contiguous fp32 rows); dW stays in TMEM for the CTA's lifetime and is
//   flushed once with fp32 atomics.
//
// Replaces frost_pw_fused_bwd_apply + frost_pw_dgrad_tc + frost_pw_wgrad_tc (aten::native_batch_norm_backward ->
// threshold_backward -> fake_quantize backward -> convolution_backward in the reference's autograd, SURVEY.md 8a' 1-5)
// for layers with K <= 64 and 64 < cout/ceil(cout/128) (FrostNet-L: 16->96, 24->72, 24->144, 56->168, 56->336):
// 4 B/element of HBM traffic (dy) instead of 20 (dy, dz written, dz read twice).
#include <cuda_bf16.h>
#include <algorithm>
#include <cstring>
#include "pw_tma.cuh"
#include "bn_math.cuh"

namespace frost {

constexpr int CH_NP = 128;                 // pixels per tile
constexpr int CH_EPI_WARPS = 16;
constexpr int CH_EPI_THREADS = CH_EPI_WARPS * 32;
constexpr int CH_THREADS = (CH_EPI_WARPS + 2) * 32;    // warps 0-15 epilogue, 16 TMA producer, 17 MMA
constexpr int CH_MAX_CT = 3;
constexpr int CH_TILE = 16384;             // 128 rows x 128 B
constexpr int CH_PLANE = 2 * CH_TILE;      // one dz plane: [128 ch][128 px] bf16 = two 64-pixel blocks
constexpr int CH_TAIL = 1024;
// TMEM columns: I buffers at 0 / 128; dx buffer(s) of 64 columns from 256; dW of channel tile ct behind them, 64 columns each.
// Two dx buffers (n_ct <= 2) let the dx epilogue of tile t run after the dz production of tile t+1.
constexpr int CH_COL_DX = 256;

struct PwChainParams {
  const uint8_t* x;
  int64_t M;
  int K, K16, ldx;
  int cout, bn, n_ct, depth, dx_bufs;
  const int32_t* x_zp;
  const int32_t* w_zp;
  const int32_t* wsum;
  FrostBnBackwardArgs bwd;
  float* dx;
  int accumulate;
  float* dwq;
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128_lbo(uint32_t smem_addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;         // LBO: next 64-element block along M/N
  d |= (uint64_t)(1024 >> 4) << 32;              // SBO: next 8 rows along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__device__ __forceinline__ uint32_t pack2_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}

__global__ void __launch_bounds__(CH_THREADS, 1) pw_chain_bwd_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                                    const __grid_constant__ CUtensorMap tm_w,
                                                                    const __grid_constant__ CUtensorMap tm_wt,
                                                                    const __grid_constant__ PwChainParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int n_ct = p.n_ct, K16 = p.K16, depth = p.depth;
  const int wt_blk = K16 * 128;                          // one 64-channel block of W'^T: [K16 rows][128 B]
  uint8_t* s_wi8 = smem;                                 // [n_ct][128 ch][128 B]      MMA1 A
  uint8_t* s_wt = s_wi8 + n_ct * CH_TILE;                // [n_ct][2][K16][128 B]      MMA2 B
  uint8_t* s_x = s_wt + n_ct * 2 * wt_blk;               // [2][128 px][128 B] u8      MMA1 B
  uint8_t* s_xq = s_x + 2 * CH_TILE;                     // [2][128 px][64 k] bf16     MMA3 B
  uint8_t* s_dz = s_xq + 2 * CH_TILE;                    // [depth][hi, lo][2 blocks][128 ch][64 px] bf16
  uint8_t* tail = s_dz + depth * 2 * CH_PLANE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);
  uint64_t* wfull = bars;            // 1
  uint64_t* xfull = bars + 1;        // 2
  uint64_t* xempty = bars + 3;       // 2
  uint64_t* xqfull = bars + 5;       // 2
  uint64_t* xqempty = bars + 7;      // 2
  uint64_t* tfull = bars + 9;        // 2
  uint64_t* tempty = bars + 11;      // 2
  uint64_t* dzfull = bars + 13;      // 2
  uint64_t* dzempty = bars + 15;     // 2
  uint64_t* dxfull = bars + 17;      // 2
  uint64_t* dxempty = bars + 19;     // 2
  uint64_t* gfull = bars + 21;       // 1
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
  const int dx_bufs = p.dx_bufs;
  const int col_g = CH_COL_DX + 64 * dx_bufs;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t n_ptiles = (p.M + CH_NP - 1) / CH_NP;
  const int nt = (int)((n_ptiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles of this CTA (blockIdx.x < n_ptiles)
  const int n_units = nt * n_ct;

  if (threadIdx.x == 0) {
    mbar_init(wfull, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&xfull[i], 1); mbar_init(&xempty[i], 1);
      mbar_init(&xqfull[i], CH_EPI_WARPS); mbar_init(&xqempty[i], 1);
      mbar_init(&tfull[i], 1); mbar_init(&tempty[i], CH_EPI_WARPS);
      mbar_init(&dzfull[i], CH_EPI_WARPS); mbar_init(&dzempty[i], 1);
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&dxfull[i], 1); mbar_init(&dxempty[i], CH_EPI_WARPS); }
    mbar_init(gfull, 1);
    mbar_fence_init();
  }
  if (warp == CH_EPI_WARPS + 1) tmem_alloc<512>(tmem_slot);
  // rows of channels that do not exist in a tile (c_local >= n_valid) and pixels past the end are never written: zero once
  for (int i = threadIdx.x * 16; i < depth * 2 * CH_PLANE; i += CH_THREADS * 16) *reinterpret_cast<uint4*>(s_dz + i) = make_uint4(0u, 0u, 0u, 0u);
  fence_proxy_async();
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  pdl_trigger();
  const uint32_t tmem_base = *tmem_slot;
  // (the TMA producer needs neither zero point: it must not wait an L2 round trip for them before its first load)
  const int zp_a = warp == CH_EPI_WARPS ? 0 : *p.x_zp, zp_w = warp == CH_EPI_WARPS ? 0 : *p.w_zp;
  if (zp_w != 0 && zp_w != -128 && zp_w != 127) __trap();

  if (warp == CH_EPI_WARPS) {
    // ================================================================= TMA producer
    if (lane == 0) {
      tma_prefetch_desc(&tm_x);
      mbar_expect_tx(wfull, (uint32_t)(n_ct * (CH_TILE + 2 * wt_blk)));
      for (int ct = 0; ct < n_ct; ++ct) {
        tma_load_2d(&tm_w, wfull, smem_u32(s_wi8 + ct * CH_TILE), 0, ct * p.bn);
        for (int blk = 0; blk < 2; ++blk)
          tma_load_2d(&tm_wt, wfull, smem_u32(s_wt + (ct * 2 + blk) * wt_blk), ct * p.bn + 64 * blk, 0);
      }
      for (int t = 0; t < nt; ++t) {
        const int s = t & 1;
        mbar_wait_parked(&xempty[s], ((t >> 1) & 1) ^ 1);
        mbar_expect_tx(&xfull[s], CH_TILE);
        tma_load_2d(&tm_x, &xfull[s], smem_u32(s_x + s * CH_TILE), 0, (int)((blockIdx.x + (int64_t)t * gridDim.x) * CH_NP));
      }
    }
  } else if (warp == CH_EPI_WARPS + 1) {
    // ================================================================= MMA issuer
    const int w_fmt = zp_w == 0 ? 1 : 0;
    const uint32_t idesc1 = umma_idesc(2 /*S32*/, w_fmt, 0, 128, CH_NP);                     // I = W * X^T
    const uint32_t idesc2 = umma_idesc(1 /*F32*/, 1, 1, 128, K16) | (1u << 15);             // dx: A (dz) MN-major, B (W'^T) K-major
    const uint32_t idesc3 = umma_idesc(1 /*F32*/, 1, 1, 128, K16) | (1u << 16);             // dW: A (dz) K-major, B (X') MN-major
    const int nk1 = (p.K + 31) / 32;
    mbar_wait_parked(wfull, 0);
    tc_fence_after();
    // the dz-consuming MMAs of unit u-1 are issued after MMA1 of unit u: the tensor pipe always has the next accumulator
    // in flight while the epilogue turns the previous one into dz
    auto issue_dz_mmas = [&](int u) {
      const int t = u / n_ct, ct = u - t * n_ct;
      const int slot = u % depth;
      mbar_wait_parked(&dzfull[slot], (u / depth) & 1);
      const int db = t % dx_bufs;                        // dx buffer of this tile
      if (ct == 0) {
        mbar_wait_parked(&xqfull[t & 1], (t >> 1) & 1);
        mbar_wait_parked(&dxempty[db], ((t / dx_bufs) & 1) ^ 1);
      }
      tc_fence_after();
      if (lane == 0) {
        const int nkc = (min(p.bn, p.cout - ct * p.bn) + 15) / 16;      // k-steps over the channels that exist
        for (int pl = 0; pl < 2; ++pl) {
          const uint32_t dz = smem_u32(s_dz + (slot * 2 + pl) * CH_PLANE);
          // dx[px][k] += dz[px][ch] * W'[ch][k]
          const uint64_t a_mn = umma_desc_mn_sw128_lbo(dz, CH_TILE);
          for (int j = 0; j < nkc; ++j) {
            const uint64_t bdesc = umma_desc_sw128(smem_u32(s_wt + (ct * 2 + (j >> 2)) * wt_blk)) + (uint64_t)(2 * (j & 3));
            umma_f16(tmem_base + CH_COL_DX + 64 * db, a_mn + (uint64_t)((2048 * j) >> 4), bdesc, idesc2, (ct | pl | j) != 0 ? 1u : 0u);
          }
          // dW[ch][k] += dz^T[ch][px] * X'[px][k]
          const uint64_t b_mn = umma_desc_mn_sw128_lbo(smem_u32(s_xq + (t & 1) * CH_TILE), CH_TILE);
          for (int j = 0; j < 8; ++j) {
            const uint64_t adesc = umma_desc_sw128(dz + (j >> 2) * CH_TILE) + (uint64_t)(2 * (j & 3));
            umma_f16(tmem_base + col_g + 64 * ct, adesc, b_mn + (uint64_t)((2048 * j) >> 4), idesc3, (t | pl | j) != 0 ? 1u : 0u);
          }
        }
        umma_commit(&dzempty[slot]);
        if (ct == n_ct - 1) {
          umma_commit(&dxfull[db]);
          umma_commit(&xqempty[t & 1]);
          umma_commit(&xempty[t & 1]);       // MMA1 of every channel tile has completed, and X' was converted (xqfull)
        }
      }
      __syncwarp();
    };
    for (int u = 0; u < n_units; ++u) {
      const int t = u / n_ct, ct = u - t * n_ct;
      const int ib = u & 1;
      if (ct == 0) {
        mbar_wait_parked(&xfull[t & 1], (t >> 1) & 1);
      }
      mbar_wait_parked(&tempty[ib], ((u >> 1) & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t adesc = umma_desc_sw128(smem_u32(s_wi8 + ct * CH_TILE));
        const uint64_t bdesc = umma_desc_sw128(smem_u32(s_x + (t & 1) * CH_TILE));
        for (int k4 = 0; k4 < nk1; ++k4)
          umma_i8(tmem_base + ib * CH_NP, adesc + (uint64_t)(2 * k4), bdesc + (uint64_t)(2 * k4), idesc1, k4 != 0 ? 1u : 0u);
        umma_commit(&tfull[ib]);
      }
      __syncwarp();
      if (u >= 1) issue_dz_mmas(u - 1);
    }
    if (n_units > 0) issue_dz_mmas(n_units - 1);
    if (lane == 0) umma_commit(gfull);
    __syncwarp();
  } else {
    // ================================================================= epilogue (16 warps)
    const int quarter = warp & 3, grp = warp >> 2;
    const int tid = threadIdx.x;
    const int c_local = quarter * 32 + lane;               // channel-per-thread: the lane is the channel row of the tile
    const int wsign = (zp_w == 127) ? -1 : 1;
    const FrostBnBackwardArgs& b = p.bwd;
    const int relu = b.relu;
    const float inv = __fdiv_rn(1.0f, *b.out_scale), zpf = (float)*b.out_zp;
    const int cout = p.cout;
    // ---- per (thread, channel tile) constants
    int k_corr[CH_MAX_CT], k_lo[CH_MAX_CT];
    unsigned k_w[CH_MAX_CT];
    float k_P[CH_MAX_CT], k_Q[CH_MAX_CT], k_R[CH_MAX_CT];
    bool k_act[CH_MAX_CT];
#pragma unroll
    for (int ct = 0; ct < CH_MAX_CT; ++ct) {
      k_act[ct] = false; k_corr[ct] = 0; k_lo[ct] = 0; k_w[ct] = 0u; k_P[ct] = k_Q[ct] = k_R[ct] = 0.f;
      if (ct < n_ct) {
        const int n_valid = min(p.bn, cout - ct * p.bn);
        if (c_local < n_valid) {
          const int c = ct * p.bn + c_local;
          k_act[ct] = true;
          const int ws = p.wsum[c];
          const int ws_eff = zp_w == 0 ? ws : (zp_w == -128 ? ws + 128 * p.K : 127 * p.K - ws);
          k_corr[ct] = wsign * zp_a * ws_eff;
          const float A = b.A[c], B = b.B[c], mean = b.mean_I[c];
          const MaskInterval mk = bn_mask_interval(A, B, relu, inv, zpf);
          k_lo[ct] = mk.lo;
          k_w[ct] = mk.width;
          const double sa_sw = (double)(*b.x_scale) * (double)(*b.w_scale);
          const BnBwdChannel r = bn_bwd_channel(__ldcg(b.sums + 2 * c), __ldcg(b.sums + 2 * c + 1), (double)b.M, sa_sw, A, b.kfac[c], mean,
                                                b.gamma[c], b.sf[c], b.eps, b.frozen ? 0 : 1);
          k_P[ct] = r.c1;
          k_R[ct] = -r.c1 * r.a1;
          k_Q[ct] = r.c1 * (r.a1 * mean - r.a0);
          if (blockIdx.x == 0 && grp == 0) {
            b.dgamma_bn[c] = r.dgamma_bn;
            b.dbeta[c] = r.dbeta;
            b.dsf_bn[c] = r.dsf_bn;
          }
        }
      }
    }
    const float s_w = *b.w_scale, s_a = *b.x_scale;
    const float zp_magic = 8388608.0f + (float)zp_a;          // byte -> float without I2F (0x4B0000bb = 2^23 + bb)
    const int64_t row_b = (int64_t)cout * 4;
    const int col0 = grp * 32;                                // this warp's 32 pixel columns of the tile: two chunks of 16

    // dx of pixel tile t: pixel-per-thread, group g takes input channels [16g, 16g+16)
    auto dx_epilogue = [&](int t) {
      const int64_t pt = blockIdx.x + (int64_t)t * gridDim.x;
      const int px_valid = (int)min((int64_t)CH_NP, p.M - pt * CH_NP);
      const int db = t % dx_bufs;
      mbar_wait_parked(&dxfull[db], (t / dx_bufs) & 1);
      tc_fence_after();
      if (grp * 16 < K16) {
        uint32_t v[16];
        tmem_ld_32x16(tmem_base + CH_COL_DX + 64 * db + grp * 16 + ((uint32_t)(quarter * 32) << 16), v);
        const int pix = quarter * 32 + lane;
        if (pix < px_valid) {
          float* dst = p.dx + (pt * CH_NP + pix) * (int64_t)p.K + grp * 16;
          const int nk = min(16, p.K - grp * 16);               // K % 8 == 0: 8 or 16
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            if (j4 * 4 < nk) {
              float4 val = make_float4(__uint_as_float(v[4 * j4]) * s_w, __uint_as_float(v[4 * j4 + 1]) * s_w,
                                       __uint_as_float(v[4 * j4 + 2]) * s_w, __uint_as_float(v[4 * j4 + 3]) * s_w);
              float4* d4 = reinterpret_cast<float4*>(dst) + j4;
              if (p.accumulate) {
                const float4 old = ld_cg(d4);
                val.x += old.x; val.y += old.y; val.z += old.z; val.w += old.w;
              }
              *d4 = val;
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&dxempty[db]);
    };

    for (int t = 0; t < nt; ++t) {
      const int64_t pt = blockIdx.x + (int64_t)t * gridDim.x;
      const int px_valid = (int)min((int64_t)CH_NP, p.M - pt * CH_NP);
      const int s = t & 1;
      // ---- (a) X' = (x - zp_a) as bf16, [px][k] (MN-major B operand of the weight-gradient MMA)
      mbar_wait_parked(&xfull[s], (t >> 1) & 1);
      mbar_wait_parked(&xqempty[s], ((t >> 1) & 1) ^ 1);
      {
        const int px = tid >> 2, kg = tid & 3;
        if (kg * 16 < K16) {
          const uint4 raw = *reinterpret_cast<const uint4*>(s_x + s * CH_TILE + sw128_offset(px, kg));
          const unsigned w4[4] = {raw.x, raw.y, raw.z, raw.w};
          float f[16];
#pragma unroll
          for (int e = 0; e < 16; ++e) f[e] = __uint_as_float(__byte_perm(w4[e >> 2], 0x4B000000u, 0x7650u + (e & 3))) - zp_magic;
          uint8_t* dst = s_xq + s * CH_TILE;
          *reinterpret_cast<uint4*>(dst + sw128_offset(px, 2 * kg)) =
              make_uint4(pack2_bf16(f[0], f[1]), pack2_bf16(f[2], f[3]), pack2_bf16(f[4], f[5]), pack2_bf16(f[6], f[7]));
          *reinterpret_cast<uint4*>(dst + sw128_offset(px, 2 * kg + 1)) =
              make_uint4(pack2_bf16(f[8], f[9]), pack2_bf16(f[10], f[11]), pack2_bf16(f[12], f[13]), pack2_bf16(f[14], f[15]));
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&xqfull[s]);
      // ---- (b) dz of every channel tile -> shared memory
#pragma unroll
      for (int ct = 0; ct < CH_MAX_CT; ++ct) {
        if (ct < n_ct) {
          const int u = t * n_ct + ct;
          const int ib = u & 1, slot = u % depth;
          const bool act = k_act[ct];
          const int c = ct * p.bn + c_local;
          // both chunks of dy are requested before the barriers are waited for
          float dy[32];
          if (act) {
            const char* q = reinterpret_cast<const char*>(b.dy + (pt * CH_NP + col0) * (int64_t)cout + c);
            if (col0 + 32 <= px_valid) {
#pragma unroll
              for (int j = 0; j < 32; ++j) { dy[j] = ld_cg_chain(q); q += row_b; }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) { dy[j] = (col0 + j < px_valid) ? ld_cg_chain(q) : 0.0f; q += row_b; }
            }
          }
          mbar_wait_parked(&tfull[ib], (u >> 1) & 1);
          mbar_wait_parked(&dzempty[slot], ((u / depth) & 1) ^ 1);
          tc_fence_after();
          uint8_t* plane_hi = s_dz + (slot * 2 + 0) * CH_PLANE;
          uint8_t* plane_lo = s_dz + (slot * 2 + 1) * CH_PLANE;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            const int colc = col0 + 16 * hh;
            uint32_t v[16];
            tmem_ld_32x16(tmem_base + ib * CH_NP + colc + ((uint32_t)(quarter * 32) << 16), v);
            if (act) {
              float o[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int I = wsign * (int)v[j] - k_corr[ct];
                const float base = fmaf(k_R[ct], (float)I, k_Q[ct]);
                const float val = ((unsigned)(I - k_lo[ct]) < k_w[ct]) ? fmaf(k_P[ct], dy[16 * hh + j], base) : base;
                o[j] = (colc + j < px_valid) ? val : 0.0f;        // pixels past the end must not reach the weight gradient
              }
              uint32_t hi[8], lo[8];
#pragma unroll
              for (int j2 = 0; j2 < 8; ++j2) {
                const __nv_bfloat162 h2 = __floats2bfloat162_rn(o[2 * j2], o[2 * j2 + 1]);
                const float2 hf = __bfloat1622float2(h2);
                hi[j2] = *reinterpret_cast<const uint32_t*>(&h2);
                lo[j2] = pack2_bf16(o[2 * j2] - hf.x, o[2 * j2 + 1] - hf.y);
              }
              // row = my channel; 16 pixels = two 16-byte chunks of the 64-pixel block
              const int blk = colc >> 6, c16 = (colc & 63) >> 3;
              const uint32_t off0 = blk * CH_TILE + sw128_offset(c_local, c16), off1 = blk * CH_TILE + sw128_offset(c_local, c16 + 1);
              *reinterpret_cast<uint4*>(plane_hi + off0) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
              *reinterpret_cast<uint4*>(plane_hi + off1) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
              *reinterpret_cast<uint4*>(plane_lo + off0) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
              *reinterpret_cast<uint4*>(plane_lo + off1) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
            }
          }
          tc_fence_before();
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            mbar_arrive(&tempty[ib]);
            mbar_arrive(&dzfull[slot]);
          }
        }
      }
      // ---- (c) dx of the PREVIOUS tile (two dx buffers): its dgrad MMAs had this tile's whole dz production to finish
      if (dx_bufs == 2) {
        if (t > 0) dx_epilogue(t - 1);
      } else {
        dx_epilogue(t);
      }
    }
    if (dx_bufs == 2 && nt > 0) dx_epilogue(nt - 1);
    // ---- weight gradient: dW[ch][k] of every channel tile, accumulated over all of this CTA's pixels
    mbar_wait_parked(gfull, 0);
    tc_fence_after();
#pragma unroll
    for (int ct = 0; ct < CH_MAX_CT; ++ct) {
      if (ct < n_ct && grp * 16 < K16) {
        uint32_t v[16];
        tmem_ld_32x16(tmem_base + col_g + 64 * ct + grp * 16 + ((uint32_t)(quarter * 32) << 16), v);
        if (k_act[ct] && nt > 0) {
          float* dst = p.dwq + (int64_t)(ct * p.bn + c_local) * p.K + grp * 16;
          const int nk = min(16, p.K - grp * 16);
#pragma unroll
          for (int j = 0; j < 16; ++j)
            if (j < nk) atomicAdd(dst + j, __uint_as_float(v[j]) * s_a);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == CH_EPI_WARPS + 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace frost

using namespace frost;

extern "C" int frost_pw_chain_supported(int K, int cout) {
  if (K <= 0 || cout <= 0 || K % 8 != 0 || cout % 8 != 0 || K > 64) return 0;
  const int n_ct = (cout + 127) / 128;
  if (n_ct > CH_MAX_CT) return 0;
  const int bn = (((cout + n_ct - 1) / n_ct) + 15) & ~15;
  return bn > 64 ? 1 : 0;          // narrower tiles replicate weight rows in the other kernels; not needed here
}

extern "C" int frost_pw_chain_backward(const FrostPwChainArgs* a, void* stream) {
  const char* who = "frost_pw_chain_backward";
  FROST_REQUIRE(a, "%s: null args", who);
  const FrostPwOperands& o = a->op;
  const FrostBnBackwardArgs& b = a->bn;
  FROST_REQUIRE(o.x && o.w_mma && o.x_zp && o.w_zp && o.wsum && a->wt_bf16 && a->dwq, "%s: null pointer", who);
  FROST_REQUIRE(b.dy && b.A && b.B && b.mean_I && b.kfac && b.gamma && b.sf && b.x_scale && b.w_scale && b.out_scale && b.out_zp &&
                    b.sums && b.dgamma_bn && b.dbeta && b.dsf_bn,
                "%s: null pointer", who);
  FROST_REQUIRE(frost_pw_chain_supported(o.K, o.cout), "%s: K=%d cout=%d is outside the chained kernel's range (K <= 64, wide cout)", who, o.K, o.cout);
  FROST_REQUIRE(o.M > 0 && o.M < ((int64_t)1 << 31) && b.C == o.cout && b.M == o.M, "%s: bad problem size", who);
  FROST_REQUIRE(o.ldx >= o.K && o.ldx % 16 == 0 && o.ldw >= o.K && o.ldw % 16 == 0, "%s: row pitches must be multiples of 16", who);
  FROST_REQUIRE((reinterpret_cast<uintptr_t>(o.x) & 15) == 0 && (reinterpret_cast<uintptr_t>(o.w_mma) & 15) == 0 &&
                    (reinterpret_cast<uintptr_t>(a->wt_bf16) & 15) == 0 && (reinterpret_cast<uintptr_t>(b.dy) & 15) == 0 &&
                    (!a->dx || (reinterpret_cast<uintptr_t>(a->dx) & 15) == 0),
                "%s: operands must be 16-byte aligned", who);
  if (!encode_fn()) {
    set_error("%s: cuTensorMapEncodeTiled is not available from this driver", who);
    return FROST_ENOSUP;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(a->dwq, 0, sizeof(float) * (size_t)o.K * o.cout, st) != cudaSuccess) {
    set_error("%s: memset failed", who);
    return FROST_ECUDA;
  }
  PwChainParams p;
  memset(&p, 0, sizeof(p));
  p.x = o.x; p.M = o.M; p.K = o.K; p.ldx = o.ldx; p.cout = o.cout;
  p.K16 = (o.K + 15) & ~15;
  p.n_ct = (o.cout + 127) / 128;
  p.bn = (((o.cout + p.n_ct - 1) / p.n_ct) + 15) & ~15;
  p.x_zp = o.x_zp; p.w_zp = o.w_zp; p.wsum = o.wsum;
  p.bwd = b;
  p.dx = a->dx;
  p.accumulate = a->accumulate;
  p.dwq = a->dwq;
  const size_t fixed = 1024 + (size_t)p.n_ct * (CH_TILE + 2 * p.K16 * 128) + 4 * CH_TILE + CH_TAIL;
  p.depth = (fixed + 2 * 2 * CH_PLANE <= 227 * 1024) ? 2 : 1;
  p.dx_bufs = p.n_ct <= 2 ? 2 : 1;                 // TMEM: 256 (I) + 64 * dx_bufs + 64 * n_ct <= 512
  const size_t smem = fixed + (size_t)p.depth * 2 * CH_PLANE;
  FROST_REQUIRE(smem <= 227 * 1024, "%s: shared memory budget exceeded (%zu bytes)", who, smem);
  FROST_REQUIRE(a->dx, "%s: dx is required (layers without an input gradient use the unfused path)", who);
  CUtensorMap tm_x, tm_w, tm_wt;
  if (!make_map_u8(&tm_x, o.x, (uint64_t)o.K, (uint64_t)o.M, (uint64_t)o.ldx, 128, CH_NP) ||
      !make_map_u8(&tm_w, o.w_mma, (uint64_t)o.K, (uint64_t)o.cout, (uint64_t)o.ldw, 128, 128) ||
      !make_map_b16(&tm_wt, a->wt_bf16, (uint64_t)o.cout, (uint64_t)o.K, (uint64_t)o.cout * 2, 64, (uint32_t)p.K16)) {
    set_error("%s: cuTensorMapEncodeTiled failed", who);
    return FROST_ECUDA;
  }
  if (first_use_on_device(reinterpret_cast<const void*>(&pw_chain_bwd_kernel))) {
    cudaError_t e = cudaFuncSetAttribute(pw_chain_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      set_error("%s: cudaFuncSetAttribute failed: %s", who, cudaGetErrorString(e));
      return FROST_ECUDA;
    }
  }
  const int64_t n_ptiles = ceil_div(o.M, CH_NP);
  const int gx = (int)std::min<int64_t>(n_ptiles, kNumSMs);
  cudaError_t e = launch_pdl(pw_chain_bwd_kernel, dim3(gx), dim3(CH_THREADS), smem, st, tm_x, tm_w, tm_wt, p);
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", who, cudaGetErrorString(e));
    return FROST_ECUDA;
  }
  FROST_LAUNCH_CHECK(who);
  return FROST_OK;
}
