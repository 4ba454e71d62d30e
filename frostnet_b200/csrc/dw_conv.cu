// dw_conv.cu - depthwise kxk convolution (k in {3,5}, stride in {1,2}) forward / dgrad / wgrad on NHWC
// quantize indices: frostnet.py:116-118 (conv2 of the Frost bottleneck) after fuse + prepare_qat, i.e. the
// F.conv2d(groups=C) at conv_fused.py:155 and its aten::convolution_backward.
//
// Common shape: a thread owns ONE 4-channel group (one 32-bit word of uint8 indices / one float4 of
// gradient per pixel) for its whole lifetime and walks strips of 4 pixels along W.  A block covers
// `cgb` channel groups (<= 64) x several strips; the per-channel-group weights live in shared memory.
// Per kernel row the thread issues all (predicated, coalesced) loads of the input window first, then
// the MACs - out-of-image taps read the zero-point index, i.e. contribute (q - zp_a) = 0 exactly like
// the reference's zero padding of the dequantised tensor.
#include "common.cuh"

namespace frost {

struct SmemStat {
  long long sum;
  unsigned long long sq;
  int mn, mx;
};

constexpr int DW_TW = 4;
constexpr int DW_MAX_CGB = 64;

// channel-group chunking shared by the per-channel kernels: threads per block is a multiple of the number
// of 4-channel groups handled by the block, so a thread's channel group never changes.
void dw_launch_shape(int C, int max_cgb, int* cg_per_block, int* nchunks, int* threads) {
  const int CG = C / 4;
  int chunks = (CG + max_cgb - 1) / max_cgb;
  while (CG % chunks != 0 && chunks < CG) ++chunks;
  const int cgb = CG / chunks;
  *cg_per_block = cgb;
  *nchunks = chunks;
  *threads = cgb * (256 / cgb);
}

__device__ __forceinline__ int sext_byte(unsigned w, int ch) {
  // PRMT with the sign-replicate bit: byte `ch` sign-extended to 32 bits in one instruction
  // (inline PTX: __byte_perm documents only 3 selector bits per nibble)
  const unsigned sel = (unsigned)ch | ((8u | ch) << 4) | ((8u | ch) << 8) | ((8u | ch) << 12);
  int d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(0u), "r"(sel));
  return d;
}
__device__ __forceinline__ int zext_byte(unsigned w, int ch) { return (int)((w >> (8 * ch)) & 0xffu); }

// Strip walker: strip index -> (strip column sw, row oh, image n) advanced by a fixed step with adds and
// carries only (a 64-bit div/mod pair per strip costs more instructions than a 3x3 strip's arithmetic).
struct StripWalker {
  int sw, oh, n;
  int d_sw, d_oh, d_n;
  int strips_w, rows;
  __device__ __forceinline__ void init(int64_t first, int64_t step, int strips_w_, int rows_) {
    strips_w = strips_w_;
    rows = rows_;
    sw = (int)(first % strips_w);
    const int64_t t1 = first / strips_w;
    oh = (int)(t1 % rows);
    n = (int)(t1 / rows);
    d_sw = (int)(step % strips_w);
    const int64_t s1 = step / strips_w;
    d_oh = (int)(s1 % rows);
    d_n = (int)(s1 / rows);
  }
  __device__ __forceinline__ void next() {
    sw += d_sw;
    int carry = 0;
    if (sw >= strips_w) { sw -= strips_w; carry = 1; }
    oh += d_oh + carry;
    carry = 0;
    if (oh >= rows) { oh -= rows; carry = 1; }
    n += d_n + carry;
  }
};

// ================================================================= forward (dp4a)
// Per kernel row the thread loads the strip's input window (one 32-bit word = 4 channels per pixel),
// transposes 4x4 byte blocks in registers (8 PRMT) so that one word holds 4 CONSECUTIVE PIXELS of one channel,
// and then takes 4 taps per dp4a against per-channel packed weight words (taps 0-3; tap 4 in a second word):
// ~1 instruction per MAC instead of ~2.3 with byte unpacking.  Everything stays on raw indices:
//   I = sum x*w - zp_a*sum_taps(w) - zp_w*sum_taps(x) + K*zp_a*zp_w     (out-of-image taps read x = zp_a).
__device__ __forceinline__ int dp4a_us(unsigned a, unsigned b, int c) {
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int dp4a_uu(unsigned a, unsigned b, int c) {
  int d;
  asm("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
// rows = pixels (a,b,c,d), bytes = channels  ->  T[ch] = (a.ch, b.ch, c.ch, d.ch)
__device__ __forceinline__ void transpose4x4(unsigned a, unsigned b, unsigned c, unsigned d, unsigned (&T)[4]) {
  const unsigned t0 = __byte_perm(a, b, 0x5140), t1 = __byte_perm(c, d, 0x5140);
  const unsigned t2 = __byte_perm(a, b, 0x7362), t3 = __byte_perm(c, d, 0x7362);
  T[0] = __byte_perm(t0, t1, 0x5410);
  T[1] = __byte_perm(t0, t1, 0x7632);
  T[2] = __byte_perm(t2, t3, 0x5410);
  T[3] = __byte_perm(t2, t3, 0x7632);
}

template <int KS, int S>
__global__ void __launch_bounds__(256, 2) dw_conv_fwd_kernel(const uint8_t* xq, const int32_t* x_zp_p,
                                                            const int8_t* wq, const int32_t* w_zp_p,
                                                            int N, int H, int W, int C, int ldx, int Ho, int Wo, int cgb,
                                                            int32_t* acc_out, FrostChanStats* stats) {
  extern __shared__ __align__(16) unsigned char dw_smem[];
  pdl_enter();
  constexpr int PAD = (KS - 1) / 2;
  constexpr int IW = (DW_TW - 1) * S + KS;       // input pixels touched by a strip
  constexpr int NB = (IW + 3) / 4;               // 4-pixel blocks
  unsigned* s_wA = reinterpret_cast<unsigned*>(dw_smem);          // [KS][4][cgb] taps 0-3 of (row, channel)
  unsigned* s_wB = s_wA + KS * 4 * DW_MAX_CGB;                    // [KS][4][cgb] tap 4 (KS == 5)
  int* s_wsum = reinterpret_cast<int*>(s_wB + KS * 4 * DW_MAX_CGB);  // [cgb*4] sum over all taps of w
  SmemStat* s_stat = reinterpret_cast<SmemStat*>(s_wsum + DW_MAX_CGB * 4);  // [cgb*4]
  const int zp_a = *x_zp_p, zp_w = *w_zp_p;
  if (zp_w != 0) return;                     // one-signed weights: dw_conv_fwd_generic_kernel takes over
  const unsigned zp4 = (unsigned)zp_a * 0x01010101u;
  const int cg_local = threadIdx.x % cgb;
  const int cg = blockIdx.y * cgb + cg_local;
  const int spb = blockDim.x / cgb;
  const int strip_local = threadIdx.x / cgb;

  for (int i = threadIdx.x; i < KS * 4 * cgb; i += blockDim.x) {
    const int g = i % cgb, ch = (i / cgb) & 3, r = i / (cgb * 4);
    const int c = (blockIdx.y * cgb + g) * 4 + ch;
    unsigned wa = 0u, wb = 0u;
#pragma unroll
    for (int dx = 0; dx < KS; ++dx) {
      const unsigned byte = (unsigned)(uint8_t)wq[(int64_t)(r * KS + dx) * C + c];
      if (dx < 4) wa |= byte << (8 * dx);
      else wb |= byte;
    }
    s_wA[(r * 4 + ch) * cgb + g] = wa;
    s_wB[(r * 4 + ch) * cgb + g] = wb;
  }
  for (int i = threadIdx.x; i < cgb * 4; i += blockDim.x) {
    const int c = blockIdx.y * cgb * 4 + i;
    int sum = 0;
    for (int t = 0; t < KS * KS; ++t) sum += (int)wq[(int64_t)t * C + c];
    s_wsum[i] = sum;
    s_stat[i].sum = 0; s_stat[i].sq = 0; s_stat[i].mn = INT_MAX; s_stat[i].mx = INT_MIN;
  }
  __syncthreads();
  int corr[4];
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) corr[ch] = -zp_a * s_wsum[cg_local * 4 + ch] + KS * KS * zp_a * zp_w;

  long long st_sum[4] = {0, 0, 0, 0};
  unsigned long long st_sq[4] = {0, 0, 0, 0};
  int st_mn[4] = {INT_MAX, INT_MAX, INT_MAX, INT_MAX};
  int st_mx[4] = {INT_MIN, INT_MIN, INT_MIN, INT_MIN};

  const int strips_w = (Wo + DW_TW - 1) / DW_TW;
  StripWalker wk;
  wk.init((int64_t)blockIdx.x * spb + strip_local, (int64_t)gridDim.x * spb, strips_w, Ho);
  for (; wk.n < N; wk.next()) {
    const int oh = wk.oh, n = wk.n;
    const int ow0 = wk.sw * DW_TW;
    int acc[DW_TW][4];
#pragma unroll
    for (int t = 0; t < DW_TW; ++t)
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) acc[t][ch] = 0;
    // all KS rows of the window are requested before any arithmetic: one memory latency per strip, not KS.
    // One 64-bit base per strip; every load adds a 32-bit (row, column) offset to it (a 64-bit multiply per load
    // was ~7 address instructions for each of the 24-60 loads of a strip - these kernels are issue-bound).
    unsigned xw_all[KS][NB * 4];
    const int ih0 = oh * S - PAD, iw0 = ow0 * S - PAD;
    const uint8_t* pbase = xq + (((int64_t)n * H + ih0) * W + iw0) * ldx + cg * 4;   // only dereferenced inside the image
    const int rowstep = W * ldx;                 // input pixels are ldx bytes apart (ldx == C: dense NHWC)
#pragma unroll
    for (int r = 0; r < KS; ++r) {
      const bool rok = (unsigned)(ih0 + r) < (unsigned)H;
      const uint8_t* row = pbase + r * rowstep;
#pragma unroll
      for (int j = 0; j < NB * 4; ++j) {
        const bool ok = (j < IW) && rok && ((unsigned)(iw0 + j) < (unsigned)W);
        xw_all[r][j] = ok ? ld_cg(reinterpret_cast<const unsigned*>(row + j * ldx)) : zp4;
      }
    }
#pragma unroll
    for (int r = 0; r < KS; ++r) {
      const unsigned* xw = xw_all[r];
      unsigned T[NB + 1][4];
#pragma unroll
      for (int b4 = 0; b4 < NB; ++b4) transpose4x4(xw[4 * b4], xw[4 * b4 + 1], xw[4 * b4 + 2], xw[4 * b4 + 3], T[b4]);
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) T[NB][ch] = 0u;
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) {
        const unsigned wa = s_wA[(r * 4 + ch) * cgb + cg_local];
        const unsigned wb = (KS == 5) ? s_wB[(r * 4 + ch) * cgb + cg_local] : 0u;
#pragma unroll
        for (int t = 0; t < DW_TW; ++t) {
          constexpr int dummy = 0; (void)dummy;
          const int start = t * S;                               // first input pixel of output t (compile-time)
          const int b0 = start >> 2, o0 = start & 3;
          const unsigned win = (o0 == 0) ? T[b0][ch] : __byte_perm(T[b0][ch], T[b0 + 1][ch], 0x3210u + 0x1111u * o0);
          acc[t][ch] = dp4a_us(win, wa, acc[t][ch]);
          unsigned x4 = 0u;
          if (KS == 5) {
            const int s4 = start + 4, b1 = s4 >> 2, o1 = s4 & 3;
            x4 = (o1 == 0) ? T[b1][ch] : __byte_perm(T[b1][ch], 0u, (unsigned)o1);
            acc[t][ch] = dp4a_us(x4, wb, acc[t][ch]);
          }
        }
      }
    }
    int32_t* obase = acc_out + (((int64_t)n * Ho + oh) * Wo + ow0) * C + cg * 4;
#pragma unroll
    for (int t = 0; t < DW_TW; ++t) {
      const int ow = ow0 + t;
      if (ow < Wo) {
        int I[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) I[ch] = acc[t][ch] + corr[ch];
        *reinterpret_cast<int4*>(obase + t * C) = make_int4(I[0], I[1], I[2], I[3]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          st_sum[ch] += I[ch];
          st_sq[ch] += (unsigned long long)((long long)I[ch] * (long long)I[ch]);
          st_mn[ch] = min(st_mn[ch], I[ch]);
          st_mx[ch] = max(st_mx[ch], I[ch]);
        }
      }
    }
  }
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    if (st_mn[ch] <= st_mx[ch]) {
      SmemStat* s = &s_stat[cg_local * 4 + ch];
      atomicAdd(reinterpret_cast<unsigned long long*>(&s->sum), (unsigned long long)st_sum[ch]);
      atomicAdd(&s->sq, st_sq[ch]);
      atomicMin(&s->mn, st_mn[ch]);
      atomicMax(&s->mx, st_mx[ch]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cgb * 4; i += blockDim.x) {
    const SmemStat s = s_stat[i];
    chan_stats_flush(stats + blockIdx.y * cgb * 4 + i, s.sum, s.sq, s.mn, s.mx);
  }
}

// ================================================================= forward, shared-memory tiles (dp4a)
// The register-gather kernel above keeps one strip's window in flight per thread and is latency-bound (13 warps
// per SM, one L2 round trip per strip: 5-8x off the HBM time on the 14x14 / 7x7 layers).  Here a CTA owns CGB
// channel groups and walks tiles of (one image, TH output rows, full width): all 256 threads first copy the
// input window of the tile - zero-point bytes where it leaves the image - into shared memory, many independent
// loads in flight per thread, then every thread computes strips of 4 output pixels from shared memory with the
// same transposed-dp4a arithmetic.
// Tile layout: 32-bit words (4 channels of one pixel) as [row][channel group][col], col = iw + PAD, pitch = 4 mod 32
// words: a strip's window is NB aligned 16-byte reads, conflict-free across the channel groups of a quarter warp,
// and the fill (lanes = CGB groups x 32/CGB pixels) hits 32 distinct banks.
template <int KS, int S, int CGB>
__global__ void __launch_bounds__(256, 2) dw_conv_fwd_tiled_kernel(const uint8_t* xq, const int32_t* x_zp_p, const int8_t* wq,
                                                                  const int32_t* w_zp_p, int N, int H, int W, int C, int ldx, int Ho, int Wo,
                                                                  int TH, int PITCH, int32_t* acc_out, FrostChanStats* stats) {
  extern __shared__ __align__(16) unsigned char dw_smem[];
  pdl_enter();
  constexpr int PAD = (KS - 1) / 2;
  constexpr int IW = (DW_TW - 1) * S + KS;
  constexpr int NB = (IW + 3) / 4;
  constexpr int SLOTS = 256 / CGB;
  unsigned* s_wA = reinterpret_cast<unsigned*>(dw_smem);            // [KS][4][CGB]
  unsigned* s_wB = s_wA + KS * 4 * CGB;                             // [KS][4][CGB]
  int* s_wsum = reinterpret_cast<int*>(s_wB + KS * 4 * CGB);        // [CGB*4]
  SmemStat* s_stat = reinterpret_cast<SmemStat*>(s_wsum + CGB * 4); // [CGB*4]
  unsigned* s_tile = reinterpret_cast<unsigned*>(s_stat + CGB * 4); // [IH][CGB][PITCH]
  const int zp_a = *x_zp_p, zp_w = *w_zp_p;
  if (zp_w != 0) return;                     // one-signed weights: dw_conv_fwd_generic_kernel takes over
  const unsigned zp4 = (unsigned)zp_a * 0x01010101u;
  const int CG = C >> 2;
  const int cg0 = blockIdx.y * CGB;
  const int ncg = min(CGB, CG - cg0);
  const int cg_l = threadIdx.x % CGB;
  const int slot = threadIdx.x / CGB;
  const bool cg_ok = cg_l < ncg;
  const int c_first = (cg0 + cg_l) * 4;

  for (int i = threadIdx.x; i < KS * 4 * CGB; i += blockDim.x) {
    const int g = i % CGB, ch = (i / CGB) & 3, r = i / (CGB * 4);
    const int c = (cg0 + g) * 4 + ch;
    unsigned wa = 0u, wb = 0u;
    if (g < ncg) {
#pragma unroll
      for (int dx = 0; dx < KS; ++dx) {
        const unsigned byte = (unsigned)(uint8_t)wq[(int64_t)(r * KS + dx) * C + c];
        if (dx < 4) wa |= byte << (8 * dx);
        else wb |= byte;
      }
    }
    s_wA[(r * 4 + ch) * CGB + g] = wa;
    s_wB[(r * 4 + ch) * CGB + g] = wb;
  }
  for (int i = threadIdx.x; i < CGB * 4; i += blockDim.x) {
    int sum = 0;
    if ((i >> 2) < ncg)
      for (int t = 0; t < KS * KS; ++t) sum += (int)wq[(int64_t)t * C + cg0 * 4 + i];
    s_wsum[i] = sum;
    s_stat[i].sum = 0; s_stat[i].sq = 0; s_stat[i].mn = INT_MAX; s_stat[i].mx = INT_MIN;
  }
  __syncthreads();
  int corr[4];
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) corr[ch] = -zp_a * s_wsum[cg_l * 4 + ch];

  long long st_sum[4] = {0, 0, 0, 0};
  unsigned long long st_sq[4] = {0, 0, 0, 0};
  int st_mn[4] = {INT_MAX, INT_MAX, INT_MAX, INT_MAX};
  int st_mx[4] = {INT_MIN, INT_MIN, INT_MIN, INT_MIN};

  const int strips_w = (Wo + DW_TW - 1) / DW_TW;
  const int fill_w = (strips_w - 1) * DW_TW * S + NB * 4;     // tile columns any strip may read (<= PITCH)
  const int tiles_h = (Ho + TH - 1) / TH;
  const int d_row = SLOTS / strips_w, d_sw = SLOTS % strips_w;
  for (int tile = blockIdx.x; tile < N * tiles_h; tile += gridDim.x) {
    const int n = tile / tiles_h;
    const int oh0 = (tile - n * tiles_h) * TH;
    const int th = min(TH, Ho - oh0);
    const int ih_rows = (th - 1) * S + KS;
    const int ih0 = oh0 * S - PAD;
    __syncthreads();                                          // the previous tile has been consumed
    // ---- fill: thread = (channel group, column slot); 2 rows x 4 columns of loads in flight
    for (int row = 0; row < ih_rows; row += 2) {
      unsigned v[2][4];
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        const int ih = ih0 + row + rr;
        const bool rok = cg_ok && (row + rr < ih_rows) && ((unsigned)ih < (unsigned)H);
        const uint8_t* src = xq + (((int64_t)n * H + (rok ? ih : 0)) * W) * ldx + c_first;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = slot + q * SLOTS;
          const int iw = col - PAD;
          v[rr][q] = (rok && col < fill_w && (unsigned)iw < (unsigned)W) ? ld_cg(reinterpret_cast<const unsigned*>(src + (int64_t)iw * ldx)) : zp4;
        }
      }
#pragma unroll
      for (int rr = 0; rr < 2; ++rr) {
        if (row + rr < ih_rows) {
          unsigned* dst = s_tile + ((size_t)(row + rr) * CGB + cg_l) * PITCH;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int col = slot + q * SLOTS;
            if (col < fill_w) dst[col] = v[rr][q];
          }
        }
      }
      for (int col = slot + 4 * SLOTS; col < fill_w; col += SLOTS) {   // wider than 4*SLOTS columns (never for W <= 112)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int ih = ih0 + row + rr;
          if (row + rr >= ih_rows) continue;
          const int iw = col - PAD;
          const bool ok = cg_ok && ((unsigned)ih < (unsigned)H) && ((unsigned)iw < (unsigned)W);
          s_tile[((size_t)(row + rr) * CGB + cg_l) * PITCH + col] =
              ok ? ld_cg(reinterpret_cast<const unsigned*>(xq + (((int64_t)n * H + ih) * W + iw) * ldx + c_first)) : zp4;
        }
      }
    }
    __syncthreads();
    // ---- compute: strips of 4 output pixels, window rows read from shared memory
    if (cg_ok) {
      int orow = slot / strips_w, sw = slot - orow * strips_w;
      for (; orow < th; ) {
        int acc[DW_TW][4];
#pragma unroll
        for (int t = 0; t < DW_TW; ++t)
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) acc[t][ch] = 0;
#pragma unroll
        for (int r = 0; r < KS; ++r) {
          const uint4* src = reinterpret_cast<const uint4*>(s_tile + ((size_t)(orow * S + r) * CGB + cg_l) * PITCH + sw * DW_TW * S);
          unsigned T[NB + 1][4];
#pragma unroll
          for (int b4 = 0; b4 < NB; ++b4) {
            const uint4 w4 = src[b4];
            transpose4x4(w4.x, w4.y, w4.z, w4.w, T[b4]);
          }
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) T[NB][ch] = 0u;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const unsigned wa = s_wA[(r * 4 + ch) * CGB + cg_l];
            const unsigned wb = (KS == 5) ? s_wB[(r * 4 + ch) * CGB + cg_l] : 0u;
#pragma unroll
            for (int t = 0; t < DW_TW; ++t) {
              const int start = t * S;
              const int b0 = start >> 2, o0 = start & 3;
              const unsigned win = (o0 == 0) ? T[b0][ch] : __byte_perm(T[b0][ch], T[b0 + 1][ch], 0x3210u + 0x1111u * o0);
              acc[t][ch] = dp4a_us(win, wa, acc[t][ch]);
              if (KS == 5) {
                const int s4 = start + 4, b1 = s4 >> 2, o1 = s4 & 3;
                const unsigned x4 = (o1 == 0) ? T[b1][ch] : __byte_perm(T[b1][ch], 0u, (unsigned)o1);
                acc[t][ch] = dp4a_us(x4, wb, acc[t][ch]);
              }
            }
          }
        }
        const int oh = oh0 + orow, ow0 = sw * DW_TW;
#pragma unroll
        for (int t = 0; t < DW_TW; ++t) {
          const int ow = ow0 + t;
          if (ow < Wo) {
            int I[4];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) I[ch] = acc[t][ch] + corr[ch];
            *reinterpret_cast<int4*>(acc_out + (((int64_t)n * Ho + oh) * Wo + ow) * C + c_first) = make_int4(I[0], I[1], I[2], I[3]);
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              st_sum[ch] += I[ch];
              st_sq[ch] += (unsigned long long)((long long)I[ch] * (long long)I[ch]);
              st_mn[ch] = min(st_mn[ch], I[ch]);
              st_mx[ch] = max(st_mx[ch], I[ch]);
            }
          }
        }
        sw += d_sw;
        orow += d_row;
        if (sw >= strips_w) { sw -= strips_w; ++orow; }
      }
    }
  }
  if (cg_ok) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
      if (st_mn[ch] <= st_mx[ch]) {
        SmemStat* sst = &s_stat[cg_l * 4 + ch];
        atomicAdd(reinterpret_cast<unsigned long long*>(&sst->sum), (unsigned long long)st_sum[ch]);
        atomicAdd(&sst->sq, st_sq[ch]);
        atomicMin(&sst->mn, st_mn[ch]);
        atomicMax(&sst->mx, st_mx[ch]);
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < ncg * 4; i += blockDim.x) {
    const SmemStat t = s_stat[i];
    chan_stats_flush(stats + cg0 * 4 + i, t.sum, t.sq, t.mn, t.mx);
  }
}

// ================================================================= forward, generic zero-points (byte unpacking)
// Only runs when the weight zero-point is not 0 (one-signed weight tensor: SURVEY K5); exits otherwise.
template <int KS, int S>
__global__ void __launch_bounds__(256, 2) dw_conv_fwd_generic_kernel(const uint8_t* xq, const int32_t* x_zp_p,
                                                            const int8_t* wq, const int32_t* w_zp_p,
                                                            int N, int H, int W, int C, int ldx, int Ho, int Wo, int cgb,
                                                            int32_t* acc_out, FrostChanStats* stats) {
  extern __shared__ __align__(16) unsigned char dw_smem[];
  pdl_enter();
  unsigned* s_w = reinterpret_cast<unsigned*>(dw_smem);                               // [KS*KS][cgb] packed int8x4
  SmemStat* s_stat = reinterpret_cast<SmemStat*>(dw_smem + sizeof(unsigned) * KS * KS * DW_MAX_CGB);  // [cgb*4]
  constexpr int PAD = (KS - 1) / 2;
  constexpr int IW = (DW_TW - 1) * S + KS;
  const int zp_a = *x_zp_p, zp_w = *w_zp_p;
  if (zp_w == 0) return;                     // the dp4a kernel handles the symmetric case
  const unsigned zp4 = (unsigned)zp_a * 0x01010101u;
  const int cg_local = threadIdx.x % cgb;
  const int cg = blockIdx.y * cgb + cg_local;
  const int spb = blockDim.x / cgb;
  const int strip_local = threadIdx.x / cgb;

  for (int i = threadIdx.x; i < KS * KS * cgb; i += blockDim.x) {
    const int t = i / cgb, g = i % cgb;
    s_w[t * cgb + g] = ld_cg(reinterpret_cast<const unsigned*>(wq + (int64_t)t * C + (blockIdx.y * cgb + g) * 4));
  }
  for (int i = threadIdx.x; i < cgb * 4; i += blockDim.x) {
    s_stat[i].sum = 0; s_stat[i].sq = 0; s_stat[i].mn = INT_MAX; s_stat[i].mx = INT_MIN;
  }
  __syncthreads();

  long long st_sum[4] = {0, 0, 0, 0};
  unsigned long long st_sq[4] = {0, 0, 0, 0};
  int st_mn[4] = {INT_MAX, INT_MAX, INT_MAX, INT_MAX};
  int st_mx[4] = {INT_MIN, INT_MIN, INT_MIN, INT_MIN};

  const int strips_w = (Wo + DW_TW - 1) / DW_TW;
  StripWalker wk;
  wk.init((int64_t)blockIdx.x * spb + strip_local, (int64_t)gridDim.x * spb, strips_w, Ho);
  for (; wk.n < N; wk.next()) {
    const int oh = wk.oh, n = wk.n;
    const int ow0 = wk.sw * DW_TW;
    int acc[DW_TW][4];
#pragma unroll
    for (int t = 0; t < DW_TW; ++t)
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) acc[t][ch] = 0;
#pragma unroll
    for (int r = 0; r < KS; ++r) {
      const int ih = oh * S - PAD + r, iw0 = ow0 * S - PAD;
      const bool rok = (unsigned)ih < (unsigned)H;
      const uint8_t* row = xq + (((int64_t)n * H + ih) * W + iw0) * ldx + cg * 4;    // only dereferenced inside the image
      unsigned xw[IW];
#pragma unroll
      for (int j = 0; j < IW; ++j) {
        const bool ok = rok && ((unsigned)(iw0 + j) < (unsigned)W);
        xw[j] = ok ? ld_cg(reinterpret_cast<const unsigned*>(row + j * ldx)) : zp4;
      }
      int wr[KS][4];
#pragma unroll
      for (int dx = 0; dx < KS; ++dx) {
        const unsigned pk = s_w[(r * KS + dx) * cgb + cg_local];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) wr[dx][ch] = sext_byte(pk, ch) - zp_w;
      }
#pragma unroll
      for (int j = 0; j < IW; ++j) {
        int xa[4];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) xa[ch] = zext_byte(xw[j], ch) - zp_a;
#pragma unroll
        for (int t = 0; t < DW_TW; ++t) {
          const int dx = j - t * S;
          if (dx >= 0 && dx < KS) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) acc[t][ch] += xa[ch] * wr[dx][ch];
          }
        }
      }
    }
#pragma unroll
    for (int t = 0; t < DW_TW; ++t) {
      const int ow = ow0 + t;
      if (ow < Wo) {
        *reinterpret_cast<int4*>(acc_out + (((int64_t)n * Ho + oh) * Wo + ow) * C + cg * 4) =
            make_int4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const int I = acc[t][ch];
          st_sum[ch] += I;
          st_sq[ch] += (unsigned long long)((long long)I * (long long)I);
          st_mn[ch] = min(st_mn[ch], I);
          st_mx[ch] = max(st_mx[ch], I);
        }
      }
    }
  }
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    if (st_mn[ch] <= st_mx[ch]) {
      SmemStat* s = &s_stat[cg_local * 4 + ch];
      atomicAdd(reinterpret_cast<unsigned long long*>(&s->sum), (unsigned long long)st_sum[ch]);
      atomicAdd(&s->sq, st_sq[ch]);
      atomicMin(&s->mn, st_mn[ch]);
      atomicMax(&s->mx, st_mx[ch]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < cgb * 4; i += blockDim.x) {
    const SmemStat s = s_stat[i];
    chan_stats_flush(stats + blockIdx.y * cgb * 4 + i, s.sum, s.sq, s.mn, s.mx);
  }
}

// ================================================================= dgrad
// dx[n][ih][iw][c] (+)= sum_{r,s} dz[n][oh][ow][c] * wf[r][s][c],  oh*S - PAD + r == ih, ow*S - PAD + s == iw.
// Strips start at multiples of 4 along W, so for stride 2 the set of (pixel, tap) pairs that hit an
// output column is known at compile time; valid kernel rows are a run-time (per-thread) stride-S loop.
template <int KS, int S>
__global__ void __launch_bounds__(256, 4) dw_dgrad_kernel(const float* dz, const int8_t* wq,
                                                         const float* w_scale_p, const int32_t* w_zp_p,
                                                         int N, int H, int W, int C, int Ho, int Wo, int cgb,
                                                         float* dx, int accumulate) {
  extern __shared__ __align__(16) unsigned char dw_smem[];
  pdl_enter();
  float4* s_w = reinterpret_cast<float4*>(dw_smem);  // [KS*KS][cgb] dequantised weights
  constexpr int PAD = (KS - 1) / 2;
  // window of output columns touched by a 4-pixel input strip starting at iw0 (iw0 % 4 == 0):
  //   ow = (iw0 + t + PAD - s) / S  for the (t, s) with (t + PAD - s) % S == 0
  constexpr int JLO = -(PAD / S);                 // min of (t+PAD-s)/S over the valid pairs
  constexpr int JHI = (DW_TW - 1 + PAD) / S;      // max
  constexpr int NJ = JHI - JLO + 1;
  const int cg_local = threadIdx.x % cgb;
  const int cg = blockIdx.y * cgb + cg_local;
  const int spb = blockDim.x / cgb;
  const int strip_local = threadIdx.x / cgb;
  {
    const float zp_w = (float)*w_zp_p, s_wt = *w_scale_p;
    for (int i = threadIdx.x; i < KS * KS * cgb; i += blockDim.x) {
      const int t = i / cgb, g = i % cgb;
      const unsigned pk = ld_cg(reinterpret_cast<const unsigned*>(wq + (int64_t)t * C + (blockIdx.y * cgb + g) * 4));
      s_w[t * cgb + g] = make_float4(((float)sext_byte(pk, 0) - zp_w) * s_wt, ((float)sext_byte(pk, 1) - zp_w) * s_wt,
                                     ((float)sext_byte(pk, 2) - zp_w) * s_wt, ((float)sext_byte(pk, 3) - zp_w) * s_wt);
    }
  }
  __syncthreads();
  const int strips_w = (W + DW_TW - 1) / DW_TW;
  StripWalker wk;
  wk.init((int64_t)blockIdx.x * spb + strip_local, (int64_t)gridDim.x * spb, strips_w, H);
  for (; wk.n < N; wk.next()) {
    const int ih = wk.oh, n = wk.n;
    const int iw0 = wk.sw * DW_TW;
    const int ow_base = iw0 / S;          // iw0 % 4 == 0 -> exact
    float acc[DW_TW][4];
#pragma unroll
    for (int t = 0; t < DW_TW; ++t)
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) acc[t][ch] = 0.0f;
    const int r0 = (ih + PAD) % S;
    for (int r = r0; r < KS; r += S) {
      const int th = ih + PAD - r;
      if (th < 0) break;
      const int oh = th / S;
      if (oh >= Ho) continue;
      const float* row = dz + (((int64_t)n * Ho + oh) * Wo + (ow_base + JLO)) * C + cg * 4;   // only dereferenced in range
      float4 d[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const int ow = ow_base + JLO + j;
        d[j] = ((unsigned)ow < (unsigned)Wo) ? ld_cg(reinterpret_cast<const float4*>(row + j * C))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int s = 0; s < KS; ++s) {
        const float4 w = s_w[(r * KS + s) * cgb + cg_local];
#pragma unroll
        for (int t = 0; t < DW_TW; ++t) {
          if ((t + PAD - s) % S == 0) {                       // compile-time
            const int jj = (t + PAD - s) / S - JLO;           // compile-time; exact because S divides (t+PAD-s)
            acc[t][0] = fmaf(d[jj].x, w.x, acc[t][0]);
            acc[t][1] = fmaf(d[jj].y, w.y, acc[t][1]);
            acc[t][2] = fmaf(d[jj].z, w.z, acc[t][2]);
            acc[t][3] = fmaf(d[jj].w, w.w, acc[t][3]);
          }
        }
      }
    }
    float* obase = dx + (((int64_t)n * H + ih) * W + iw0) * C + cg * 4;
#pragma unroll
    for (int t = 0; t < DW_TW; ++t) {
      const int iw = iw0 + t;
      if (iw < W) {
        float4* o = reinterpret_cast<float4*>(obase + t * C);
        float4 v = make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
        if (accumulate) {
          const float4 old = ld_cg(o);
          v.x += old.x; v.y += old.y; v.z += old.z; v.w += old.w;
        }
        *o = v;
      }
    }
  }
}

// ================================================================= wgrad
// dwq[r*KS+s][c] += s_a * sum_{n,oh,ow} dz[n][oh][ow][c] * (x[n][oh*S-PAD+r][ow*S-PAD+s][c] - zp_a)
// One thread: one channel group and ONE kernel row r (KS taps x 4 channels = 20 accumulators, so three
// CTAs fit per SM), strips of 4 output pixels.  The KS threads that share a strip read the same dz words
// (L1 hits); each reads its own input row.
template <int KS, int S>
__global__ void __launch_bounds__(256, 3) dw_wgrad_kernel(const float* dz, const uint8_t* xq,
                                                         const float* x_scale_p, const int32_t* x_zp_p,
                                                         int N, int H, int W, int C, int ldx, int Ho, int Wo, int cgb,
                                                         float* dwq) {
  extern __shared__ __align__(16) unsigned char dw_smem[];
  pdl_enter();
  float* s_acc = reinterpret_cast<float*>(dw_smem);  // [KS*KS][cgb*4]
  constexpr int PAD = (KS - 1) / 2;
  constexpr int IW = (DW_TW - 1) * S + KS;
  const int zp_a = *x_zp_p;
  const unsigned zp4 = (unsigned)zp_a * 0x01010101u;
  const float zpf_magic = 8388608.0f + (float)zp_a;
  const int cg_local = threadIdx.x % cgb;
  const int r = (threadIdx.x / cgb) % KS;
  const int strip_local = threadIdx.x / (cgb * KS);
  const int spb = blockDim.x / (cgb * KS);
  const int cg = blockIdx.y * cgb + cg_local;
  for (int i = threadIdx.x; i < KS * KS * cgb * 4; i += blockDim.x) s_acc[i] = 0.0f;
  __syncthreads();
  float acc[KS][4];
#pragma unroll
  for (int t = 0; t < KS; ++t)
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) acc[t][ch] = 0.0f;
  const int strips_w = (Wo + DW_TW - 1) / DW_TW;
  StripWalker wk;
  wk.init((int64_t)blockIdx.x * spb + strip_local, (int64_t)gridDim.x * spb, strips_w, Ho);
  for (; wk.n < N; wk.next()) {
    const int oh = wk.oh, n = wk.n;
    const int ow0 = wk.sw * DW_TW;
    const int ih = oh * S - PAD + r;
    if ((unsigned)ih >= (unsigned)H) continue;       // this kernel row falls outside the image for this strip
    const int iw0 = ow0 * S - PAD;
    const float* drow = dz + (((int64_t)n * Ho + oh) * Wo + ow0) * C + cg * 4;
    const uint8_t* row = xq + (((int64_t)n * H + ih) * W + iw0) * ldx + cg * 4;      // only dereferenced inside the image
    float4 d[DW_TW];
    unsigned xw[IW];
#pragma unroll
    for (int t = 0; t < DW_TW; ++t)
      d[t] = (ow0 + t < Wo) ? ld_cg(reinterpret_cast<const float4*>(drow + t * C)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int j = 0; j < IW; ++j) {
      xw[j] = ((unsigned)(iw0 + j) < (unsigned)W) ? ld_cg(reinterpret_cast<const unsigned*>(row + j * ldx)) : zp4;
    }
#pragma unroll
    for (int j = 0; j < IW; ++j) {
      float xa[4];
      // byte -> float without I2F: 0x4B0000bb is the float 2^23 + bb; subtracting 2^23 + zp is exact
#pragma unroll
      for (int ch = 0; ch < 4; ++ch) xa[ch] = __uint_as_float(__byte_perm(xw[j], 0x4B000000u, 0x7650u + ch)) - zpf_magic;
#pragma unroll
      for (int t = 0; t < DW_TW; ++t) {
        const int s = j - t * S;
        if (s >= 0 && s < KS) {
          acc[s][0] = fmaf(d[t].x, xa[0], acc[s][0]);
          acc[s][1] = fmaf(d[t].y, xa[1], acc[s][1]);
          acc[s][2] = fmaf(d[t].z, xa[2], acc[s][2]);
          acc[s][3] = fmaf(d[t].w, xa[3], acc[s][3]);
        }
      }
    }
  }
#pragma unroll
  for (int t = 0; t < KS; ++t)
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) atomicAdd(&s_acc[((r * KS + t) * cgb + cg_local) * 4 + ch], acc[t][ch]);
  __syncthreads();
  const float s_a = *x_scale_p;
  // one 16-byte vector atomic per (tap, channel group): 4x fewer same-sector L2 atomics than scalar adds
  for (int i = threadIdx.x; i < KS * KS * cgb; i += blockDim.x) {
    const int t = i / cgb, g = i % cgb;
    const float4 v = *reinterpret_cast<const float4*>(s_acc + (size_t)i * 4);
    atomicAdd(reinterpret_cast<float4*>(dwq + (int64_t)t * C + (blockIdx.y * cgb + g) * 4),
              make_float4(v.x * s_a, v.y * s_a, v.z * s_a, v.w * s_a));
  }
}

}  // namespace frost

using namespace frost;

static bool dw_shape_ok(int C, int k, int stride) { return C > 0 && C % 4 == 0 && (k == 3 || k == 5) && (stride == 1 || stride == 2); }

extern "C" int frost_dw_conv_forward(const uint8_t* xq, int ldx, const int32_t* x_zp, const int8_t* wq, const int32_t* w_zp,
                                     int N, int H, int W, int C, int k, int stride, int32_t* acc,
                                     FrostChanStats* stats, void* stream) {
  FROST_REQUIRE(xq && x_zp && wq && w_zp && acc && stats, "frost_dw_conv_forward: null pointer");
  FROST_REQUIRE(ldx >= C && ldx % 4 == 0, "frost_dw_conv_forward: ldx=%d must be >= C and a multiple of 4", ldx);
  FROST_REQUIRE(N > 0 && H > 0 && W > 0 && dw_shape_ok(C, k, stride),
                "frost_dw_conv_forward: bad shape (C%%4==0, k in {3,5}, stride in {1,2})");
  cudaStream_t st = (cudaStream_t)stream;
  const int pad = (k - 1) / 2;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  int cgb, chunks, threads;
  dw_launch_shape(C, DW_MAX_CGB, &cgb, &chunks, &threads);
  const int spb = threads / cgb;
  const int strips_w = (Wo + DW_TW - 1) / DW_TW;
  const int64_t total_strips = (int64_t)N * Ho * strips_w;
  const size_t smem_g = sizeof(unsigned) * k * k * DW_MAX_CGB + sizeof(SmemStat) * cgb * 4;
  // ---- shared-memory tiled kernel (symmetric weights, the normal case)
  const int IW = (DW_TW - 1) * stride + k, NB = (IW + 3) / 4;
  const int fill_w = (strips_w - 1) * DW_TW * stride + NB * 4;
  int pitch = (fill_w + 31) / 32 * 32 + 4;
  if (pitch - 32 >= fill_w) pitch -= 32;
  const int tcgb = (C >= 256 && Ho * strips_w <= 64) ? 16 : 8;      // small planes: 16 groups x 16 strip slots
  const int max_ih = (int)(44 * 1024 / ((size_t)tcgb * pitch * 4));
  // measured on B200 (tools/microbench_ops.py dw): the tiles win only on the widest stride-1 planes (112x112: 161 vs
  // 189 us); elsewhere the gather kernel is ALU-bound, not latency-bound, and the fill/compute barrier costs more
  // than it hides.  FROST_TUNE_DW_FWD_TILED: 1 = this rule, 2 = never, 3 = wherever a tile fits.
  const int tmode = tunable(FROST_TUNE_DW_FWD_TILED);
  bool tiled = max_ih >= k && (tmode == 3 || (tmode == 1 && stride == 1 && Wo >= 100));
  if (tiled) {
    int th = std::min(Ho, (max_ih - k) / stride + 1);
    const int tiles_h = (Ho + th - 1) / th;
    th = (Ho + tiles_h - 1) / tiles_h;
    const int ih = (th - 1) * stride + k;
    const size_t smem_t = sizeof(unsigned) * (2 * k * 4 * tcgb + 4 * tcgb) + sizeof(SmemStat) * 4 * tcgb + sizeof(unsigned) * (size_t)ih * tcgb * pitch;
    const int tchunks = (C / 4 + tcgb - 1) / tcgb;
    const int64_t wave = std::max<int64_t>(1, (int64_t)kNumSMs * tunable(FROST_TUNE_DW_FWD_CTAS_PER_SM) / tchunks);
    const dim3 tgrid((unsigned)std::min<int64_t>((int64_t)N * tiles_h, wave), tchunks);
#define LT(KS, S, G)                                                                                                     \
  launch_pdl(dw_conv_fwd_tiled_kernel<KS, S, G>, tgrid, dim3(256), smem_t, st, xq, x_zp, wq, w_zp, N, H, W, C, ldx, Ho, Wo, th, \
             pitch, acc, stats)
#define LTG(KS, S)                 \
  do {                             \
    if (tcgb == 16) LT(KS, S, 16); \
    else LT(KS, S, 8);             \
  } while (0)
    if (k == 3 && stride == 1) LTG(3, 1);
    else if (k == 3 && stride == 2) LTG(3, 2);
    else if (k == 5 && stride == 1) LTG(5, 1);
    else LTG(5, 2);
#undef LTG
#undef LT
  }
  // one resident wave: every CTA ends with one set of atomics per channel on the same 32-byte records, and
  // same-sector atomics serialise in L2 - 16 waves of CTAs cost more in that tail than the convolution itself
  const int64_t wave = std::max<int64_t>(1, (int64_t)kNumSMs * tunable(FROST_TUNE_DW_FWD_CTAS_PER_SM) / chunks);
  const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total_strips, (int64_t)spb * 2), wave));
  dim3 grid(gx, chunks);
  const size_t smem = sizeof(unsigned) * (2 * k * 4 * DW_MAX_CGB + DW_MAX_CGB * 4) + sizeof(SmemStat) * cgb * 4;
#define L(KS, S)                                                                                                        \
  do {                                                                                                                  \
    if (!tiled) launch_pdl(dw_conv_fwd_kernel<KS, S>, grid, dim3(threads), smem, st, xq, x_zp, wq, w_zp, N, H, W, C, ldx, Ho, Wo, cgb, acc, stats);     \
    launch_pdl(dw_conv_fwd_generic_kernel<KS, S>, grid, dim3(threads), smem_g, st, xq, x_zp, wq, w_zp, N, H, W, C, ldx, Ho, Wo, cgb, acc, stats); \
  } while (0)
  if (k == 3 && stride == 1) L(3, 1);
  else if (k == 3 && stride == 2) L(3, 2);
  else if (k == 5 && stride == 1) L(5, 1);
  else L(5, 2);
#undef L
  FROST_LAUNCH_CHECK("dw_conv_fwd");
  return FROST_OK;
}

namespace frost {
int dw_dgrad_tma_launch(const float* dz, const int8_t* wq, const float* w_scale, const int32_t* w_zp, int N, int H, int W, int C,
                        int k, float* dx, int accumulate, cudaStream_t st);      // dw_tma.cu
}

extern "C" int frost_dw_dgrad(const float* dz, const int8_t* wq, const float* w_scale, const int32_t* w_zp, int N, int H,
                              int W, int C, int k, int stride, float* dx, int accumulate, void* stream) {
  FROST_REQUIRE(dz && wq && w_scale && w_zp && dx, "frost_dw_dgrad: null pointer");
  FROST_REQUIRE(N > 0 && H > 0 && W > 0 && dw_shape_ok(C, k, stride), "frost_dw_dgrad: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  // stride 1: shared-memory tiles filled by TMA (dw_tma.cu); FROST_TUNE_DW_DGRAD_TILED = 2 forces the gather kernel
  if (stride == 1 && tunable(FROST_TUNE_DW_DGRAD_TILED) == 1 && (reinterpret_cast<uintptr_t>(dz) & 15) == 0) {
    const int rc = dw_dgrad_tma_launch(dz, wq, w_scale, w_zp, N, H, W, C, k, dx, accumulate, st);
    if (rc == FROST_OK) {
      FROST_LAUNCH_CHECK("dw_dgrad_tma");
      return FROST_OK;
    }
    if (rc != FROST_ENOSUP) return rc;
  }
  const int pad = (k - 1) / 2;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  int cgb, chunks, threads;
  dw_launch_shape(C, DW_MAX_CGB, &cgb, &chunks, &threads);
  const int spb = threads / cgb;
  const int64_t total_strips = (int64_t)N * H * ((W + DW_TW - 1) / DW_TW);
  const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total_strips, (int64_t)spb * 2),
                                                              (int64_t)kNumSMs * tunable(FROST_TUNE_DW_DGRAD_CTAS_PER_SM) / chunks + 1));
  dim3 grid(gx, chunks);
  const size_t smem = sizeof(float4) * k * k * cgb;
#define L(KS, S) launch_pdl(dw_dgrad_kernel<KS, S>, grid, dim3(threads), smem, st, dz, wq, w_scale, w_zp, N, H, W, C, Ho, Wo, cgb, dx, accumulate)
  if (k == 3 && stride == 1) L(3, 1);
  else if (k == 3 && stride == 2) L(3, 2);
  else if (k == 5 && stride == 1) L(5, 1);
  else L(5, 2);
#undef L
  FROST_LAUNCH_CHECK("dw_dgrad");
  return FROST_OK;
}

extern "C" int frost_dw_wgrad(const float* dz, const uint8_t* xq, int ldx, const float* x_scale, const int32_t* x_zp, int N, int H,
                              int W, int C, int k, int stride, float* dwq, void* stream) {
  FROST_REQUIRE(dz && xq && x_scale && x_zp && dwq, "frost_dw_wgrad: null pointer");
  FROST_REQUIRE(ldx >= C && ldx % 4 == 0, "frost_dw_wgrad: ldx=%d must be >= C and a multiple of 4", ldx);
  FROST_REQUIRE(N > 0 && H > 0 && W > 0 && dw_shape_ok(C, k, stride), "frost_dw_wgrad: bad shape");
  FROST_REQUIRE((reinterpret_cast<uintptr_t>(dwq) & 15) == 0, "frost_dw_wgrad: dwq must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int pad = (k - 1) / 2;
  const int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
  if (cudaMemsetAsync(dwq, 0, sizeof(float) * (size_t)k * k * C, st) != cudaSuccess) {
    set_error("frost_dw_wgrad: memset failed");
    return FROST_ECUDA;
  }
  int cgb, chunks, threads;
  dw_launch_shape(C, 256 / (k * 2) > 48 ? 48 : 256 / k, &cgb, &chunks, &threads);   // cgb * k <= 256
  const int spb = 256 / (cgb * k);
  threads = cgb * k * spb;
  const int64_t total_strips = (int64_t)N * Ho * ((Wo + DW_TW - 1) / DW_TW);
  // every thread should see >= 8 strips so that the final smem/global reduction is amortised
  const int64_t wave = std::max<int64_t>(1, (int64_t)kNumSMs * tunable(FROST_TUNE_DW_WGRAD_CTAS_PER_SM) / chunks);
  const int gx = (int)std::max<int64_t>(1, std::min<int64_t>(ceil_div(total_strips, (int64_t)spb * 8), wave));
  dim3 grid(gx, chunks);
  const size_t smem = sizeof(float) * k * k * cgb * 4;
#define L(KS, S) launch_pdl(dw_wgrad_kernel<KS, S>, grid, dim3(threads), smem, st, dz, xq, x_scale, x_zp, N, H, W, C, ldx, Ho, Wo, cgb, dwq)
  if (k == 3 && stride == 1) L(3, 1);
  else if (k == 3 && stride == 2) L(3, 2);
  else if (k == 5 && stride == 1) L(5, 1);
  else L(5, 2);
#undef L
  FROST_LAUNCH_CHECK("dw_wgrad");
  return FROST_OK;
}
