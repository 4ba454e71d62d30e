// dw_tma.cu - depthwise dgrad (stride 1) from shared-memory tiles that TMA fills.
//
// The gather kernel of dw_conv.cu reads every dz element k*k / 2 times (k kernel rows x the column overlap of adjacent
// 4-pixel strips) through L2 - L1 cannot be used under programmatic dependent launch (common.cuh) - and that traffic, not
// the FMAs or DRAM, bounds it: 10 x 125 MB over the L2 fabric for the 14x14x624 5x5 layer = the measured 140 us.
// Here a CTA owns 32 channels and walks tiles of (image, TH input rows, full width); one thread issues ONE 4-D TMA box per
// tile - {32 channels, W + 2*PAD columns, TH + 2*PAD rows, 1 image} of dz, out-of-image rows / columns / channels
// zero-filled by the hardware, which IS the zero padding of the transposed convolution - into one of two buffers, and the
// FMAs of tile t run against shared memory while tile t+1 is in flight.  Every dz element crosses L2 once per tile (+ halo).
//   dx[n][ih][iw][c] (+)= sum_{r,s} dz[n][ih+PAD-r][iw+PAD-s][c] * w[r][s][c]          (frost_dw_dgrad, stride 1)
#include <cuda.h>
#include <algorithm>
#include "pw_tma.cuh"

namespace frost {

constexpr int DT_THREADS = 512;
constexpr int DT_CB = 32;                 // channels per CTA (8 groups of 4 = one 128-byte line per pixel)
constexpr int DT_TW = 4;                  // pixels per strip

__device__ __forceinline__ void tma_load_4d(const CUtensorMap* tm, uint64_t* bar, uint32_t dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

__device__ __forceinline__ int dt_sext_byte(unsigned w, int ch) { return (int)(signed char)((w >> (8 * ch)) & 0xffu); }

template <int KS>
__global__ void __launch_bounds__(DT_THREADS, 1) dw_dgrad_tma_kernel(const __grid_constant__ CUtensorMap tm_dz, const int8_t* wq,
                                                                    const float* w_scale_p, const int32_t* w_zp_p, int N, int H,
                                                                    int W, int C, int TH, int NB, int tile_bytes, float* dx, int accumulate) {
  extern __shared__ uint8_t dt_smem_raw[];
  uint8_t* smem = dt_smem_raw + ((128u - (smem_u32(dt_smem_raw) & 127u)) & 127u);
  constexpr int PAD = (KS - 1) / 2;
  const int BW = W + 2 * PAD, BH = TH + 2 * PAD;
  uint8_t* tiles = smem;                                               // [2][NB images][BH][BW][32 ch] fp32
  float4* s_w = reinterpret_cast<float4*>(smem + 2 * tile_bytes);      // [8 groups][KS*KS]
  uint64_t* full = reinterpret_cast<uint64_t*>(s_w + 8 * KS * KS);     // [2]
  const int c0 = blockIdx.y * DT_CB;
  const int g = threadIdx.x & 7, slot = threadIdx.x >> 3;              // channel group, strip slot (64 per CTA)
  const bool g_ok = c0 + g * 4 < C;
  if (threadIdx.x == 0) {
    mbar_init(&full[0], 1);
    mbar_init(&full[1], 1);
    mbar_fence_init();
    tma_prefetch_desc(&tm_dz);
  }
  pdl_wait();
  {
    const float zp_w = (float)*w_zp_p, s_wt = *w_scale_p;
    for (int i = threadIdx.x; i < 8 * KS * KS; i += DT_THREADS) {
      const int gg = i / (KS * KS), t = i - gg * (KS * KS);
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c0 + gg * 4 < C) {
        const unsigned pk = ld_cg(reinterpret_cast<const unsigned*>(wq + (int64_t)t * C + c0 + gg * 4));
        w = make_float4(((float)dt_sext_byte(pk, 0) - zp_w) * s_wt, ((float)dt_sext_byte(pk, 1) - zp_w) * s_wt,
                        ((float)dt_sext_byte(pk, 2) - zp_w) * s_wt, ((float)dt_sext_byte(pk, 3) - zp_w) * s_wt);
      }
      s_w[i] = w;
    }
  }
  __syncthreads();
  pdl_trigger();
  const float4* my_w = s_w + g * (KS * KS);
  const int bands = (H + TH - 1) / TH;
  const int n_groups = (N + NB - 1) / NB;                              // NB images per tile (small planes: amortise the TMA latency)
  const int n_tiles = n_groups * bands;
  auto issue = [&](int tile, int buf) {                                // thread 0 only
    const int ng = tile / bands, ih0 = (tile - ng * bands) * TH;
    mbar_expect_tx(&full[buf], (uint32_t)tile_bytes);
    tma_load_4d(&tm_dz, &full[buf], smem_u32(tiles + buf * tile_bytes), c0, -PAD, ih0 - PAD, ng * NB);
  };
  if (threadIdx.x == 0 && (int)blockIdx.x < n_tiles) issue(blockIdx.x, 0);
  const int strips_w = (W + DT_TW - 1) / DT_TW;
  int it = 0;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
    const int buf = it & 1;
    // the buffer the next tile goes into was read during iteration it - 1; the barrier at the end of that iteration freed it
    if (threadIdx.x == 0 && tile + (int)gridDim.x < n_tiles) issue(tile + gridDim.x, buf ^ 1);
    mbar_wait(&full[buf], (it >> 1) & 1);
    const int ng = tile / bands, ih0 = (tile - ng * bands) * TH;
    const int rows = min(TH, H - ih0);
    const int imgs = min(NB, N - ng * NB);
    const uint32_t row_b = (uint32_t)BW * 128u, img_b = (uint32_t)BH * row_b;
    const int per_img = rows * strips_w;
    for (int sidx = slot; sidx < imgs * per_img; sidx += DT_THREADS / 8) {
      const int im = sidx / per_img, rem = sidx - im * per_img;
      const int n = ng * NB + im;
      const uint32_t tbase = smem_u32(tiles + buf * tile_bytes) + (uint32_t)im * img_b + (uint32_t)g * 16u;
      const int tr = rem / strips_w, iw0 = (rem - tr * strips_w) * DT_TW;
      float acc[DT_TW][4];
#pragma unroll
      for (int t = 0; t < DT_TW; ++t)
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) acc[t][ch] = 0.0f;
      // dz row ih + PAD - r sits at tile row tr + 2*PAD - r; column iw + PAD - s at tile column iw + 2*PAD - s
      uint32_t rp = tbase + (uint32_t)(tr + 2 * PAD) * row_b + (uint32_t)iw0 * 128u;
#pragma unroll
      for (int r = 0; r < KS; ++r, rp -= row_b) {
        float4 d[DT_TW + KS - 1];                                      // tile columns iw0 .. iw0 + 3 + 2*PAD
#pragma unroll
        for (int j = 0; j < DT_TW + KS - 1; ++j)
          asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(d[j].x), "=f"(d[j].y), "=f"(d[j].z), "=f"(d[j].w) : "r"(rp + j * 128u));
#pragma unroll
        for (int s = 0; s < KS; ++s) {
          const float4 w = my_w[r * KS + s];
#pragma unroll
          for (int t = 0; t < DT_TW; ++t) {
            const float4 v = d[t + 2 * PAD - s];
            acc[t][0] = fmaf(v.x, w.x, acc[t][0]);
            acc[t][1] = fmaf(v.y, w.y, acc[t][1]);
            acc[t][2] = fmaf(v.z, w.z, acc[t][2]);
            acc[t][3] = fmaf(v.w, w.w, acc[t][3]);
          }
        }
      }
      if (g_ok) {
        float* obase = dx + (((int64_t)n * H + ih0 + tr) * W + iw0) * C + c0 + g * 4;
#pragma unroll
        for (int t = 0; t < DT_TW; ++t) {
          if (iw0 + t < W) {
            float4* o = reinterpret_cast<float4*>(obase + (int64_t)t * C);
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
    __syncthreads();                                                   // everybody is done with `buf`: it may be refilled
  }
}

// Returns FROST_ENOSUP when the shape does not fit (the caller then takes the gather kernel).
int dw_dgrad_tma_launch(const float* dz, const int8_t* wq, const float* w_scale, const int32_t* w_zp, int N, int H, int W, int C,
                        int k, float* dx, int accumulate, cudaStream_t st) {
  // Measured on B200 (tools/microbench_ops.py dw): the tiles win where a pixel's channels ARE the 128-byte line the box
  // fetches (C == 32: 112x112, 180 -> 143 us = 5.7 TB/s); with wider tensors every box row is a separate 128-byte fragment
  // of a 288 ... 6912-byte pixel and TMA moves those slower than the gather kernel's coalesced loads (56x56x72: 197 vs 121 us,
  // 14x14x624: 172 vs 142 us, 7x7x1440: 98 vs 87 us) - those shapes stay on the gather kernel.
  if (!encode_fn() || (k != 3 && k != 5) || C != DT_CB) return FROST_ENOSUP;
  const int pad = (k - 1) / 2, BW = W + 2 * pad;
  if (BW > 256) return FROST_ENOSUP;
  // rows per tile: two buffers of (TH + 2 pad) x BW x 128 B within ~200 KB
  const size_t budget = 192 * 1024;
  int TH = std::min(H, 32);
  while (TH > 1 && 2 * (size_t)(TH + 2 * pad) * BW * 128 > budget) --TH;
  if (2 * (size_t)(TH + 2 * pad) * BW * 128 > budget) return FROST_ENOSUP;
  // whole planes that are small: several images per tile, up to ~96 KB per buffer
  int NB = 1;
  if (TH == H) NB = (int)std::max<size_t>(1, std::min<size_t>(16, (budget / 2) / ((size_t)(TH + 2 * pad) * BW * 128)));
  const int tile_bytes = NB * (TH + 2 * pad) * BW * 128;
  EncodeTiledFn fn = encode_fn();
  CUtensorMap tm;
  const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
  const cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  const cuuint32_t box[4] = {DT_CB, (cuuint32_t)BW, (cuuint32_t)(TH + 2 * pad), (cuuint32_t)NB};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  if (fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(dz), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return FROST_ENOSUP;
  const size_t smem = 128 + 2 * (size_t)tile_bytes + sizeof(float4) * 8 * k * k + 64;
  const int chunks = (C + DT_CB - 1) / DT_CB;
  const int bands = (H + TH - 1) / TH;
  const int n_tiles = ((N + NB - 1) / NB) * bands;
  const int gx = std::max(1, std::min(n_tiles, kNumSMs / chunks + (kNumSMs % chunks ? 1 : 0)));
  cudaError_t e;
  if (k == 3) {
    if (first_use_on_device(reinterpret_cast<const void*>(&dw_dgrad_tma_kernel<3>)))
      cudaFuncSetAttribute(dw_dgrad_tma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    e = launch_pdl(dw_dgrad_tma_kernel<3>, dim3(gx, chunks), dim3(DT_THREADS), smem, st, tm, wq, w_scale, w_zp, N, H, W, C, TH, NB, tile_bytes, dx,
                   accumulate);
  } else {
    if (first_use_on_device(reinterpret_cast<const void*>(&dw_dgrad_tma_kernel<5>)))
      cudaFuncSetAttribute(dw_dgrad_tma_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    e = launch_pdl(dw_dgrad_tma_kernel<5>, dim3(gx, chunks), dim3(DT_THREADS), smem, st, tm, wq, w_scale, w_zp, N, H, W, C, TH, NB, tile_bytes, dx,
                   accumulate);
  }
  if (e != cudaSuccess) {
    set_error("frost_dw_dgrad (tiled): launch failed: %s", cudaGetErrorString(e));
    return FROST_ECUDA;
  }
  return FROST_OK;
}

}  // namespace frost
