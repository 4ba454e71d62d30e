// se.cu - the element-wise pieces of the quantization-aware squeeze-and-excite block (SEModule,
// Classification/models/imagenet/mobilenetv3.py:85-102) that the other entry points do not cover:
//   ReLU between the two quantised Linear layers (nniqat.LinearReLU), the broadcast product x * gate.expand_as(x)
//   (FloatFunctional.mul - its observer / fake-quant is frost_fq_forward on the product) and their gradients.
// The pool is frost_pool_dropout_forward, the Linear layers frost_linear_forward / _backward on fake-quantised weights,
// the gate frost_hsigmoid_forward.  fp32 NCHW throughout: a row is one (image, channel) plane of HW values.
#include "common.cuh"

namespace frost {

// y = max(x, 0), mask = [x > 0]      (aten::relu and the mask threshold_backward uses)
__global__ void __launch_bounds__(256) relu_fwd_kernel(const float* x, int64_t n, float* y,
                                                       uint8_t* mask) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = x[i];
    y[i] = v > 0.0f ? v : 0.0f;
    if (mask) mask[i] = v > 0.0f ? 1 : 0;
  }
}

// y[r][i] = x[r][i] * g[r]      one warp per row, rows strided over the grid
__global__ void __launch_bounds__(256) bcast_mul_fwd_kernel(const float* x, const float* g, int64_t rows, int hw,
                                                            float* y) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += n_warps) {
    const float gr = g[r];
    const float* xr = x + r * hw;
    float* yr = y + r * hw;
    for (int i = lane; i < hw; i += 32) yr[i] = __fmul_rn(xr[i], gr);
  }
}

// dx[r][i] = dy[r][i] * g[r] ;  dg[r] = sum_i dy[r][i] * x[r][i]
__global__ void __launch_bounds__(256) bcast_mul_bwd_kernel(const float* dy, const float* x,
                                                            const float* g, int64_t rows, int hw, float* dx,
                                                            float* dg) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < rows; r += n_warps) {
    const float gr = g[r];
    const float* dyr = dy + r * hw;
    const float* xr = x + r * hw;
    float* dxr = dx + r * hw;
    float acc = 0.0f;
    for (int i = lane; i < hw; i += 32) {
      const float d = dyr[i];
      dxr[i] = __fmul_rn(d, gr);
      acc = fmaf(d, xr[i], acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) dg[r] = acc;
  }
}

}  // namespace frost

using namespace frost;

extern "C" int frost_relu_forward(const float* x, int64_t n, float* y, uint8_t* mask, void* stream) {
  FROST_REQUIRE(x && y && n > 0, "frost_relu_forward: bad args");
  relu_fwd_kernel<<<grid_for(n, 256 * 4, kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(x, n, y, mask);
  FROST_LAUNCH_CHECK("relu_fwd");
  return FROST_OK;
}

extern "C" int frost_bcast_mul_forward(const float* x, const float* gate, int64_t rows, int hw, float* y, void* stream) {
  FROST_REQUIRE(x && gate && y && rows > 0 && hw > 0, "frost_bcast_mul_forward: bad args");
  bcast_mul_fwd_kernel<<<grid_for(rows, 8, kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(x, gate, rows, hw, y);
  FROST_LAUNCH_CHECK("bcast_mul_fwd");
  return FROST_OK;
}

extern "C" int frost_bcast_mul_backward(const float* dy, const float* x, const float* gate, int64_t rows, int hw, float* dx,
                                        float* dgate, void* stream) {
  FROST_REQUIRE(dy && x && gate && dx && dgate && rows > 0 && hw > 0, "frost_bcast_mul_backward: bad args");
  bcast_mul_bwd_kernel<<<grid_for(rows, 8, kNumSMs * 8), 256, 0, (cudaStream_t)stream>>>(dy, x, gate, rows, hw, dx, dgate);
  FROST_LAUNCH_CHECK("bcast_mul_bwd");
  return FROST_OK;
}
