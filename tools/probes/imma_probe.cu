// Throughput probe: mma.sync.m16n8k32 (u8 x s8 -> s32) issue rate per SM on sm_100a, with and without shared-memory
// operand loads.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_probe imma_probe.cu && ./imma_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void imma(int (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int ACCS>
__global__ void __launch_bounds__(256) probe(int iters, int* out) {
  int d[ACCS][4] = {};
  unsigned a[4] = {threadIdx.x, threadIdx.x * 3u, 7u, 9u}, b[2] = {threadIdx.x ^ 5u, 11u};
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < ACCS; ++k) imma(d[k], a, b);
  }
  int s = 0;
  for (int k = 0; k < ACCS; ++k) s += d[k][0] + d[k][1] + d[k][2] + d[k][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) probe_lds(int iters, int* out) {
  __shared__ unsigned tile[4096];
  for (int i = threadIdx.x; i < 4096; i += 256) tile[i] = i * 2654435761u;
  __syncthreads();
  int d[4][4] = {};
  unsigned b[2] = {threadIdx.x ^ 5u, 11u};
  const unsigned* base = tile + (threadIdx.x & 31) * 2;
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      unsigned a[4];
      a[0] = base[k * 64 + (i & 15) * 128];
      a[1] = base[k * 64 + 16 + (i & 15) * 128];
      a[2] = base[k * 64 + 32 + (i & 15) * 128];
      a[3] = base[k * 64 + 48 + (i & 15) * 128];
      imma(d[k], a, b);
    }
  }
  int s = 0;
  for (int k = 0; k < 4; ++k) s += d[k][0] + d[k][1] + d[k][2] + d[k][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int* out;
  cudaMalloc(&out, 148 * 8 * 256 * sizeof(int));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 20000;
  for (int rep = 0; rep < 2; ++rep) {
    for (int mode = 0; mode < 3; ++mode) {
      const int ctas = 148 * 4;
      cudaEventRecord(e0);
      if (mode == 0) probe<4><<<ctas, 256>>>(iters, out);
      if (mode == 1) probe<8><<<ctas, 256>>>(iters / 2, out);
      if (mode == 2) probe_lds<<<ctas, 256>>>(iters, out);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      const double mmas = (double)ctas * 8 * iters * 4;
      printf("mode %d: %.3f ms, %.1f G mma/s, %.2f mma/clk/SM (at 1.965 GHz), %.1f dense TOPS\n", mode, ms, mmas / ms / 1e6,
             mmas / (ms * 1e-3) / 148 / 1.965e9, mmas * 16 * 8 * 32 * 2 / (ms * 1e-3) / 1e12);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
