// Throughput probe: tcgen05.ld (TMEM -> registers) bytes per clock per SM on sm_100a, by load shape, number of loads in
// flight before tcgen05.wait::ld, and number of warps.  One CTA per SM, 512 TMEM columns, no MMA (the values do not matter).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld_probe tmem_ld_probe.cu && ./tmem_ld_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NCOL>
__device__ __forceinline__ void ld(uint32_t taddr, uint32_t (&v)[NCOL]);
template <>
__device__ __forceinline__ void ld<16>(uint32_t t, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(t) : "memory");
}
template <>
__device__ __forceinline__ void ld<32>(uint32_t t, uint32_t (&v)[32]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                 "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                 "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(t) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// NCOL columns per load, INFLIGHT loads issued before each wait
template <int NCOL, int INFLIGHT>
__global__ void __launch_bounds__(512, 1) probe(int iters, int nwarps, unsigned* out, long long* clocks) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  unsigned acc = 0;
  const long long t0 = clock64();
  if (warp < nwarps) {
    for (int i = 0; i < iters; ++i) {
      uint32_t v[INFLIGHT][NCOL];
#pragma unroll
      for (int k = 0; k < INFLIGHT; ++k) ld<NCOL>(base + (uint32_t)(((i * INFLIGHT + k) * NCOL + (warp >> 2) * 64) & 511 & ~(NCOL - 1)), v[k]);
      wait_ld();
#pragma unroll
      for (int k = 0; k < INFLIGHT; ++k)
#pragma unroll
        for (int j = 0; j < NCOL; ++j) acc ^= v[k][j];
    }
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(slot) : "memory");
}

template <int NCOL, int INFLIGHT>
static void run(const char* name, int nwarps) {
  const int grid = 148, iters = 4096;
  unsigned* out;
  long long* clk;
  cudaMalloc(&out, grid * 512 * sizeof(unsigned));
  cudaMalloc(&clk, grid * sizeof(long long));
  probe<NCOL, INFLIGHT><<<grid, 512>>>(64, nwarps, out, clk);
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  cudaEventRecord(a);
  probe<NCOL, INFLIGHT><<<grid, 512>>>(iters, nwarps, out, clk);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  long long c0 = 0;
  cudaMemcpy(&c0, clk, sizeof(c0), cudaMemcpyDeviceToHost);
  const double bytes = (double)iters * INFLIGHT * NCOL * 32 * 4 * nwarps;   // per SM
  printf("%-22s warps %2d  %8.1f us  %7.1f B/clk/SM (clock64 of warp 0: %lld)  %s\n", name, nwarps, ms * 1e3, bytes / (double)c0, c0,
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(clk);
}

int main() {
  for (int nw : {1, 4, 8, 16}) {
    run<16, 1>("32x32b.x16, 1 per wait", nw);
    run<16, 4>("32x32b.x16, 4 per wait", nw);
    run<32, 1>("32x32b.x32, 1 per wait", nw);
    run<32, 2>("32x32b.x32, 2 per wait", nw);
  }
  return 0;
}
