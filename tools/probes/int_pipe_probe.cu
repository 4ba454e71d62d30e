// Issue-rate probe for the integer instructions of the fused forward's statistics loop on sm_100a: IMAD, IMAD.WIDE with a
// 64-bit accumulator, VIMNMX3, IADD3, and the loop's own mix.  16 warps per SM (4 per sub-partition), 8 independent chains.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipe_probe int_pipe_probe.cu && ./int_pipe_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(512, 1) probe(int iters, int a, int b, long long* out, long long* clocks) {
  int x[8];
  long long w[8];
  for (int k = 0; k < 8; ++k) { x[k] = threadIdx.x * (k + 3) + a; w[k] = k; }
  int mn = 1 << 30, mx = -(1 << 30), s = 0;
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (OP == 0) x[k] = x[k] * a - b;                                   // IMAD
      if (OP == 1) w[k] += (long long)x[k] * (long long)(x[k] ^ i);       // IMAD.WIDE (64-bit accumulate) + LOP
      if (OP == 2) { mn = min(mn, min(x[k] ^ i, x[(k + 1) & 7] + i)); }   // VIMNMX3 + 2 ALU
      if (OP == 3) s += (x[k] ^ i) + (x[(k + 1) & 7] | i);                // IADD3 + 2 ALU
      if (OP == 4) {                                                      // the statistics loop's mix per element
        const int I = x[k] * a - b;
        x[k] = I ^ i;
        s += I;
        w[k & 1] += (long long)I * (long long)I;
        mn = min(mn, I);
        mx = max(mx, I);
      }
      if (OP == 5) {                                                      // the same with the squares in two 32-bit halves
        const int I = x[k] * a - b;
        x[k] = I ^ i;
        s += I;
        const unsigned lo = (unsigned)I * (unsigned)I;
        const unsigned hi = __umulhi((unsigned)abs(I), (unsigned)abs(I));
        w[k & 1] += lo;
        w[2 + (k & 1)] += hi;
        mn = min(mn, I);
        mx = max(mx, I);
      }
    }
  }
  const long long t1 = clock64();
  long long r = s + mn + mx;
  for (int k = 0; k < 8; ++k) r += x[k] + w[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
  if (threadIdx.x == 0) clocks[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(const char* name, int per_iter) {
  const int grid = 148, iters = 2048;
  long long *out, *clk;
  cudaMalloc(&out, grid * 512 * sizeof(long long));
  cudaMalloc(&clk, grid * sizeof(long long));
  probe<OP><<<grid, 512>>>(16, 3, 5, out, clk);
  probe<OP><<<grid, 512>>>(iters, 3, 5, out, clk);
  cudaError_t e = cudaDeviceSynchronize();
  long long c0 = 0;
  cudaMemcpy(&c0, clk, sizeof(c0), cudaMemcpyDeviceToHost);
  // 4 warps per sub-partition, per_iter "units" per thread and iteration
  printf("%-44s %6.2f clk per unit and sub-partition (4 warps)  %s\n", name, (double)c0 / ((double)iters * per_iter * 4),
         e == cudaSuccess ? "" : cudaGetErrorString(e));
  cudaFree(out);
  cudaFree(clk);
}

int main() {
  run<0>("IMAD (unit = 1 instruction)", 8);
  run<1>("IMAD.WIDE 64-bit acc + 1 LOP3 (unit = pair)", 8);
  run<2>("VIMNMX3 + 2 ALU (unit = triple)", 8);
  run<3>("IADD3 + 2 ALU (unit = triple)", 8);
  run<4>("statistics mix (unit = element)", 8);
  run<5>("statistics mix, split squares (unit = element)", 8);
  return 0;
}
