// FP64 pipe microbenchmark for sm_100a: DFMA (CUDA cores) vs DMMA m8n8k4 (tensor) issue rates, and the
// two together.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE>  // 0: DFMA only, 1: DMMA only, 2: both interleaved
__global__ void __launch_bounds__(128) k(double *out, int iters, double a, double b) {
  double f[8], c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { f[i] = threadIdx.x + i; c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE != 1) { f[i] = fma(f[i], a, b); f[i] = fma(f[i], a, b); f[i] = fma(f[i], a, b); f[i] = fma(f[i], a, b); }
      if (MODE != 0) dmma(c[i][0], c[i][1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += f[i] + c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int blocks_per_sm) {
  int sm; cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  double *out; cudaMalloc(&out, sizeof(double) * sm * blocks_per_sm * 128);
  const int iters = 20000;
  k<MODE><<<sm * blocks_per_sm, 128>>>(out, 100, 1.0000001, 1e-9);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<sm * blocks_per_sm, 128>>>(out, iters, 1.0000001, 1e-9);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warps = (double)sm * blocks_per_sm * 4;
  const double dfma = MODE != 1 ? warps * iters * 32.0 : 0, dm = MODE != 0 ? warps * iters * 8.0 : 0;
  printf("%-10s %2d warps/SM: %.3f ms  DFMA %.2f TFLOP/s  DMMA %.2f TFLOP/s (m8n8k4 = 512 flop)\n", name, blocks_per_sm * 4, ms,
         dfma * 64.0 / ms / 1e9, dm * 512.0 / ms / 1e9);
  cudaFree(out);
}

int main() {
  for (int b : {1, 2, 4, 8}) { run<0>("dfma", b); run<1>("dmma", b); run<2>("both", b); }
  return 0;
}
