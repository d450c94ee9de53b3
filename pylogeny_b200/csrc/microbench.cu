// Pipe microbenchmarks (roofline denominators of the walk kernels, bench.py): issue rate of
// DFMA (CUDA cores) and of DMMA m8n8k4 (FP64 tensor instructions) on this device, measured in the
// run with CUDA events.  On B200 both land on ONE pipe (interleaving them gives the same total),
// so the walk kernels are bounded by max(DFMA, DMMA) flop/s, not by their sum.
#include "common.h"

namespace plg {

__device__ __forceinline__ void mb_dmma(double &d0, double &d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE>  // 0: DFMA only, 1: DMMA only, 2: both interleaved
__global__ void __launch_bounds__(128) fp64_peak_kernel(double *out, int iters, double a, double b) {
  double f[8], c[8][2];
#pragma unroll
  for (int i = 0; i < 8; i++) { f[i] = threadIdx.x + i; c[i][0] = i; c[i][1] = -i; }
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE != 1) { f[i] = fma(f[i], a, b); f[i] = fma(f[i], a, b); f[i] = fma(f[i], a, b); f[i] = fma(f[i], a, b); }
      if (MODE != 0) mb_dmma(c[i][0], c[i][1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += f[i] + c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// INT32 logic pipe: eight independent LOP3 chains per thread (the Fitch walks' roofline denominator)
__global__ void __launch_bounds__(128) lop3_peak_kernel(unsigned int *out, int iters, unsigned int a, unsigned int b) {
  unsigned int x[8];
#pragma unroll
  for (int i = 0; i < 8; i++) x[i] = threadIdx.x * 2654435761u + i;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      x[i] = (x[i] & a) ^ (x[i] | b);  // one LOP3 each
      x[i] = (x[i] | a) ^ (x[i] & b);
      x[i] = (x[i] & a) ^ (x[i] | b);
      x[i] = (x[i] | a) ^ (x[i] & b);
    }
  }
  unsigned int s = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

static double lop3_peak_run(int blocks_per_sm, int iters) {
  int dev = 0, sm = 0;
  PLG_CUDA(cudaGetDevice(&dev));
  PLG_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
  DevBuf<unsigned int> out;
  out.alloc((size_t)sm * blocks_per_sm * 128);
  lop3_peak_kernel<<<sm * blocks_per_sm, 128>>>(out.p, 100, 0x5a5a5a5au, 0x0f0f3c3cu);
  cudaEvent_t e0, e1;
  PLG_CUDA(cudaEventCreate(&e0));
  PLG_CUDA(cudaEventCreate(&e1));
  PLG_CUDA(cudaEventRecord(e0));
  lop3_peak_kernel<<<sm * blocks_per_sm, 128>>>(out.p, iters, 0x5a5a5a5au, 0x0f0f3c3cu);
  PLG_CUDA(cudaEventRecord(e1));
  PLG_CUDA(cudaEventSynchronize(e1));
  count_launch(2);
  float ms = 0;
  PLG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return (double)sm * blocks_per_sm * 128 * (double)iters * 32.0 / ((double)ms * 1e-3) / 1e12;  // lane-ops / s
}

template <int MODE>
static double fp64_peak_run(int blocks_per_sm, int iters) {
  int dev = 0, sm = 0;
  PLG_CUDA(cudaGetDevice(&dev));
  PLG_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
  DevBuf<double> out;
  out.alloc((size_t)sm * blocks_per_sm * 128);
  fp64_peak_kernel<MODE><<<sm * blocks_per_sm, 128>>>(out.p, 100, 1.0000001, 1e-9);
  cudaEvent_t e0, e1;
  PLG_CUDA(cudaEventCreate(&e0));
  PLG_CUDA(cudaEventCreate(&e1));
  PLG_CUDA(cudaEventRecord(e0));
  fp64_peak_kernel<MODE><<<sm * blocks_per_sm, 128>>>(out.p, iters, 1.0000001, 1e-9);
  PLG_CUDA(cudaEventRecord(e1));
  PLG_CUDA(cudaEventSynchronize(e1));
  count_launch(2);
  float ms = 0;
  PLG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double warps = (double)sm * blocks_per_sm * 4;
  const double flop = (MODE != 1 ? warps * iters * 32.0 * 64.0 : 0.0) + (MODE != 0 ? warps * iters * 8.0 * 512.0 : 0.0);
  return flop / ((double)ms * 1e-3) / 1e12;
}

}  // namespace plg

using namespace plg;

extern "C" int plg_fp64_peak(int mode, double *tflops) {
  PLG_API_BEGIN_LOCKED
  if (!tflops) PLG_FAIL(PLG_ERR_ARG, "null argument");
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0)
    PLG_FAIL(PLG_ERR_CUDA, "no CUDA device: pylogeny_b200 has no CPU fallback");
  double best = 0.0;
  for (int rep = 0; rep < 3; rep++) {  // best of three, 16 warps per SM, ~3-6 ms each
    double t;
    if (mode == 0) t = fp64_peak_run<0>(4, 20000);
    else if (mode == 1) t = fp64_peak_run<1>(4, 20000);
    else if (mode == 2) t = fp64_peak_run<2>(4, 20000);
    else if (mode == 3) t = lop3_peak_run(8, 20000);
    else PLG_FAIL(PLG_ERR_ARG, "mode: 0 DFMA, 1 DMMA m8n8k4, 2 both interleaved, 3 LOP3 (INT32 logic)");
    best = t > best ? t : best;
  }
  *tflops = best;
  PLG_API_END
}
