// FP64 peak probes for the roofline denominators of the Chebyshev (DMMA) cost path:
//   (1) DFMA chain throughput (vector FP64 pipe), (2) mma.sync.m8n8k4.f64 throughput (DMMA),
//   (3) both interleaved (do they share a pipe?).  Timed with CUDA events, best of 5.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int MODE>  // 0 dfma, 1 dmma, 2 both
__global__ void __launch_bounds__(256) probe(double* out, int iters, double seed) {
  double f[8], c[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = seed + i + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = seed * i;
  const double a = seed * 1.0000001, b = seed * 0.9999999;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fma(f[i], a, b);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int i = 0; i < 8; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
double run(int sms, int iters) {
  double* out;
  const int grid = sms * 8, block = 256;
  cudaMalloc(&out, sizeof(double) * grid * block);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    cudaEventRecord(e0);
    probe<MODE><<<grid, block>>>(out, iters, 1.0 + 1e-9 * rep);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaFree(out);
  const double warps = (double)grid * block / 32;
  double flops = 0;
  if (MODE == 0 || MODE == 2) flops += warps * iters * 8.0 * 32 * 2;        // 8 DFMA per lane
  if (MODE == 1 || MODE == 2) flops += warps * iters * 8.0 * (8 * 8 * 4 * 2);  // 8 m8n8k4
  return flops / (best * 1e-3) / 1e12;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  const int iters = 20000;
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"dfma_tflops\": %.3f, \"dmma_m8n8k4_tflops\": %.3f, \"dfma_plus_dmma_tflops\": %.3f}\n",
         p.name, p.multiProcessorCount, run<0>(p.multiProcessorCount, iters), run<1>(p.multiProcessorCount, iters),
         run<2>(p.multiProcessorCount, iters));
  return 0;
}
