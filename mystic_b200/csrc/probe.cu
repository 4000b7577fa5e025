// FP64 peak probes: the roofline denominator of the Chebyshev / polynomial-fit contraction
// (de_cheb_mma.cuh) measured on the device the bench runs on, in the same process, instead of a
// constant.  (1) a DFMA chain per thread (vector FP64 pipe), (2) mma.sync.m8n8k4.f64 (SASS
// DMMA.8x8x4, the shape every f64 MMA is lowered to on sm_100a).  Timed with CUDA events on the
// caller's stream, best of 5 after one warm-up launch.
#include <cuda_runtime.h>

#include "../../include/mystic_b200.h"

namespace {

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

template <bool DMMA>
__global__ void __launch_bounds__(256) fp64_probe_kernel(double* out, int iters, double seed) {
  double f[8], c[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = seed + i + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = seed * i;
  const double a = seed * 1.0000001, b = seed * 0.9999999;
  for (int it = 0; it < iters; ++it) {
    if (!DMMA) {
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = fma(f[i], a, b);
    } else {
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

template <bool DMMA>
cudaError_t run_probe(int sms, int iters, cudaStream_t s, double* out, double* tflops) {
  const int grid = sms * 8, block = 256;
  cudaEvent_t e0, e1;
  cudaError_t err = cudaEventCreate(&e0);
  if (err != cudaSuccess) return err;
  err = cudaEventCreate(&e1);
  if (err != cudaSuccess) return err;
  float best = 1e30f;
  for (int rep = 0; rep < 6 && err == cudaSuccess; ++rep) {
    cudaEventRecord(e0, s);
    fp64_probe_kernel<DMMA><<<grid, block, 0, s>>>(out, iters, 1.0 + 1e-9 * rep);
    cudaEventRecord(e1, s);
    err = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  const double warps = (double)grid * block / 32;
  const double flops = DMMA ? warps * iters * 8.0 * (8 * 8 * 4 * 2) : warps * iters * 8.0 * 32 * 2;
  *tflops = flops / (best * 1e-3) / 1e12;
  return err;
}

}  // namespace

extern "C" int mb2_probe_fp64_peak(int device, int32_t iters, double* dfma_tflops, double* dmma_tflops, void* stream) {
  if (iters < 1 || dfma_tflops == nullptr || dmma_tflops == nullptr) return MB2_EINVAL;
  int prev = -1;
  if (cudaGetDevice(&prev) != cudaSuccess || cudaSetDevice(device) != cudaSuccess) return MB2_ECUDA;
  int sms = 0;
  double* out = nullptr;
  cudaError_t err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (err == cudaSuccess) err = cudaMalloc(&out, sizeof(double) * (size_t)sms * 8 * 256);
  if (err == cudaSuccess) err = run_probe<false>(sms, iters, (cudaStream_t)stream, out, dfma_tflops);
  if (err == cudaSuccess) err = run_probe<true>(sms, iters, (cudaStream_t)stream, out, dmma_tflops);
  if (out != nullptr) cudaFree(out);
  cudaSetDevice(prev);
  return err == cudaSuccess ? MB2_OK : MB2_ECUDA;
}
