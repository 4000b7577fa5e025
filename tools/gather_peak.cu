// Memory-system ceiling for the access pattern of the short-row DE generation (BASELINE config 4):
// stream NP rows of D doubles in, gather K random contiguous segments of `seg` 4-element chunks
// from random rows, stream the rows out.  No arithmetic besides a checksum: what this kernel
// reaches is what the fused kernel can reach at most for the same bytes.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_peak tools/gather_peak.cu
//   run:   tools/gather_peak [np=1048576] [dim=64]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ double2 ld_stream2(const double* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream2(double* p, double a, double b) {
  asm volatile("st.global.L1::no_allocate.v2.f64 [%0], {%1, %2};" ::"l"(p), "d"(a), "d"(b) : "memory");
}
__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// 4 lanes per row, 8 rows per warp; lane `sub` copies the 16-byte pieces sub, sub+4, ... and
// gathers chunk (start + sub) of each donor when sub < seg
template <int K>
__global__ void __launch_bounds__(256) gather_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t np, int D,
                                                     int seg, uint32_t salt) {
  const int lane = threadIdx.x & 31, grp = lane >> 2, sub = lane & 3;
  const int64_t gw = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int NQ = D >> 2;
  for (int64_t ch = gw; ch * 8 < np; ch += nw) {
    const int64_t c = ch * 8 + grp;
    if (c >= np) continue;
    const double* prow = in + c * D;
    double2 p[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int i = 2 * (t * 4 + sub);
      p[t] = (i < D) ? ld_stream2(prow + i) : make_double2(0., 0.);
    }
    double acc = 0.0;
    if (K > 0) {
      const uint32_t h = mix((uint32_t)c * 0x9E3779B9u + salt);
      const int q0 = (int)(h % (uint32_t)NQ);
      int q = q0 + sub;
      if (q >= NQ) q -= NQ;
      double2 d[K > 0 ? K : 1][2];
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int64_t r = (int64_t)(mix(h + 0x51ED270Bu * (j + 1)) % (uint32_t)np);
        if (sub < seg) {
          d[j][0] = ld_stream2(in + r * D + 4 * q);
          d[j][1] = ld_stream2(in + r * D + 4 * q + 2);
        } else {
          d[j][0] = d[j][1] = make_double2(0., 0.);
        }
      }
#pragma unroll
      for (int j = 0; j < K; ++j) acc += d[j][0].x + d[j][0].y + d[j][1].x + d[j][1].y;
    }
    double* nrow = out + c * D;
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      const int i = 2 * (t * 4 + sub);
      if (i < D) st_stream2(nrow + i, p[t].x + acc, p[t].y);
    }
  }
}

template <int K>
static void run(const double* in, double* out, int64_t np, int D, int seg, int ctas_per_sm) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int grid = 148 * ctas_per_sm;
  float best = 1e30f;
  for (int rep = 0; rep < 12; ++rep) {
    cudaEventRecord(e0);
    gather_kernel<K><<<grid, 256>>>(in, out, np, D, seg, 17u * rep + 1u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 2 && ms < best) best = ms;
  }
  const double bytes = (double)np * (16.0 * D + (double)K * seg * 32.0);
  printf("{\"donors\": %d, \"chunks_per_donor\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"requested_GB\": %.4f, \"GBps\": %.1f}\n", K, seg,
         ctas_per_sm, best, bytes / 1e9, bytes / 1e6 / best);
}

int main(int argc, char** argv) {
  const int64_t np = argc > 1 ? atoll(argv[1]) : 1048576;
  const int D = argc > 2 ? atoi(argv[2]) : 64;
  double *a, *b;
  cudaMalloc(&a, sizeof(double) * np * D);
  cudaMalloc(&b, sizeof(double) * np * D);
  cudaMemset(a, 0, sizeof(double) * np * D);
  cudaMemset(b, 0, sizeof(double) * np * D);
  for (int occ : {2, 4, 8}) {
    run<0>(a, b, np, D, 0, occ);
    run<3>(a, b, np, D, 1, occ);
    run<3>(a, b, np, D, 3, occ);
    run<3>(a, b, np, D, 4, occ);
    run<2>(a, b, np, D, 4, occ);
    run<5>(a, b, np, D, 4, occ);
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("error: %s\n", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}
