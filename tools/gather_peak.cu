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


// ---- the same pattern through the bulk-copy engine (cp.async.bulk, SASS UBLKCP): a warp owns `ST` stages
// of 8 rows; one warp instruction issues the 8 parent copies (lanes 0..7) and the 8*K donor-segment copies
// (lanes 8..8+8K-1) of a stage onto its mbarrier, the rows leave with 8 bulk stores.  No registers are held
// while the data flies and the next stage is requested before this one is touched.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n.reg .pred p;\nWAIT_LOOP:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 1000000;\n@!p bra WAIT_LOOP;\n}\n" ::"r"(
          smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(smem_u32(src)), "r"(bytes) : "memory");
}

template <int K, int ST, int WARPS, int R = 8>
__global__ void __launch_bounds__(WARPS * 32) bulk_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t np, int D,
                                                          int seg, uint32_t salt) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[WARPS * ST];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned rowb = (unsigned)D * 8u, segb = (unsigned)seg * 32u;
  const size_t stage_bytes = R * (size_t)rowb + R * (size_t)K * segb;
  unsigned char* my = smem + (size_t)warp * ST * stage_bytes;
  uint64_t* bar = bars + warp * ST;
  if (lane == 0)
    for (int s = 0; s < ST; ++s) mbar_init(bar + s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int64_t gw = (int64_t)blockIdx.x * WARPS + warp, nw = (int64_t)gridDim.x * WARPS;
  const int NQ = D >> 2;
  auto issue = [&](int64_t ch, int s) {
    unsigned char* st = my + (size_t)s * stage_bytes;
    if (lane == 0) mbar_expect_tx(bar + s, (unsigned)stage_bytes);
    __syncwarp();
    if (lane < R) {
      bulk_g2s(st + lane * rowb, in + (ch * R + lane) * D, rowb, bar + s);
    } else if (lane < R + R * K) {
      const int row = (lane - R) % R, j = (lane - R) / R;
      const int64_t c = ch * R + row;
      const uint32_t h = mix((uint32_t)c * 0x9E3779B9u + salt);
      int q0 = (int)(h % (uint32_t)NQ);
      if (q0 + seg > NQ) q0 = NQ - seg;  // no wrap in the probe
      const int64_t r = (int64_t)(mix(h + 0x51ED270Bu * (j + 1)) % (uint32_t)np);
      bulk_g2s(st + R * rowb + (size_t)(j * R + row) * segb, in + r * D + 4 * q0, segb, bar + s);
    }
  };
  int64_t nch = np / R;
  for (int s = 0; s < ST; ++s)
    if (gw + s * nw < nch) issue(gw + s * nw, s);
  int s = 0;
  unsigned parity = 0;
  for (int64_t ch = gw; ch < nch; ch += nw) {
    unsigned char* st = my + (size_t)s * stage_bytes;
    mbar_wait(bar + s, parity);
    // touch: add the first donor element of each segment to the first element of its row
    if (K > 0 && lane < R) {
      double a = 0.0;
      for (int j = 0; j < K; ++j) a += *reinterpret_cast<double*>(st + R * rowb + (size_t)(j * R + lane) * segb);
      *reinterpret_cast<double*>(st + lane * rowb) += a;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane < R) {
      bulk_s2g(out + (ch * R + lane) * D, st + lane * rowb, rowb);
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
    if (ch + (int64_t)ST * nw < nch) issue(ch + (int64_t)ST * nw, s);
    if (++s == ST) {
      s = 0;
      parity ^= 1u;
    }
  }
  if (lane < R) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

template <int K, int ST, int WARPS, int R = 8>
static void run_bulk(const double* in, double* out, int64_t np, int D, int seg) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const size_t smem = (size_t)WARPS * ST * (R * (size_t)D * 8 + R * (size_t)K * seg * 32);
  cudaFuncSetAttribute(bulk_kernel<K, ST, WARPS, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bulk_kernel<K, ST, WARPS, R>, WARPS * 32, smem);
  if (occ < 1) {
    printf("{\"bulk\": true, \"error\": \"does not fit\", \"smem\": %zu}\n", smem);
    return;
  }
  const int grid = 148 * occ;
  float best = 1e30f;
  for (int rep = 0; rep < 12; ++rep) {
    cudaEventRecord(e0);
    bulk_kernel<K, ST, WARPS, R><<<grid, WARPS * 32, smem>>>(in, out, np, D, seg, 17u * rep + 1u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 2 && ms < best) best = ms;
  }
  const double bytes = (double)np * (16.0 * D + (double)K * seg * 32.0);
  printf("{\"bulk\": true, \"donors\": %d, \"chunks_per_donor\": %d, \"stages\": %d, \"warps\": %d, \"rows_per_stage\": %d, \"ctas_per_sm\": %d, \"ms\": %.4f, \"requested_GB\": %.4f, \"GBps\": %.1f}\n",
         K, seg, ST, WARPS, R, occ, best, bytes / 1e9, bytes / 1e6 / best);
}

// Whole-row reads of an IN-PLACE generation of the full-row strategies (BASELINE configs 2 and 5): per candidate the
// parent row and K random donor rows come in by bulk copies, nothing goes out.  A warp owns a ring of ST stages of
// (1 + K) rows -- the fused kernel's groups x stages -- and touches one element per row.
template <int K, int ST, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) rows_in_kernel(const double* __restrict__ in, double* __restrict__ out, int64_t np, int D,
                                                             uint32_t salt) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[WARPS * ST];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned rowb = (unsigned)D * 8u;
  const size_t stage_bytes = (size_t)(1 + K) * rowb;
  unsigned char* my = smem + (size_t)warp * ST * stage_bytes;
  uint64_t* bar = bars + warp * ST;
  if (lane == 0)
    for (int s = 0; s < ST; ++s) mbar_init(bar + s, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  const int64_t gw = (int64_t)blockIdx.x * WARPS + warp, nw = (int64_t)gridDim.x * WARPS;
  auto issue = [&](int64_t c, int s) {
    unsigned char* st = my + (size_t)s * stage_bytes;
    if (lane == 0) mbar_expect_tx(bar + s, (unsigned)stage_bytes);
    __syncwarp();
    if (lane <= K) {
      const uint32_t h = mix((uint32_t)c * 0x9E3779B9u + salt);
      const int64_t r = lane == 0 ? c : (int64_t)(mix(h + 0x51ED270Bu * lane) % (uint32_t)np);
      bulk_g2s(st + (size_t)lane * rowb, in + r * D, rowb, bar + s);
    }
  };
  for (int s = 0; s < ST; ++s)
    if (gw + s * nw < np) issue(gw + s * nw, s);
  int s = 0;
  unsigned parity = 0;
  double acc = 0.0;
  for (int64_t c = gw; c < np; c += nw) {
    unsigned char* st = my + (size_t)s * stage_bytes;
    mbar_wait(bar + s, parity);
    if (lane <= K) acc += *reinterpret_cast<double*>(st + (size_t)lane * rowb);
    __syncwarp();
    if (c + (int64_t)ST * nw < np) issue(c + (int64_t)ST * nw, s);
    if (++s == ST) {
      s = 0;
      parity ^= 1u;
    }
  }
  if (acc == 12345.678) out[0] = acc;
}

template <int K, int ST, int WARPS>
static void run_rows_in(const double* in, double* out, int64_t np, int D) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const size_t smem = (size_t)WARPS * ST * (1 + K) * (size_t)D * 8;
  cudaFuncSetAttribute(rows_in_kernel<K, ST, WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int occ = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, rows_in_kernel<K, ST, WARPS>, WARPS * 32, smem);
  if (occ < 1) {
    printf("{\"rows_in\": true, \"error\": \"does not fit\", \"smem\": %zu}\n", smem);
    return;
  }
  const int grid = 148 * occ;
  float best = 1e30f;
  for (int rep = 0; rep < 12; ++rep) {
    cudaEventRecord(e0);
    rows_in_kernel<K, ST, WARPS><<<grid, WARPS * 32, smem>>>(in, out, np, D, 17u * rep + 1u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep >= 2 && ms < best) best = ms;
  }
  const double bytes = (double)np * (1.0 + K) * D * 8.0;
  printf("{\"rows_in\": true, \"donors\": %d, \"stages\": %d, \"warps\": %d, \"ctas_per_sm\": %d, \"rows_in_flight_per_sm\": %d, \"ms\": %.4f, \"requested_GB\": %.4f, \"GBps\": %.1f}\n",
         K, ST, WARPS, occ, occ * WARPS * ST * (1 + K), best, bytes / 1e9, bytes / 1e6 / best);
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
  if (argc > 3 && argv[3][0] == 'r') {  // `rows`: whole-row reads of an in-place generation (parent + 2 donors in, nothing out)
    if (D >= 512) {
      run_rows_in<2, 1, 8>(a, b, np, D);    // the fused kernel's ring at config 2: 8 groups x 1 stage x 3 rows
      run_rows_in<2, 1, 6>(a, b, np, D);
      run_rows_in<2, 1, 4>(a, b, np, D);
      run_rows_in<2, 2, 4>(a, b, np, D);
      run_rows_in<0, 1, 24>(a, b, np, D);   // contiguous rows only: the streaming-read ceiling of the mechanism
      run_rows_in<0, 2, 12>(a, b, np, D);
    } else {
      run_rows_in<2, 2, 16>(a, b, np, D);   // config 5: 16 groups x 2 stages x 3 rows
      run_rows_in<2, 1, 32>(a, b, np, D);
      run_rows_in<2, 3, 10>(a, b, np, D);
      run_rows_in<0, 4, 24>(a, b, np, D);
    }
    cudaDeviceSynchronize();
    return 0;
  }
  if (D >= 512) {  // long rows (x2: Rosenbrock 1024-D, exponential crossover): one row per stage
    run_bulk<0, 1, 16, 1>(a, b, np, D, 0);
    run_bulk<0, 2, 8, 1>(a, b, np, D, 0);
    run_bulk<3, 1, 16, 1>(a, b, np, D, 3);
    run_bulk<3, 2, 8, 1>(a, b, np, D, 3);
    run_bulk<3, 2, 12, 1>(a, b, np, D, 3);
    run_bulk<3, 1, 24, 1>(a, b, np, D, 3);
    run_bulk<2, 2, 12, 1>(a, b, np, D, 3);
    cudaDeviceSynchronize();
    return 0;
  }
  for (int occ : {2, 4, 8}) {
    run<0>(a, b, np, D, 0, occ);
    run<3>(a, b, np, D, 1, occ);
    run<3>(a, b, np, D, 3, occ);
    run<3>(a, b, np, D, 4, occ);
    run<2>(a, b, np, D, 4, occ);
    run<5>(a, b, np, D, 4, occ);
  }
  run_bulk<0, 2, 8>(a, b, np, D, 0);
  run_bulk<0, 4, 8>(a, b, np, D, 0);
  run_bulk<0, 2, 16>(a, b, np, D, 0);
  run_bulk<3, 1, 8>(a, b, np, D, 3);
  run_bulk<3, 2, 8>(a, b, np, D, 3);
  run_bulk<3, 3, 8>(a, b, np, D, 3);
  run_bulk<3, 2, 4>(a, b, np, D, 3);
  run_bulk<3, 2, 16>(a, b, np, D, 3);
  run_bulk<3, 4, 4>(a, b, np, D, 3);
  run_bulk<2, 2, 8>(a, b, np, D, 4);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("error: %s\n", cudaGetErrorString(e));
    return 1;
  }
  return 0;
}
