// Kernel instantiations, split over translation units so the build runs in parallel.
#include "../../include/mystic_b200.h"
#include "de_tma_pool.cuh"

namespace mb2 {
typedef void (*PoolKernel)(const StepParams, int, int, int);

// K2p: <= 8 consumer groups of 2 warps + the producer warp = 544 threads (120 registers)
PoolKernel get_pool_kernel4(int cost);  // groups of 4 warps: inst_tma_pool4.cu
PoolKernel get_pool_kernel1(int cost);  // one-warp groups (rows under 768 dimensions): inst_tma_pool1.cu
PoolKernel get_pool_kernel(int cost, int gwarps) {
  if (gwarps == 4) return get_pool_kernel4(cost);
  if (gwarps == 1) return get_pool_kernel1(cost);
  if (cost >= MB2_COST_SPHERE && cost <= MB2_COST_POWERS) return de_step_best1bin_pool_kernel<COST_SEPARABLE, 2, 544>;
  switch (cost) {
    case MB2_COST_ROSEN: return de_step_best1bin_pool_kernel<COST_ROSEN, 2, 544>;
    case MB2_COST_CORANA: return de_step_best1bin_pool_kernel<COST_CORANA, 2, 544>;
    case MB2_COST_GRIEWANGK: return de_step_best1bin_pool_kernel<COST_GRIEWANGK, 2, 544>;
    case MB2_COST_ZIMMERMANN: return de_step_best1bin_pool_kernel<COST_ZIMMERMANN, 2, 544>;
    case MB2_COST_CHEBYSHEV: return de_step_best1bin_pool_kernel<COST_CHEBYSHEV, 2, 544>;
  }
  return nullptr;
}
}  // namespace mb2
