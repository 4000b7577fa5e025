// Kernel instantiations, split over translation units so the build runs in parallel.
#include "../../include/mystic_b200.h"
#include "de_tma_pool.cuh"

namespace mb2 {
typedef void (*PoolKernel)(const StepParams, int, int, int);

// K2p: <= 4 consumer groups of 4 warps + the producer warp = 544 threads (120 registers)
PoolKernel get_pool_kernel4(int cost) {
  if (cost >= MB2_COST_SPHERE && cost <= MB2_COST_POWERS) return de_step_best1bin_pool_kernel<COST_SEPARABLE, 4, 544>;
  switch (cost) {
    case MB2_COST_ROSEN: return de_step_best1bin_pool_kernel<COST_ROSEN, 4, 544>;
    case MB2_COST_CORANA: return de_step_best1bin_pool_kernel<COST_CORANA, 4, 544>;
    case MB2_COST_GRIEWANGK: return de_step_best1bin_pool_kernel<COST_GRIEWANGK, 4, 544>;
    case MB2_COST_ZIMMERMANN: return de_step_best1bin_pool_kernel<COST_ZIMMERMANN, 4, 544>;
    case MB2_COST_CHEBYSHEV: return de_step_best1bin_pool_kernel<COST_CHEBYSHEV, 4, 544>;
  }
  return nullptr;
}
}  // namespace mb2
