// Kernel instantiations, split over translation units so the build runs in parallel.
#include "../../include/mystic_b200.h"
#include "de_tma_pool.cuh"

namespace mb2 {
typedef void (*PoolKernel)(const StepParams, int, int, int);

// K2p: <= 18 one-warp consumer groups + two producer warps = 640 threads (5 warps per scheduler: 96 registers)
PoolKernel get_pool_kernel1(int cost) {
  if (cost >= MB2_COST_SPHERE && cost <= MB2_COST_POWERS) return de_step_best1bin_pool_kernel<COST_SEPARABLE, 1, 640>;
  switch (cost) {
    case MB2_COST_ROSEN: return de_step_best1bin_pool_kernel<COST_ROSEN, 1, 640>;
    case MB2_COST_CORANA: return de_step_best1bin_pool_kernel<COST_CORANA, 1, 640>;
    case MB2_COST_GRIEWANGK: return de_step_best1bin_pool_kernel<COST_GRIEWANGK, 1, 640>;
    case MB2_COST_ZIMMERMANN: return de_step_best1bin_pool_kernel<COST_ZIMMERMANN, 1, 640>;
    case MB2_COST_CHEBYSHEV: return de_step_best1bin_pool_kernel<COST_CHEBYSHEV, 1, 640>;
  }
  return nullptr;
}
}  // namespace mb2
