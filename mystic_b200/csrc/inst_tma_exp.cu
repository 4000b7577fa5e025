// Kernel instantiations, split over translation units so the build runs in parallel.
#include "../../include/mystic_b200.h"
#include "de_tma_exp.cuh"

namespace mb2 {
typedef void (*ExpTmaKernel)(const StepParams, int, int, int, int);

template <int GW>
static ExpTmaKernel exp_tma_cost(int cost) {
  if (cost >= MB2_COST_SPHERE && cost <= MB2_COST_POWERS) return de_step_exp_tma_kernel<COST_SEPARABLE, GW, 1024>;
  switch (cost) {
    case MB2_COST_ROSEN: return de_step_exp_tma_kernel<COST_ROSEN, GW, 1024>;
    case MB2_COST_CORANA: return de_step_exp_tma_kernel<COST_CORANA, GW, 1024>;
    case MB2_COST_GRIEWANGK: return de_step_exp_tma_kernel<COST_GRIEWANGK, GW, 1024>;
  }
  return nullptr;
}

// K2e: exponential crossover over TMA-staged long rows (all strategies but Best1Bin)
ExpTmaKernel get_exp_tma_kernel(int cost, int gwarps) {
  switch (gwarps) {
    case 1: return exp_tma_cost<1>(cost);
    case 2: return exp_tma_cost<2>(cost);
    case 4: return exp_tma_cost<4>(cost);
    case 8: return exp_tma_cost<8>(cost);
  }
  return nullptr;
}
}  // namespace mb2
