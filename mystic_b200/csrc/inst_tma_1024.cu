// Kernel instantiations, split over translation units so the build runs in parallel.
#include "../../include/mystic_b200.h"
#include "de_tma.cuh"

namespace mb2 {
typedef void (*TmaKernel)(const StepParams, int);

template <int GW>
static TmaKernel tma_cost(int cost) {
  if (cost >= MB2_COST_SPHERE && cost <= MB2_COST_POWERS) return de_step_best1bin_tma_kernel<COST_SEPARABLE, GW, 1024>;
  switch (cost) {
    case MB2_COST_ROSEN: return de_step_best1bin_tma_kernel<COST_ROSEN, GW, 1024>;
    case MB2_COST_CORANA: return de_step_best1bin_tma_kernel<COST_CORANA, GW, 1024>;
    case MB2_COST_GRIEWANGK: return de_step_best1bin_tma_kernel<COST_GRIEWANGK, GW, 1024>;
    case MB2_COST_ZIMMERMANN: return de_step_best1bin_tma_kernel<COST_ZIMMERMANN, GW, 1024>;
    case MB2_COST_CHEBYSHEV: return de_step_best1bin_tma_kernel<COST_CHEBYSHEV, GW, 1024>;
  }
  return nullptr;
}

// CTAs of up to 512 threads use the 128-register build, larger ones the 64-register build
TmaKernel get_tma_kernel_1024(int cost, int gwarps) {
  switch (gwarps) {
    case 1: return tma_cost<1>(cost);
    case 2: return tma_cost<2>(cost);
    case 4: return tma_cost<4>(cost);
    case 8: return tma_cost<8>(cost);
  }
  return nullptr;
}
}  // namespace mb2
