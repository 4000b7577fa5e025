// Kernel instantiations, split over translation units so the build runs in parallel.
#include "../../include/mystic_b200.h"
#include "de_tma_exp.cuh"

namespace mb2 {
typedef void (*ExpTmaKernel)(const StepParams, int, int, int, int);

#ifndef MB2_EXP_TMA_ODD
#define MB2_EXP_TMA_ODD false
#define MB2_EXP_TMA_GETTER get_exp_tma_kernel_even
#endif

template <int GW, int MAXT>
static ExpTmaKernel exp_tma_cost(int cost) {
  if (cost >= MB2_COST_SPHERE && cost <= MB2_COST_POWERS) return de_step_exp_tma_kernel<COST_SEPARABLE, GW, MAXT, MB2_EXP_TMA_ODD>;
  switch (cost) {
    case MB2_COST_ROSEN: return de_step_exp_tma_kernel<COST_ROSEN, GW, MAXT, MB2_EXP_TMA_ODD>;
    case MB2_COST_CORANA: return de_step_exp_tma_kernel<COST_CORANA, GW, MAXT, MB2_EXP_TMA_ODD>;
    case MB2_COST_GRIEWANGK: return de_step_exp_tma_kernel<COST_GRIEWANGK, GW, MAXT, MB2_EXP_TMA_ODD>;
  }
  return nullptr;
}

// K2e: exponential crossover over TMA-staged long rows (all strategies but Best1Bin).  CTAs of up to 512 threads
// (groups of one or two warps) use the 128-register build, larger ones the 64-register build.
// (this file is compiled twice: rows of even length here, rows of odd length through inst_tma_exp_odd.cu)
ExpTmaKernel MB2_EXP_TMA_GETTER(int cost, int gwarps, int threads) {
  if (threads <= 512 && gwarps <= 2) return gwarps == 1 ? exp_tma_cost<1, 512>(cost) : exp_tma_cost<2, 512>(cost);
  switch (gwarps) {
    case 1: return exp_tma_cost<1, 1024>(cost);
    case 2: return exp_tma_cost<2, 1024>(cost);
    case 4: return exp_tma_cost<4, 1024>(cost);
    case 8: return exp_tma_cost<8, 1024>(cost);
  }
  return nullptr;
}
}  // namespace mb2
