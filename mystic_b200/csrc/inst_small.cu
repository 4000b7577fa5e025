// Kernel instantiations, split over translation units so the build runs in parallel.
#include "../../include/mystic_b200.h"
#include "de_cheb_mma.cuh"
#include "de_small.cuh"

namespace mb2 {
typedef void (*TmaKernel)(const StepParams, int);
typedef void (*ThreadRowKernel)(const StepParams, int, int, int);
TmaKernel get_tma_kernel_512(int cost, int gwarps);
TmaKernel get_tma_kernel_1024(int cost, int gwarps);
TmaKernel get_tma_kernel_768(int cost, int gwarps);

TmaKernel get_tma_kernel(int cost, int gwarps, int threads) {
  if (threads > 512 && threads <= 768) {
    TmaKernel k = get_tma_kernel_768(cost, gwarps);
    if (k != nullptr) return k;
  }
  return threads > 512 ? get_tma_kernel_1024(cost, gwarps) : get_tma_kernel_512(cost, gwarps);
}

ThreadRowKernel get_small_kernel(int cost) {
  switch (cost) {
    case MB2_COST_ROSEN: return de_step_small_kernel<COST_ROSEN>;
    case MB2_COST_CORANA: return de_step_small_kernel<COST_CORANA>;
    case MB2_COST_GRIEWANGK: return de_step_small_kernel<COST_GRIEWANGK>;
    case MB2_COST_ZIMMERMANN: return de_step_small_kernel<COST_ZIMMERMANN>;
    case MB2_COST_CHEBYSHEV: return de_step_small_kernel<COST_CHEBYSHEV>;
  }
  return nullptr;
}

ThreadRowKernel get_cheb_mma_kernel(int ksteps, bool resid) {
  if (resid)
    return ksteps == 1 ? de_step_cheb_mma_kernel<1, true>
                       : (ksteps == 2 ? de_step_cheb_mma_kernel<2, true> : de_step_cheb_mma_kernel<4, true>);
  return ksteps == 1 ? de_step_cheb_mma_kernel<1> : (ksteps == 2 ? de_step_cheb_mma_kernel<2> : de_step_cheb_mma_kernel<4>);
}
}  // namespace mb2
