// Kernel instantiations for the data-fit residual costs (polynomial least squares, lorentzian),
// split from the other translation units so the build runs in parallel.  Rows are short (the
// model's parameters), so only the scalar-load variants are built.
#include "../../include/mystic_b200.h"
#include "de_kernels.cuh"
#include "de_small.cuh"
#include "de_subwarp.cuh"

namespace mb2 {
typedef void (*StepKernel)(const StepParams);
typedef void (*ThreadRowKernel)(const StepParams, int, int, int);

template <int COST>
static StepKernel fit_ldg(int base, int xover) {
  if (xover == XOVER_BIN) return (StepKernel)de_step_kernel<BASE_BEST1, XOVER_BIN, COST, false>;
  switch (base) {
    case BASE_BEST1: return (StepKernel)de_step_kernel<BASE_BEST1, XOVER_EXP, COST, false>;
    case BASE_RAND1: return (StepKernel)de_step_kernel<BASE_RAND1, XOVER_EXP, COST, false>;
    case BASE_RANDTOBEST1: return (StepKernel)de_step_kernel<BASE_RANDTOBEST1, XOVER_EXP, COST, false>;
    case BASE_BEST2: return (StepKernel)de_step_kernel<BASE_BEST2, XOVER_EXP, COST, false>;
    case BASE_RAND2: return (StepKernel)de_step_kernel<BASE_RAND2, XOVER_EXP, COST, false>;
  }
  return nullptr;
}

StepKernel get_datafit_ldg_kernel(int base, int xover, int cost) {
  return cost == MB2_COST_POLYFIT ? fit_ldg<COST_POLYFIT>(base, xover)
                                  : (cost == MB2_COST_LORENTZIAN ? fit_ldg<COST_LORENTZIAN>(base, xover) : nullptr);
}

ThreadRowKernel get_datafit_small_kernel(int cost) {
  return cost == MB2_COST_POLYFIT ? (ThreadRowKernel)de_step_small_kernel<COST_POLYFIT>
                                  : (cost == MB2_COST_LORENTZIAN ? (ThreadRowKernel)de_step_small_kernel<COST_LORENTZIAN> : nullptr);
}

ThreadRowKernel get_datafit_subwarp_kernel(int cost) {  // 4 lanes per row (dim <= 32)
  return cost == MB2_COST_POLYFIT
             ? (ThreadRowKernel)de_step_subwarp_kernel<COST_POLYFIT, 4, false>
             : (cost == MB2_COST_LORENTZIAN ? (ThreadRowKernel)de_step_subwarp_kernel<COST_LORENTZIAN, 4, false> : nullptr);
}

ThreadRowKernel get_datafit_subwarp_islands_kernel(int cost) {
  return cost == MB2_COST_POLYFIT
             ? (ThreadRowKernel)de_step_subwarp_islands_kernel<COST_POLYFIT, 4, false>
             : (cost == MB2_COST_LORENTZIAN ? (ThreadRowKernel)de_step_subwarp_islands_kernel<COST_LORENTZIAN, 4, false> : nullptr);
}
}  // namespace mb2
