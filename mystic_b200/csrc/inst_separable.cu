// Kernel instantiations for the separable test functions (sphere, rastrigin, schwefel, ackley:
// one kernel body, the function is picked at run time), split from the other translation units
// so the build runs in parallel.
#include "../../include/mystic_b200.h"
#include "de_kernels.cuh"
#include "de_small.cuh"
#include "de_subwarp.cuh"

namespace mb2 {
typedef void (*StepKernel)(const StepParams);
typedef void (*ThreadRowKernel)(const StepParams, int, int, int);

template <bool VEC>
static StepKernel sep_ldg(int base, int xover) {
  if (xover == XOVER_BIN) return (StepKernel)de_step_kernel<BASE_BEST1, XOVER_BIN, COST_SEPARABLE, VEC>;
  switch (base) {
    case BASE_BEST1: return (StepKernel)de_step_kernel<BASE_BEST1, XOVER_EXP, COST_SEPARABLE, VEC>;
    case BASE_RAND1: return (StepKernel)de_step_kernel<BASE_RAND1, XOVER_EXP, COST_SEPARABLE, VEC>;
    case BASE_RANDTOBEST1: return (StepKernel)de_step_kernel<BASE_RANDTOBEST1, XOVER_EXP, COST_SEPARABLE, VEC>;
    case BASE_BEST2: return (StepKernel)de_step_kernel<BASE_BEST2, XOVER_EXP, COST_SEPARABLE, VEC>;
    case BASE_RAND2: return (StepKernel)de_step_kernel<BASE_RAND2, XOVER_EXP, COST_SEPARABLE, VEC>;
  }
  return nullptr;
}

StepKernel get_separable_ldg_kernel(int base, int xover, bool vec) { return vec ? sep_ldg<true>(base, xover) : sep_ldg<false>(base, xover); }

ThreadRowKernel get_separable_small_kernel() { return (ThreadRowKernel)de_step_small_kernel<COST_SEPARABLE>; }

template <int LPR>
static ThreadRowKernel sep_sub(bool vec) {
  return vec ? (ThreadRowKernel)de_step_subwarp_kernel<COST_SEPARABLE, LPR, true>
             : (ThreadRowKernel)de_step_subwarp_kernel<COST_SEPARABLE, LPR, false>;
}

ThreadRowKernel get_separable_subwarp_kernel(int lanes_per_row, bool vec) {
  switch (lanes_per_row) {
    case 4: return sep_sub<4>(vec);
    case 8: return sep_sub<8>(vec);
    case 16: return sep_sub<16>(vec);
    case 32: return sep_sub<32>(vec);
  }
  return nullptr;
}

template <int LPR>
static ThreadRowKernel sep_isl(bool vec) {
  return vec ? (ThreadRowKernel)de_step_subwarp_islands_kernel<COST_SEPARABLE, LPR, true>
             : (ThreadRowKernel)de_step_subwarp_islands_kernel<COST_SEPARABLE, LPR, false>;
}

ThreadRowKernel get_separable_subwarp_islands_kernel(int lanes_per_row, bool vec) {
  switch (lanes_per_row) {
    case 4: return sep_isl<4>(vec);
    case 8: return sep_isl<8>(vec);
    case 16: return sep_isl<16>(vec);
    case 32: return sep_isl<32>(vec);
  }
  return nullptr;
}
}  // namespace mb2
