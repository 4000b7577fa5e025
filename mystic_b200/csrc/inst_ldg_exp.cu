// Kernel instantiations, split over translation units so the build runs in parallel.
#include "../../include/mystic_b200.h"
#include "de_kernels.cuh"

namespace mb2 {
typedef void (*StepKernel)(const StepParams);

template <int BASE, int COST>
static StepKernel exp_vec(bool vec) {
  return vec ? (StepKernel)de_step_kernel<BASE, XOVER_EXP, COST, true> : (StepKernel)de_step_kernel<BASE, XOVER_EXP, COST, false>;
}

template <int BASE>
static StepKernel exp_cost(int cost, bool vec) {
  switch (cost) {
    case MB2_COST_ROSEN: return exp_vec<BASE, COST_ROSEN>(vec);
    case MB2_COST_CORANA: return exp_vec<BASE, COST_CORANA>(vec);
    case MB2_COST_GRIEWANGK: return exp_vec<BASE, COST_GRIEWANGK>(vec);
    case MB2_COST_ZIMMERMANN: return exp_vec<BASE, COST_ZIMMERMANN>(vec);
    case MB2_COST_CHEBYSHEV: return exp_vec<BASE, COST_CHEBYSHEV>(vec);
  }
  return nullptr;
}

StepKernel get_ldg_exp_kernel_a(int base, int cost, bool vec) {
  switch (base) {
    case BASE_BEST1: return exp_cost<BASE_BEST1>(cost, vec);
    case BASE_RAND1: return exp_cost<BASE_RAND1>(cost, vec);
    case BASE_RANDTOBEST1: return exp_cost<BASE_RANDTOBEST1>(cost, vec);
  }
  return nullptr;
}
}  // namespace mb2
