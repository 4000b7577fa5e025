// Kernel instantiations, split over translation units so the build runs in parallel.
#include "../../include/mystic_b200.h"
#include "de_kernels.cuh"

namespace mb2 {
typedef void (*StepKernel)(const StepParams);
StepKernel get_ldg_exp_kernel_a(int base, int cost, bool vec);

template <int BASE, int XOVER, int COST>
static StepKernel pick_vec(bool vec) {
  return vec ? (StepKernel)de_step_kernel<BASE, XOVER, COST, true> : (StepKernel)de_step_kernel<BASE, XOVER, COST, false>;
}

template <int BASE, int XOVER>
static StepKernel pick_cost(int cost, bool vec) {
  switch (cost) {
    case MB2_COST_ROSEN: return pick_vec<BASE, XOVER, COST_ROSEN>(vec);
    case MB2_COST_CORANA: return pick_vec<BASE, XOVER, COST_CORANA>(vec);
    case MB2_COST_GRIEWANGK: return pick_vec<BASE, XOVER, COST_GRIEWANGK>(vec);
    case MB2_COST_ZIMMERMANN: return pick_vec<BASE, XOVER, COST_ZIMMERMANN>(vec);
    case MB2_COST_CHEBYSHEV: return pick_vec<BASE, XOVER, COST_CHEBYSHEV>(vec);
  }
  return nullptr;
}

// mystic/strategy.py: only Best1Bin is binomial; every other name runs the exponential loop
StepKernel get_ldg_kernel(int base, int xover, int cost, bool vec) {
  if (xover == XOVER_BIN) return pick_cost<BASE_BEST1, XOVER_BIN>(cost, vec);
  switch (base) {
    case BASE_BEST2: return pick_cost<BASE_BEST2, XOVER_EXP>(cost, vec);
    case BASE_RAND2: return pick_cost<BASE_RAND2, XOVER_EXP>(cost, vec);
  }
  return get_ldg_exp_kernel_a(base, cost, vec);
}
}  // namespace mb2
