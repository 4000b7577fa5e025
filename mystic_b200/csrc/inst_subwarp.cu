// Kernel instantiations, split over translation units so the build runs in parallel.
#include "../../include/mystic_b200.h"
#include "de_subwarp.cuh"

namespace mb2 {
typedef void (*ThreadRowKernel)(const StepParams, int, int, int);

template <int COST, int LPR>
static ThreadRowKernel sub_vec(bool vec) {
  return vec ? (ThreadRowKernel)de_step_subwarp_kernel<COST, LPR, true> : (ThreadRowKernel)de_step_subwarp_kernel<COST, LPR, false>;
}

template <int COST>
static ThreadRowKernel sub_lpr(int lpr, bool vec) {
  switch (lpr) {
    case 4: return sub_vec<COST, 4>(vec);
    case 8: return sub_vec<COST, 8>(vec);
    case 16: return sub_vec<COST, 16>(vec);
    case 32: return sub_vec<COST, 32>(vec);
  }
  return nullptr;
}

template <int COST, int LPR>
static ThreadRowKernel isl_vec(bool vec) {
  return vec ? (ThreadRowKernel)de_step_subwarp_islands_kernel<COST, LPR, true>
             : (ThreadRowKernel)de_step_subwarp_islands_kernel<COST, LPR, false>;
}

template <int COST>
static ThreadRowKernel isl_lpr(int lpr, bool vec) {
  switch (lpr) {
    case 4: return isl_vec<COST, 4>(vec);
    case 8: return isl_vec<COST, 8>(vec);
    case 16: return isl_vec<COST, 16>(vec);
    case 32: return isl_vec<COST, 32>(vec);
  }
  return nullptr;
}

// batch of independent solvers: blockIdx.y = island
ThreadRowKernel get_subwarp_islands_kernel(int cost, int lanes_per_row, bool vec) {
  switch (cost) {
    case MB2_COST_ROSEN: return isl_lpr<COST_ROSEN>(lanes_per_row, vec);
    case MB2_COST_CORANA: return isl_lpr<COST_CORANA>(lanes_per_row, vec);
    case MB2_COST_GRIEWANGK: return isl_lpr<COST_GRIEWANGK>(lanes_per_row, vec);
    case MB2_COST_ZIMMERMANN: return isl_lpr<COST_ZIMMERMANN>(lanes_per_row, vec);
    case MB2_COST_CHEBYSHEV: return isl_lpr<COST_CHEBYSHEV>(lanes_per_row, vec);
  }
  return nullptr;
}

ThreadRowKernel get_subwarp_kernel(int cost, int lanes_per_row, bool vec) {
  switch (cost) {
    case MB2_COST_ROSEN: return sub_lpr<COST_ROSEN>(lanes_per_row, vec);
    case MB2_COST_CORANA: return sub_lpr<COST_CORANA>(lanes_per_row, vec);
    case MB2_COST_GRIEWANGK: return sub_lpr<COST_GRIEWANGK>(lanes_per_row, vec);
    case MB2_COST_ZIMMERMANN: return sub_lpr<COST_ZIMMERMANN>(lanes_per_row, vec);
    case MB2_COST_CHEBYSHEV: return sub_lpr<COST_CHEBYSHEV>(lanes_per_row, vec);
  }
  return nullptr;
}
}  // namespace mb2
