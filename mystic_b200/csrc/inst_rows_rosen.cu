// Kernel instantiations, split over translation units so the build runs in parallel.
#include "../../include/mystic_b200.h"
#include "de_rows.cuh"

namespace mb2 {
typedef void (*ThreadRowKernel)(const StepParams, int, int, int);

template <int LPR>
static ThreadRowKernel rows_k(int k) {
  switch (k) {
    case 2: return (ThreadRowKernel)de_step_rows_kernel<COST_ROSEN, LPR, 2>;
    case 3: return (ThreadRowKernel)de_step_rows_kernel<COST_ROSEN, LPR, 3>;
    case 4: return (ThreadRowKernel)de_step_rows_kernel<COST_ROSEN, LPR, 4>;
    case 5: return (ThreadRowKernel)de_step_rows_kernel<COST_ROSEN, LPR, 5>;
  }
  return nullptr;
}

ThreadRowKernel get_rows_kernel_rosen(int lanes_per_row, int ndonors) {
  switch (lanes_per_row) {
    case 4: return rows_k<4>(ndonors);
    case 8: return rows_k<8>(ndonors);
  }
  return nullptr;
}
}  // namespace mb2
