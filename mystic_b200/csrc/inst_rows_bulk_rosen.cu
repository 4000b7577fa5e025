// Kernel instantiations, split over translation units so the build runs in parallel.
#include "../../include/mystic_b200.h"
#include "de_rows_bulk.cuh"

namespace mb2 {
typedef void (*BulkRowsKernel)(const StepParams, int, const BulkRowsLayout);

template <int LPR>
static BulkRowsKernel rows_bulk_k(int k) {
  switch (k) {
    case 2: return (BulkRowsKernel)de_step_rows_bulk_kernel<COST_ROSEN, LPR, 2>;
    case 3: return (BulkRowsKernel)de_step_rows_bulk_kernel<COST_ROSEN, LPR, 3>;
    case 4: return (BulkRowsKernel)de_step_rows_bulk_kernel<COST_ROSEN, LPR, 4>;
    case 5: return (BulkRowsKernel)de_step_rows_bulk_kernel<COST_ROSEN, LPR, 5>;
  }
  return nullptr;
}

// K2f: short rows, exponential crossover, every byte moved by cp.async.bulk (launch: base, shared-memory layout)
BulkRowsKernel get_rows_bulk_kernel_rosen(int lanes_per_row, int ndonors) {
  switch (lanes_per_row) {
    case 4: return rows_bulk_k<4>(ndonors);
    case 8: return rows_bulk_k<8>(ndonors);
  }
  return nullptr;
}
}  // namespace mb2
