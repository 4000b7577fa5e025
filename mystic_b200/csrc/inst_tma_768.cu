// Kernel instantiations, split over translation units so the build runs in parallel.
// One-warp groups in CTAs of 513..768 threads (85 registers per thread: the griewangk body spills at the 64 of the
// 1024-thread build).
#include "../../include/mystic_b200.h"
#include "de_tma.cuh"

namespace mb2 {
typedef void (*TmaKernel)(const StepParams, int);

TmaKernel get_tma_kernel_768(int cost, int gwarps) {
  if (gwarps != 1) return nullptr;
  switch (cost) {
    case MB2_COST_ROSEN: return de_step_best1bin_tma_kernel<COST_ROSEN, 1, 768>;
    case MB2_COST_GRIEWANGK: return de_step_best1bin_tma_kernel<COST_GRIEWANGK, 1, 768>;
  }
  return nullptr;
}
}  // namespace mb2
