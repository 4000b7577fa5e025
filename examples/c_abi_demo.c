/* The drop-in boundary from plain C (no Python, no torch): Rosenbrock 1024-D, NP = 4096, Best1Bin with a tight
 * strict range -- SetRandomInitialPoints / SetStrictRanges / Step of the reference's solver as the calls of
 * include/mystic_b200.h.  Build and run (from the repository root):
 *   gcc -std=c99 -O2 examples/c_abi_demo.c -Iinclude -I/usr/local/cuda/include -Lmystic_b200/_lib -lmystic_b200 \
 *       -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/mystic_b200/_lib -o /tmp/c_abi_demo && /tmp/c_abi_demo */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>

#include "mystic_b200.h"

#define CHECK(call)                                                                  \
  do {                                                                               \
    int rc_ = (call);                                                                \
    if (rc_ != 0) {                                                                  \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, mb2_last_error());        \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

int main(void) {
  const int D = 1024;
  const long long NP = 4096;
  double *pop_a, *pop_b, *e_a, *e_b;
  if (cudaMalloc((void**)&pop_a, sizeof(double) * NP * D) || cudaMalloc((void**)&pop_b, sizeof(double) * NP * D) ||
      cudaMalloc((void**)&e_a, sizeof(double) * NP) || cudaMalloc((void**)&e_b, sizeof(double) * NP)) {
    fprintf(stderr, "cudaMalloc failed\n");
    return 1;
  }
  double* lo = (double*)malloc(sizeof(double) * D);
  double* hi = (double*)malloc(sizeof(double) * D);
  double* best = (double*)malloc(sizeof(double) * D);
  for (int i = 0; i < D; ++i) {
    lo[i] = -5.0;
    hi[i] = 5.0;
  }
  mb2_ctx* ctx = NULL;
  CHECK(mb2_create(0, NP, D, NP, 0, &ctx));                     /* one shard = the whole population */
  CHECK(mb2_bind_buffers(ctx, pop_a, pop_b, e_a, e_b));         /* caller-owned device memory */
  CHECK(mb2_set_cost(ctx, MB2_COST_ROSEN, NULL, 0));
  CHECK(mb2_set_bounds(ctx, MB2_BOUNDS_CLAMP, lo, hi));         /* SetStrictRanges(lo, hi, tight=True) */
  CHECK(mb2_set_strategy(ctx, MB2_BEST1BIN, 0.8, 0.9));         /* ScalingFactor, CrossProbability */
  CHECK(mb2_set_rng_philox(ctx, 0x5EEDull));
  CHECK(mb2_init_population(ctx, 0x5EEDull, lo, hi, NULL));     /* SetRandomInitialPoints on the device */
  CHECK(mb2_eval_initial(ctx, NULL));                           /* generation 0 */
  mb2_result r;
  CHECK(mb2_read_result(ctx, &r, NULL, NULL));
  const double e0 = r.best_energy;
  for (int g = 1; g <= 200; ++g) CHECK(mb2_step(ctx, NULL));    /* asynchronous on the (default) stream */
  CHECK(mb2_read_result(ctx, &r, best, NULL));                  /* synchronises */
  printf("abi %d: generation %lld, bestEnergy %.6g -> %.6g, best member %lld, x[0..2] = %.4f %.4f %.4f, kernel family %d\n",
         mb2_abi_version(), (long long)r.generation, e0, r.best_energy, (long long)r.best_index, best[0], best[1], best[2],
         mb2_kernel_family(ctx));
  if (mb2_set_cost(ctx, 99, NULL, 0) == 0) return 1;            /* errors are return codes + a message */
  printf("expected error: %s\n", mb2_last_error());
  CHECK(mb2_destroy(ctx));
  cudaFree(pop_a); cudaFree(pop_b); cudaFree(e_a); cudaFree(e_b);
  free(lo); free(hi); free(best);
  return !(r.generation == 200 && r.best_energy < e0);
}
