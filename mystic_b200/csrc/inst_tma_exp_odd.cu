// K2e for rows of odd length (16-byte aligned covers addressed by per-row phases, de_tma_exp.cuh): the same
// instantiation list as inst_tma_exp.cu with ODD = true, in its own translation unit so the two compile in parallel.
#define MB2_EXP_TMA_ODD true
#define MB2_EXP_TMA_GETTER get_exp_tma_kernel_odd
#include "inst_tma_exp.cu"
