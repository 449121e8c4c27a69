// EP kernels for (N, k) = (1024, 2): params_sqrd_lvl_1 / lvl_4 (pbs_level 2) and lvl_256 (pbs_level 4) — the reference's test-only sets
#define TAC_N 1024
#define TAC_K 2
#define TAC_SHAPE_FN shape_ops_n1024_k2
#define TAC_PBS_LEVELS(X) X(2) X(4)
#include "kernels_shape.inl"
