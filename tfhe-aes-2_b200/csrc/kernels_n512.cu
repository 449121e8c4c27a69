// EP kernels for (N, k) = (512, 4): params_sqrd_lvl_64, the shipped set (pbs_level 3)
#define TAC_N 512
#define TAC_K 4
#define TAC_SHAPE_FN shape_ops_n512_k4
#define TAC_PBS_LEVELS(X) X(3)
#include "kernels_shape.inl"
