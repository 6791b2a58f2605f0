// tensor-core kernels, embedding dimension 128
#define GQE_DIM 128
#include "gqe_tc_inst.cuh"
