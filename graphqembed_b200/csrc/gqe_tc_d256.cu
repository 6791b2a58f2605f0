// tensor-core kernels, embedding dimension 256
#define GQE_DIM 256
#include "gqe_tc_inst.cuh"
