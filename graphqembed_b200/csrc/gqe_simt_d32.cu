// exact-fp32 kernels, embedding dimension 32
#define GQE_DIM 32
#include "gqe_simt_inst.cuh"
