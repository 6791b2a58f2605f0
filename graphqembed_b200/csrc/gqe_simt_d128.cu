// exact-fp32 kernels, embedding dimension 128
#define GQE_DIM 128
#include "gqe_simt_inst.cuh"
