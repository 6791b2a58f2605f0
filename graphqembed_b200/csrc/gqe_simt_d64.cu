// exact-fp32 kernels, embedding dimension 64
#define GQE_DIM 64
#include "gqe_simt_inst.cuh"
