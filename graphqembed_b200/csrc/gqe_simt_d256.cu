// exact-fp32 kernels, embedding dimension 256
#define GQE_DIM 256
#include "gqe_simt_inst.cuh"
