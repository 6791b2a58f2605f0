// gqe_launch.h -- host-callable launchers of the kernels, one translation unit
// per embedding dimension behind them.
#pragma once
#include "gqe_params.h"

namespace gqe {

#define GQE_DECLARE_DIM(D)                                                                                      \
  cudaError_t launch_fused_simt_d##D(int structure, const LaunchParams& lp, int64_t grid, cudaStream_t st);     \
  cudaError_t launch_op_simt_d##D(const OpParams& op, cudaStream_t st);
GQE_DECLARE_DIM(32)
GQE_DECLARE_DIM(64)
GQE_DECLARE_DIM(128)
GQE_DECLARE_DIM(256)
#undef GQE_DECLARE_DIM

// structure < 0 selects the grouped kernel
inline cudaError_t launch_fused_simt(int d, int structure, const LaunchParams& lp, int64_t grid, cudaStream_t st) {
  switch (d) {
    case 32: return launch_fused_simt_d32(structure, lp, grid, st);
    case 64: return launch_fused_simt_d64(structure, lp, grid, st);
    case 128: return launch_fused_simt_d128(structure, lp, grid, st);
    case 256: return launch_fused_simt_d256(structure, lp, grid, st);
    default: return cudaErrorInvalidValue;
  }
}
inline cudaError_t launch_op_simt(int d, const OpParams& op, cudaStream_t st) {
  switch (d) {
    case 32: return launch_op_simt_d32(op, st);
    case 64: return launch_op_simt_d64(op, st);
    case 128: return launch_op_simt_d128(op, st);
    case 256: return launch_op_simt_d256(op, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace gqe
