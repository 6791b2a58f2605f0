// gqe_simt_inst.cuh -- instantiates the exact-fp32 kernels for ONE embedding
// dimension (GQE_DIM) so that the dimensions compile in parallel.
#pragma once
#include "gqe_launch.h"
#include "gqe_simt.cuh"

namespace gqe {

// cudaFuncSetAttribute is per device: remember which devices were configured.
static int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev & 63;
}

template <int D, int STRUCT>
static cudaError_t launch_one(const LaunchParams& lp, int64_t grid, cudaStream_t st) {
  static bool configured[64] = {false};
  auto kern = gqe_fused_simt<D, STRUCT>;
  const int dev = current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem<D>));
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  kern<<<(unsigned)grid, kThreads, sizeof(TileSmem<D>), st>>>(lp);
  return cudaGetLastError();
}

template <int D>
static cudaError_t launch_struct(int structure, const LaunchParams& lp, int64_t grid, cudaStream_t st) {
  switch (structure) {
    case GQE_CHAIN1: return launch_one<D, GQE_CHAIN1>(lp, grid, st);
    case GQE_CHAIN2: return launch_one<D, GQE_CHAIN2>(lp, grid, st);
    case GQE_CHAIN3: return launch_one<D, GQE_CHAIN3>(lp, grid, st);
    case GQE_INTER2: return launch_one<D, GQE_INTER2>(lp, grid, st);
    case GQE_INTER3: return launch_one<D, GQE_INTER3>(lp, grid, st);
    case GQE_INTER_CHAIN3: return launch_one<D, GQE_INTER_CHAIN3>(lp, grid, st);
    case GQE_CHAIN_INTER3: return launch_one<D, GQE_CHAIN_INTER3>(lp, grid, st);
    default: return launch_one<D, -1>(lp, grid, st);
  }
}

template <int D>
static cudaError_t launch_op_d(const OpParams& op, cudaStream_t st) {
  static bool configured[64] = {false};
  auto kern = gqe_op_simt<D>;
  const int dev = current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileSmem<D>));
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  const int64_t grid = (op.n + kTileRows - 1) / kTileRows;
  kern<<<(unsigned)grid, kThreads, sizeof(TileSmem<D>), st>>>(op);
  return cudaGetLastError();
}


#define GQE_CAT2(a, b) a##b
#define GQE_CAT(a, b) GQE_CAT2(a, b)
cudaError_t GQE_CAT(launch_fused_simt_d, GQE_DIM)(int structure, const LaunchParams& lp, int64_t grid, cudaStream_t st) {
  return launch_struct<GQE_DIM>(structure, lp, grid, st);
}
cudaError_t GQE_CAT(launch_op_simt_d, GQE_DIM)(const OpParams& op, cudaStream_t st) { return launch_op_d<GQE_DIM>(op, st); }

}  // namespace gqe
