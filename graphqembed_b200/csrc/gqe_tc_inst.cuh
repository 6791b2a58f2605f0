// gqe_tc_inst.cuh -- instantiates the tensor-core kernels for ONE embedding
// dimension (GQE_DIM = 128 or 256).
#pragma once
#include <cstdlib>

#include "gqe_launch.h"
#include "gqe_tc.cuh"

namespace gqe {

static int tc_current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev & 63;
}

// CTA-pair launch: clusters of two, as many as fit the device at once (tiles are dealt round-robin)
template <int D, int STRUCT>
static cudaError_t tc_launch_pair(const LaunchParams& lp, int64_t tiles, int slots, cudaStream_t st) {
  static bool configured[64] = {false};
  static int clusters[64] = {0};
  auto kern = tc::gqe_fused_tc<D, STRUCT, true>;
  const int dev = tc_current_device();
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(tc::Cfg<D>::kThreads);
  cfg.dynamicSmemBytes = tc::Cfg<D>::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  static const int pdl = [] { const char* e = std::getenv("GQE_PDL"); return e ? std::atoi(e) : 1; }();
  attr[1].val.programmaticStreamSerializationAllowed = pdl;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<D>::kSmemBytes);
    if (e != cudaSuccess) return e;
    cfg.gridDim = dim3((unsigned)slots);
    int n = 0;
    cfg.numAttrs = 1;                       // (the occupancy query looks at the cluster shape only)
    e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    cfg.numAttrs = 2;
    if (e != cudaSuccess) return e;
    if (n <= 0) return cudaErrorInvalidConfiguration;
    clusters[dev] = n < slots / 2 ? n : slots / 2;
    configured[dev] = true;
  }
  const int64_t pairs = tiles / 2;          // (tiles is even: every segment was padded)
  const int64_t nc = pairs < clusters[dev] ? pairs : clusters[dev];
  cfg.gridDim = dim3((unsigned)(2 * nc));
  return cudaLaunchKernelEx(&cfg, kern, lp);
}

// Persistent launch: one CTA per SM slot (SMs x CTAs/SM), never more CTAs than tiles.
template <int D, int STRUCT>
static cudaError_t tc_launch_one(const LaunchParams& lp, int64_t tiles, cudaStream_t st) {
  static bool configured[64] = {false};
  static int slots[64] = {0};
  auto kern = tc::gqe_fused_tc<D, STRUCT>;
  const int dev = tc_current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<D>::kSmemBytes);
    if (e != cudaSuccess) return e;
    int sms = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    slots[dev] = sms * tc::Cfg<D>::kCtasPerSm;
    configured[dev] = true;
  }
  if (lp.pair == 1) return tc_launch_pair<D, STRUCT>(lp, tiles, slots[dev], st);
  const int64_t grid = tiles < slots[dev] ? tiles : slots[dev];
  if constexpr (D == 256 && STRUCT < 0) {
    // node-type-sharded tables: the instantiation with two helper warps per CTA that stage remote rows
    if (lp.stage_on) {
      auto skern = tc::gqe_fused_tc<D, STRUCT, false, true>;
      static bool sconf[64] = {false};
      if (!sconf[dev]) {
        cudaError_t e = cudaFuncSetAttribute(skern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::Cfg<D>::kSmemBytes);
        if (e != cudaSuccess) return e;
        sconf[dev] = true;
      }
      cudaLaunchConfig_t scfg = {};
      scfg.gridDim = dim3((unsigned)grid);
      scfg.blockDim = dim3(tc::Cfg<D>::kThreads + 32 * tc::Cfg<D>::kHelperWarps);
      scfg.dynamicSmemBytes = tc::Cfg<D>::kSmemBytes;
      scfg.stream = st;
      cudaLaunchAttribute sattr[1];
      sattr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      sattr[0].val.programmaticStreamSerializationAllowed = 1;
      scfg.attrs = sattr;
      scfg.numAttrs = 1;
      return cudaLaunchKernelEx(&scfg, skern, lp);
    }
  }
  // programmatic dependent launch on gqe_pack (see tc::producer); GQE_PDL=0 in the environment
  // turns it off (diagnostics: measured 2.6 us per call on the benchmark mix)
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(tc::Cfg<D>::kThreads);
  cfg.dynamicSmemBytes = tc::Cfg<D>::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  static const int pdl = [] { const char* e = std::getenv("GQE_PDL"); return e ? std::atoi(e) : 1; }();
  attr[0].val.programmaticStreamSerializationAllowed = pdl;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, lp);
}

template <int D>
static cudaError_t tc_launch_struct(int structure, const LaunchParams& lp, int64_t grid, cudaStream_t st) {
  switch (structure) {
    case GQE_CHAIN1: return tc_launch_one<D, GQE_CHAIN1>(lp, grid, st);
    case GQE_CHAIN2: return tc_launch_one<D, GQE_CHAIN2>(lp, grid, st);
    case GQE_CHAIN3: return tc_launch_one<D, GQE_CHAIN3>(lp, grid, st);
    case GQE_INTER2: return tc_launch_one<D, GQE_INTER2>(lp, grid, st);
    case GQE_INTER3: return tc_launch_one<D, GQE_INTER3>(lp, grid, st);
    case GQE_INTER_CHAIN3: return tc_launch_one<D, GQE_INTER_CHAIN3>(lp, grid, st);
    case GQE_CHAIN_INTER3: return tc_launch_one<D, GQE_CHAIN_INTER3>(lp, grid, st);
    default: return tc_launch_one<D, -1>(lp, grid, st);
  }
}

#if GQE_DIM == 256
int score_col_src_host(int n) { return tc::score_col_src(n); }
int tc_use_pair(int d, int64_t tiles) {   // experiment, see tc::producer<PAIR>: only with GQE_PAIR=1 in the environment
  const char* e = std::getenv("GQE_PAIR");
  return (e && std::atoi(e) == 1 && d == 256 && tiles >= 2 * 148) ? 1 : 0;
}
#endif

#define GQE_CAT2(a, b) a##b
#define GQE_CAT(a, b) GQE_CAT2(a, b)
cudaError_t GQE_CAT(launch_fused_tc_d, GQE_DIM)(int structure, const LaunchParams& lp, int64_t grid, cudaStream_t st) {
  return tc_launch_struct<GQE_DIM>(structure, lp, grid, st);
}
cudaError_t GQE_CAT(launch_pack_d, GQE_DIM)(const PackParams& pp, int n_entries, cudaStream_t st) {
  if (n_entries <= 0) return cudaSuccess;
  const dim3 grid((GQE_DIM * GQE_DIM / 8 + 255) / 256, (unsigned)n_entries);
  tc::gqe_pack<GQE_DIM><<<grid, 256, 0, st>>>(pp);
  return cudaGetLastError();
}

}  // namespace gqe
