// Microbenchmark: the ceiling of a random ROW gather from local HBM (rows of D fp32, the embedding
// tables' access pattern) next to a streaming read of the same bytes -- what the gather-bound
// kernels (gqe_fused_vec, the gathers of gqe_fused_tc) can be held against.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hbm_gather hbm_gather.cu && ./hbm_gather [table_rows]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int D = 256;  // floats per row (1 KiB)

// one warp per row batch: U rows in flight per warp
template <int U>
__global__ void __launch_bounds__(256) gather(const float* __restrict__ table, const int32_t* __restrict__ rows, int64_t n,
                                              float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float acc = 0.f;
  for (int64_t r0 = warp * U; r0 < n; r0 += nw * U) {
    float4 v[U][2];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u < n ? r0 + u : n - 1;
      const float4* src = reinterpret_cast<const float4*>(table + (size_t)rows[r] * D);
#pragma unroll
      for (int j = 0; j < 2; ++j) v[u][j] = __ldg(src + lane + 32 * j);
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < 2; ++j) acc += v[u][j].x + v[u][j].y + v[u][j].z + v[u][j].w;
  }
  if (acc == 12345.678f) out[0] = acc;
}


// ---- the same gather through the TMA: every warp owns a ring of P slots of R rows in shared memory, lanes 0..R-1
// issue one 1 KiB cp.async.bulk per row (completion on the slot's mbarrier), the warp reads the rows back with
// LDS.128 -- the data path of the streaming kernel (gqe_vec.cu)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int R, int P>
__global__ void __launch_bounds__(1024, 1) gather_tma(const float* __restrict__ table, const int32_t* __restrict__ rows,
                                                      int64_t n, float* __restrict__ out) {
  extern __shared__ __align__(128) float slots[];   // [warps][P][R][D]
  __shared__ uint64_t bars[32][P];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * W + wib, nw = (int64_t)gridDim.x * W;
  if (lane == 0)
    for (int s = 0; s < P; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bars[wib][s])));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  float* my = slots + (size_t)wib * P * R * D;
  const int64_t items = (n / R - warp + nw - 1) / nw;   // this warp's row groups: warp, warp + nw, ...
  auto issue = [&](int64_t it) {
    if (it >= items) return;
    const int s = (int)(it % P);
    const uint32_t bar = smem_u32(&bars[wib][s]);
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(R * D * 4) : "memory");
    __syncwarp();
    if (lane < R) {
      const int64_t g = (warp + it * nw) * R + lane;
      const float* src = table + (size_t)rows[g] * D;
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_u32(my + ((size_t)s * R + lane) * D)),
                   "l"(src), "r"(D * 4), "r"(bar)
                   : "memory");
    }
  };
  for (int i = 0; i < P; ++i) issue(i);
  float acc = 0.f;
  for (int64_t it = 0; it < items; ++it) {
    const int s = (int)(it % P);
    const uint32_t bar = smem_u32(&bars[wib][s]), parity = (uint32_t)((it / P) & 1);
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    const float4* src = reinterpret_cast<const float4*>(my + (size_t)s * R * D);
#pragma unroll
    for (int u = 0; u < R * D / 128; ++u) { const float4 v = src[lane + 32 * u]; acc += v.x + v.y + v.z + v.w; }
    __syncwarp();
    issue(it + P);
  }
  if (acc == 12345.678f) out[0] = acc;
}

// and through cp.async (LDGSTS, 16 bytes per lane)
template <int R, int P>
__global__ void __launch_bounds__(1024, 1) gather_ldgsts(const float* __restrict__ table, const int32_t* __restrict__ rows,
                                                         int64_t n, float* __restrict__ out) {
  extern __shared__ __align__(128) float slots[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, W = blockDim.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * W + wib, nw = (int64_t)gridDim.x * W;
  float* my = slots + (size_t)wib * P * R * D;
  const int64_t items = (n / R - warp + nw - 1) / nw;
  auto issue = [&](int64_t it) {
    if (it < items) {
      const int s = (int)(it % P);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float* src = table + (size_t)rows[(warp + it * nw) * R + r] * D;
#pragma unroll
        for (int j = 0; j < 2; ++j)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(my + ((size_t)s * R + r) * D + 4 * (lane + 32 * j))),
                       "l"(src + 4 * (lane + 32 * j)) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  for (int i = 0; i < P; ++i) issue(i);
  float acc = 0.f;
  for (int64_t it = 0; it < items; ++it) {
    const int s = (int)(it % P);
    asm volatile("cp.async.wait_group %0;" ::"n"(P - 1) : "memory");
    __syncwarp();
    const float4* src = reinterpret_cast<const float4*>(my + (size_t)s * R * D);
#pragma unroll
    for (int u = 0; u < R * D / 128; ++u) { const float4 v = src[lane + 32 * u]; acc += v.x + v.y + v.z + v.w; }
    __syncwarp();
    issue(it + P);
  }
  if (acc == 12345.678f) out[0] = acc;
}

__global__ void __launch_bounds__(256) stream_read(const float4* __restrict__ src, int64_t n4, float* __restrict__ out) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(src + i);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 12345.678f) out[0] = acc;
}

int main(int argc, char** argv) {
  const int64_t table_rows = argc > 1 ? atoll(argv[1]) : 10000000;   // 10 M rows x 1 KiB = 10 GB
  const int64_t n = 4 << 20;                                          // rows gathered per launch (4 GiB)
  float* table; float* out; int32_t* rows; float* flush;
  CK(cudaMalloc(&table, (size_t)table_rows * D * 4));
  CK(cudaMemset(table, 0, (size_t)table_rows * D * 4));
  CK(cudaMalloc(&out, 4));
  CK(cudaMalloc(&flush, 256 << 20));
  std::vector<int32_t> h(n);
  uint64_t s = 88172645463325252ull;
  for (auto& x : h) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; x = (int32_t)(s % (uint64_t)table_rows); }
  CK(cudaMalloc(&rows, n * 4));
  CK(cudaMemcpy(rows, h.data(), n * 4, cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto time = [&](auto launch, const char* name, double bytes) {
    float best = 1e30f;
    for (int it = 0; it < 6; ++it) {
      cudaMemsetAsync(flush, it, 256 << 20);
      cudaEventRecord(e0); launch(); cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (it > 0 && ms < best) best = ms;
    }
    printf("%-44s %8.3f ms  %8.1f GB/s\n", name, best, bytes / best / 1e6);
  };
  const double gbytes = (double)n * D * 4;
  printf("table %lld rows x %d B = %.1f GB; %lld random rows per launch\n", (long long)table_rows, D * 4,
         table_rows * D * 4 / 1e9, (long long)n);
  for (int bps : {2, 4, 8}) {
    char name[96];
    snprintf(name, sizeof name, "gather 1 row/warp in flight, %d CTAs/SM", bps);
    time([&] { gather<1><<<148 * bps, 256>>>(table, rows, n, out); }, name, gbytes);
    snprintf(name, sizeof name, "gather 2 rows/warp in flight, %d CTAs/SM", bps);
    time([&] { gather<2><<<148 * bps, 256>>>(table, rows, n, out); }, name, gbytes);
    snprintf(name, sizeof name, "gather 4 rows/warp in flight, %d CTAs/SM", bps);
    time([&] { gather<4><<<148 * bps, 256>>>(table, rows, n, out); }, name, gbytes);
    snprintf(name, sizeof name, "gather 8 rows/warp in flight, %d CTAs/SM", bps);
    time([&] { gather<8><<<148 * bps, 256>>>(table, rows, n, out); }, name, gbytes);
  }

  {
    auto run = [&](auto kern, const char* what, int warps, int R, int P) {
      const size_t smem = (size_t)warps * P * R * D * 4;
      CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      char name[96];
      snprintf(name, sizeof name, "%s: %d warps/SM, %d slots x %d rows", what, warps, P, R);
      time([&] { kern<<<148, warps * 32, smem>>>(table, rows, n, out); CK(cudaGetLastError()); }, name, gbytes);
    };
    for (int warps : {8, 12, 16}) run(gather_tma<4, 3>, "TMA bulk 1 KiB rows", warps, 4, 3);
    run(gather_tma<4, 2>, "TMA bulk 1 KiB rows", 16, 4, 2);
    run(gather_tma<4, 4>, "TMA bulk 1 KiB rows", 12, 4, 4);
    run(gather_tma<4, 6>, "TMA bulk 1 KiB rows", 8, 4, 6);
    for (int warps : {8, 12, 16}) run(gather_ldgsts<4, 3>, "cp.async 16 B/lane", warps, 4, 3);
  }
  const int64_t n4 = std::min<int64_t>((int64_t)table_rows * D / 4, (int64_t)1 << 28);
  time([&] { stream_read<<<148 * 8, 256>>>(reinterpret_cast<const float4*>(table), n4, out); }, "streaming read (4 GiB)", n4 * 16.0);
  return 0;
}
