// Microbenchmark: random 1 KB-row gather from PEER memory over NVLink, by access mechanism.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o peer_gather peer_gather.cu && ./peer_gather
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

constexpr int D = 256;  // floats per row

template <int MODE, int U>
__global__ void __launch_bounds__(512) gather(const float* __restrict__ table, const int32_t* __restrict__ rows, int64_t n,
                                              float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  float acc = 0.f;
  for (int64_t r0 = warp * U; r0 < n; r0 += nw * U) {
    float4 v[U][2];
    if (MODE == 2) {
      if (lane < U && r0 + lane < n) {
        const float* p = table + (size_t)rows[r0 + lane] * D;
        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(D * 4) : "memory");
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t r = r0 + u < n ? r0 + u : n - 1;
      const float4* src = reinterpret_cast<const float4*>(table + (size_t)rows[r] * D);
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (MODE == 1) {
          asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u][j].x), "=f"(v[u][j].y), "=f"(v[u][j].z), "=f"(v[u][j].w) : "l"(src + lane + 32 * j));
        } else if (MODE == 3) {
          asm volatile("ld.global.relaxed.sys.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[u][j].x), "=f"(v[u][j].y), "=f"(v[u][j].z), "=f"(v[u][j].w) : "l"(src + lane + 32 * j));
        } else {
          v[u][j] = __ldg(src + lane + 32 * j);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int j = 0; j < 2; ++j) acc += v[u][j].x + v[u][j].y + v[u][j].z + v[u][j].w;
  }
  if (acc == 12345.678f) out[0] = acc;
}

// TMA bulk copy global(peer) -> smem, one row per copy, 16 rows in flight per CTA
__global__ void __launch_bounds__(128) gather_tma(const float* __restrict__ table, const int32_t* __restrict__ rows, int64_t n,
                                                  float* __restrict__ out, int rows_per_cta) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ uint64_t bar;
  const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  uint32_t phase = 0;
  float acc = 0.f;
  for (int64_t r0 = (int64_t)blockIdx.x * rows_per_cta; r0 < n; r0 += (int64_t)gridDim.x * rows_per_cta) {
    const int cnt = (int)((n - r0) < rows_per_cta ? (n - r0) : rows_per_cta);
    if (threadIdx.x == 0)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(cnt * D * 4));
    __syncthreads();
    if (threadIdx.x < cnt) {
      const float* src = table + (size_t)rows[r0 + threadIdx.x] * D;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(sm + threadIdx.x * D * 4);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                   "l"(src), "r"(D * 4), "r"(bar_a) : "memory");
    }
    uint32_t done = 0;
    while (!done)
      asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar_a), "r"(phase));
    phase ^= 1;
    acc += reinterpret_cast<float*>(sm)[threadIdx.x];
    __syncthreads();
  }
  if (acc == 12345.678f) out[0] = acc;
}

int main() {
  int nd = 0;
  CK(cudaGetDeviceCount(&nd));
  printf("devices: %d\n", nd);
  const int64_t table_rows = 1250002, n = 240000;
  std::vector<int32_t> h(n);
  srand(1);
  for (auto& x : h) x = (int32_t)(((int64_t)rand() * 32768 + rand()) % table_rows);
  for (int owner = 0; owner < (nd > 1 ? 2 : 1); ++owner) {
    float* table;
    CK(cudaSetDevice(owner));
    CK(cudaMalloc(&table, table_rows * D * 4));
    CK(cudaMemset(table, 0, table_rows * D * 4));
    CK(cudaSetDevice(0));
    if (owner != 0) {
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, 0, owner));
      printf("can access peer: %d\n", can);
      CK(cudaDeviceEnablePeerAccess(owner, 0));
    }
    int32_t* rows;
    float* out;
    CK(cudaMalloc(&rows, n * 4));
    CK(cudaMalloc(&out, 1024));
    CK(cudaMemcpy(rows, h.data(), n * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    auto time = [&](const char* name, auto launch) {
      launch();
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      for (int i = 0; i < 5; ++i) launch();
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      ms /= 5;
      printf("%s table on gpu%d: %-28s %8.3f ms  %8.1f GB/s\n", owner ? "PEER " : "LOCAL", owner, name, ms, n * D * 4 / ms / 1e6);
    };
    time("ldg.nc U=4 148x512", [&] { gather<0, 4><<<148, 512>>>(table, rows, n, out); });
    time("ldg.nc U=8 296x512", [&] { gather<0, 8><<<296, 512>>>(table, rows, n, out); });
    time("ld.global U=4 148x512", [&] { gather<1, 4><<<148, 512>>>(table, rows, n, out); });
    time("ld.global U=8 592x512", [&] { gather<1, 8><<<592, 512>>>(table, rows, n, out); });
    time("ld.relaxed.sys U=4", [&] { gather<3, 4><<<148, 512>>>(table, rows, n, out); });
    time("prefetch.L2 + ldg U=4", [&] { gather<2, 4><<<148, 512>>>(table, rows, n, out); });
    CK(cudaFuncSetAttribute(gather_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    time("TMA bulk 64 rows/CTA x148", [&] { gather_tma<<<148, 128, 64 * 1024>>>(table, rows, n, out, 64); });
    time("TMA bulk 64 rows/CTA x444", [&] { gather_tma<<<444, 128, 64 * 1024>>>(table, rows, n, out, 64); });
    time("TMA bulk 128 rows/CTA x148", [&] { gather_tma<<<148, 128, 128 * 1024>>>(table, rows, n, out, 128); });
    CK(cudaFree(rows));
    CK(cudaFree(out));
  }
  return 0;
}
