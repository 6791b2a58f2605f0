// gqe_rows.cu -- raw embedding-row gather for the sharded-table exchange.
//
// The owner of a node-type shard copies the requested rows of its table,
// un-normalised, into a dense [n, d] block that is then shipped to the scoring
// rank (NCCL) -- the staged alternative to reading the rows in place over
// NVLink.  Pure HBM traffic: n * d * 4 bytes read + written, 128-bit accesses,
// one warp per row, grid sized to the SM count.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gqe_launch.h"

namespace gqe {

__global__ void __launch_bounds__(256) gqe_gather_rows(const float* __restrict__ table, const int32_t* __restrict__ rows,
                                                       int64_t n, int d4, float* __restrict__ out, int64_t table_rows,
                                                       unsigned long long* err) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r0 = warp0 * 4; r0 < n; r0 += n_warps * 4) {
    // 4 rows in flight per warp
    const float4* src[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t r = r0 + u;
      int64_t row = r < n ? (int64_t)__ldg(rows + r) : 0;
      if (row < 0 || row >= table_rows) {   // a request outside the shard reads row 0 and is reported
        report_index(err, IDX_ERR_ROW_RANGE, 0, (int32_t)row);
        row = 0;
      }
      src[u] = r < n ? reinterpret_cast<const float4*>(table) + (size_t)row * d4 : nullptr;
    }
    for (int c = lane; c < d4; c += 32) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (src[u]) v[u] = __ldg(src[u] + c);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (src[u]) reinterpret_cast<float4*>(out)[(size_t)(r0 + u) * d4 + c] = v[u];
    }
  }
}

cudaError_t launch_gather_rows(const float* table, const int32_t* rows, int64_t n, int d, float* out, int64_t table_rows,
                               unsigned long long* err, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (n + 31) / 32;  // 8 warps x 4 rows per CTA pass
  const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
  gqe_gather_rows<<<grid, 256, 0, st>>>(table, rows, n, d / 4, out, table_rows, err);
  return cudaGetLastError();
}

}  // namespace gqe
