// gqe_rows.cu -- the small data movers around the scoring kernels.
//
//   gqe_gather_rows    raw embedding-row gather for the staged sharded-table exchange: the owner of a
//                      node-type shard copies the requested rows of its table, un-normalised, into a dense
//                      [n, d] block that is then shipped to the scoring rank (NCCL) -- the comparison point of
//                      reading the rows over NVLink inside the fused kernel.  Pure HBM traffic: n * d * 4 bytes
//                      read + written, 128-bit accesses, one warp per row, grid sized to the SM count.
//   gqe_fetch_indices  index arrays of a *_host call: pinned host memory -> device staging over PCIe, every
//                      range in one launch (instead of one copy-engine operation per range).
//   gqe_store_batch    batches out of a device-resident query store: slices gathered, one negative drawn per
//                      query (netquery/train_helpers.py:95-107, netquery/model.py:113-120).
#include <cuda_runtime.h>
#include <stdint.h>

#include "gqe_launch.h"
#include "gqe_ptx.cuh"

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

// ---- index arrays of a *_host call ------------------------------------------------------------------
// The caller's int32 arrays live in pinned host memory.  A copy-engine transfer costs ~6-10 us of fixed
// latency EACH (three anchor slots + the targets = four of them in front of a 125 us kernel); here the SMs
// read all ranges over PCIe in one launch instead (16 bytes per thread per request, every thread's requests
// in flight together), and the scoring kernel -- launched programmatically dependent on this one -- sets
// itself up meanwhile (LaunchParams::wait_prev).  blockIdx.y = range.
__global__ void __launch_bounds__(256) gqe_fetch_indices(const __grid_constant__ FetchParams p) {
  ptx::griddep_launch_dependents();
  const int r = blockIdx.y;
  const int32_t* __restrict__ src = p.src[r];
  int32_t* __restrict__ dst = p.dst[r];
  const int64_t n = p.n[r];
  const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
  const uintptr_t sa = reinterpret_cast<uintptr_t>(src), da = reinterpret_cast<uintptr_t>(dst);
  if (((sa ^ da) & 15u) != 0) {     // differently aligned: word by word
    for (int64_t i = tid; i < n; i += nth) dst[i] = __ldcv(src + i);
    return;
  }
  int64_t head = (int64_t)(((16u - (unsigned)(da & 15u)) & 15u) >> 2);
  if (head > n) head = n;
  const int64_t n4 = (n - head) >> 2, tail = head + 4 * n4;
  if (tid < head) dst[tid] = __ldcv(src + tid);
  if (tid < n - tail) dst[tail + tid] = __ldcv(src + tail + tid);
  const int4* __restrict__ s4 = reinterpret_cast<const int4*>(src + head);
  int4* __restrict__ d4 = reinterpret_cast<int4*>(dst + head);
  int64_t i = tid;
  for (; i + 3 * nth < n4; i += 4 * nth) {     // four requests in flight per thread
    const int4 a = __ldcv(s4 + i), b = __ldcv(s4 + i + nth), c = __ldcv(s4 + i + 2 * nth), d = __ldcv(s4 + i + 3 * nth);
    d4[i] = a; d4[i + nth] = b; d4[i + 2 * nth] = c; d4[i + 3 * nth] = d;
  }
  for (; i < n4; i += nth) d4[i] = __ldcv(s4 + i);
}

cudaError_t launch_fetch_indices(const FetchParams& fp, cudaStream_t st) {
  if (fp.count <= 0) return cudaSuccess;
  int64_t most = 0;
  for (int i = 0; i < fp.count; ++i) most = fp.n[i] > most ? fp.n[i] : most;
  // 16 bytes per thread and pass; a range of 256 K words (1 MiB) gets 64 blocks
  int64_t bx = (most / 4 + 1023) / 1024;
  bx = bx < 1 ? 1 : (bx > 96 ? 96 : bx);
  gqe_fetch_indices<<<dim3((unsigned)bx, (unsigned)fp.count), 256, 0, st>>>(fp);
  return cudaGetLastError();
}

// ---- batches out of a device-resident query store ------------------------------------------------
// blockIdx.y = slice.  Query i of a slice: its anchors and its positive target are copied into the call's
// concatenated index arrays, its negative is element floor(u * len) of its stored list, u = the upper 32 bits
// of splitmix64(seed, position of the query in the call) / 2^32 (bias len / 2^32).  The reference draws with
// random.choice per query (model.py:116-120): the same uniform distribution over the same lists.
__device__ __forceinline__ uint32_t draw32(unsigned long long seed, unsigned long long pos) {
  unsigned long long z = seed + 0x9E3779B97F4A7C15ull * (pos + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (uint32_t)(z >> 32);
}
__global__ void __launch_bounds__(256) gqe_store_batch(const __grid_constant__ StoreBatchParams p) {
  const StoreSliceDev& s = p.s[blockIdx.y];
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < s.n; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t q = s.start + i, out = s.out_q0 + i;
    for (int k = 0; k < s.n_anchor; ++k) p.anchors_out[(int64_t)k * p.nq_total + out] = __ldg(s.anchors + (int64_t)k * s.block_q + q);
    const int32_t pos = __ldg(s.targets + q);
    const uint32_t u = draw32(p.seed, (unsigned long long)out);
    int32_t neg = pos;
    if (s.neg_ptr) {
      const int64_t lo = __ldg(s.neg_ptr + q), len = __ldg(s.neg_ptr + q + 1) - lo;
      if (len > 0) neg = __ldg(s.negs + lo + (int64_t)(((unsigned long long)u * (unsigned long long)len) >> 32));
      else report_index(p.err, IDX_ERR_NO_NEGATIVE, s.tgt_mode, (int32_t)out);
    } else if (s.pool_n > 0) {
      neg = __ldg(s.negs + (int64_t)(((unsigned long long)u * (unsigned long long)s.pool_n) >> 32));
    } else {
      report_index(p.err, IDX_ERR_NO_NEGATIVE, s.tgt_mode, (int32_t)out);
    }
    reinterpret_cast<int2*>(p.pairs_out)[out] = make_int2(pos, neg);
  }
}

cudaError_t launch_store_batch(const StoreBatchParams& sp, int64_t max_n, cudaStream_t st) {
  if (sp.count <= 0 || max_n <= 0) return cudaSuccess;
  int64_t bx = (max_n + 255) / 256;
  bx = bx > 64 ? 64 : bx;
  gqe_store_batch<<<dim3((unsigned)bx, (unsigned)sp.count), 256, 0, st>>>(sp);
  return cudaGetLastError();
}

}  // namespace gqe
