// gqe_sage.cu -- the two device steps of the GraphSAGE-style Encoder (reference
// netquery/encoders.py:47-129 with MeanAggregator, netquery/aggregators.py:17-68; the --depth > 0
// path of netquery/utils.py:93-126).
//
//   k_segment_mean   out[i, :] = mean_{j in [ptr[i], ptr[i+1])} src[cols[j], :]
//                    -- MeanAggregator.forward after its neighbour sampling: the reference builds a
//                    dense [batch, unique neighbours] mask and multiplies it with the gathered feature
//                    matrix (aggregators.py:55-67); here the mask is the CSR it really is and the mean
//                    is a gather-reduce, one warp per output row, 128-bit loads.  HBM-gather-bound.
//   k_linear         out[M, n] = act(W[M, K] . X[K, n])
//                    -- the compress step, compress_params[mode].mm(combined) + ReLU
//                    (encoders.py:118-123): K = d * (1 + outgoing relation types) is not square, so
//                    the d x d kernels do not apply; 64 x 64 output tiles, fp32 FFMA, exact.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gqe_launch.h"

namespace gqe {
namespace {

__global__ void __launch_bounds__(256) k_segment_mean(const float* __restrict__ src, int64_t n_src, int d4, int64_t n,
                                                      const int64_t* __restrict__ ptr, const int32_t* __restrict__ cols,
                                                      float* __restrict__ out, unsigned long long* err) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp0; i < n; i += n_warps) {
    const int64_t b = __ldg(ptr + i), e = __ldg(ptr + i + 1);
    for (int c0 = 0; c0 < d4; c0 += 32) {
      const int c = c0 + lane;
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int64_t j = b; j < e; ++j) {
        int64_t row = (int64_t)__ldg(cols + j);
        if (row < 0 || row >= n_src) {
          if (lane == 0) report_index(err, IDX_ERR_ROW_RANGE, 0, (int32_t)row);
          row = 0;
        }
        if (c < d4) {
          const float4 v = __ldg(reinterpret_cast<const float4*>(src) + row * d4 + c);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
      // mask.div(num_neigh) then mask.mm(embed_matrix): every term is scaled before the sum in the
      // reference; summing first and dividing once differs by rounding only (~1e-7 relative)
      const float cnt = (float)(e - b);
      if (c < d4)
        reinterpret_cast<float4*>(out)[i * d4 + c] = make_float4(__fdiv_rn(acc.x, cnt), __fdiv_rn(acc.y, cnt),
                                                                 __fdiv_rn(acc.z, cnt), __fdiv_rn(acc.w, cnt));
    }
  }
}

constexpr int kLT = 64, kLK = 16;   // output tile 64 x 64, K step 16, 256 threads x (4 x 4) outputs

__global__ void __launch_bounds__(256) k_linear(const float* __restrict__ w, int m, int k, int64_t n,
                                                const float* __restrict__ x, int relu, float* __restrict__ out) {
  __shared__ float ws[kLK][kLT + 4];    // W tile, transposed: ws[kk][row]
  __shared__ float xs[kLK][kLT + 4];    // X tile: xs[kk][col]
  const int m0 = blockIdx.y * kLT;
  const int64_t n0 = (int64_t)blockIdx.x * kLT;
  const int tr = (threadIdx.x / 16) * 4, tc = (threadIdx.x % 16) * 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < k; k0 += kLK) {
    for (int idx = threadIdx.x; idx < kLT * kLK; idx += 256) {
      const int r = idx / kLK, kk = idx % kLK;            // W is row-major [m, k]: consecutive kk
      ws[kk][r] = (m0 + r < m && k0 + kk < k) ? __ldg(w + (size_t)(m0 + r) * k + k0 + kk) : 0.f;
      const int kx = idx / kLT, cx = idx % kLT;            // X is [k, n]: consecutive columns
      xs[kx][cx] = (k0 + kx < k && n0 + cx < n) ? __ldg(x + (size_t)(k0 + kx) * n + n0 + cx) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kLK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&ws[kk][tr]);
      const float4 b = *reinterpret_cast<const float4*>(&xs[kk][tc]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = m0 + tr + i;
      const int64_t c = n0 + tc + j;
      if (r < m && c < n) {
        const float v = acc[i][j];
        out[(size_t)r * n + c] = (relu && v < 0.f) ? 0.f : v;   // F.relu propagates NaN; so does this
      }
    }
}

}  // namespace

cudaError_t launch_segment_mean(const float* src, int64_t n_src, int d, int64_t n, const int64_t* ptr, const int32_t* cols,
                                float* out, unsigned long long* err, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (n + 7) / 8;
  const int grid = (int)(want < (int64_t)sms * 8 ? want : (int64_t)sms * 8);
  k_segment_mean<<<grid, 256, 0, st>>>(src, n_src, d / 4, n, ptr, cols, out, err);
  return cudaGetLastError();
}

cudaError_t launch_linear(const float* w, int m, int k, int64_t n, const float* x, int relu, float* out, cudaStream_t st) {
  if (n <= 0 || m <= 0) return cudaSuccess;
  const dim3 grid((unsigned)((n + kLT - 1) / kLT), (unsigned)((m + kLT - 1) / kLT));
  k_linear<<<grid, 256, 0, st>>>(w, m, k, n, x, relu, out);
  return cudaGetLastError();
}

}  // namespace gqe
