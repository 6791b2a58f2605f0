// gqe_bwd.cu -- operator-level backward kernels (fp32, CUDA cores) of the scoring path.
//
// The training step of the reference is loss.backward() through the operators of
// netquery/encoders.py:41-43 (gather + L2 normalise), decoders.py:142-150,200-208,228-236
// (relation operators), decoders.py:288-300,311-319 (intersection) and nn.CosineSimilarity
// (model.py:68).  Each kernel here is the vector-Jacobian product of ONE of those operators
// on feature-major [d, n] tensors (the reference's own layout); graphqembed_b200/autograd.py
// chains them exactly as autograd chains the reference's ops.  Parameter gradients are
// ACCUMULATED into the destination (atomicAdd: the summation order over the batch is not
// fixed, like torch's own embedding / index_add backward on CUDA).
#include <cuda_runtime.h>
#include <stdint.h>

#include "gqe_launch.h"

namespace gqe {
namespace {

constexpr int kCols = 32;       // columns (batch elements) per CTA tile
constexpr int kBwdThreads = 256;

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---- y = W x  /  y = W^T x  on a tile is the forward OP_MATMUL of gqe_simt.cuh ----------

// ---- gW += gY X^T (form 0: y = W x)   or   gW += X gY^T (form 1: y = W^T x) ---------------
// One CTA: a 64x64 tile of the d x d result over a chunk of 64 batch columns (split-K with
// atomic accumulation).  Both operands are row-contiguous along the batch, so the loads are
// coalesced; they are stored transposed so that a thread's 4 rows are one float4.
__global__ void __launch_bounds__(256) k_matmul_wgrad(const float* __restrict__ gy, const float* __restrict__ x,
                                                      float* __restrict__ gw, int d, int64_t n, int form) {
  constexpr int KC = 64;
  __shared__ float As[KC][64 + 4];  // As[c][i] = gy[i0 + i][c0 + c]
  __shared__ float Bs[KC][64 + 4];  // Bs[c][k] = x[k0 + k][c0 + c]
  const int i0 = blockIdx.y * 64, k0 = blockIdx.x * 64;
  const int64_t c0 = (int64_t)blockIdx.z * KC;
  for (int idx = threadIdx.x; idx < 64 * KC; idx += 256) {
    const int r = idx / KC, c = idx % KC;
    const bool ok = c0 + c < n;
    As[c][r] = (ok && i0 + r < d) ? __ldg(gy + (size_t)(i0 + r) * n + c0 + c) : 0.f;
    Bs[c][r] = (ok && k0 + r < d) ? __ldg(x + (size_t)(k0 + r) * n + c0 + c) : 0.f;
  }
  __syncthreads();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
#pragma unroll 4
  for (int c = 0; c < KC; ++c) {
    const float4 av = *reinterpret_cast<const float4*>(&As[c][4 * ty]);
    const float4 bv = *reinterpret_cast<const float4*>(&Bs[c][4 * tx]);
    const float a[4] = {av.x, av.y, av.z, av.w}, b[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gi = i0 + 4 * ty + i, xk = k0 + 4 * tx + j;
      if (gi >= d || xk >= d) continue;
      // form 0: W[out i][in k];  form 1 (y = W^T x): W[in k][out i]
      atomicAdd(gw + (form == 0 ? (size_t)gi * d + xk : (size_t)xk * d + gi), acc[i][j]);
    }
}

// ---- gv[i] += sum_c gy[i][c] (* x[i][c])  : TransE / DistMult parameter gradient ------------
__global__ void __launch_bounds__(256) k_rowsum(const float* __restrict__ gy, const float* __restrict__ x,
                                                float* __restrict__ gv, int64_t n) {
  __shared__ float red[8];
  const int i = blockIdx.x;
  const int64_t c_begin = (int64_t)blockIdx.y * 8192;
  const int64_t c_end = min(c_begin + 8192, n);
  float s = 0.f;
  for (int64_t c = c_begin + threadIdx.x; c < c_end; c += 256) {
    const float g = __ldg(gy + (size_t)i * n + c);
    s += x ? g * __ldg(x + (size_t)i * n + c) : g;
  }
  s = wsum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    atomicAdd(gv + i, t);
  }
}

// ---- intersection aggregator: out = agg_k act(e_k), act = relu or identity, agg = mean / min --
__global__ void __launch_bounds__(256) k_aggregate(const float* __restrict__ e1, const float* __restrict__ e2,
                                                   const float* __restrict__ e3, int relu, int use_min, int64_t total,
                                                   float* __restrict__ out) {
  const int k = e3 ? 3 : 2;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    float v[3] = {e1[i], e2[i], e3 ? e3[i] : 0.f};
    float a = 0.f;
    for (int j = 0; j < k; ++j) {
      const float z = relu ? (v[j] < 0.f ? 0.f : v[j]) : v[j];
      a = j == 0 ? z : (use_min ? ((a < z || a != a) ? a : z) : a + z);
    }
    out[i] = use_min ? a : __fdiv_rn(a, (float)k);
  }
}
// its backward: mean -> g/k to every operand; min -> g to the FIRST operand holding the minimum
// (torch.min(dim) routes the gradient to the index it returns); relu gates by e_k > 0.
__global__ void __launch_bounds__(256) k_aggregate_bwd(const float* __restrict__ e1, const float* __restrict__ e2,
                                                       const float* __restrict__ e3, int relu, int use_min, int64_t total,
                                                       const float* __restrict__ gout, float* __restrict__ g1,
                                                       float* __restrict__ g2, float* __restrict__ g3) {
  const int k = e3 ? 3 : 2;
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += (int64_t)gridDim.x * 256) {
    const float v[3] = {e1[i], e2[i], e3 ? e3[i] : 0.f};
    float z[3];
    int arg = 0;
    for (int j = 0; j < k; ++j) {
      z[j] = relu ? (v[j] < 0.f ? 0.f : v[j]) : v[j];
      if (j > 0 && z[j] < z[arg]) arg = j;
    }
    const float g = gout[i];
    float o[3];
    for (int j = 0; j < k; ++j) {
      float gj = use_min ? (j == arg ? g : 0.f) : g / (float)k;
      if (relu && !(v[j] > 0.f)) gj = 0.f;
      o[j] = gj;
    }
    g1[i] = o[0];
    g2[i] = o[1];
    if (e3) g3[i] = o[2];
  }
}

// ---- cosine(dim 0, eps) / raw dot backward ----------------------------------------------------
// A CTA owns 32 columns: tiles of x and y are parked in shared memory (coalesced along the
// batch), a warp reduces 4 columns, and gx / gy are written back coalesced.
//   cos = sum_k (x_k / max(|x|, eps)) (y_k / max(|y|, eps))
__global__ void __launch_bounds__(kBwdThreads) k_cosine_bwd(const float* __restrict__ x, const float* __restrict__ y,
                                                            const float* __restrict__ gout, int raw_dot, int d, int64_t n,
                                                            float* __restrict__ gx, float* __restrict__ gy) {
  extern __shared__ float sm[];
  float* xs = sm;                               // [d][33]
  float* ys = sm + (size_t)d * (kCols + 1);     // [d][33]
  __shared__ float coef[kCols][4];              // per column: a_x, b_x, a_y, b_y  (g = a * other + b * self)
  const int64_t c0 = (int64_t)blockIdx.x * kCols;
  const int nv = (int)min((int64_t)kCols, n - c0);
  for (int idx = threadIdx.x; idx < d * kCols; idx += kBwdThreads) {
    const int k = idx / kCols, c = idx % kCols;
    xs[k * (kCols + 1) + c] = c < nv ? x[(size_t)k * n + c0 + c] : 0.f;
    ys[k * (kCols + 1) + c] = c < nv ? y[(size_t)k * n + c0 + c] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < kCols; c += kBwdThreads / 32) {
    float dot = 0.f, sx = 0.f, sy = 0.f;
    for (int k = lane; k < d; k += 32) {
      const float a = xs[k * (kCols + 1) + c], b = ys[k * (kCols + 1) + c];
      dot = fmaf(a, b, dot); sx = fmaf(a, a, sx); sy = fmaf(b, b, sy);
    }
    dot = wsum(dot); sx = wsum(sx); sy = wsum(sy);
    if (lane == 0) {
      const float g = c < nv ? gout[c0 + c] : 0.f;
      if (raw_dot) {
        coef[c][0] = g; coef[c][1] = 0.f; coef[c][2] = g; coef[c][3] = 0.f;
      } else {
        const float nx = sqrtf(sx), ny = sqrtf(sy);
        const float cx = fmaxf(nx, kCosEps), cy = fmaxf(ny, kCosEps);
        const float inv = 1.f / (cx * cy);
        // d cos / d x = y / (cx cy) - [|x| > eps] (x.y) x / (cx^3 cy)   (a clamped norm is a constant)
        coef[c][0] = g * inv;
        coef[c][1] = nx > kCosEps ? -g * dot * inv / (cx * cx) : 0.f;
        coef[c][2] = g * inv;
        coef[c][3] = ny > kCosEps ? -g * dot * inv / (cy * cy) : 0.f;
      }
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < d * kCols; idx += kBwdThreads) {
    const int k = idx / kCols, c = idx % kCols;
    if (c >= nv) continue;
    const float a = xs[k * (kCols + 1) + c], b = ys[k * (kCols + 1) + c];
    if (gx) gx[(size_t)k * n + c0 + c] = coef[c][0] * b + coef[c][1] * a;
    if (gy) gy[(size_t)k * n + c0 + c] = coef[c][2] * a + coef[c][3] * b;
  }
}

// raw dot forward (BilinearDiag chain score, decoders.py:232)
__global__ void __launch_bounds__(kBwdThreads) k_dot(const float* __restrict__ x, const float* __restrict__ y, int d,
                                                     int64_t n, float* __restrict__ out) {
  for (int64_t c = (int64_t)blockIdx.x * kBwdThreads + threadIdx.x; c < n; c += (int64_t)gridDim.x * kBwdThreads) {
    float s = 0.f;
    for (int k = 0; k < d; ++k) s = fmaf(x[(size_t)k * n + c], y[(size_t)k * n + c], s);
    out[c] = s;
  }
}

// ---- DirectEncoder backward: out[:, c] = t / |t|, t = table[rows[c]] ---------------------------
//   g_t = (g - x_hat (x_hat . g)) / |t|,  scattered (atomicAdd) into the dense table gradient
__global__ void __launch_bounds__(kBwdThreads) k_encode_bwd(const float* __restrict__ table, const int32_t* __restrict__ rows,
                                                            const float* __restrict__ gout, int d, int64_t n,
                                                            float* __restrict__ gtable, int64_t table_rows,
                                                            unsigned long long* err) {
  extern __shared__ float sm[];
  float* gs = sm;  // [d][33]
  const int64_t c0 = (int64_t)blockIdx.x * kCols;
  const int nv = (int)min((int64_t)kCols, n - c0);
  for (int idx = threadIdx.x; idx < d * kCols; idx += kBwdThreads) {
    const int k = idx / kCols, c = idx % kCols;
    gs[k * (kCols + 1) + c] = c < nv ? gout[(size_t)k * n + c0 + c] : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = warp; c < nv; c += kBwdThreads / 32) {
    const int64_t row_i = (int64_t)__ldg(rows + c0 + c);
    if (row_i < 0 || row_i >= table_rows) {   // never scatter outside the gradient table
      if (lane == 0) report_index(err, IDX_ERR_ROW_RANGE, 0, (int32_t)row_i);
      continue;
    }
    const size_t row = (size_t)row_i;
    const float* t = table + row * d;
    float ss = 0.f, tg = 0.f;
    for (int k = lane; k < d; k += 32) {
      const float v = __ldg(t + k);
      ss = fmaf(v, v, ss);
      tg = fmaf(v, gs[k * (kCols + 1) + c], tg);
    }
    ss = wsum(ss); tg = wsum(tg);
    const float nrm = sqrtf(ss);
    const float inv = 1.f / nrm, proj = tg / (nrm * nrm * nrm);   // x_hat (x_hat.g) / |t| = t (t.g) / |t|^3
    for (int k = lane; k < d; k += 32)
      atomicAdd(gtable + row * d + k, gs[k * (kCols + 1) + c] * inv - __ldg(t + k) * proj);
  }
}

int grid_for(int64_t work, int per_cta) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (work + per_cta - 1) / per_cta;
  return (int)(want < (int64_t)sms * 8 ? (want > 0 ? want : 1) : (int64_t)sms * 8);
}

}  // namespace

cudaError_t launch_matmul_wgrad(int d, int64_t n, const float* gy, const float* x, int form, float* gw, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const dim3 grid((d + 63) / 64, (d + 63) / 64, (unsigned)((n + 63) / 64));
  k_matmul_wgrad<<<grid, 256, 0, st>>>(gy, x, gw, d, n, form);
  return cudaGetLastError();
}
cudaError_t launch_rowsum(int d, int64_t n, const float* gy, const float* x, float* gv, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_rowsum<<<dim3(d, (unsigned)((n + 8191) / 8192)), 256, 0, st>>>(gy, x, gv, n);
  return cudaGetLastError();
}
cudaError_t launch_aggregate(int64_t total, const float* e1, const float* e2, const float* e3, int relu, int use_min,
                             float* out, cudaStream_t st) {
  if (total <= 0) return cudaSuccess;
  k_aggregate<<<grid_for(total, 1024), 256, 0, st>>>(e1, e2, e3, relu, use_min, total, out);
  return cudaGetLastError();
}
cudaError_t launch_aggregate_bwd(int64_t total, const float* e1, const float* e2, const float* e3, int relu, int use_min,
                                 const float* gout, float* g1, float* g2, float* g3, cudaStream_t st) {
  if (total <= 0) return cudaSuccess;
  k_aggregate_bwd<<<grid_for(total, 1024), 256, 0, st>>>(e1, e2, e3, relu, use_min, total, gout, g1, g2, g3);
  return cudaGetLastError();
}
cudaError_t launch_cosine_bwd(int d, int64_t n, const float* x, const float* y, const float* gout, int raw_dot, float* gx,
                              float* gy, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const size_t smem = (size_t)2 * d * (kCols + 1) * sizeof(float);
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(k_cosine_bwd, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 256 * (kCols + 1) * 4);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  k_cosine_bwd<<<(unsigned)((n + kCols - 1) / kCols), kBwdThreads, smem, st>>>(x, y, gout, raw_dot, d, n, gx, gy);
  return cudaGetLastError();
}
cudaError_t launch_dot(int d, int64_t n, const float* x, const float* y, float* out, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_dot<<<grid_for(n, kBwdThreads), kBwdThreads, 0, st>>>(x, y, d, n, out);
  return cudaGetLastError();
}
cudaError_t launch_encode_bwd(int d, int64_t n, const float* table, const int32_t* rows, const float* gout, float* gtable,
                              int64_t table_rows, unsigned long long* err, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const size_t smem = (size_t)d * (kCols + 1) * sizeof(float);
  k_encode_bwd<<<(unsigned)((n + kCols - 1) / kCols), kBwdThreads, smem, st>>>(table, rows, gout, d, n, gtable, table_rows, err);
  return cudaGetLastError();
}

}  // namespace gqe
