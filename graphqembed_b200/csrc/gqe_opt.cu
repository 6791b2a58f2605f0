// gqe_opt.cu -- sparse side of the training step: row gradients of the embedding tables and a
// row-wise Adam that reproduces DENSE torch.optim.Adam.
//
// The reference's step is loss.backward(); optimizer.step() with nn.Embedding tables and dense
// Adam (netquery/bio/train.py:59-62, train_helpers.py:78-79): every step materialises an
// [N_mode + 2, d] gradient per touched table and updates all of its rows, although a batch of 512
// queries touches ~1500 of them.  Here
//   * k_encode_bwd_rows emits the gradient of the DirectEncoder (encoders.py:41-43) PER GATHERED
//     ROW -- (rows[c], g_c) pairs, no scatter, no atomics: the table gradient as a sparse tensor;
//   * k_adam_rows applies Adam to exactly those rows, and first CATCHES UP each row with the
//     zero-gradient steps dense Adam would have applied since the row was last touched (the
//     moments decay, the row keeps moving): m <- b1 m, v <- b2 v,
//     p <- p - lr/(1 - b1^t) * m / (sqrt(v)/sqrt(1 - b2^t) + eps) for every skipped step t of the
//     table.  So the trajectory is dense Adam's, at the cost of the touched rows.  Rows that
//     were never touched have m = v = 0 and never move: nothing to catch up.
// A catch-up longer than kExactSteps applies the first kExactSteps steps one by one and folds
// the rest into the moments (b1^512 ~ 4e-24: the skipped parameter movement is below fp32
// resolution).
#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>

#include "gqe_launch.h"

namespace gqe {
namespace {

constexpr int kExactSteps = 512;

__device__ __forceinline__ float osum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// out[c, :] = (g_c - x_hat (x_hat . g_c)) / |t|,  t = table[rows[c]], g_c = gout[:, c]   (one warp per column)
__global__ void __launch_bounds__(256) k_encode_bwd_rows(const float* __restrict__ table, const int32_t* __restrict__ rows,
                                                         const float* __restrict__ gout, int d, int64_t n,
                                                         float* __restrict__ out, int64_t table_rows,
                                                         unsigned long long* err) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t c = warp0; c < n; c += n_warps) {
    int64_t row = (int64_t)__ldg(rows + c);
    const bool bad = row < 0 || row >= table_rows;
    if (bad) {
      if (lane == 0) report_index(err, IDX_ERR_ROW_RANGE, 0, (int32_t)row);
      row = 0;
    }
    const float* t = table + row * d;
    float ss = 0.f, tg = 0.f;
    for (int k = lane; k < d; k += 32) {
      const float v = __ldg(t + k);
      ss = fmaf(v, v, ss);
      tg = fmaf(v, gout[(size_t)k * n + c], tg);
    }
    ss = osum(ss); tg = osum(tg);
    const float nrm = sqrtf(ss);
    const float inv = 1.f / nrm, proj = tg / (nrm * nrm * nrm);   // x_hat (x_hat.g) / |t| = t (t.g) / |t|^3
    for (int k = lane; k < d; k += 32)
      out[(size_t)c * d + k] = bad ? 0.f : gout[(size_t)k * n + c] * inv - __ldg(t + k) * proj;
  }
}

struct AdamHyper {
  float lr, beta1, beta2, eps;
  float bc1, bc2s;   // bias corrections of the step being applied: 1 - beta1^step, sqrt(1 - beta2^step) (host, double)
};

// zero-gradient Adam steps (from, to] of one row (lane-strided over d <= 256).
// 1 - beta^s is carried as om <- om * beta + (1 - beta) (no cancellation: ~1e-7 relative in fp32; the
// start value comes from one double pow per row), so the per-step work is fp32 only -- the first
// version walked beta^s in double per element, ~30x the cost on this part's fp64 rate.
__device__ __forceinline__ void adam_catch_up(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, int d,
                                              int lane, int from, int to, const AdamHyper h) {
  if (to <= from) return;
  constexpr int kMaxPer = 8;                        // d <= 256
  const int exact = min(to - from, kExactSteps);
  float om1 = 0.f, om2 = 0.f;
  if (lane == 0) om1 = (float)(1.0 - pow((double)h.beta1, (double)from));
  if (lane == 1) om2 = (float)(1.0 - pow((double)h.beta2, (double)from));
  om1 = __shfl_sync(0xffffffffu, om1, 0);
  om2 = __shfl_sync(0xffffffffu, om2, 1);
  float pm[kMaxPer], pv[kMaxPer], pp[kMaxPer];
  bool live = false;
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) {
    const int k = lane + 32 * j;
    pm[j] = k < d ? m[k] : 0.f;
    pv[j] = k < d ? v[k] : 0.f;
    pp[j] = k < d ? p[k] : 0.f;
    live = live || pm[j] != 0.f || pv[j] != 0.f;
  }
  if (!__any_sync(0xffffffffu, live)) return;       // never touched: m = v = 0, nothing moves
  const float c1 = 1.f - h.beta1, c2 = 1.f - h.beta2;
  for (int s = 0; s < exact; ++s) {
    om1 = fmaf(om1, h.beta1, c1);
    om2 = fmaf(om2, h.beta2, c2);
    const float step = h.lr / om1;
    const float inv_bc2s = rsqrtf(om2);
#pragma unroll
    for (int j = 0; j < kMaxPer; ++j) {
      pm[j] *= h.beta1;
      pv[j] *= h.beta2;
      pp[j] -= step * (pm[j] / (sqrtf(pv[j]) * inv_bc2s + h.eps));
    }
  }
  const int rest = to - from - exact;
  const float r1 = rest > 0 ? powf(h.beta1, (float)rest) : 1.f, r2 = rest > 0 ? powf(h.beta2, (float)rest) : 1.f;
#pragma unroll
  for (int j = 0; j < kMaxPer; ++j) {
    const int k = lane + 32 * j;
    if (k < d) { m[k] = pm[j] * r1; v[k] = pv[j] * r2; p[k] = pp[j]; }
  }
}

// grads == nullptr: catch the listed rows (or, rows == nullptr, ALL rows) up to `step`;
//                   duplicate rows in the list are fine (one claimant per row).
// grads != nullptr, !GSUM: `rows` are UNIQUE and grads holds one gradient row per list entry; catch each
//                   row up to step - 1, then apply Adam step `step` with its gradient row.
// grads != nullptr, GSUM: grads is a DENSE [table_rows, d] buffer into which the batch's row gradients
//                   were accumulated (k_encode_bwd); `rows` may repeat.  The first warp to claim a row
//                   (last[row] <- -step while it works) catches it up, applies Adam step `step` with
//                   the accumulated row and ZEROES that row of the buffer again; the other claimants
//                   of the row skip it.  Rows already at `step` (an earlier list of the same step) are
//                   skipped as well.
template <typename I, bool GSUM>
__global__ void __launch_bounds__(256) k_adam_rows(float* __restrict__ table, float* __restrict__ m, float* __restrict__ v,
                                                   int32_t* __restrict__ last, int64_t table_rows, int d, int64_t n,
                                                   const I* __restrict__ rows, float* __restrict__ grads,
                                                   int step, const AdamHyper h) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t i = warp0; i < n; i += n_warps) {
    const int64_t row = rows ? (int64_t)rows[i] : i;
    if (row < 0 || row >= table_rows) continue;
    float* p = table + row * d;
    float* pm = m + row * d;
    float* pv = v + row * d;
    const int target = grads ? step - 1 : step;
    int from = 0;
    bool mine = false;
    if (lane == 0) {
      from = last[row];
      if (GSUM) mine = from >= 0 && from < step && atomicCAS(last + row, from, -step) == from;
      else if (from < target) mine = atomicCAS(last + row, from, target) == from;
      else mine = grads != nullptr && from == target;
    }
    from = __shfl_sync(0xffffffffu, from, 0);
    mine = __shfl_sync(0xffffffffu, mine ? 1 : 0, 0) != 0;
    if (!mine) continue;                       // another warp of this launch owns the row
    if (from < target && from > 0) adam_catch_up(p, pm, pv, d, lane, from, target, h);
    if (grads) {
      const float bc1 = h.bc1, bc2s = h.bc2s;
      float* g = grads + (size_t)(GSUM ? row : i) * d;
      for (int k = lane; k < d; k += 32) {
        const float gk = g[k];
        if (GSUM) g[k] = 0.f;
        const float mk = pm[k] + (gk - pm[k]) * (1.f - h.beta1);     // exp_avg.lerp_(grad, 1 - beta1)
        const float vk = pv[k] * h.beta2 + (1.f - h.beta2) * gk * gk;
        pm[k] = mk;
        pv[k] = vk;
        p[k] -= (h.lr / bc1) * (mk / (sqrtf(vk) / bc2s + h.eps));
      }
      __syncwarp();
      if (lane == 0) {
        if (GSUM) __threadfence();
        last[row] = step;
      }
    }
  }
}

// torch.optim.Adam on a small dense parameter (relation matrices / vectors, DeepSets pre / post)
__global__ void __launch_bounds__(256) k_adam_dense(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                    const float* __restrict__ g, int64_t n, int step, const AdamHyper h) {
  const float bc1 = h.bc1, bc2s = h.bc2s;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gk = g[i];
    const float mk = m[i] + (gk - m[i]) * (1.f - h.beta1);
    const float vk = v[i] * h.beta2 + (1.f - h.beta2) * gk * gk;
    m[i] = mk;
    v[i] = vk;
    p[i] -= (h.lr / bc1) * (mk / (sqrtf(vk) / bc2s + h.eps));
  }
}

AdamHyper make_hyper(float lr, float beta1, float beta2, float eps, int step) {
  AdamHyper h{lr, beta1, beta2, eps, 1.f, 1.f};
  if (step > 0) {
    h.bc1 = (float)(1.0 - std::pow((double)beta1, (double)step));
    h.bc2s = std::sqrt((float)(1.0 - std::pow((double)beta2, (double)step)));
  }
  return h;
}

int opt_grid(int64_t warps_wanted) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t want = (warps_wanted + 7) / 8;
  return (int)(want < (int64_t)sms * 8 ? (want > 0 ? want : 1) : (int64_t)sms * 8);
}

}  // namespace

cudaError_t launch_encode_bwd_rows(int d, int64_t n, const float* table, const int32_t* rows, const float* gout, float* out,
                                   int64_t table_rows, unsigned long long* err, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  k_encode_bwd_rows<<<opt_grid(n), 256, 0, st>>>(table, rows, gout, d, n, out, table_rows, err);
  return cudaGetLastError();
}

cudaError_t launch_adam_rows(float* table, float* m, float* v, int32_t* last, int64_t table_rows, int d, int64_t n,
                             const int64_t* rows, const float* grads, int step, float lr, float beta1, float beta2,
                             float eps, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const AdamHyper h = make_hyper(lr, beta1, beta2, eps, step);
  k_adam_rows<int64_t, false><<<opt_grid(n), 256, 0, st>>>(table, m, v, last, table_rows, d, n, rows, const_cast<float*>(grads), step, h);
  return cudaGetLastError();
}

cudaError_t launch_adam_rows_accum(float* table, float* m, float* v, int32_t* last, int64_t table_rows, int d, int64_t n,
                                   const int32_t* rows, float* gsum, int step, float lr, float beta1, float beta2, float eps,
                                   cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const AdamHyper h = make_hyper(lr, beta1, beta2, eps, step);
  if (gsum) k_adam_rows<int32_t, true><<<opt_grid(n), 256, 0, st>>>(table, m, v, last, table_rows, d, n, rows, gsum, step, h);
  else k_adam_rows<int32_t, false><<<opt_grid(n), 256, 0, st>>>(table, m, v, last, table_rows, d, n, rows, nullptr, step, h);
  return cudaGetLastError();
}

cudaError_t launch_adam_dense(float* p, float* m, float* v, const float* g, int64_t n, int step, float lr, float beta1,
                              float beta2, float eps, cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  const AdamHyper h = make_hyper(lr, beta1, beta2, eps, step);
  const int64_t blocks = (n + 255) / 256;
  k_adam_dense<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, st>>>(p, m, v, g, n, step, h);
  return cudaGetLastError();
}

}  // namespace gqe
