// gqe_pairs.cu -- (query, target) pair scoring for many targets per query.
//
// The evaluation callers of the reference score ONE query against up to 1000
// negatives (netquery/utils.py:70-91 repeats the query once per negative, so the
// reference re-gathers and re-projects every anchor K times).  Here the fused
// kernel evaluates the query side once and leaves the query embeddings q[Q, d] in
// HBM; this kernel then scores every pair
//     score(p) = cos( t_hat(target_rows[p]), q[query_of(p)] )     model.py:97
// with t_hat = t/|t| (encoders.py:41-43).  Pure HBM gather: one table row (4d bytes)
// + 4 index bytes in, 4 bytes out per pair; the query rows are re-read from L2.
// A warp owns 32 consecutive pairs and keeps kU table rows in flight.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gqe_launch.h"

namespace gqe {

template <int D>
__global__ void __launch_bounds__(256) gqe_score_pairs(const __grid_constant__ PairParams p) {
  constexpr int NV = (D + 127) / 128;       // float4 per lane per row (d < 128: upper lanes idle)
  constexpr int LANES = D >= 128 ? 32 : D / 4;
  constexpr int kU = 4;
  const int lane = threadIdx.x & 31;
  const bool act = lane < LANES;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const bool ragged = p.target_offsets != nullptr;

  for (int s = 0; s < p.n_segs; ++s) {
    const PairSeg& sg = p.seg[s];
    const int64_t pair_begin = ragged ? 0 : sg.q_begin * p.T;
    const int64_t pair_end = ragged ? p.n_pairs : sg.q_end * p.T;
    for (int64_t p0 = pair_begin + warp * 32; p0 < pair_end; p0 += n_warps * 32) {
      const int64_t p1 = min(p0 + 32, pair_end);
      // query of the first pair of the block; later pairs advance linearly
      int64_t q;
      if (!ragged) {
        q = p0 / p.T;
      } else {
        int64_t lo = sg.q_begin, hi = sg.q_end;  // largest q with offsets[q] <= p0
        while (hi - lo > 1) {
          const int64_t mid = (lo + hi) >> 1;
          if (__ldg(p.target_offsets + mid) <= p0) lo = mid; else hi = mid;
        }
        q = lo;
      }
      int64_t q_last = ragged ? __ldg(p.target_offsets + q + 1) : (q + 1) * p.T;  // first pair of the next query
      const int32_t my_row = (p0 + lane < p1) ? resolve_index(sg.mode, 0, __ldg(p.target_rows + p0 + lane), p.index_kind, p.err) : 0;
      float4 qv[NV];
      float nq = 1.f;
      int64_t q_loaded = -1;
      for (int64_t b = p0; b < p1; b += kU) {
        float4 tv[kU][NV];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int32_t row = __shfl_sync(0xffffffffu, my_row, (int)(b - p0) + u);
          const float4* src = reinterpret_cast<const float4*>(sg.tgt_table + (size_t)row * D);
#pragma unroll
          for (int j = 0; j < NV; ++j)
            tv[u][j] = (act && b + u < p1) ? __ldg(src + lane + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
          const int64_t pr = b + u;
          if (pr >= p1) break;  // warp-uniform
          while (pr >= q_last) {  // next query (empty target lists are skipped)
            ++q;
            q_last = ragged ? __ldg(p.target_offsets + q + 1) : (q + 1) * p.T;
          }
          if (q != q_loaded) {
            const float4* qs = reinterpret_cast<const float4*>(p.q + (size_t)q * D);
            float qq = 0.f;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
              qv[j] = act ? __ldg(qs + lane + 32 * j) : make_float4(0.f, 0.f, 0.f, 0.f);
              qq = fmaf(qv[j].x, qv[j].x, qq); qq = fmaf(qv[j].y, qv[j].y, qq);
              qq = fmaf(qv[j].z, qv[j].z, qq); qq = fmaf(qv[j].w, qv[j].w, qq);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) qq += __shfl_xor_sync(0xffffffffu, qq, o);
            nq = fmaxf(sqrtf(qq), kCosEps);  // cosine eps: max(|q|, 1e-8)
            q_loaded = q;
          }
          float dt = 0.f, nt = 0.f;
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            const float4 t = tv[u][j];
            dt = fmaf(qv[j].x, t.x, dt); dt = fmaf(qv[j].y, t.y, dt); dt = fmaf(qv[j].z, t.z, dt); dt = fmaf(qv[j].w, t.w, dt);
            nt = fmaf(t.x, t.x, nt); nt = fmaf(t.y, t.y, nt); nt = fmaf(t.z, t.z, nt); nt = fmaf(t.w, t.w, nt);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            dt += __shfl_xor_sync(0xffffffffu, dt, o);
            nt += __shfl_xor_sync(0xffffffffu, nt, o);
          }
          // t_hat = t/|t| has unit norm: cos(t_hat, q) = (t.q/|t|) / max(|q|, eps); a zero
          // target row gives 0/0 = NaN as in the reference
          if (lane == 0) p.out_scores[pr] = __fdiv_rn(dt, sqrtf(nt)) / nq;
        }
      }
    }
  }
}

template <int D>
static cudaError_t launch_pairs_t(const PairParams& pp, int64_t n_pairs_total, cudaStream_t st) {
  if (n_pairs_total <= 0) return cudaSuccess;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t blocks = (n_pairs_total + 255) / 256;  // 8 warps x 32 pairs per CTA pass
  const int grid = (int)(blocks < (int64_t)sms * 8 ? blocks : (int64_t)sms * 8);
  gqe_score_pairs<D><<<grid, 256, 0, st>>>(pp);
  return cudaGetLastError();
}

cudaError_t launch_score_pairs(int d, const PairParams& pp, int64_t n_pairs_total, cudaStream_t st) {
  switch (d) {
    case 32: return launch_pairs_t<32>(pp, n_pairs_total, st);
    case 64: return launch_pairs_t<64>(pp, n_pairs_total, st);
    case 128: return launch_pairs_t<128>(pp, n_pairs_total, st);
    case 256: return launch_pairs_t<256>(pp, n_pairs_total, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace gqe
