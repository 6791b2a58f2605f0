// gqe_vec.cu -- streaming kernel of the contraction-free operators.
//
// With the TransE or DistMult metapath decoder (reference netquery/decoders.py:181-208,
// 211-236) every relation operator is a vector add / multiply, and with
// SimpleSetIntersection (decoders.py:302-319) the intersection is an element-wise mean / min:
// the whole forward path of a query -- gather, L2 normalisation (encoders.py:41-43), the
// operators, the cosine (raw dot for DistMult chains) against the positive and the negative
// target and the hinge (model.py:112-127) -- touches (A + T) table rows of 4d bytes and does
// O(d) flops on them.  It is bound by the HBM gather, so there is no tile, no shared memory
// and no block-level synchronisation here: ONE WARP owns one query, issues the 128-bit loads
// of all its rows at once (up to 5 rows x d/128 float4 per lane in flight, 24 warps per SM),
// keeps everything in registers and reduces with warp shuffles.  The indices of the warp's
// next query are fetched (and mapped through the node map) while the rows of the current one
// are in flight.
//
// Same arithmetic, in the same order per element, as the tile kernels of gqe_simt.cuh (which
// remain the path for the ragged / many-target layouts and for DeepSets intersections on top
// of these decoders); only the order of the d-long reductions differs.
#include <cuda_runtime.h>
#include <stdint.h>

#include "gqe_launch.h"

namespace gqe {
namespace {

constexpr int kVecThreads = 256;
constexpr int kVecWarps = kVecThreads / 32;

__device__ __forceinline__ float vsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double vsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float vmin_nan(float a, float b) { return (a < b || a != a) ? a : b; }
__device__ __forceinline__ float sq4(const float4 v, float s) {
  s = fmaf(v.x, v.x, s); s = fmaf(v.y, v.y, s); s = fmaf(v.z, v.z, s); return fmaf(v.w, v.w, s);
}
__device__ __forceinline__ float dot4(const float4 a, const float4 b, float s) {
  s = fmaf(a.x, b.x, s); s = fmaf(a.y, b.y, s); s = fmaf(a.z, b.z, s); return fmaf(a.w, b.w, s);
}
__device__ __forceinline__ float4 div4(const float4 v, float n) {
  return make_float4(__fdiv_rn(v.x, n), __fdiv_rn(v.y, n), __fdiv_rn(v.z, n), __fdiv_rn(v.w, n));
}
// TransE: x + v (decoders.py:203,208); DistMult: x * v (decoders.py:231,236)
__device__ __forceinline__ float4 rel4(const float4 x, const float4 v, bool mul) {
  return mul ? make_float4(x.x * v.x, x.y * v.y, x.z * v.z, x.w * v.w)
             : make_float4(x.x + v.x, x.y + v.y, x.z + v.z, x.w + v.w);
}

template <int D>
__global__ void __launch_bounds__(kVecThreads, 3) gqe_fused_vec(const __grid_constant__ LaunchParams p) {
  constexpr int NV = (D + 127) / 128;          // float4 per lane per row
  constexpr int LANES = D >= 128 ? 32 : D / 4; // d < 128: the upper lanes idle
  __shared__ double red[kVecWarps];
  __shared__ int last;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const bool act = lane < LANES;
  const int64_t warp = (int64_t)blockIdx.x * kVecWarps + wib;
  const int64_t n_warps = (int64_t)gridDim.x * kVecWarps;
  const bool mul = p.decoder == GQE_DEC_DISTMULT;
  const bool use_min = p.inter == GQE_INTER_DEEPSETS_MIN || p.inter == GQE_INTER_SIMPLE_MIN;
  const int T = p.T;
  const int ik = p.index_kind;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  double local = 0.0;

  for (int si = 0; si < p.n_segs; ++si) {
    const SegDev& s = p.seg[si];
    const int structure = s.structure;
    const bool chain = structure <= GQE_CHAIN3;
    const int na = s.n_anchor;
    // lanes 0..na-1 hold the anchor indices of a query, lanes 8..8+T-1 its target indices.  Three
    // stages, one loop iteration apart, so that no dependent load is ever waited for: the raw index
    // of query i+2, the node-map lookup of query i+1, the bounds check + row loads of query i.
    const bool is_anc = lane < na, is_tgt = lane >= 8 && lane < 8 + T;
    const int my_mode = is_anc ? s.anc_mode[lane] : s.tgt_mode;
    const ModeDev& my_md = p.mode[my_mode];
    auto fetch_raw = [&](int64_t q) -> int32_t {
      if (q >= s.q_end) return 0;
      if (is_anc) return __ldg(p.anchor_rows + (int64_t)lane * p.anchor_stride + q);
      if (is_tgt) return __ldg(p.target_rows + q * T + (lane - 8));
      return 0;
    };
    int64_t q = s.q_begin + warp;
    int32_t raw_cur = fetch_raw(q), raw_nxt = fetch_raw(q + n_warps);
    int32_t cand = (is_anc || is_tgt) ? index_lookup(my_md, raw_cur, ik) : 0;
    for (; q < s.q_end; q += n_warps) {
      const int32_t idx = (is_anc || is_tgt) ? index_check(my_md, my_mode, cand, raw_cur, ik, p.err) : 0;
      // ---- all row loads of this query
      float4 a[GQE_MAX_ANCHORS][NV], t[2][NV];
#pragma unroll
      for (int b = 0; b < GQE_MAX_ANCHORS; ++b) {
        const int32_t row = __shfl_sync(0xffffffffu, idx, b);
        const float4* src = reinterpret_cast<const float4*>(s.anc_table[b < na ? b : 0] + (size_t)row * D);
#pragma unroll
        for (int j = 0; j < NV; ++j) a[b][j] = (act && b < na) ? __ldg(src + lane + 32 * j) : zero4;
      }
#pragma unroll
      for (int tt = 0; tt < 2; ++tt) {
        const int32_t row = __shfl_sync(0xffffffffu, idx, 8 + tt);
        const float4* src = reinterpret_cast<const float4*>(s.tgt_table + (size_t)row * D);
#pragma unroll
        for (int j = 0; j < NV; ++j) t[tt][j] = (act && tt < T) ? __ldg(src + lane + 32 * j) : zero4;
      }
      // ---- the later queries' indices travel while the rows do
      const int32_t raw_nxt2 = fetch_raw(q + 2 * n_warps);
      const int32_t cand_nxt = ((is_anc || is_tgt) && q + n_warps < s.q_end) ? index_lookup(my_md, raw_nxt, ik) : 0;

      float sc[2] = {0.f, 0.f};
      if (chain) {
        // cos(a_hat, t_hat (+|*) v_r1 ... v_rn), raw dot for DistMult (decoders.py:200-205,228-233)
        float sa = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) sa = sq4(a[0][j], sa);
        const float na_ = sqrtf(vsum(sa));
        float4 ah[NV];
        float sah = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) { ah[j] = div4(a[0][j], na_); sah = sq4(ah[j], sah); }
        sah = vsum(sah);
        const int hops = structure + 1;
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          if (tt >= T) break;
          float st = 0.f;
#pragma unroll
          for (int j = 0; j < NV; ++j) st = sq4(t[tt][j], st);
          const float nt = sqrtf(vsum(st));
          float dot = 0.f, sy = 0.f;
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            float4 y = div4(t[tt][j], nt);
            for (int h = 0; h < hops; ++h)
              y = rel4(y, act ? __ldg(reinterpret_cast<const float4*>(s.rel[h]) + lane + 32 * j) : zero4, mul);
            if (!act) y = zero4;
            dot = dot4(y, ah[j], dot);
            sy = sq4(y, sy);
          }
          dot = vsum(dot);
          if (mul) {
            sc[tt] = dot;
          } else {
            sy = vsum(sy);
            sc[tt] = dot / (fmaxf(sqrtf(sy), kCosEps) * fmaxf(sqrtf(sah), kCosEps));
          }
        }
      } else {
        // q = agg_b project(a_hat_b) [projected once more for 3-chain_inter]; cos(t_hat, q)  (model.py:77-109)
        float4 qv[NV];
#pragma unroll
        for (int b = 0; b < GQE_MAX_ANCHORS; ++b) {
          if (b >= na) break;
          float sa = 0.f;
#pragma unroll
          for (int j = 0; j < NV; ++j) sa = sq4(a[b][j], sa);
          const float nb = sqrtf(vsum(sa));
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            float4 e = div4(a[b][j], nb);
            auto vec = [&](int r) { return act ? __ldg(reinterpret_cast<const float4*>(s.rel[r]) + lane + 32 * j) : zero4; };
            if (structure == GQE_INTER_CHAIN3 && b == 1) {
              e = rel4(e, vec(1), mul);   // reverse(r2b) first (model.py:85)
              e = rel4(e, vec(2), mul);   // then reverse(r2a)
            } else {
              e = rel4(e, vec(b), mul);
            }
            float4 g;
            if (b == 0) g = e;
            else if (use_min) g = make_float4(vmin_nan(qv[j].x, e.x), vmin_nan(qv[j].y, e.y), vmin_nan(qv[j].z, e.z), vmin_nan(qv[j].w, e.w));
            else g = make_float4(qv[j].x + e.x, qv[j].y + e.y, qv[j].z + e.z, qv[j].w + e.w);
            if (!use_min && b == na - 1) g = div4(g, (float)na);   // torch.mean over the stack
            qv[j] = g;
          }
        }
        float sq = 0.f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
          if (structure == GQE_CHAIN_INTER3)   // model.py:107
            qv[j] = rel4(qv[j], act ? __ldg(reinterpret_cast<const float4*>(s.rel[2]) + lane + 32 * j) : zero4, mul);
          if (!act) qv[j] = zero4;
          sq = sq4(qv[j], sq);
        }
        const float nq = fmaxf(sqrtf(vsum(sq)), kCosEps);
#pragma unroll
        for (int tt = 0; tt < 2; ++tt) {
          if (tt >= T) break;
          float st = 0.f, dq = 0.f;
#pragma unroll
          for (int j = 0; j < NV; ++j) { st = sq4(t[tt][j], st); dq = dot4(t[tt][j], qv[j], dq); }
          st = vsum(st);
          dq = vsum(dq);
          // t_hat = t/|t| has unit norm: cos(t_hat, q) = (t.q/|t|) / max(|q|, eps); a zero target row
          // gives 0/0 = NaN as in the reference
          sc[tt] = __fdiv_rn(dq, sqrtf(st)) / nq;
        }
      }
      if (lane == 0) {
        if (p.out_scores) {
          p.out_scores[q * T] = sc[0];
          if (T > 1) p.out_scores[q * T + 1] = sc[1];
        }
        if (p.out_loss) {
          const float h = p.margin - (sc[0] - sc[1]);
          local += (double)(h < 0.f ? 0.f : h);
        }
      }
      raw_cur = raw_nxt;
      raw_nxt = raw_nxt2;
      cand = cand_nxt;
    }
  }

  if (!p.out_loss) return;
  // deterministic reduction: the query -> warp assignment is static, warps are added in order
  local = vsum(local);
  if (lane == 0) red[wib] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double sum = 0.0;
    for (int w = 0; w < kVecWarps; ++w) sum += red[w];
    p.partials[blockIdx.x] = sum;
    __threadfence();
    last = atomicAdd(p.ticket, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (last && wib == 0) {
    __threadfence();
    double sum = 0.0;
    for (unsigned int i = lane; i < gridDim.x; i += 32) sum += __ldcg(p.partials + i);
    sum = vsum(sum);
    if (lane == 0) {
      const double acc = *p.loss_acc + sum;
      *p.loss_acc = p.final_launch ? 0.0 : acc;
      *p.out_loss = (float)(acc * p.inv_q);
      *p.ticket = 0u;
    }
  }
}

template <int D>
cudaError_t launch_vec_t(const LaunchParams& lp, cudaStream_t st) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // never more CTAs than 64-row tiles (the margin-loss partials are sized by the tile count)
  const int64_t cap = (int64_t)sms * 3;
  const int grid = (int)(lp.n_tiles < cap ? (lp.n_tiles > 0 ? lp.n_tiles : 1) : cap);
  gqe_fused_vec<D><<<grid, kVecThreads, 0, st>>>(lp);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_fused_vec(int d, const LaunchParams& lp, cudaStream_t st) {
  switch (d) {
    case 32: return launch_vec_t<32>(lp, st);
    case 64: return launch_vec_t<64>(lp, st);
    case 128: return launch_vec_t<128>(lp, st);
    case 256: return launch_vec_t<256>(lp, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace gqe
