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
// x / n as x * (1 / n): ONE IEEE division per row instead of one per element (each of those is a
// MUFU + Newton steps + a checked slow path: they made this kernel issue-bound at 0.25 of the HBM
// roofline).  <= 1 ulp from the reference's true division; n = 0 gives 0 * inf = NaN like 0 / 0.
__device__ __forceinline__ float4 scale4(const float4 v, float inv) {
  return make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
}
// TransE: x + v (decoders.py:203,208); DistMult: x * v (decoders.py:231,236)
__device__ __forceinline__ float4 rel4(const float4 x, const float4 v, bool mul) {
  return mul ? make_float4(x.x * v.x, x.y * v.y, x.z * v.z, x.w * v.w)
             : make_float4(x.x + v.x, x.y + v.y, x.z + v.z, x.w + v.w);
}

// butterflies of several values in lockstep (their shuffle latencies overlap)
template <int N>
__device__ __forceinline__ void vsum_n(float (&v)[N]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
}

// One query, everything in registers.  `idx`: lanes 0..NA-1 hold the anchor rows, lanes 8, 9 the
// target rows (lane 9 = lane 8 when the query has one target).  Returns the two scores.
template <int D, int NA, bool CHAIN>
__device__ __forceinline__ void score_query(const SegDev& s, int32_t idx, int lane, bool mul, bool use_min, float (&sc)[2]) {
  constexpr int NV = (D + 127) / 128;          // float4 per lane per row
  constexpr int LANES = D >= 128 ? 32 : D / 4; // d < 128: the upper lanes idle
  const bool act = LANES == 32 || lane < LANES;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto ld = [&](const float* base, int j) {
    return act ? __ldg(reinterpret_cast<const float4*>(base) + lane + 32 * j) : zero4;
  };
  // ---- all row loads of this query, back to back
  float4 a[NA][NV], t[2][NV];
#pragma unroll
  for (int b = 0; b < NA; ++b) {
    const float* src = s.anc_table[b] + (size_t)__shfl_sync(0xffffffffu, idx, b) * D;
#pragma unroll
    for (int j = 0; j < NV; ++j) a[b][j] = ld(src, j);
  }
#pragma unroll
  for (int tt = 0; tt < 2; ++tt) {
    const float* src = s.tgt_table + (size_t)__shfl_sync(0xffffffffu, idx, 8 + tt) * D;
#pragma unroll
    for (int j = 0; j < NV; ++j) t[tt][j] = ld(src, j);
  }
  if (CHAIN) {
    // cos(a_hat, t_hat (+|*) v_r1 ... v_rn), raw dot for DistMult (decoders.py:200-205,228-233).
    // a_hat has unit norm, so the cosine's own |a_hat| is 1 (to 1e-7) and is not recomputed.
    float n2[3] = {0.f, 0.f, 0.f};               // |a|^2, |t0|^2, |t1|^2
#pragma unroll
    for (int j = 0; j < NV; ++j) { n2[0] = sq4(a[0][j], n2[0]); n2[1] = sq4(t[0][j], n2[1]); n2[2] = sq4(t[1][j], n2[2]); }
    vsum_n(n2);
    // x / |x| as x * (1 / |x|): one IEEE reciprocal per row instead of a division per element
    // (<= 1 ulp from the reference's true division; |x| = 0 gives 0 * inf = NaN like 0 / 0)
    const float ia = __frcp_rn(sqrtf(n2[0])), it0 = __frcp_rn(sqrtf(n2[1])), it1 = __frcp_rn(sqrtf(n2[2]));
    const int hops = s.structure + 1;
    float r[4] = {0.f, 0.f, 0.f, 0.f};           // y0.a_hat, |y0|^2, y1.a_hat, |y1|^2
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float4 y0 = scale4(t[0][j], it0), y1 = scale4(t[1][j], it1);
      for (int h = 0; h < hops; ++h) {
        const float4 v = ld(s.rel[h], j);
        y0 = rel4(y0, v, mul);
        y1 = rel4(y1, v, mul);
      }
      if (!act) { y0 = zero4; y1 = zero4; }
      const float4 ah = scale4(a[0][j], ia);
      r[0] = dot4(y0, ah, r[0]); r[1] = sq4(y0, r[1]);
      r[2] = dot4(y1, ah, r[2]); r[3] = sq4(y1, r[3]);
    }
    vsum_n(r);
    sc[0] = mul ? r[0] : r[0] / fmaxf(sqrtf(r[1]), kCosEps);
    sc[1] = mul ? r[2] : r[2] / fmaxf(sqrtf(r[3]), kCosEps);
  } else {
    // q = agg_b project(a_hat_b) [projected once more for 3-chain_inter]; cos(t_hat, q)  (model.py:77-109)
    float n2[NA];
#pragma unroll
    for (int b = 0; b < NA; ++b) {
      n2[b] = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) n2[b] = sq4(a[b][j], n2[b]);
    }
    vsum_n(n2);
    float ib[NA];
#pragma unroll
    for (int b = 0; b < NA; ++b) ib[b] = __frcp_rn(sqrtf(n2[b]));
    const int structure = s.structure;
    float r[5] = {0.f, 0.f, 0.f, 0.f, 0.f};      // |q|^2, |t0|^2, t0.q, |t1|^2, t1.q
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float4 q;
#pragma unroll
      for (int b = 0; b < NA; ++b) {
        float4 e = scale4(a[b][j], ib[b]);
        if (NA == 2 && b == 1 && structure == GQE_INTER_CHAIN3) {
          e = rel4(e, ld(s.rel[1], j), mul);     // reverse(r2b) first (model.py:85)
          e = rel4(e, ld(s.rel[2], j), mul);     // then reverse(r2a)
        } else {
          e = rel4(e, ld(s.rel[b], j), mul);
        }
        if (b == 0) q = e;
        else if (use_min) q = make_float4(vmin_nan(q.x, e.x), vmin_nan(q.y, e.y), vmin_nan(q.z, e.z), vmin_nan(q.w, e.w));
        else q = make_float4(q.x + e.x, q.y + e.y, q.z + e.z, q.w + e.w);
      }
      if (!use_min) q = scale4(q, NA == 2 ? 0.5f : 1.f / 3.f);                    // torch.mean over the stack
      if (NA == 2 && structure == GQE_CHAIN_INTER3) q = rel4(q, ld(s.rel[2], j), mul);   // model.py:107
      if (!act) q = zero4;
      r[0] = sq4(q, r[0]);
      r[1] = sq4(t[0][j], r[1]); r[2] = dot4(t[0][j], q, r[2]);
      r[3] = sq4(t[1][j], r[3]); r[4] = dot4(t[1][j], q, r[4]);
    }
    vsum_n(r);
    const float nq = fmaxf(sqrtf(r[0]), kCosEps);
    // t_hat = t/|t| has unit norm: cos(t_hat, q) = (t.q/|t|) / max(|q|, eps); a zero target row gives
    // 0/0 = NaN as in the reference
    sc[0] = __fdiv_rn(r[2], sqrtf(r[1])) / nq;
    sc[1] = __fdiv_rn(r[4], sqrtf(r[3])) / nq;
  }
}

#ifndef GQE_VEC_BLOCKS_D256
#define GQE_VEC_BLOCKS_D256 3
#endif
template <int D>
__global__ void __launch_bounds__(kVecThreads, (D >= 256 ? GQE_VEC_BLOCKS_D256 : 4)) gqe_fused_vec(const __grid_constant__ LaunchParams p) {
  __shared__ double red[kVecWarps];
  __shared__ int last;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t warp = (int64_t)blockIdx.x * kVecWarps + wib;
  const int64_t n_warps = (int64_t)gridDim.x * kVecWarps;
  const bool mul = p.decoder == GQE_DEC_DISTMULT;
  const bool use_min = p.inter == GQE_INTER_DEEPSETS_MIN || p.inter == GQE_INTER_SIMPLE_MIN;
  const int T = p.T;
  const int ik = p.index_kind;
  double local = 0.0;

  // The queries of ALL segments form one flat iteration space g = 0 .. total-1 (segment after
  // segment); warp w takes g = w, w + n_warps, ...  so that the software pipeline below never
  // drains at a formula boundary and the last wave is one query per warp, not one per formula.
  int64_t total = 0;
  for (int i = 0; i < p.n_segs; ++i) total += p.seg[i].q_end - p.seg[i].q_begin;
  struct Cursor { int si; int64_t base; };                 // g -> (segment, query); only moves forward
  auto locate = [&](Cursor& c, int64_t g) -> int64_t {    // the query index, -1 past the end
    if (g >= total) return -1;
    while (g >= c.base + (p.seg[c.si].q_end - p.seg[c.si].q_begin)) {
      c.base += p.seg[c.si].q_end - p.seg[c.si].q_begin;
      ++c.si;
    }
    return p.seg[c.si].q_begin + (g - c.base);
  };
  // lanes 0..na-1 hold the anchor indices of a query, lanes 8, 9 its target indices (lane 9 repeats
  // target 0 when T == 1).  Three stages, one loop iteration apart, so that no dependent load is ever
  // waited for: the raw index of query i+2, the node-map lookup of query i+1, the bounds check + row
  // loads of query i.
  const int tslot = lane == 9 && T > 1 ? 1 : 0;
  auto lane_mode = [&](const SegDev& sg) -> int {        // this lane's node type in a query of `sg`, -1: none
    if (lane < sg.n_anchor) return sg.anc_mode[lane];
    if (lane == 8 || lane == 9) return sg.tgt_mode;
    return -1;
  };
  auto fetch_raw = [&](int64_t q, const Cursor& c) -> int32_t {
    if (q < 0) return 0;
    if (lane < p.seg[c.si].n_anchor) return __ldg(p.anchor_rows + (int64_t)lane * p.anchor_stride + q);
    if (lane == 8 || lane == 9) return __ldg(p.target_rows + q * T + tslot);
    return 0;
  };
  Cursor c0{0, 0}, c1{0, 0}, c2{0, 0};
  int64_t g = warp;
  int64_t q = locate(c0, g), q1 = locate(c1, g + n_warps);
  int32_t raw_cur = fetch_raw(q, c0), raw_nxt = fetch_raw(q1, c1);
  int32_t cand = 0;
  if (q >= 0) {
    const int m = lane_mode(p.seg[c0.si]);
    if (m >= 0) cand = index_lookup(p.mode[m], raw_cur, ik);
  }
  for (; q >= 0; g += n_warps) {
    const SegDev& s = p.seg[c0.si];
    int32_t idx = 0;
    {
      const int my_mode = lane_mode(s);
      if (my_mode >= 0) idx = index_check(p.mode[my_mode], my_mode, cand, raw_cur, ik, p.err);
    }
    // the later queries' indices travel while this query's rows do
    const int64_t q2 = locate(c2, g + 2 * n_warps);
    const int32_t raw_nxt2 = fetch_raw(q2, c2);
    int32_t cand_nxt = 0;
    if (q1 >= 0) {
      const int m = lane_mode(p.seg[c1.si]);
      if (m >= 0) cand_nxt = index_lookup(p.mode[m], raw_nxt, ik);
    }
    float sc[2];
    if (s.structure <= GQE_CHAIN3) score_query<D, 1, true>(s, idx, lane, mul, use_min, sc);
    else if (s.n_anchor == 2) score_query<D, 2, false>(s, idx, lane, mul, use_min, sc);
    else score_query<D, 3, false>(s, idx, lane, mul, use_min, sc);
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
    q = q1;
    q1 = q2;
    c0 = c1;
    c1 = c2;
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
  const int64_t cap = (int64_t)sms * (D >= 256 ? GQE_VEC_BLOCKS_D256 : 4);
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
