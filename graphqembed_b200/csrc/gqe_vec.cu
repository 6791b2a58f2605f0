// gqe_vec.cu -- streaming kernel of the contraction-free operators.
//
// With the TransE or DistMult metapath decoder (reference netquery/decoders.py:181-208,
// 211-236) every relation operator is a vector add / multiply, and with
// SimpleSetIntersection (decoders.py:302-319) the intersection is an element-wise mean / min:
// the whole forward path of a query -- gather, L2 normalisation (encoders.py:41-43), the
// operators, the cosine (raw dot for DistMult chains) against the positive and the negative
// target and the hinge (model.py:112-127) -- touches (A + T) table rows of 4d bytes and does
// O(d) flops on them.  It is bound by the HBM gather, so there is no tile, no shared memory
// and no block-level synchronisation here: a 16-lane HALF WARP owns one query (a warp scores two
// queries of the same formula side by side, so every shuffle butterfly, index lookup and scalar
// epilogue instruction is shared by two queries -- the kernel was issue-bound with a warp per
// query: 626 warp instructions per query, 59 % issue-slot use at 0.44 of the HBM roofline),
// issues the 128-bit loads of all its rows at once (up to 5 rows x d/64 float4 per lane in
// flight), keeps everything in registers and reduces with shuffles inside the half.  Blocks walk
// over groups of 16 consecutive queries of one formula, so the structure branches are uniform;
// the indices of the next groups are fetched (and mapped through the node map) while the rows of
// the current one are in flight.  Norms and cosines use MUFU.RSQ forms (~2^-22 relative) instead
// of IEEE sqrt / division, whose checked slow paths cost ~10 instructions each.
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

__device__ __forceinline__ double vsum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
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

// x / max(|x|, eps) and friends, branch-free (no IEEE slow paths: a sqrtf / division each costs a
// checked call sequence, ~10 instructions, and this kernel is issue-bound).  ~2^-22 relative.
//   unit_dot(d, nn)  = d / sqrt(nn)        (nn = 0 -> 0 * inf = NaN, the reference's 0 / 0)
//   clamped_norm(nn) = max(sqrt(nn), eps)  (keeps NaN)
__device__ __forceinline__ float unit_dot(float d, float nn) { return d * rsqrtf(nn); }
__device__ __forceinline__ float clamped_norm(float nn) {
  const float n = nn > 0.f ? nn * rsqrtf(nn) : nn;
  return n >= kCosEps ? n : (n != n ? n : kCosEps);
}
// torch.min over the stacked operands propagates NaN: one instruction
__device__ __forceinline__ float min_nan(float a, float b) {
  float r;
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}

constexpr int kHalf = 16;                 // lanes per query
constexpr int kQW = 32 / kHalf;           // queries per warp
constexpr int kQB = kVecWarps * kQW;      // queries per block and iteration (a "group")

// butterflies of several values inside each 16-lane half, in lockstep
template <int N>
__device__ __forceinline__ void hsum_n(float (&v)[N]) {
#pragma unroll
  for (int o = kHalf / 2; o > 0; o >>= 1) {
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
  }
}

// One query per 16-lane half of the warp (two queries of the SAME formula per warp), everything in
// registers.  `idx`: sub-lanes 0..NA-1 of each half hold the anchor rows of its query, sub-lanes 8, 9
// the target rows (9 = 8 when the query has one target).  Returns the half's two scores.
template <int D, int NA, bool CHAIN>
__device__ __forceinline__ void score_query(const SegDev& s, const float (*rel_sm)[D], int32_t idx, int sl, bool mul,
                                            bool use_min, float (&sc)[2]) {
  constexpr int LANES = D / 4 < kHalf ? D / 4 : kHalf;   // d = 32: the upper sub-lanes idle
  constexpr int NV = D / 4 / LANES;                      // float4 per lane per row
  const bool act = LANES == kHalf || sl < LANES;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  auto ld = [&](const float* base, int j) {
    return act ? __ldg(reinterpret_cast<const float4*>(base) + sl + LANES * j) : zero4;
  };
  // table rows are read once: keep them out of L1, so that what IS re-read (relation vectors, node maps) stays
  auto ldrow = [&](const float* base, int j) {
    float4 v = zero4;
    if (act)
      asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                   : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                   : "l"(reinterpret_cast<const float4*>(base) + sl + LANES * j));
    return v;
  };
  // relation vector h of the formula, from the block's shared-memory copy (see the kernel)
  auto ldr = [&](int h, int j) {
    return act ? *(reinterpret_cast<const float4*>(rel_sm[h]) + sl + LANES * j) : zero4;
  };
  // ---- all row loads of both queries, back to back
  float4 a[NA][NV], t[2][NV];
#pragma unroll
  for (int b = 0; b < NA; ++b) {
    const float* src = s.anc_table[b] + (size_t)__shfl_sync(0xffffffffu, idx, b, kHalf) * D;
#pragma unroll
    for (int j = 0; j < NV; ++j) a[b][j] = ldrow(src, j);
  }
#pragma unroll
  for (int tt = 0; tt < 2; ++tt) {
    const float* src = s.tgt_table + (size_t)__shfl_sync(0xffffffffu, idx, 8 + tt, kHalf) * D;
#pragma unroll
    for (int j = 0; j < NV; ++j) t[tt][j] = ldrow(src, j);
  }
  if (CHAIN) {
    // cos(a_hat, t_hat (+|*) v_r1 ... v_rn), raw dot for DistMult (decoders.py:200-205,228-233).
    // a_hat has unit norm, so the cosine's own |a_hat| is 1 (to 1e-7) and is not recomputed.
    float n2[3] = {0.f, 0.f, 0.f};               // |a|^2, |t0|^2, |t1|^2
#pragma unroll
    for (int j = 0; j < NV; ++j) { n2[0] = sq4(a[0][j], n2[0]); n2[1] = sq4(t[0][j], n2[1]); n2[2] = sq4(t[1][j], n2[2]); }
    hsum_n(n2);
    // x / |x| as x * rsqrt(|x|^2): |x| = 0 gives 0 * inf = NaN like the reference's 0 / 0
    const float ia = rsqrtf(n2[0]), it0 = rsqrtf(n2[1]), it1 = rsqrtf(n2[2]);
    const int hops = s.structure + 1;
    float r[4] = {0.f, 0.f, 0.f, 0.f};           // y0.a_hat, |y0|^2, y1.a_hat, |y1|^2
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float4 y0 = scale4(t[0][j], it0), y1 = scale4(t[1][j], it1);
      for (int h = 0; h < hops; ++h) {
        const float4 v = ldr(h, j);
        y0 = rel4(y0, v, mul);
        y1 = rel4(y1, v, mul);
      }
      if (!act) { y0 = zero4; y1 = zero4; }
      const float4 ah = scale4(a[0][j], ia);
      r[0] = dot4(y0, ah, r[0]); r[1] = sq4(y0, r[1]);
      r[2] = dot4(y1, ah, r[2]); r[3] = sq4(y1, r[3]);
    }
    hsum_n(r);
    sc[0] = mul ? r[0] : __fdividef(r[0], clamped_norm(r[1]));
    sc[1] = mul ? r[2] : __fdividef(r[2], clamped_norm(r[3]));
  } else {
    // q = agg_b project(a_hat_b) [projected once more for 3-chain_inter]; cos(t_hat, q)  (model.py:77-109)
    float n2[NA];
#pragma unroll
    for (int b = 0; b < NA; ++b) {
      n2[b] = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) n2[b] = sq4(a[b][j], n2[b]);
    }
    hsum_n(n2);
    float ib[NA];
#pragma unroll
    for (int b = 0; b < NA; ++b) ib[b] = rsqrtf(n2[b]);
    const int structure = s.structure;
    float r[5] = {0.f, 0.f, 0.f, 0.f, 0.f};      // |q|^2, |t0|^2, t0.q, |t1|^2, t1.q
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float4 q;
#pragma unroll
      for (int b = 0; b < NA; ++b) {
        float4 e = scale4(a[b][j], ib[b]);
        if (NA == 2 && b == 1 && structure == GQE_INTER_CHAIN3) {
          e = rel4(e, ldr(1, j), mul);     // reverse(r2b) first (model.py:85)
          e = rel4(e, ldr(2, j), mul);     // then reverse(r2a)
        } else {
          e = rel4(e, ldr(b, j), mul);
        }
        if (b == 0) q = e;
        else if (use_min) q = make_float4(min_nan(q.x, e.x), min_nan(q.y, e.y), min_nan(q.z, e.z), min_nan(q.w, e.w));
        else q = make_float4(q.x + e.x, q.y + e.y, q.z + e.z, q.w + e.w);
      }
      if (!use_min) q = scale4(q, NA == 2 ? 0.5f : 1.f / 3.f);                    // torch.mean over the stack
      if (NA == 2 && structure == GQE_CHAIN_INTER3) q = rel4(q, ldr(2, j), mul);   // model.py:107
      if (!act) q = zero4;
      r[0] = sq4(q, r[0]);
      r[1] = sq4(t[0][j], r[1]); r[2] = dot4(t[0][j], q, r[2]);
      r[3] = sq4(t[1][j], r[3]); r[4] = dot4(t[1][j], q, r[4]);
    }
    hsum_n(r);
    const float nq = clamped_norm(r[0]);
    // t_hat = t/|t| has unit norm: cos(t_hat, q) = (t.q/|t|) / max(|q|, eps); a zero target row gives
    // 0/0 = NaN as in the reference
    sc[0] = __fdividef(unit_dot(r[2], r[1]), nq);
    sc[1] = __fdividef(unit_dot(r[4], r[3]), nq);
  }
}

#ifndef GQE_VEC_BLOCKS_D256
#define GQE_VEC_BLOCKS_D256 2
#endif
template <int D>
__global__ void __launch_bounds__(kVecThreads, (D >= 256 ? GQE_VEC_BLOCKS_D256 : 3)) gqe_fused_vec(const __grid_constant__ LaunchParams p) {
  __shared__ double red[kVecWarps];
  __shared__ int last;
  // the (<= 3) relation vectors of the formula the block is working on: every query of the formula reads
  // them in the middle of its dependent arithmetic, where an L1 miss (the streaming rows evict them) was
  // 12 % of all stall samples; reloaded when the block moves on to another formula
  // (two buffers: the next formula's vectors are fetched into registers while the current group is scored)
  __shared__ __align__(16) float rel_sm[2][GQE_MAX_RELS][D];
  constexpr int kRelPer = (GQE_MAX_RELS * D + kVecThreads - 1) / kVecThreads;   // values per thread
  int rel_of = -1, rel_buf = 0;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int sl = lane & (kHalf - 1);
  const int qslot = wib * kQW + lane / kHalf;     // this half-warp's query inside a group
  const bool mul = p.decoder == GQE_DEC_DISTMULT;
  const bool use_min = p.inter == GQE_INTER_DEEPSETS_MIN || p.inter == GQE_INTER_SIMPLE_MIN;
  const int T = p.T;
  const int ik = p.index_kind;
  double local = 0.0;

  // Iteration space: GROUPS of kQB consecutive queries of one formula (the last group of a formula
  // may be partly empty).  Block b takes the group PAIRS b, b + gridDim.x, ... (two consecutive groups
  // mostly share their formula: half as many relation-vector switches as with single groups).  Which
  // formula a group belongs to depends on blockIdx and the loop counter only, so every branch on the
  // query structure below is uniform across the block: no divergence, the warp shuffles need no
  // convergence barriers.
  // Per-segment facts the bookkeeping needs every iteration, as 32-bit words in shared memory (the
  // launch parameters live in the constant bank, where a dynamically indexed 64-bit field costs a
  // dependent LDC each: 16 % of the stall samples went there).
  __shared__ int sg_gbase[kMaxSegs + 1];   // first group of segment i (prefix sum), [n_segs] = total
  __shared__ int sg_q0[kMaxSegs], sg_nq[kMaxSegs];
  __shared__ int sg_mode[kMaxSegs];        // anchor modes in bytes 0..2, target mode in byte 3, n_anchor in bits 28..31
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < p.n_segs; ++i) {
      const SegDev& sd = p.seg[i];
      sg_gbase[i] = acc;
      sg_q0[i] = (int)sd.q_begin;
      sg_nq[i] = (int)(sd.q_end - sd.q_begin);
      acc += (sg_nq[i] + kQB - 1) / kQB;
      sg_mode[i] = (sd.anc_mode[0] & 15) | ((sd.anc_mode[1] & 15) << 4) | ((sd.anc_mode[2] & 15) << 8) |
                   ((sd.tgt_mode & 15) << 12) | (sd.n_anchor << 28);
    }
    sg_gbase[p.n_segs] = acc;
  }
  __syncthreads();
  const int total = sg_gbase[p.n_segs];
  const int G = (int)gridDim.x;
  auto group_of = [&](int t) -> int { return (((t >> 1) * G + (int)blockIdx.x) << 1) | (t & 1); };   // t-th group of this block
  struct Cursor { int si; };                                // segment of a group; only moves forward
  auto locate = [&](Cursor& c, int g) -> bool {            // false past the end
    if (g >= total) return false;
    while (g >= sg_gbase[c.si + 1]) ++c.si;
    return true;
  };
  // this half-warp's query of group g (-1: none)
  auto query_of = [&](const Cursor& c, int g, bool in) -> int64_t {
    if (!in) return -1;
    const int ql = (g - sg_gbase[c.si]) * kQB + qslot;
    return ql < sg_nq[c.si] ? (int64_t)(sg_q0[c.si] + ql) : -1;
  };
  // sub-lanes 0..na-1 of a half hold the anchor indices of its query, sub-lanes 8, 9 its target indices
  // (9 repeats target 0 when T == 1).
  const int tslot = sl == 9 && T > 1 ? 1 : 0;
  auto lane_mode = [&](int si) -> int {                    // this lane's node type in a query of segment si, -1: none
    const int w = sg_mode[si];
    if (sl < (int)((unsigned)w >> 28)) return (w >> (4 * sl)) & 15;
    if (sl == 8 || sl == 9) return (w >> 12) & 15;
    return -1;
  };
  auto fetch_raw = [&](int64_t q, int si) -> int32_t {
    if (q < 0) return 0;
    if (sl < (int)((unsigned)sg_mode[si] >> 28)) return __ldg(p.anchor_rows + (int64_t)sl * p.anchor_stride + q);
    if (sl == 8 || sl == 9) return __ldg(p.target_rows + q * T + tslot);
    return 0;
  };
  // Three stages, one loop iteration apart, so that no dependent load is ever waited for: the raw
  // index of group i+2, the node-map lookup of group i+1, the bounds check + row loads of group i.
  // [Measured, not kept: requesting the NEXT groups' rows into L2 (prefetch.global.L2 right behind
  // the lookup, or one iteration later with a four-stage pipeline) is 4-8 % slower on both the
  // 10 GB and the L2-resident tables -- profiles/r02_vec_experiments.md.]
  Cursor c0{0}, c1{0}, c2{0};
  int t = 0;
  bool in0 = locate(c0, group_of(0)), in1 = locate(c1, group_of(1));
  int64_t q = query_of(c0, group_of(0), in0), q1 = query_of(c1, group_of(1), in1);
  int32_t raw_cur = in0 ? fetch_raw(q, c0.si) : 0, raw_nxt = in1 ? fetch_raw(q1, c1.si) : 0;
  int32_t cand = 0;
  if (q >= 0) {
    const int m = lane_mode(c0.si);
    if (m >= 0) cand = index_lookup(p.mode[m], raw_cur, ik);
  }
  for (; in0; ++t) {
    const SegDev& s = p.seg[c0.si];
    int32_t idx = 0;
    if (q >= 0) {
      const int my_mode = lane_mode(c0.si);
      if (my_mode >= 0) idx = index_check(p.mode[my_mode], my_mode, cand, raw_cur, ik, p.err);
    }
    // the later groups' indices travel while this group's rows do
    const int g2 = group_of(t + 2);
    const bool in2 = locate(c2, g2);
    const int64_t q2 = query_of(c2, g2, in2);
    const int32_t raw_nxt2 = in2 ? fetch_raw(q2, c2.si) : 0;
    int32_t cand_nxt = 0;
    if (q1 >= 0) {
      const int m = lane_mode(c1.si);
      if (m >= 0) cand_nxt = index_lookup(p.mode[m], raw_nxt, ik);
    }
    if (rel_of != c0.si) {             // the block's first group (block-uniform: c0 depends on blockIdx and the loop counter only)
      for (int i = threadIdx.x; i < GQE_MAX_RELS * D; i += kVecThreads) {
        const float* r = s.rel[i / D];
        rel_sm[rel_buf][i / D][i % D] = r ? __ldg(r + i % D) : 0.f;
      }
      __syncthreads();
      rel_of = c0.si;
    }
    // the NEXT group belongs to another formula: its vectors travel while this group is scored
    const bool rel_next = in1 && c1.si != rel_of;
    float rel_pf[kRelPer];
    if (rel_next) {
#pragma unroll
      for (int u = 0; u < kRelPer; ++u) {
        const int i = threadIdx.x + u * kVecThreads;
        const float* r = i < GQE_MAX_RELS * D ? p.seg[c1.si].rel[i / D] : nullptr;
        rel_pf[u] = r ? __ldg(r + i % D) : 0.f;
      }
    }
    float sc[2];
    if (s.structure <= GQE_CHAIN3) score_query<D, 1, true>(s, rel_sm[rel_buf], idx, sl, mul, use_min, sc);
    else if (s.n_anchor == 2) score_query<D, 2, false>(s, rel_sm[rel_buf], idx, sl, mul, use_min, sc);
    else score_query<D, 3, false>(s, rel_sm[rel_buf], idx, sl, mul, use_min, sc);
    if (rel_next) {                    // (the other buffer was last read before the previous switch's barrier)
#pragma unroll
      for (int u = 0; u < kRelPer; ++u) {
        const int i = threadIdx.x + u * kVecThreads;
        if (i < GQE_MAX_RELS * D) rel_sm[rel_buf ^ 1][i / D][i % D] = rel_pf[u];
      }
      __syncthreads();
      rel_buf ^= 1;
      rel_of = c1.si;
    }
    if (sl == 0 && q >= 0) {
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
    in0 = in1;
    in1 = in2;
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
  for (int i = 0; i < lp.n_segs; ++i)    // (the kernel keeps query positions as 32-bit words)
    if (lp.seg[i].q_end >= 0x7fffffffLL) return cudaErrorInvalidValue;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // never more CTAs than 64-row tiles (the margin-loss partials are sized by the tile count)
  const int64_t cap = (int64_t)sms * (D >= 256 ? GQE_VEC_BLOCKS_D256 : 3);
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
