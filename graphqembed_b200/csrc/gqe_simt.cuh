// gqe_simt.cuh -- fp32 CUDA-core kernels of the conjunctive-query scorer.
//
// One CTA owns a tile of 64 rows (queries, or (query,target) pairs for chain
// structures) and keeps every intermediate embedding of that tile on chip, in
// the reference's own feature-major orientation X[k][row] (reference
// netquery/encoders.py:41 transposes to [d, B]; every projection is
// mats[rel].mm(embeds), decoders.py:150).  Gather, L2 normalisation, the chained
// relation operators, the DeepSets / elementwise intersection, the cosine
// against every target and the hinge loss are all evaluated inside that one
// launch; nothing but indices is read and nothing but scores (or the scalar
// loss) is written.
//
// This file is the exact-fp32 path: FFMA for the d x d contractions, used for
// every decoder/intersection kind and every supported dimension.  The
// tensor-core (tcgen05) path for the Bilinear/DeepSets contractions lives in
// gqe_tc.cuh and is checked against this one.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "gqe_params.h"

namespace gqe {

template <int D>
struct __align__(16) TileSmem {
  float X[D][kRowStride];
  float A[D][kRowStride];
  float P[2][kPanelK][D + 1];
  float score[kTileRows];
  double red[kWarps];
  int last;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// torch semantics: relu / min / clamp propagate NaN (fmaxf/fminf would not).
__device__ __forceinline__ float relu_nan(float x) { return x < 0.f ? 0.f : x; }
__device__ __forceinline__ float min_nan(float a, float b) { return (a < b || a != a) ? a : b; }

// ---------------------------------------------------------------------------
// DirectEncoder (reference netquery/encoders.py:41-43): gather up to 64 table
// rows, divide each by its L2 norm (true division, no epsilon -> a zero row
// gives NaN exactly like the reference) and park them feature-major in smem.
// A warp reads one row per instruction group, fully coalesced.
template <int D>
__device__ __forceinline__ void gather_normalise(float (*X)[kRowStride], const float* __restrict__ table,
                                                 const int32_t* __restrict__ rows, int n_valid, const ModeDev& md,
                                                 int mode, int index_kind, unsigned long long* err) {
  constexpr int TC = D / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int rr = 0; rr < kRowsPerWarp; rr += 4) {
    float v[4][TC];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = warp * kRowsPerWarp + rr + u;
      const bool ok = r < n_valid;
      const size_t row = ok ? (size_t)resolve_index(md, mode, __ldg(rows + r), index_kind, err) : 0;
      const float* src = table + row * D + lane;
#pragma unroll
      for (int j = 0; j < TC; ++j) v[u][j] = ok ? __ldg(src + 32 * j) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int r = warp * kRowsPerWarp + rr + u;
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < TC; ++j) ss = fmaf(v[u][j], v[u][j], ss);
      ss = warp_sum(ss);
      const float nrm = sqrtf(ss);
      const bool ok = r < n_valid;
#pragma unroll
      for (int j = 0; j < TC; ++j) X[lane + 32 * j][r] = ok ? __fdiv_rn(v[u][j], nrm) : 0.f;
    }
  }
}

// ---------------------------------------------------------------------------
// In-place tile contraction  X[n][r] <- epi( sum_k B(k,n) * X[k][r] ).
//   TRANS == false : B(k,n) = M[k*D+n]   act.mm(M) of the chain decoder (decoders.py:145)
//   TRANS == true  : B(k,n) = M[n*D+k]   M.mm(embeds) of project / pre / post
//                                         (decoders.py:150,289,299)
// Thread (warp w, lane l) owns rows 8w..8w+7 and columns l+32j.  The operand
// matrix streams through a double-buffered 16-deep smem panel; it is shared by
// every CTA and stays L2-resident.
enum { EPI_NONE = 0, EPI_RELU = 1 };

template <int D, bool TRANS, int EPI>
__device__ __forceinline__ void tile_matmul(float (*X)[kRowStride], float (*P)[kPanelK][D + 1],
                                            const float* __restrict__ M) {
  constexpr int TC = D / 32;
  constexpr int PER_T = kPanelK * D / kThreads;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc[kRowsPerWarp][TC];
#pragma unroll
  for (int i = 0; i < kRowsPerWarp; ++i)
#pragma unroll
    for (int j = 0; j < TC; ++j) acc[i][j] = 0.f;

  float stage[PER_T];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < PER_T; ++i) {
      const int idx = threadIdx.x + i * kThreads;
      if (!TRANS) {
        stage[i] = __ldg(M + (size_t)(k0 + idx / D) * D + (idx % D));
      } else {
        stage[i] = __ldg(M + (size_t)(idx / kPanelK) * D + k0 + (idx % kPanelK));
      }
    }
  };
  auto park = [&](int buf) {
#pragma unroll
    for (int i = 0; i < PER_T; ++i) {
      const int idx = threadIdx.x + i * kThreads;
      if (!TRANS) P[buf][idx / D][idx % D] = stage[i];
      else        P[buf][idx % kPanelK][idx / kPanelK] = stage[i];
    }
  };

  fetch(0);
  park(0);
  __syncthreads();
  for (int k0 = 0; k0 < D; k0 += kPanelK) {
    const int buf = (k0 / kPanelK) & 1;
    const bool more = k0 + kPanelK < D;
    if (more) fetch(k0 + kPanelK);
#pragma unroll
    for (int kk = 0; kk < kPanelK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&X[k0 + kk][warp * kRowsPerWarp]);
      const float4 a1 = *reinterpret_cast<const float4*>(&X[k0 + kk][warp * kRowsPerWarp + 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[TC];
#pragma unroll
      for (int j = 0; j < TC; ++j) b[j] = P[buf][kk][lane + 32 * j];
#pragma unroll
      for (int i = 0; i < kRowsPerWarp; ++i)
#pragma unroll
        for (int j = 0; j < TC; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) park(buf ^ 1);
    __syncthreads();
  }
  // every warp has finished reading X (barrier above): overwrite in place
#pragma unroll
  for (int j = 0; j < TC; ++j) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = (EPI == EPI_RELU) ? relu_nan(acc[i][j]) : acc[i][j];
    float* dst = &X[lane + 32 * j][warp * kRowsPerWarp];
    *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
  __syncthreads();
}

// TransE / DistMult relation operators: X[k][r] (+|*)= v[k]
// (decoders.py:203,208 / decoders.py:231,236).
template <int D>
__device__ __forceinline__ void tile_vec_op(float (*X)[kRowStride], const float* __restrict__ v, bool mul) {
  for (int idx = threadIdx.x; idx < D * kTileRows; idx += kThreads) {
    const int k = idx / kTileRows, r = idx % kTileRows;
    const float w = __ldg(v + k);
    X[k][r] = mul ? X[k][r] * w : X[k][r] + w;
  }
  __syncthreads();
}

// project() of whichever decoder the model uses.
template <int D>
__device__ __forceinline__ void tile_project(float (*X)[kRowStride], float (*P)[kPanelK][D + 1],
                                             const float* __restrict__ param, int decoder) {
  if (decoder == GQE_DEC_BILINEAR) tile_matmul<D, true, EPI_NONE>(X, P, param);
  else tile_vec_op<D>(X, param, decoder == GQE_DEC_DISTMULT);
}

__device__ __forceinline__ int64_t query_of_pair(const LaunchParams& p, int64_t pair, int64_t n_queries) {
  if (p.target_offsets == nullptr) return pair / p.T;
  int64_t lo = 0, hi = n_queries;  // largest q with offsets[q] <= pair
  while (hi - lo > 1) {
    const int64_t mid = (lo + hi) >> 1;
    if (__ldg(p.target_offsets + mid) <= pair) lo = mid; else hi = mid;
  }
  return lo;
}

// Hinge of model.py:124-125 on (positive, negative) score pairs, NaN-propagating
// like torch.clamp; returns this thread's contribution.
__device__ __forceinline__ float hinge(float margin, float pos, float neg) {
  const float h = margin - (pos - neg);
  return h < 0.f ? 0.f : h;
}

// Deterministic grid reduction of the per-CTA hinge sums (model.py:126 mean).
template <int D>
__device__ __forceinline__ void loss_reduce(const LaunchParams& p, TileSmem<D>& sm, double local) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  local = warp_sum(local);
  if (lane == 0) sm.red[warp] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kWarps; ++w) s += sm.red[w];
    p.partials[blockIdx.x] = s;
    __threadfence();
    const unsigned int t = atomicAdd(p.ticket, 1u);
    sm.last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (sm.last && warp == 0) {
    __threadfence();
    double s = 0.0;
    for (unsigned int i = lane; i < gridDim.x; i += 32) s += __ldcg(p.partials + i);
    s = warp_sum(s);
    if (lane == 0) {
      const double acc = *p.loss_acc + s;
      *p.loss_acc = p.final_launch ? 0.0 : acc;
      *p.out_loss = (float)(acc * p.inv_q);
      *p.ticket = 0u;
    }
  }
}

// ---------------------------------------------------------------------------
// 1/2/3-chain (model.py:71-76 -> path_dec.forward): rows are (query, target)
// pairs; the relation operators act on the TARGET side.
template <int D>
__device__ __forceinline__ void chain_tile(const LaunchParams& p, const SegDev& s, int n_hops,
                                           int64_t tile_in_seg, TileSmem<D>& sm) {
  constexpr int TC = D / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool ragged = p.target_offsets != nullptr;
  const int64_t pair_begin = ragged ? 0 : s.q_begin * p.T;
  const int64_t pair_end = ragged ? p.n_pairs : s.q_end * p.T;
  const int64_t base = pair_begin + tile_in_seg * kTileRows;
  const int n_valid = (int)min((int64_t)kTileRows, pair_end - base);

  gather_normalise<D>(sm.X, s.tgt_table, p.target_rows + base, n_valid, p.mode[s.tgt_mode], s.tgt_mode, p.index_kind, p.err);
  __syncthreads();
  for (int h = 0; h < n_hops; ++h) {
    if (p.decoder == GQE_DEC_BILINEAR) tile_matmul<D, false, EPI_NONE>(sm.X, sm.P, s.rel[h]);
    else tile_vec_op<D>(sm.X, s.rel[h], p.decoder == GQE_DEC_DISTMULT);
  }

  // score each pair against its query's (normalised) anchor
#pragma unroll 2
  for (int rr = 0; rr < kRowsPerWarp; ++rr) {
    const int r = warp * kRowsPerWarp + rr;
    if (r >= n_valid) break;
    const int64_t pair = base + r;
    const int64_t q = s.q_begin + query_of_pair(p, pair - pair_begin, s.q_end - s.q_begin);
    const size_t arow = (size_t)resolve_index(p.mode[s.anc_mode[0]], s.anc_mode[0], __ldg(p.anchor_rows + q), p.index_kind, p.err);
    const float* src = s.anc_table[0] + arow * D + lane;
    float a[TC], y[TC];
    float sa = 0.f;
#pragma unroll
    for (int j = 0; j < TC; ++j) { a[j] = __ldg(src + 32 * j); sa = fmaf(a[j], a[j], sa); }
    sa = warp_sum(sa);
    const float na = sqrtf(sa);
    float dot = 0.f, sy = 0.f, sah = 0.f;
#pragma unroll
    for (int j = 0; j < TC; ++j) {
      const float ah = __fdiv_rn(a[j], na);
      y[j] = sm.X[lane + 32 * j][r];
      dot = fmaf(y[j], ah, dot);
      sy = fmaf(y[j], y[j], sy);
      sah = fmaf(ah, ah, sah);
    }
    dot = warp_sum(dot);
    float score;
    if (p.decoder == GQE_DEC_DISTMULT) {
      score = dot;  // raw dot product, no cosine (decoders.py:232)
    } else {
      sy = warp_sum(sy);
      sah = warp_sum(sah);
      score = dot / (fmaxf(sqrtf(sy), kCosEps) * fmaxf(sqrtf(sah), kCosEps));
    }
    if (lane == 0) {
      if (p.out_scores) p.out_scores[pair] = score;
      sm.score[r] = score;
    }
  }
  if (p.out_loss) {
    __syncthreads();
    double local = 0.0;
    if (threadIdx.x < kTileRows / 2 && 2 * (int)threadIdx.x + 1 < n_valid)
      local = (double)hinge(p.margin, sm.score[2 * threadIdx.x], sm.score[2 * threadIdx.x + 1]);
    loss_reduce<D>(p, sm, local);
  }
}

// ---------------------------------------------------------------------------
// 2-inter / 3-inter / 3-inter_chain / 3-chain_inter (model.py:77-109): rows are
// queries; the query embedding is built once and scored against every target.
template <int D>
__device__ __forceinline__ void inter_tile(const LaunchParams& p, const SegDev& s, int structure,
                                           int64_t tile_in_seg, TileSmem<D>& sm) {
  constexpr int TC = D / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t q0 = s.q_begin + tile_in_seg * kTileRows;
  const int n_valid = (int)min((int64_t)kTileRows, s.q_end - q0);
  const bool deepsets = p.inter == GQE_INTER_DEEPSETS_MEAN || p.inter == GQE_INTER_DEEPSETS_MIN;
  const bool use_min = p.inter == GQE_INTER_DEEPSETS_MIN || p.inter == GQE_INTER_SIMPLE_MIN;
  const int n_branch = s.n_anchor;

  for (int b = 0; b < n_branch; ++b) {
    gather_normalise<D>(sm.X, s.anc_table[b], p.anchor_rows + (int64_t)b * p.anchor_stride + q0, n_valid,
                        p.mode[s.anc_mode[b]], s.anc_mode[b], p.index_kind, p.err);
    __syncthreads();
    if (structure == GQE_INTER_CHAIN3 && b == 1) {
      tile_project<D>(sm.X, sm.P, s.rel[1], p.decoder);  // reverse(r2b) first (model.py:85)
      tile_project<D>(sm.X, sm.P, s.rel[2], p.decoder);  // then reverse(r2a)
    } else {
      tile_project<D>(sm.X, sm.P, s.rel[b], p.decoder);
    }
    if (deepsets) tile_matmul<D, true, EPI_RELU>(sm.X, sm.P, s.pre);  // relu(pre.mm(e)) decoders.py:289-292
    for (int idx = threadIdx.x; idx < D * kTileRows; idx += kThreads) {
      const int k = idx / kTileRows, r = idx % kTileRows;
      const float x = sm.X[k][r];
      float a;
      if (b == 0) a = x;
      else if (use_min) a = min_nan(sm.A[k][r], x);
      else a = sm.A[k][r] + x;
      if (!use_min && b == n_branch - 1) a = __fdiv_rn(a, (float)n_branch);  // torch.mean over the stack
      sm.A[k][r] = a;
    }
    __syncthreads();
  }
  if (deepsets) tile_matmul<D, true, EPI_NONE>(sm.A, sm.P, s.post);  // post.mm(combined) decoders.py:299
  if (structure == GQE_CHAIN_INTER3) tile_project<D>(sm.A, sm.P, s.rel[2], p.decoder);  // model.py:107

  if (p.q_out) {  // many targets per query: hand the query embedding rows to gqe_score_pairs
    for (int idx = threadIdx.x; idx < D * kTileRows; idx += kThreads) {
      const int r = idx / D, k = idx % D;
      if (r < n_valid) p.q_out[(size_t)(q0 + r) * D + k] = sm.A[k][r];
    }
    return;
  }
  double local = 0.0;
#pragma unroll 1
  for (int rr = 0; rr < kRowsPerWarp; ++rr) {
    const int r = warp * kRowsPerWarp + rr;
    if (r >= n_valid) break;
    const int64_t q = q0 + r;
    float qv[TC];
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < TC; ++j) { qv[j] = sm.A[lane + 32 * j][r]; sq = fmaf(qv[j], qv[j], sq); }
    sq = warp_sum(sq);
    const float nq = fmaxf(sqrtf(sq), kCosEps);
    int64_t t0, t1;
    if (p.target_offsets) { t0 = __ldg(p.target_offsets + q); t1 = __ldg(p.target_offsets + q + 1); }
    else { t0 = q * p.T; t1 = t0 + p.T; }
    float s0 = 0.f, s1 = 0.f;
    for (int64_t t = t0; t < t1; ++t) {
      const size_t trow = (size_t)resolve_index(p.mode[s.tgt_mode], s.tgt_mode, __ldg(p.target_rows + t), p.index_kind, p.err);
      const float* src = s.tgt_table + trow * D + lane;
      float st = 0.f, dq = 0.f;
#pragma unroll
      for (int j = 0; j < TC; ++j) {
        const float x = __ldg(src + 32 * j);
        st = fmaf(x, x, st);
        dq = fmaf(x, qv[j], dq);
      }
      st = warp_sum(st);
      dq = warp_sum(dq);
      // t_hat = t/|t| has unit norm, so cos(t_hat, q) = (t.q/|t|) / max(|q|, eps);
      // a zero target row gives 0/0 = NaN as in the reference.
      const float score = __fdiv_rn(dq, sqrtf(st)) / nq;
      if (lane == 0 && p.out_scores) p.out_scores[t] = score;
      if (t == t0) s0 = score; else if (t == t0 + 1) s1 = score;
    }
    if (p.out_loss && lane == 0) local += (double)hinge(p.margin, s0, s1);
  }
  if (p.out_loss) loss_reduce<D>(p, sm, local);
}

// ---------------------------------------------------------------------------
// Kernels.  STRUCT >= 0: the single-formula kernel of that query structure
// (one kernel per structure); STRUCT < 0: the grouped kernel, which looks the
// structure of its segment up at run time (CTA-uniform branch).
template <int D, int STRUCT>
__global__ void __launch_bounds__(kThreads, (D <= 128 ? 2 : 1)) gqe_fused_simt(const __grid_constant__ LaunchParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<D>& sm = *reinterpret_cast<TileSmem<D>*>(smem_raw);
  int si = 0;
  if (STRUCT < 0) {
    for (int i = 1; i < p.n_segs; ++i)
      if ((int64_t)blockIdx.x >= p.seg[i].tile_begin) si = i;
  }
  const SegDev& s = p.seg[si];
  const int structure = STRUCT >= 0 ? STRUCT : s.structure;
  const int64_t tile_in_seg = (int64_t)blockIdx.x - s.tile_begin;
  if (structure <= GQE_CHAIN3) chain_tile<D>(p, s, structure + 1, tile_in_seg, sm);
  else inter_tile<D>(p, s, structure, tile_in_seg, sm);
}

// ---------------------------------------------------------------------------
// Operator-level kernels (the un-fused reference surface) on feature-major
// [d, n] tensors.  Same device building blocks, one tile of 64 columns per CTA.
template <int D>
__device__ __forceinline__ void load_fm(float (*X)[kRowStride], const float* __restrict__ src, int64_t n,
                                        int64_t c0, int n_valid) {
  for (int idx = threadIdx.x; idx < D * kTileRows; idx += kThreads) {
    const int k = idx / kTileRows, r = idx % kTileRows;
    X[k][r] = r < n_valid ? src[(size_t)k * n + c0 + r] : 0.f;
  }
}
template <int D>
__device__ __forceinline__ void store_fm(float* __restrict__ dst, float (*X)[kRowStride], int64_t n, int64_t c0,
                                         int n_valid) {
  for (int idx = threadIdx.x; idx < D * kTileRows; idx += kThreads) {
    const int k = idx / kTileRows, r = idx % kTileRows;
    if (r < n_valid) dst[(size_t)k * n + c0 + r] = X[k][r];
  }
}

// cosine_similarity(dim=0, eps) of two smem tiles, one column per warp pass
template <int D>
__device__ __forceinline__ void tile_cosine(float (*U)[kRowStride], float (*V)[kRowStride], float* out,
                                            int64_t c0, int n_valid, bool raw_dot) {
  constexpr int TC = D / 32;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int rr = 0; rr < kRowsPerWarp; ++rr) {
    const int r = warp * kRowsPerWarp + rr;
    if (r >= n_valid) break;
    float dot = 0.f, su = 0.f, sv = 0.f;
#pragma unroll
    for (int j = 0; j < TC; ++j) {
      const float u = U[lane + 32 * j][r], v = V[lane + 32 * j][r];
      dot = fmaf(u, v, dot); su = fmaf(u, u, su); sv = fmaf(v, v, sv);
    }
    dot = warp_sum(dot); su = warp_sum(su); sv = warp_sum(sv);
    const float sc = raw_dot ? dot : dot / (fmaxf(sqrtf(su), kCosEps) * fmaxf(sqrtf(sv), kCosEps));
    if (lane == 0) out[c0 + r] = sc;
  }
}

template <int D>
__global__ void __launch_bounds__(kThreads, (D <= 128 ? 2 : 1)) gqe_op_simt(const __grid_constant__ OpParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  TileSmem<D>& sm = *reinterpret_cast<TileSmem<D>*>(smem_raw);
  const int64_t c0 = (int64_t)blockIdx.x * kTileRows;
  const int n_valid = (int)min((int64_t)kTileRows, p.n - c0);
  switch (p.op) {
    case OP_ENCODE:
      {
        ModeDev md;
        md.lut = nullptr; md.base = 0; md.len = 0; md.rows = (uint32_t)p.table_rows; md.pad_ = 0;
        gather_normalise<D>(sm.X, p.table, p.rows + c0, n_valid, md, 0, 0, p.err);
      }
      __syncthreads();
      store_fm<D>(p.out, sm.X, p.n, c0, n_valid);
      break;
    case OP_PROJECT:
      load_fm<D>(sm.X, p.in0, p.n, c0, n_valid);
      __syncthreads();
      tile_project<D>(sm.X, sm.P, p.rel[0], p.decoder);
      store_fm<D>(p.out, sm.X, p.n, c0, n_valid);
      break;
    case OP_PATH_SCORE:
      load_fm<D>(sm.X, p.io0, p.n, c0, n_valid);
      load_fm<D>(sm.A, p.in1, p.n, c0, n_valid);
      __syncthreads();
      for (int h = 0; h < p.n_rels; ++h) {
        if (p.decoder == GQE_DEC_BILINEAR) tile_matmul<D, false, EPI_NONE>(sm.X, sm.P, p.rel[h]);
        else tile_vec_op<D>(sm.X, p.rel[h], p.decoder == GQE_DEC_DISTMULT);
      }
      if (p.mutate) store_fm<D>(p.io0, sm.X, p.n, c0, n_valid);
      tile_cosine<D>(sm.X, sm.A, p.out, c0, n_valid, p.decoder == GQE_DEC_DISTMULT);
      break;
    case OP_INTERSECT: {
      const bool deepsets = p.inter == GQE_INTER_DEEPSETS_MEAN || p.inter == GQE_INTER_DEEPSETS_MIN;
      const bool use_min = p.inter == GQE_INTER_DEEPSETS_MIN || p.inter == GQE_INTER_SIMPLE_MIN;
      const float* parts[3] = {p.in0, p.in1, p.in2};
      const int n_parts = p.in2 ? 3 : 2;
      for (int b = 0; b < n_parts; ++b) {
        load_fm<D>(sm.X, parts[b], p.n, c0, n_valid);
        __syncthreads();
        if (deepsets) tile_matmul<D, true, EPI_RELU>(sm.X, sm.P, p.pre);
        for (int idx = threadIdx.x; idx < D * kTileRows; idx += kThreads) {
          const int k = idx / kTileRows, r = idx % kTileRows;
          const float x = sm.X[k][r];
          float a;
          if (b == 0) a = x;
          else if (use_min) a = min_nan(sm.A[k][r], x);
          else a = sm.A[k][r] + x;
          if (!use_min && b == n_parts - 1) a = __fdiv_rn(a, (float)n_parts);
          sm.A[k][r] = a;
        }
        __syncthreads();
      }
      if (deepsets) tile_matmul<D, true, EPI_NONE>(sm.A, sm.P, p.post);
      store_fm<D>(p.out, sm.A, p.n, c0, n_valid);
      break;
    }
    case OP_MATMUL:  // out = W in (mutate == 0) or W^T in (mutate != 0), any [d,d] matrix (rel[0])
      load_fm<D>(sm.X, p.in0, p.n, c0, n_valid);
      __syncthreads();
      if (p.mutate) tile_matmul<D, false, EPI_NONE>(sm.X, sm.P, p.rel[0]);
      else tile_matmul<D, true, EPI_NONE>(sm.X, sm.P, p.rel[0]);
      store_fm<D>(p.out, sm.X, p.n, c0, n_valid);
      break;
    case OP_COSINE:
      load_fm<D>(sm.X, p.in0, p.n, c0, n_valid);
      load_fm<D>(sm.A, p.in1, p.n, c0, n_valid);
      __syncthreads();
      tile_cosine<D>(sm.X, sm.A, p.out, c0, n_valid, false);
      break;
    default:
      break;
  }
}

}  // namespace gqe
