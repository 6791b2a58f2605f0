// gqe_tc.cuh -- tensor-core (tcgen05 / TMEM / TMA) kernels of the
// conjunctive-query scorer for the Bilinear decoder at d = 128 / 256.
//
// One CTA owns a tile of 128 rows (queries, or (query,target) pairs for chain
// structures) = one UMMA M.  Every d x d contraction of the path -- the chained
// relation projections (reference netquery/decoders.py:142-150) and the DeepSets
// pre / post matrices (decoders.py:288-300) -- is a [128 x d] x [d x d] product on
// the 5th-generation tensor cores:
//
//   * the activations live in shared memory as two bf16 planes (hi, lo with
//     x ~= hi + lo) in the UMMA canonical K-major SWIZZLE_128B layout;
//   * the relation matrix is pre-split and pre-swizzled by gqe_pack (same layout,
//     hi and lo planes per 64-wide K block) and streamed L2 -> smem by TMA bulk
//     copies through a 3-stage mbarrier ring;
//   * one elected thread issues tcgen05.mma (kind::f16, M=128, N=d, K=16) three
//     times per K step -- hi*hi, lo*hi, hi*lo -- accumulating in fp32 in TMEM.
//     Dropping lo*lo leaves a relative error of ~2^-17 per product, two orders
//     below the 1e-4 parity bound on the cosine score (tests/test_gpu_parity.py);
//   * the worker warps read the accumulator back with tcgen05.ld (thread == row),
//     apply the epilogue (re-split for the next hop, ReLU + mean/min aggregation
//     kept in a second TMEM region, or the cosine / margin loss against the gathered
//     target rows) and never write an intermediate to global memory;
//   * chain tiles are scored straight from TMEM in the 16x256b fragment layout
//     (score_frag), one tile late, in the shadow of the next tile's contraction; the two
//     TMEM regions swap roles from tile to tile;
//   * the kernel is launched programmatically dependent on gqe_pack (only the TMA
//     producer waits for it).
//
// Warp roles: warps [0, W) workers (W = d/16: four TMEM lane quarters x d/64 column
// groups, thread == one row x 64 columns), warp W = TMA producer, warp W+1 = MMA
// issuer + TMEM owner; in the STAGE instantiation (d = 256, grouped kernel, tables sharded
// over several GPUs) warp W+2 = helper: fetches the operand rows that live in a PEER GPU's
// HBM by TMA, up to two tiles ahead of the workers, into a local staging area (helper()).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "gqe_params.h"
#include "gqe_ptx.cuh"

namespace gqe {
namespace tc {

constexpr int kRows = 128;     // rows per tile = UMMA M
constexpr int kKBlock = 64;    // bf16 per 128-byte swizzle row
constexpr int kStages = 3;     // weight ring depth

template <int D>
struct Cfg {
  static constexpr int kColsPerThread = 64;    // each worker thread owns 64 columns of one row
  static constexpr int kColGroups = D / kColsPerThread;
  static constexpr int kWorkerWarps = 4 * kColGroups;  // 4 TMEM lane quarters x column groups
  static constexpr int kWorkerThreads = kWorkerWarps * 32;
  static constexpr int kThreads = kWorkerThreads + 64;
  // STAGE instantiation (node-type-sharded tables, d = 256): one more warp -- a CTA's 18 warps are allocated
  // as 20 anyway -- that fetches the rows of the NEXT tiles from PEER GPUs into a local staging area while the
  // current tile is worked on (helper()).  It moves rows with TMA bulk copies through the third weight stage's
  // 32 KiB of shared memory (the weight ring of this instantiation has two stages): kStageRingRows rows in flight
  static constexpr int kHelperWarps = 1;
  static constexpr int kStageRingRows = 32;
  static_assert(kStageRingRows * D * 4 <= D * 128, "the helper's row ring lives in one weight stage");
  static constexpr int kStageOps = 5;      // staged operands of a tile: first gather, anchors 1, 2, scoring rows 0, 1
  static constexpr int kStageDepth = 4;    // tiles a staged row set stays valid for (ring look-ahead 2 + deferred chain score)
  static constexpr size_t kStageFloatsPerCta = (size_t)kStageDepth * kStageOps * kRows * D;
  static constexpr int kKB = D / kKBlock;
  static constexpr int kABlockBytes = kRows * 128;     // one 64-wide K block of the A tile
  static constexpr int kAPlaneBytes = kKB * kABlockBytes;
  static constexpr int kStageBytes = D * 128;          // one plane of one K block of the weights
  static constexpr int kOffAhi = 0;
  static constexpr int kOffAlo = kAPlaneBytes;
  static constexpr int kOffB = 2 * kAPlaneBytes;
  static constexpr int kOffCtl = kOffB + kStages * kStageBytes;
  static constexpr int kCtlBytes = 1024;
  static constexpr int kOffScratch = kOffCtl + kCtlBytes;  // [128 rows][3] fp32 partial sums (score_frag)
  static constexpr int kScratchBytes = kRows * 3 * 4;
  static constexpr int kSmemBytes = kOffScratch + kScratchBytes;
  static_assert(kSmemBytes <= 232448, "dynamic shared memory per CTA");
  static constexpr int kTmemCols = 2 * D;              // accumulator + aggregation region
  static constexpr int kPackedBytes = 4 * D * D;       // one packed matrix (2 planes x bf16)
  static constexpr int kCtasPerSm = (D <= 128) ? 2 : 1;
};

struct Ctl {
  uint64_t full[kStages + 1];    // (the STAGE instantiation runs a ring of four half stages)
  uint64_t empty[kStages + 1];
  uint64_t a_ready;
  uint64_t acc_full;
  uint64_t sched_full[2];   // tile-id ring: scheduler (producer thread) -> workers, MMA issuer
  uint64_t sched_empty[2];
  double red[16];
  int64_t tile_id[2];
  uint32_t tmem_base;
  int last;
  uint8_t bad[32];          // per worker warp: bit 0 = a lane saw a bad index in this tile, bit 1 = in the carried row
  uint64_t stage_ready[4][5];   // STAGE: helper warp -> workers: staged operand o of tile (k mod 4) has landed
  uint64_t row_bar[32];         // STAGE: TMA completion of the helper's row-ring slots
};
static_assert(sizeof(Ctl) <= 1024, "control block");

// (the per-structure program -- Prog, build_program, the G_* / M_* / E_* / F_* codes -- lives in
// gqe_params.h: the host builds it once per formula into SegDev::prog)
__device__ __forceinline__ const uint8_t* step_matrix(const SegDev& s, int mat) {
  const float* p = mat <= M_REL2 ? s.rel[mat] : (mat == M_PRE ? s.pre : s.post);
  return reinterpret_cast<const uint8_t*>(p);  // packed bf16 planes on this path
}

// 16 bytes of a table row: read once per tile, so kept out of L1 (what IS re-read through L1 -- index
// arrays, node maps -- then stays there).  Works on peer (NVLink-mapped) addresses like a plain load.
__device__ __forceinline__ float4 ld_row(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// the same from the staging area the helper warps fill while this kernel runs: through L2, never the
// non-coherent path, never L1 (a staging slot is rewritten every four tiles)
__device__ __forceinline__ float4 ld_row_staged(const float4* p) {
  float4 v;
  asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
// Which operands of a tile are fetched from a peer GPU by the helper warps (bit o of the result; the
// workers and the helpers evaluate this for every tile of the CTA, in the same order):
//   o = 0 first gather (chain: targets, else anchor 0)   1, 2 anchors 1, 2   3, 4 scoring rows (chain:
//   anchor 0; else the two targets)
// The CTA's FIRST tile (kseq = 0) is gathered in place: nothing could have been fetched ahead of it, and the
// workers' own 64 KiB in flight beat the helper's 32 on a single operand.
__device__ __forceinline__ uint32_t staged_ops(uint32_t kseq, bool chain, uint32_t rm, int n_branch, int T, bool q_out) {
  if (kseq == 0) return 0u;
  if (chain) return ((rm & 8u) ? 1u : 0u) | ((rm & 1u) ? 8u : 0u);
  uint32_t m = (rm & 1u) | (rm & 2u) | ((n_branch > 2 && (rm & 4u)) ? 4u : 0u);
  if ((rm & 8u) && !q_out && T <= 2) m |= 8u | (T > 1 ? 16u : 0u);
  return m;
}

// ---- small math helpers -----------------------------------------------------------
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float relu_nan_(float x) { return x < 0.f ? 0.f : x; }
__device__ __forceinline__ float min_nan_(float a, float b) { return (a < b || a != a) ? a : b; }

// x0,x1 -> packed bf16 hi pair and packed bf16 lo pair with x ~= hi + lo
__device__ __forceinline__ void split2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
  const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - __low2float(h), x1 - __high2float(h));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// byte offset of (row r, 16-byte chunk c of K block kb) inside an A plane
__device__ __forceinline__ uint32_t a_chunk_off(int r, int kb, int c) {
  return (uint32_t)(kb * (kRows * 128) + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4));
}

// ---- DirectEncoder: gather, L2-normalise, split, park as the A operand ---------------
// (reference netquery/encoders.py:41-43; true division, no epsilon.)  A warp reads
// one table row per instruction group with 128-bit loads, fully coalesced.
// `my_row` holds, in lane i, the table row of tile row wid*RPW + i (-1: past the end);
// the index loads were issued at tile start so only the row loads are exposed here.
// (Measured: 8 rows in flight at d = 256 -- as a non-inlined function, so that the loads do not
// spill -- is SLOWER, 4.4 us per gather against 3.6: the phase is not bound by the loads a warp
// has in flight.)
template <int D, bool STAGED = false>
__device__ __forceinline__ void gather_to_a(uint8_t* smem, const float* __restrict__ table, int32_t my_row, int wid,
                                            int lane) {
  using C = Cfg<D>;
  constexpr int RPW = kRows / C::kWorkerWarps;
  constexpr int NV = D / 128;
  constexpr int U = D >= 256 ? 4 : 8;  // rows in flight per warp (register budget)
#pragma unroll 1
  for (int r0 = 0; r0 < RPW; r0 += U) {
    float4 v[U][NV];
    float ss[U];
    bool okv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int32_t row = __shfl_sync(0xffffffffu, my_row, r0 + u);
      okv[u] = row >= 0;
      const float4* src = reinterpret_cast<const float4*>(table + (size_t)(okv[u] ? row : 0) * D);
#pragma unroll
      for (int j = 0; j < NV; ++j)
        v[u][j] = okv[u] ? (STAGED ? ld_row_staged(src + lane + 32 * j) : ld_row(src + lane + 32 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      ss[u] = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        ss[u] = fmaf(v[u][j].x, v[u][j].x, ss[u]);
        ss[u] = fmaf(v[u][j].y, v[u][j].y, ss[u]);
        ss[u] = fmaf(v[u][j].z, v[u][j].z, ss[u]);
        ss[u] = fmaf(v[u][j].w, v[u][j].w, ss[u]);
      }
    }
    // the U butterflies run in lockstep so that their shuffle latencies overlap
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
      for (int u = 0; u < U; ++u) ss[u] += __shfl_xor_sync(0xffffffffu, ss[u], o);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int r = wid * RPW + r0 + u;
      // x / |x| as x * rsqrt(|x|^2): branch-free (MUFU.RSQ, ~2^-22 relative, far below the
      // bf16x3 error); |x| = 0 -> inf -> 0 * inf = NaN, the reference's 0/0
      const float inv = okv[u] ? rsqrtf(ss[u]) : 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        uint2 hi, lo;
        split2(v[u][j].x * inv, v[u][j].y * inv, hi.x, lo.x);
        split2(v[u][j].z * inv, v[u][j].w * inv, hi.y, lo.y);
        // element k = 128 j + 4 lane: K block 2j + lane/16, chunk (lane%16)/2, half (lane&1)
        const uint32_t off = a_chunk_off(r, 2 * j + (lane >> 4), (lane & 15) >> 1) + ((lane & 1) << 3);
        *reinterpret_cast<uint2*>(smem + C::kOffAhi + off) = hi;
        *reinterpret_cast<uint2*>(smem + C::kOffAlo + off) = lo;
      }
    }
  }
}

// Branch-free cosine pieces of the scoring phase (the compiler cannot interleave rows across
// the slow-path branches of sqrtf / IEEE division).  All ~2^-22 relative, far below 1e-4.
//   unit_dot(d, nn) = d / sqrt(nn)   (nn = 0 -> 0 * inf = NaN, the reference's 0/0)
//   clamped_norm(nn) = max(sqrt(nn), eps)
__device__ __forceinline__ float unit_dot(float d, float nn) { return d * rsqrtf(nn); }
__device__ __forceinline__ float clamped_norm(float nn) {
  const float n = nn > 0.f ? nn * rsqrtf(nn) : nn;  // keeps 0 -> 0 and NaN -> NaN
  return n >= kCosEps ? n : (n != n ? n : kCosEps);
}

// 16 fp32 values of (row r, columns col0..col0+15) -> the A operand planes
template <int D>
__device__ __forceinline__ void store_a16(uint8_t* smem, int r, int col0, const float (&x)[16]) {
  using C = Cfg<D>;
  const int kb = col0 >> 6, c0 = (col0 & 63) >> 3;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    uint4 hi, lo;
    split2(x[8 * j + 0], x[8 * j + 1], hi.x, lo.x);
    split2(x[8 * j + 2], x[8 * j + 3], hi.y, lo.y);
    split2(x[8 * j + 4], x[8 * j + 5], hi.z, lo.z);
    split2(x[8 * j + 6], x[8 * j + 7], hi.w, lo.w);
    const uint32_t off = a_chunk_off(r, kb, c0 + j);
    *reinterpret_cast<uint4*>(smem + C::kOffAhi + off) = hi;
    *reinterpret_cast<uint4*>(smem + C::kOffAlo + off) = lo;
  }
}

__device__ __forceinline__ float hinge_(float margin, float pos, float neg) {
  const float h = margin - (pos - neg);
  return h < 0.f ? 0.f : h;
}

// Deterministic two-level reduction of the per-CTA hinge sums (model.py:126 mean);
// executed by the worker threads only.
// ---- tile bookkeeping -------------------------------------------------------------------
// The kernel is PERSISTENT: one CTA per SM slot walks over tiles handed out by a dynamic
// scheduler (heaviest structures first, see gqe_capi.cu).  A tile id maps to its segment by
// a scan of the (<= 32) tile_begin offsets.
template <int STRUCT>
__device__ __forceinline__ int seg_of_tile(const LaunchParams& p, int64_t tile) {
  int si = 0;
  if (STRUCT < 0) {
    for (int i = 1; i < p.n_segs; ++i)
      if (tile >= p.seg[i].tile_begin) si = i;
  }
  return si;
}

// What travels through the tile ring: the tile id in the low 40 bits, its segment above.  The scheduler
// thread does the segment scan ONCE per tile (it needs the segment itself); the 576 consumers would
// otherwise each repeat it -- up to 32 dynamically indexed constant-bank loads per tile and thread.
__device__ __forceinline__ int64_t ring_pack(int64_t tile, int seg) { return tile | ((int64_t)seg << 40); }
__device__ __forceinline__ int64_t ring_tile(int64_t v) { return v & ((1ll << 40) - 1); }
__device__ __forceinline__ int ring_seg(int64_t v) { return (int)(v >> 40); }

struct TileRing {  // consumer side of the tile-id ring
  uint32_t k = 0;
  __device__ __forceinline__ int64_t peek(Ctl* ctl, uint32_t ahead) const {
    const uint32_t kk = k + ahead;
    ptx::mbar_wait(ptx::smem_u32(&ctl->sched_full[kk & 1]), (kk >> 1) & 1);
    return *reinterpret_cast<volatile int64_t*>(&ctl->tile_id[kk & 1]);
  }
  __device__ __forceinline__ int64_t take(Ctl* ctl) {
    const int64_t t = peek(ctl, 0);
    ptx::mbar_arrive(ptx::smem_u32(&ctl->sched_empty[k & 1]));
    ++k;
    return t;
  }
};

// Final margin-loss reduction, once per CTA after its last tile: per-tile hinge sums were
// written to p.partials[tile]; the last CTA to finish adds them up in tile order, so the
// result does not depend on which CTA ran which tile (model.py:126 mean).
template <int D>
__device__ __forceinline__ void loss_finish(const LaunchParams& p, Ctl* ctl, uint8_t* smem, int wid, int lane) {
  using C = Cfg<D>;
  // (launched programmatically dependent on the previous kernel in the stream -- with the weights
  // cached that is the previous call's fused kernel, whose last CTA reset the words read here)
  if (wid == 0) ptx::griddep_wait();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int t = atomicAdd(p.ticket, 1u);
    ctl->last = (t == gridDim.x - 1);
  }
  ptx::named_bar_sync(1, C::kWorkerThreads);
  if (!ctl->last) return;
  __threadfence();
  // every worker thread of the last CTA fetches its share of the per-tile sums (one or two independent
  // loads each instead of n_tiles / 32 dependent ones per lane of one warp), fixed assignment and order
  double* red2 = reinterpret_cast<double*>(smem);   // the A planes are dead by now
  if (p.out_loss) {
    double s = 0.0;
    for (int64_t i = threadIdx.x; i < p.n_tiles; i += C::kWorkerThreads) s += __ldcg(p.partials + i);
    red2[threadIdx.x] = s;
    ptx::named_bar_sync(1, C::kWorkerThreads);
  }
  if (wid == 0) {
    if (p.out_loss) {
      double s = 0.0;
      for (int i = lane; i < C::kWorkerThreads; i += 32) s += red2[i];
      s = warp_sum_d(s);
      if (lane == 0) {
        const double acc = *p.loss_acc + s;
        *p.loss_acc = p.final_launch ? 0.0 : acc;
        *p.out_loss = (float)(acc * p.inv_q);
      }
    }
    if (lane == 0) {  // leave the scheduler state ready for the next launch
      *p.ticket = 0u;
      *p.tile_counter = 0u;
      if (p.err_host) {   // *_host calls: the first bad index (if any) travels with the result
        p.err_host[1] = __ldcg(p.err + 1);
        p.err_host[0] = __ldcg(p.err);
      }
    }
  }
}

// float4 chunk c of row r of the transposed query tile (row stride D floats, XOR-swizzled so
// both the thread-per-row writes and the warp-per-row reads are bank-conflict free)
template <int D>
__device__ __forceinline__ float4* q_chunk(float* qsm, int r, int c) {
  return reinterpret_cast<float4*>(qsm + (size_t)r * D) + (c ^ (r & 7));
}

// ---- chain scoring straight from the accumulator ---------------------------------------
// The final accumulator row of a chain tile is the projected target y; it is scored against
// the anchor row of its query (decoders.py:146):
//     s = (y . a / |a|) / max(|y|, eps)
// The accumulator is read from TMEM in the 16x256b fragment layout (4 lanes share a row) and
// the anchor rows with 128-bit loads in the matching layout (score_col_src: every 32-byte
// sector fetched is fully used), so nothing goes through shared memory: the A planes stay free
// and the tile is scored while the NEXT tile's first contraction is on the tensor pipe.  Warp
// (quarter q, half hh, column half ch) owns TMEM lanes 32q+16hh .. +15 and 128 columns; at
// d = 256 the two column halves of a row are combined through `scratch` in a fixed order
// (deterministic).  [The same scheme for the intersections' two target rows per query is
// slower than the transposed tile: 8 distinct lines per load instruction, measured 8-9 us per
// tile against 7.4.]
//
// Column permutation of a scored accumulator (applied to the weights by gqe_pack): lane t owns
// accumulator columns 8j + 2(t%4) + {0,1}; with accumulator column 16b + 8j' + 2m + e holding
// output column 16b + 4m + 2j' + e, the four columns a lane owns in repeats (2b, 2b+1) are the
// CONTIGUOUS output columns 16b + 4m .. +3.
__host__ __device__ constexpr int score_col_src(int n) {
  return (n & ~15) | (((n >> 1) & 3) << 2) | (((n >> 3) & 1) << 1) | (n & 1);
}

struct PendTile {
  int64_t tile;
  int64_t row_begin;
  const float* table;   // the anchor table
  int32_t idx[2];       // anchor rows of this thread's two accumulator rows
  int32_t n_valid;
  uint32_t region;      // TMEM region (0 / 1) holding the tile's final accumulator
  bool valid;
  uint32_t stage_info;  // STAGE: 0, or `table` is the tile's staging slot and idx = positions: bit 0 set, bit 1 = the parity
                        // of stage_ready[slot][3] to wait for, bits 2-3 = slot
};

// this thread's two accumulator rows in the fragment layout
__device__ __forceinline__ int frag_row(int wid, int lane, int rs) {
  return 32 * (wid & 3) + 16 * ((wid >> 2) & 1) + (lane >> 2) + 8 * rs;
}

template <int D, bool STAGE = false>
__device__ __forceinline__ void score_frag(const LaunchParams& p, Ctl* ctl, float* scratch, uint32_t tmem_base,
                                           const PendTile& pt, int wid, int lane) {
  using C = Cfg<D>;
  const bool st_rows = STAGE && pt.stage_info != 0;
  if (st_rows) ptx::mbar_wait(ptx::smem_u32(&ctl->stage_ready[(pt.stage_info >> 2) & 3u][3]), (pt.stage_info >> 1) & 1u);
  constexpr int NCH = C::kColGroups / 2;   // warps sharing one 16-lane half (1 at d=128, 2 at d=256)
  constexpr int COLS = D / NCH;            // columns per warp (128)
  const int ch = wid >> 3;
  const int g = lane >> 2, c4 = 4 * (lane & 3);
  int rr[2];
  bool ok[2];
  const float* pr[2];
#pragma unroll
  for (int rs = 0; rs < 2; ++rs) {
    rr[rs] = frag_row(wid, lane, rs);
    ok[rs] = rr[rs] < pt.n_valid;
    pr[rs] = pt.table + (size_t)(ok[rs] ? pt.idx[rs] : 0) * D + COLS * ch + c4;
  }
  float yy[2] = {0.f, 0.f}, dt[2] = {0.f, 0.f}, nn[2] = {0.f, 0.f};
  const uint32_t taddr = tmem_base + pt.region * D + ((uint32_t)(32 * (wid & 3) + 16 * ((wid >> 2) & 1)) << 16) + COLS * ch;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int c = 0; c < COLS / 32; c += 2) {   // two 32-column rounds of loads in flight
    float4 av[2][2][2];  // [round][row slot][16-column block]
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int rs = 0; rs < 2; ++rs)
#pragma unroll
        for (int b = 0; b < 2; ++b)
          av[r][rs][b] = ok[rs] ? (st_rows ? ld_row_staged(reinterpret_cast<const float4*>(pr[rs] + 32 * (c + r) + 16 * b))
                                           : ld_row(reinterpret_cast<const float4*>(pr[rs] + 32 * (c + r) + 16 * b)))
                                : zero4;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      uint32_t raw[16];
      ptx::tmem_ld_16x256b_x4(taddr + 32 * (c + r), raw);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int rs = 0; rs < 2; ++rs) {
          // repeats 2b, 2b+1 of this row slot = output columns 16b + c4 .. +3 (score_col_src)
          const float y0 = __uint_as_float(raw[8 * b + 2 * rs]), y1 = __uint_as_float(raw[8 * b + 2 * rs + 1]);
          const float y2 = __uint_as_float(raw[8 * b + 4 + 2 * rs]), y3 = __uint_as_float(raw[8 * b + 4 + 2 * rs + 1]);
          const float4 a = av[r][rs][b];
          dt[rs] = fmaf(y0, a.x, dt[rs]); dt[rs] = fmaf(y1, a.y, dt[rs]);
          dt[rs] = fmaf(y2, a.z, dt[rs]); dt[rs] = fmaf(y3, a.w, dt[rs]);
          nn[rs] = fmaf(a.x, a.x, nn[rs]); nn[rs] = fmaf(a.y, a.y, nn[rs]);
          nn[rs] = fmaf(a.z, a.z, nn[rs]); nn[rs] = fmaf(a.w, a.w, nn[rs]);
          yy[rs] = fmaf(y0, y0, yy[rs]); yy[rs] = fmaf(y1, y1, yy[rs]);
          yy[rs] = fmaf(y2, y2, yy[rs]); yy[rs] = fmaf(y3, y3, yy[rs]);
        }
    }
  }
  // the four lanes of a row
#pragma unroll
  for (int o = 1; o <= 2; o <<= 1)
#pragma unroll
    for (int rs = 0; rs < 2; ++rs) {
      yy[rs] += __shfl_xor_sync(0xffffffffu, yy[rs], o);
      dt[rs] += __shfl_xor_sync(0xffffffffu, dt[rs], o);
      nn[rs] += __shfl_xor_sync(0xffffffffu, nn[rs], o);
    }
  if (NCH > 1) {  // second column half -> scratch -> first column half (fixed order)
    if (ch == 1 && (lane & 3) == 0) {
#pragma unroll
      for (int rs = 0; rs < 2; ++rs) {
        float* dst = scratch + 3 * rr[rs];
        dst[0] = yy[rs]; dst[1] = dt[rs]; dst[2] = nn[rs];
      }
    }
    ptx::named_bar_sync(1, C::kWorkerThreads);
    if (ch == 0) {
#pragma unroll
      for (int rs = 0; rs < 2; ++rs) {
        const float* src = scratch + 3 * rr[rs];
        yy[rs] += src[0]; dt[rs] += src[1]; nn[rs] += src[2];
      }
    }
  }
  double local = 0.0;
#pragma unroll
  for (int rs = 0; rs < 2; ++rs) {
    // a_hat = a/|a| has unit norm: cos = (y.a/|a|) / max(|y|, eps); a zero anchor row gives
    // 0/0 = NaN as in the reference
    const float s0 = __fdividef(unit_dot(dt[rs], nn[rs]), clamped_norm(yy[rs]));
    const float s0_next = __shfl_xor_sync(0xffffffffu, s0, 4);   // accumulator row + 1 (same slot, next lane group)
    if (ch == 0 && (lane & 3) == 0) {
      if (ok[rs] && p.out_scores) p.out_scores[pt.row_begin + rr[rs]] = s0;
      // rows (2i, 2i+1) = (pos, neg) of one query (T == 2 whenever the loss is on)
      if (!(g & 1) && rr[rs] + 1 < pt.n_valid) local += (double)hinge_(p.margin, s0, s0_next);
    }
  }
  if (p.out_loss) {
    local = warp_sum_d(local);
    if (lane == 0) ctl->red[wid] = local;
    ptx::named_bar_sync(1, C::kWorkerThreads);
    if (threadIdx.x == 0) {
      double sum = 0.0;
      for (int w = 0; w < C::kWorkerWarps; ++w) sum += ctl->red[w];
      p.partials[pt.tile] = sum;
    }
  } else if (NCH == 1) {
    // every thread's TMEM reads are done before the next tile's epilogue may overwrite the region
    ptx::named_bar_sync(1, C::kWorkerThreads);
  }
}

// diagnostics: CTA-level stamps (%globaltimer ns, clock64) in the LAST 256 records of the phase
// log: record phase_cap-1-blockIdx.x, slots 2i / 2i+1 for point i (0 entry, 1 set-up done, 2 first
// tile taken, 3 tile loop left, 4 exit)
__device__ __forceinline__ void cta_stamp(const LaunchParams& p, int point) {
  if (!p.phase_log || p.phase_cap < 512 || blockIdx.x >= 256) return;
  unsigned long long* rec = p.phase_log + (size_t)(p.phase_cap - 1 - blockIdx.x) * kPhaseSlots;
  unsigned long long gt;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
  rec[2 * point] = gt;
  rec[2 * point + 1] = (unsigned long long)clock64();
}

// L2 prefetch of the table rows a warp will need later (lane i < RPW holds the row of tile row
// wid*RPW + i, -1: none).  One 128-byte line per lane and instruction through the load/store
// path: the warp's RPW rows are RPW * D/32 = 64 lines = two instructions with all lanes active.
// [The bulk-copy prefetch (cp.async.bulk.prefetch.L2, one per row) is issued lane by lane --
// ~0.5 us of warp time per instruction -- and queues in the TMA unit in front of the weight
// stages of the running contraction.]
template <int D>
__device__ __forceinline__ void prefetch_rows_l2(const float* __restrict__ table, int32_t my_row, int lane) {
  constexpr int LINES = D / 32;   // 128-byte lines per row
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int idx = lane + 32 * i;
    const int32_t row = __shfl_sync(0xffffffffu, my_row, idx / LINES);
    if (row >= 0) ptx::prefetch_l2_line(table + (size_t)row * D + (idx % LINES) * 32);
  }
}

// Cold path of the index checks.  The hot path only notes THAT a lane saw an index outside its
// table / node map (and substitutes row 0); this re-resolves every index of the tile rows of one
// warp, slowly, to report which one it was (LaunchParams::err).
template <int D>
__device__ __noinline__ void diagnose_indices(const LaunchParams& p, const SegDev& s, bool chain, int64_t row_begin,
                                              int n_valid, int wid, int lane) {
  constexpr int RPW = kRows / Cfg<D>::kWorkerWarps;
  const int T = p.T, ik = p.index_kind;
  const int r = wid * RPW + lane;
  if (lane >= RPW || r >= n_valid) return;
  const int64_t row = row_begin + r;
  if (chain) {
    resolve_index(p.mode[s.tgt_mode], s.tgt_mode, __ldg(p.target_rows + row), ik, p.err);
    resolve_index(p.mode[s.anc_mode[0]], s.anc_mode[0], __ldg(p.anchor_rows + (T == 2 ? row >> 1 : row / T)), ik, p.err);
  } else {
    for (int k = 0; k < s.n_anchor; ++k)
      resolve_index(p.mode[s.anc_mode[k]], s.anc_mode[k], __ldg(p.anchor_rows + k * p.anchor_stride + row), ik, p.err);
    if (!p.q_out)
      for (int t = 0; t < T; ++t)
        resolve_index(p.mode[s.tgt_mode], s.tgt_mode, __ldg(p.target_rows + row * T + t), ik, p.err);
  }
}

// ---- worker warps ---------------------------------------------------------------------
template <int D, int STRUCT, bool STAGE = false>
__device__ __forceinline__ void worker(const LaunchParams& p, uint8_t* smem, Ctl* ctl) {
  using C = Cfg<D>;
  constexpr int NCH = C::kColsPerThread / 16;      // 16-column TMEM chunks per thread
  constexpr int RPW = kRows / C::kWorkerWarps;     // tile rows per warp in the warp-per-row phases
  constexpr int NV = D / 128;                      // float4 per lane per table row
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = 32 * (wid & 3) + lane;           // TMEM lane == tile row (thread-per-row phases)
  const int col_base = (wid >> 2) * C::kColsPerThread;
  const bool deepsets = p.inter == GQE_INTER_DEEPSETS_MEAN || p.inter == GQE_INTER_DEEPSETS_MIN;
  const bool use_min = p.inter == GQE_INTER_DEEPSETS_MIN || p.inter == GQE_INTER_SIMPLE_MIN;
  const int T = p.T;
  const uint32_t tmem_base = ctl->tmem_base;
  const uint32_t t_lane = tmem_base + ((uint32_t)(32 * (wid & 3)) << 16) + col_base;
  float* scratch = reinterpret_cast<float*>(smem + C::kOffScratch);
  const uint32_t bar_a_ready = ptx::smem_u32(&ctl->a_ready);
  const uint32_t bar_acc_full = ptx::smem_u32(&ctl->acc_full);
  const int my_r = wid * RPW + lane;               // meaningful for lane < RPW
  float* qsm = reinterpret_cast<float*>(smem);     // overlays the A planes (exactly 128 x D fp32)

  TileRing ring;
  uint32_t gs = 0;                                 // steps issued so far (a_ready / acc_full parity)
  PendTile pend;                                   // tile whose scoring was deferred into the next tile
  pend.valid = false;
  int32_t carry_row = -1;                          // this lane's first-gather row of the NEXT tile, resolved one tile ahead
  if (lane == 0) ctl->bad[wid] = 0;
  for (;;) {
    const int64_t tile_v = ring.take(ctl);
    const int64_t tile = ring_tile(tile_v);
    if (tile >= p.n_tiles) break;
    const SegDev& s = p.seg[ring_seg(tile_v)];
    const int structure = STRUCT >= 0 ? STRUCT : s.structure;
    const bool chain = structure <= GQE_CHAIN3;
    const int n_branch = s.n_anchor;
    // TMEM: two regions of D columns, accumulator + running aggregate of the intersections.
    // The roles alternate from tile to tile (the MMA issuer applies the same rule), so that the
    // final accumulator of tile k can be scored while tile k+1's first contraction runs (its
    // aggregate region, where tile k's result lives, is first written after that contraction).
    const uint32_t region = (ring.k - 1) & 1u;
    // chains are scored straight from TMEM (score_frag), intersections through the transposed
    // tile in shared memory
    const bool frag = chain;
    const uint32_t t_acc = t_lane + region * D;
    const uint32_t t_agg = t_lane + (region ^ 1u) * D;

    // rows of this tile
    const int64_t row_begin = (chain ? s.q_begin * T : s.q_begin) + (tile - s.tile_begin) * kRows;
    const int64_t row_end = chain ? s.q_end * T : s.q_end;
    const int n_valid = (int)min((int64_t)kRows, row_end - row_begin);

    const Prog& pg = s.prog;

    // diagnostics: per-tile phase stamps of worker thread 0 (tools/phase_report.py)
    int n_stamp = 0;
    unsigned long long* plog = nullptr;
    if (threadIdx.x == 0 && ring.k == 1) cta_stamp(p, 2);
    if (p.phase_log && threadIdx.x == 0 && tile < p.phase_cap - 256) {
      plog = p.phase_log + (size_t)tile * kPhaseSlots;
      uint32_t smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      plog[kPhaseSlots - 1] = smid;
    }
    auto stamp = [&](unsigned long long tag) {
      if (plog && n_stamp < kPhaseSlots - 1)
        plog[n_stamp++] = (tag << 56) | ((unsigned long long)clock64() & 0x00ffffffffffffffull);
    };
    stamp(1 + 16 * (unsigned long long)structure);

    // ---- every index this warp will need (lane i <-> tile row wid*RPW + i) in three stages kept
    // apart by long phases, so that no dependent load is waited for:
    //   here            the RAW index loads (caller rows or node ids), all independent;
    //   after the first A operand is handed over: the node-map lookups (index_lookup, one load
    //                   each, only when the caller passed node ids) ...
    //   ... and, after the deferred score of the previous tile, the bounds checks and the L2
    //                   prefetches of the LOCAL table rows behind them.
    // The first gather's own rows were resolved the same way while the PREVIOUS tile ran and
    // arrive in a register (carry_row).  (Peer shards are never prefetched: a bulk L2 prefetch of
    // a PEER address is ~100x slower than the NVLink loads it would hide --
    // profiles/r01_peer_gather_micro.md)
    const bool mine = lane < RPW && my_r < n_valid;
    const uint32_t rm = s.remote_mask;
    const int ik = p.index_kind;
    // STAGE: the operands of this tile that the helper warps fetch from peer GPUs into the CTA's staging slot
    // (tile number kseq of this CTA; the helpers count the same way) -- read from there, by POSITION in the tile
    // (all of it recomputed where it is used: nothing here stays live across the tile.)  The helper arrives on
    // EVERY stage_ready[k mod 4][o] once per tile, staged or not, so the parity to wait for is bit 2 of k.
    auto staged = [&](int o) -> bool {
      return STAGE && p.stage_on == 1 && (((staged_ops(ring.k - 1, chain, rm, n_branch, T, p.q_out != nullptr) & p.stage_mask) >> o) & 1u);
    };
    auto stage_ptr = [&](int o) -> const float* {
      return p.stage + (((size_t)blockIdx.x * C::kStageDepth + ((ring.k - 1) & 3u)) * C::kStageOps + (size_t)o) * (size_t)(kRows * D);
    };
    auto stage_wait = [&](int o) {
      ptx::mbar_wait(ptx::smem_u32(&ctl->stage_ready[(ring.k - 1) & 3u][o]), ((ring.k - 1) >> 2) & 1u);
    };
#define m_tgt (s.tgt_mode)
#define m_a0 (s.anc_mode[0])
#define m_a1 (s.anc_mode[1])
#define m_a2 (s.anc_mode[2])
    const int32_t* const p_s0 = chain ? p.anchor_rows + (T == 2 ? (row_begin + my_r) >> 1 : (row_begin + my_r) / T)
                                      : p.target_rows + (row_begin + my_r) * T;
    const int32_t* const p_g1 = p.anchor_rows + p.anchor_stride + row_begin + my_r;
    // bounds check of a row candidate of `mode` (one unsigned compare: a negative candidate wraps);
    // a miss reads row 0 and is only NOTED here (diagnose_indices reports it after the tile).
    // Lanes without a row keep -1.
    auto checked = [&](bool on, int32_t cand, int mode, int bit) -> int32_t {
      const bool ok = (uint32_t)cand < p.mode[mode].rows;
      if (on && !ok) ctl->bad[wid] = (uint8_t)bit;      // cold
      return on ? (ok ? cand : 0) : -1;
    };
    // gather sources (anchors 0..2, or the targets of a chain) and scoring rows (the anchor of a
    // chain, or targets 0, 1): raw index -> candidate -> checked row, in place
    int32_t gsrc0 = -1, gsrc1 = 0, gsrc2 = 0, ssrc0 = 0, ssrc1 = 0;
    const bool has_s = chain || !p.q_out, has_g1 = !chain, has_g2 = !chain && n_branch > 2, has_s1 = !chain && !p.q_out && T > 1;
    if (mine) {
      if (ring.k > 1) {          // every tile but the CTA's first was announced one tile ahead
        gsrc0 = carry_row;
      } else {   // this CTA's first tile: the one exposed index -> node map -> row chain
        const int m0 = chain ? m_tgt : m_a0;
        gsrc0 = checked(true, index_lookup(p.mode[m0], __ldg((chain ? p.target_rows : p.anchor_rows) + row_begin + my_r), ik), m0, 1);
      }
      if (has_s) ssrc0 = __ldg(p_s0);
      if (has_s1) ssrc1 = __ldg(p_s0 + 1);
      if (has_g1) gsrc1 = __ldg(p_g1);
      if (has_g2) gsrc2 = __ldg(p_g1 + p.anchor_stride);
    }
    PendTile cur;
    cur.valid = false;
    cur.stage_info = 0;
    if (frag) {
      cur.tile = tile; cur.row_begin = row_begin; cur.n_valid = n_valid; cur.region = region;
      cur.valid = true;
      cur.table = s.anc_table[0];
      if (staged(3)) {     // the anchor rows come from the staging slot (awaited in score_frag)
        cur.table = stage_ptr(3);
        cur.stage_info = 1u | ((((ring.k - 1) >> 2) & 1u) << 1) | (((ring.k - 1) & 3u) << 2);
      }
#pragma unroll
      for (int rs = 0; rs < 2; ++rs) {
        const int r = frag_row(wid, lane, rs);
        cur.idx[rs] = r < n_valid ? __ldg(p.anchor_rows + (T == 2 ? (row_begin + r) >> 1 : (row_begin + r) / T)) : 0;
      }
    }

    for (int st = 0; st < pg.n; ++st, ++gs) {
      const int g = pg.gather[st];
      if (g != G_NONE) {
        const int so = g == G_TARGET ? 0 : g;      // staging operand of this gather
        if (staged(so)) {
          stage_wait(so);
          gather_to_a<D, true>(smem, stage_ptr(so), mine ? my_r : -1, wid, lane);
        } else if (g == G_TARGET) gather_to_a<D>(smem, s.tgt_table, gsrc0, wid, lane);
        else gather_to_a<D>(smem, s.anc_table[g], g == 0 ? gsrc0 : (g == 1 ? gsrc1 : gsrc2), wid, lane);
        stamp(2);
      }
      // A operand complete (generic-proxy stores -> async proxy) and this thread's
      // TMEM reads of the previous accumulator retired: hand over to the MMA issuer
      ptx::fence_proxy_async_smem();
      ptx::tc_fence_before_sync();
      ptx::mbar_arrive(bar_a_ready);

      if (st == 0) {
        // while the first contraction runs: the raw first-gather index of the NEXT tile of this CTA
        // (its id was published by the scheduler while this tile's first gather ran) ...
        const int64_t nt_v = ring.peek(ctl, 0);
        const int64_t nt = ring_tile(nt_v);
        int32_t raw_n = 0, m2 = 0;
        bool has_n = false, remote2 = true;
        const float* tab2 = nullptr;

        if (nt < p.n_tiles) {
          const SegDev& s2 = p.seg[ring_seg(nt_v)];
          const bool chain2 = (STRUCT >= 0 ? STRUCT : s2.structure) <= GQE_CHAIN3;
          const int64_t rb2 = (chain2 ? s2.q_begin * T : s2.q_begin) + (nt - s2.tile_begin) * kRows;
          const int64_t re2 = chain2 ? s2.q_end * T : s2.q_end;
          remote2 = (s2.remote_mask & (chain2 ? 8u : 1u)) != 0;
          tab2 = chain2 ? s2.tgt_table : s2.anc_table[0];
          m2 = chain2 ? s2.tgt_mode : s2.anc_mode[0];
          has_n = lane < RPW && rb2 + my_r < re2;
          if (has_n) raw_n = __ldg((chain2 ? p.target_rows : p.anchor_rows) + rb2 + my_r);
        }
        // ... the node-map lookups of every later index of this tile (independent loads, nothing
        // waits for them yet; identity when the caller passed rows) ...
        if (ik) {
          if (mine) {
            if (has_s) ssrc0 = index_lookup(p.mode[chain ? m_a0 : m_tgt], ssrc0, ik);
            if (has_s1) ssrc1 = index_lookup(p.mode[m_tgt], ssrc1, ik);
            if (has_g1) gsrc1 = index_lookup(p.mode[m_a1], gsrc1, ik);
            if (has_g2) gsrc2 = index_lookup(p.mode[m_a2], gsrc2, ik);
          }
          if (frag) {
#pragma unroll
            for (int rs = 0; rs < 2; ++rs)
              if (frag_row(wid, lane, rs) < n_valid) cur.idx[rs] = index_lookup(p.mode[m_a0], cur.idx[rs], ik);
          }
        }
        // ... the score of the previous tile, straight from its TMEM region ...
        if (pend.valid) {
          stamp(9);
          score_frag<D, STAGE>(p, ctl, scratch, tmem_base, pend, wid, lane);
          pend.valid = false;
          stamp(8);
        }
        // ... then the bounds checks and the L2 prefetches of this tile's later rows (issued only
        // now, so that they do not queue in front of the first gather's own loads)
        ssrc0 = checked(mine && has_s, ssrc0, chain ? m_a0 : m_tgt, 1);
        ssrc1 = checked(mine && has_s1, ssrc1, m_tgt, 1);
        gsrc1 = checked(mine && has_g1, gsrc1, m_a1, 1);
        gsrc2 = checked(mine && has_g2, gsrc2, m_a2, 1);
        if (chain) {
          if (!(rm & 1u)) prefetch_rows_l2<D>(s.anc_table[0], ssrc0, lane);
        } else {
          if (!(rm & 2u)) prefetch_rows_l2<D>(s.anc_table[1], gsrc1, lane);
          if (n_branch > 2 && !(rm & 4u)) prefetch_rows_l2<D>(s.anc_table[2], gsrc2, lane);
          if (!(rm & 8u) && !p.q_out) {
            prefetch_rows_l2<D>(s.tgt_table, ssrc0, lane);
            if (T > 1) prefetch_rows_l2<D>(s.tgt_table, ssrc1, lane);
          }
        }
        if (frag) {
#pragma unroll
          for (int rs = 0; rs < 2; ++rs) {
            const int32_t c = cur.idx[rs];   // (the same anchor indices were checked through ssrc0 above)
            cur.idx[rs] = (uint32_t)c < p.mode[m_a0].rows ? c : 0;
            // staged: the position in the tile instead (rows 2i, 2i+1 of a (pos, neg) pair share the anchor of 2i)
            if (STAGE && cur.stage_info) cur.idx[rs] = T == 2 ? (frag_row(wid, lane, rs) & ~1) : frag_row(wid, lane, rs);
          }
        }
        // ... and the first-gather rows of the NEXT tile, carried into it in a register, so that
        // its only exposed gather is an L2 hit
        if (nt < p.n_tiles) {
          carry_row = checked(has_n, index_lookup(p.mode[m2], raw_n, ik), m2, 2);
          if (!remote2) prefetch_rows_l2<D>(tab2, carry_row, lane);
        }
      }

      ptx::mbar_wait(bar_acc_full, gs & 1);
      ptx::tc_fence_after_sync();
      stamp(3);

      const int epi = pg.epi[st];
      const int kind = epi & E_KIND;
      if (kind == E_SCORE) { ++gs; break; }  // scored below (or during the next tile), straight from the accumulator
      if (kind == E_NONE) continue;          // left in TMEM for the next step's epilogue
      const bool agg_read = kind == E_AGG && !(epi & F_FIRST);
      const bool agg_raw = (epi & F_AGG_RAW) != 0;
      // this step's accumulator / the running aggregate (roles swapped for F_MMA_ALT steps)
      const uint32_t t_x = (epi & F_MMA_ALT) ? t_agg : t_acc;
      const uint32_t t_a = (epi & F_MMA_ALT) ? t_acc : t_agg;
      const float inv_nb = 1.f / (float)n_branch;
      uint32_t raw[16], araw[16] = {};
      ptx::tmem_ld16(t_x, raw);
      if (agg_read) ptx::tmem_ld16(t_a, araw);
#pragma unroll
      for (int ch = 0; ch < NCH; ++ch) {
        float x[16], a[16];
        ptx::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          x[i] = __uint_as_float(raw[i]);
          a[i] = __uint_as_float(araw[i]);
        }
        if (ch + 1 < NCH) {  // next chunk streams out of TMEM while this one is processed
          ptx::tmem_ld16(t_x + 16 * (ch + 1), raw);
          if (agg_read) ptx::tmem_ld16(t_a + 16 * (ch + 1), araw);
        }
        if (kind == E_AGG) {
          if (epi & F_RELU) {
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = relu_nan_(x[i]);
          }
          if (agg_read) {
            if (agg_raw) {  // the first branch's accumulator, not yet through its relu
#pragma unroll
              for (int i = 0; i < 16; ++i) a[i] = relu_nan_(a[i]);
            }
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] = use_min ? min_nan_(a[i], x[i]) : a[i] + x[i];
          }
          if (!(epi & F_LAST)) {
            uint32_t o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(x[i]);
            ptx::tmem_st16(t_a + 16 * ch, o);
            continue;
          }
          if (!use_min) {  // torch.mean over the stacked operands
#pragma unroll
            for (int i = 0; i < 16; ++i) x[i] *= inv_nb;
          }
          if (epi & F_DEST_ACC) {  // combined embedding is the query embedding itself
            uint32_t o[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(x[i]);
            ptx::tmem_st16(t_acc + 16 * ch, o);
            continue;
          }
        }
        store_a16<D>(smem, row, col_base + 16 * ch, x);
      }
      ptx::tmem_wait_st();
      stamp(4);
      if (kind == E_AGG && (epi & F_DEST_ACC)) { ++gs; break; }
    }

    if (const int f = ctl->bad[wid]) {   // cold: which index was it?  (bit 1 = the carried row: next tile's turn)
      if (f & 1) diagnose_indices<D>(p, s, chain, row_begin, n_valid, wid, lane);
      __syncwarp();
      if (lane == 0) ctl->bad[wid] = (uint8_t)(f >> 1);
      __syncwarp();
    }

    // ---- score ------------------------------------------------------------------------
    // Scored straight from TMEM (score_frag) -- when this CTA has a next tile, only after that
    // tile's first contraction has been handed to the tensor pipe.
    if (frag) {
      if (ring_tile(ring.peek(ctl, 0)) < p.n_tiles) {
        pend = cur;
      } else {
        score_frag<D, STAGE>(p, ctl, scratch, tmem_base, cur, wid, lane);
        stamp(6);
      }
      stamp(7);
      continue;
    }
    // Simple intersections / the eval shape: the accumulator row (the query embedding) is
    // transposed through shared memory (the A planes are dead by now) so that the target rows
    // it is scored against are read warp-per-row with coalesced 128-bit loads.
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      uint32_t raw[16];
      ptx::tmem_ld16(t_acc + 16 * ch, raw);
      ptx::tmem_wait_ld();
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *q_chunk<D>(qsm, row, (col_base + 16 * ch) / 4 + i) =
            make_float4(__uint_as_float(raw[4 * i]), __uint_as_float(raw[4 * i + 1]), __uint_as_float(raw[4 * i + 2]),
                        __uint_as_float(raw[4 * i + 3]));
    }
    ptx::named_bar_sync(1, C::kWorkerThreads);
    stamp(5);

    // Rows are scored SU at a time: all table loads of a batch are issued first and the
    // warp reductions of the batch run in lockstep (one butterfly level for all values).
    double local = 0.0;
    if (p.q_out) {
      // many targets per query: hand the query embedding rows to gqe_score_pairs
#pragma unroll 1
      for (int u = 0; u < RPW; ++u) {
        const int r = wid * RPW + u;
        if (r >= n_valid) break;  // warp-uniform
        float4* dst = reinterpret_cast<float4*>(p.q_out + (size_t)(row_begin + r) * D);
#pragma unroll
        for (int j = 0; j < NV; ++j) dst[lane + 32 * j] = *q_chunk<D>(qsm, r, lane + 32 * j);
      }
    } else {
      constexpr int SU = 4 / NV;
      static_assert(RPW % SU == 0, "score batch");
      const bool st3 = staged(3);     // the target rows come from the staging slot
      if (st3) {
        stage_wait(3);
        if (staged(4)) stage_wait(4);
      }
#pragma unroll 1
      for (int u0 = 0; u0 < RPW; u0 += SU) {
        float4 ta[SU][NV], tb[SU][NV];
        int32_t ra[SU];
#pragma unroll
        for (int u = 0; u < SU; ++u) {
          ra[u] = __shfl_sync(0xffffffffu, ssrc0, u0 + u);    // -1: row past the end of the tile (warp-uniform)
          const int32_t rb = __shfl_sync(0xffffffffu, ssrc1, u0 + u);
          const bool ok = ra[u] >= 0;
          const float4* a_src = reinterpret_cast<const float4*>(s.tgt_table + (size_t)(ok ? ra[u] : 0) * D);
          const float4* b_src = reinterpret_cast<const float4*>(s.tgt_table + (size_t)(ok ? (T > 1 ? rb : ra[u]) : 0) * D);
          if (st3) {
            const size_t pos = (size_t)(wid * RPW + u0 + u);
            a_src = reinterpret_cast<const float4*>(stage_ptr(3) + pos * D);
            b_src = reinterpret_cast<const float4*>(stage_ptr(T > 1 ? 4 : 3) + pos * D);
          }
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            ta[u][j] = ok ? (st3 ? ld_row_staged(a_src + lane + 32 * j) : ld_row(a_src + lane + 32 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
            tb[u][j] = ok ? (st3 ? ld_row_staged(b_src + lane + 32 * j) : ld_row(b_src + lane + 32 * j)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        float red[SU][5];  // |q|^2, q.a, |a|^2, q.b, |b|^2
#pragma unroll
        for (int u = 0; u < SU; ++u) {
          const int r = wid * RPW + u0 + u;
          float qq = 0.f, d0 = 0.f, d1 = 0.f, n0 = 0.f, n1 = 0.f;
#pragma unroll
          for (int j = 0; j < NV; ++j) {
            const float4 y = *q_chunk<D>(qsm, r, lane + 32 * j);
            const float4 a = ta[u][j], b = tb[u][j];
            qq = fmaf(y.x, y.x, qq); qq = fmaf(y.y, y.y, qq); qq = fmaf(y.z, y.z, qq); qq = fmaf(y.w, y.w, qq);
            d0 = fmaf(y.x, a.x, d0); d0 = fmaf(y.y, a.y, d0); d0 = fmaf(y.z, a.z, d0); d0 = fmaf(y.w, a.w, d0);
            n0 = fmaf(a.x, a.x, n0); n0 = fmaf(a.y, a.y, n0); n0 = fmaf(a.z, a.z, n0); n0 = fmaf(a.w, a.w, n0);
            d1 = fmaf(y.x, b.x, d1); d1 = fmaf(y.y, b.y, d1); d1 = fmaf(y.z, b.z, d1); d1 = fmaf(y.w, b.w, d1);
            n1 = fmaf(b.x, b.x, n1); n1 = fmaf(b.y, b.y, n1); n1 = fmaf(b.z, b.z, n1); n1 = fmaf(b.w, b.w, n1);
          }
          red[u][0] = qq; red[u][1] = d0; red[u][2] = n0; red[u][3] = d1; red[u][4] = n1;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
          for (int u = 0; u < SU; ++u)
#pragma unroll
            for (int k = 0; k < 5; ++k) red[u][k] += __shfl_xor_sync(0xffffffffu, red[u][k], o);
        }
#pragma unroll
        for (int u = 0; u < SU; ++u) {
          if (ra[u] < 0) continue;
          const int r = wid * RPW + u0 + u;
          const int64_t q = row_begin + r;
          const float nq = clamped_norm(red[u][0]);
          // t_hat = t/|t| has unit norm: cos(t_hat, q) = (t.q/|t|) / max(|q|, eps); a zero
          // target row gives 0/0 = NaN as in the reference
          const float s0 = __fdividef(unit_dot(red[u][1], red[u][2]), nq);
          const float s1 = __fdividef(unit_dot(red[u][3], red[u][4]), nq);
          if (lane == 0) {
            if (p.out_scores) {
              p.out_scores[q * T] = s0;
              if (T > 1) p.out_scores[q * T + 1] = s1;
            }
            local += (double)hinge_(p.margin, s0, s1);
          }
          // more than one (pos, neg) pair per query (the eval shape): the remaining targets
          for (int t0 = 2; t0 < T; ++t0) {
            const float4* c_src = reinterpret_cast<const float4*>(s.tgt_table + (size_t)resolve_index(p.mode[m_tgt], m_tgt, __ldg(p.target_rows + q * T + t0), ik, p.err) * D);
            float d2 = 0.f, n2 = 0.f;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
              const float4 c = __ldg(c_src + lane + 32 * j);
              const float4 y = *q_chunk<D>(qsm, r, lane + 32 * j);
              d2 = fmaf(y.x, c.x, d2); d2 = fmaf(y.y, c.y, d2); d2 = fmaf(y.z, c.z, d2); d2 = fmaf(y.w, c.w, d2);
              n2 = fmaf(c.x, c.x, n2); n2 = fmaf(c.y, c.y, n2); n2 = fmaf(c.z, c.z, n2); n2 = fmaf(c.w, c.w, n2);
            }
            d2 = warp_sum_f(d2); n2 = warp_sum_f(n2);
            if (lane == 0 && p.out_scores) p.out_scores[q * T + t0] = __fdividef(unit_dot(d2, n2), nq);
          }
        }
      }
    }
    stamp(6);
    // every warp is done with the transposed tile (the next gather overwrites it) and, for the
    // loss, the per-tile hinge sum goes to its own slot: partials[tile]
    if (p.out_loss && lane == 0) ctl->red[wid] = local;
    ptx::named_bar_sync(1, C::kWorkerThreads);
    if (p.out_loss && threadIdx.x == 0) {
      double sum = 0.0;
      for (int w = 0; w < C::kWorkerWarps; ++w) sum += ctl->red[w];
      p.partials[tile] = sum;
    }
    stamp(7);
  }
#undef m_tgt
#undef m_a0
#undef m_a1
#undef m_a2
  if (pend.valid) score_frag<D, STAGE>(p, ctl, scratch, tmem_base, pend, wid, lane);  // (never: a deferred tile has a successor)
  if (threadIdx.x == 0) cta_stamp(p, 3);
  loss_finish<D>(p, ctl, smem, wid, lane);
}

// ---- STAGE: helper warp -----------------------------------------------------------------
// Node-type-sharded tables: rows that live in a PEER GPU's HBM cost an NVLink round trip (~2-3 us) per
// dependent load, cannot be prefetched into the local L2 (peer lines bypass it), and the workers have only a
// few rows in flight -- the first remote gather of every tile was exposed.  The helper
// warp consumes the tile ring like the workers, but as early as the scheduler publishes a tile (up to two
// tiles ahead of the workers), and copies every REMOTE operand row of that tile into the CTA's staging slot in
// local memory, by position in the tile.  Registers cannot hold enough rows in flight for that (8 rows per
// warp at this kernel's 96 registers, and spilling serialises the loads), so the rows travel by TMA: lane i
// owns slot i of a 32-row ring in shared memory (the third weight stage), fetches its row with cp.async.bulk
// (peer -> shared, completion on the slot's mbarrier) and writes it on with a bulk store (shared -> the local
// staging area): 32 KiB in flight per SM, 4.7 MB per GPU -- the link's bandwidth-delay product twice over.
// The workers then gather those operands from the staging slot (L2 hits, ld.global.cg) after waiting on
// stage_ready[slot][operand].  Index resolution is the workers' own (raw index -> node map -> bounds check, a
// bad index reads row 0; reporting stays with the workers).  Chain anchors of (pos, neg) pairs are fetched
// once per pair.
template <int D, int STRUCT>
__device__ __forceinline__ void helper(const LaunchParams& p, uint8_t* smem, Ctl* ctl, int lane) {
  using C = Cfg<D>;
  constexpr int NG = kRows / 32;                    // groups of 32 rows per operand
  constexpr uint32_t kRowBytes = D * 4;
  const int T = p.T, ik = p.index_kind;
  const uint32_t my_slot = ptx::smem_u32(smem + C::kOffB + 2 * C::kStageBytes) + (uint32_t)lane * kRowBytes;
  const uint32_t my_bar = ptx::smem_u32(&ctl->row_bar[lane]);
  uint32_t my_phase = 0;
  uint32_t k = 0;
  for (;; ++k) {
    ptx::mbar_wait(ptx::smem_u32(&ctl->sched_full[k & 1]), (k >> 1) & 1);
    const int64_t tile_v = *reinterpret_cast<volatile int64_t*>(&ctl->tile_id[k & 1]);
    __syncwarp();
    if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&ctl->sched_empty[k & 1]));
    const int64_t tile = ring_tile(tile_v);
    if (tile >= p.n_tiles) break;
    const SegDev& s = p.seg[ring_seg(tile_v)];
    const bool chain = (STRUCT >= 0 ? STRUCT : s.structure) <= GQE_CHAIN3;
    const int64_t row_begin = (chain ? s.q_begin * T : s.q_begin) + (tile - s.tile_begin) * kRows;
    const int64_t row_end = chain ? s.q_end * T : s.q_end;
    const int n_valid = (int)min((int64_t)kRows, row_end - row_begin);
    const uint32_t sops = staged_ops(k, chain, s.remote_mask, s.n_anchor, T, p.q_out != nullptr) & p.stage_mask;
    float* slot = p.stage + ((size_t)blockIdx.x * C::kStageDepth + (k & 3u)) * (size_t)(C::kStageOps * kRows * D);
    // The rows of EVERY staged operand first (two dependent local loads each, all operands' in flight together),
    // then the copies.  Lane l handles tile rows l, l + 32, ...; the anchors of a chain tile, which (pos, neg)
    // pairs share, are fetched once per pair: lane l handles rows 2l, 2l + 64.
    int32_t row[C::kStageOps][NG];
#pragma unroll
    for (int o = 0; o < C::kStageOps; ++o) {
      const bool on = (sops >> o) & 1u;
      const bool from_targets = chain ? (o == 0) : (o >= 3);
      const int ak = (chain || o >= 3) ? 0 : o;
      const int mode = from_targets ? s.tgt_mode : s.anc_mode[ak];
      const bool dedupe = chain && o == 3 && T == 2;
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        const int r = dedupe ? 2 * (32 * g + lane) : 32 * g + lane;
        row[o][g] = -1;
        if (on && r < n_valid) {
          int32_t raw;
          if (chain) raw = o == 0 ? __ldg(p.target_rows + row_begin + r)
                                  : __ldg(p.anchor_rows + (T == 2 ? (row_begin + r) >> 1 : (row_begin + r) / T));
          else raw = o < 3 ? __ldg(p.anchor_rows + (int64_t)o * p.anchor_stride + row_begin + r)
                           : __ldg(p.target_rows + (row_begin + r) * T + (o - 3));
          const int32_t cand = index_lookup(p.mode[mode], raw, ik);
          row[o][g] = (uint32_t)cand < p.mode[mode].rows ? cand : 0;
        }
      }
    }
#pragma unroll
    for (int o = 0; o < C::kStageOps; ++o) {
      if (!((sops >> o) & 1u)) {   // nothing to fetch: the barrier still completes once per tile (the workers' parity rule)
        if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&ctl->stage_ready[k & 3u][o]));
        continue;
      }
      const bool from_targets = chain ? (o == 0) : (o >= 3);
      const int ak = (chain || o >= 3) ? 0 : o;
      const float* table = from_targets ? s.tgt_table : s.anc_table[ak];
      const bool dedupe = chain && o == 3 && T == 2;
      float* dst_op = slot + (size_t)o * kRows * D;
#pragma unroll
      for (int g = 0; g < NG; ++g) {
        if (row[o][g] < 0) continue;
        const int r = dedupe ? 2 * (32 * g + lane) : 32 * g + lane;
        // the bulk store that last read this lane's ring slot has finished reading it
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        ptx::mbar_arrive_expect_tx(my_bar, kRowBytes);
        ptx::tma_bulk_g2s(my_slot, table + (size_t)row[o][g] * D, kRowBytes, my_bar);
        ptx::mbar_wait(my_bar, my_phase);
        my_phase ^= 1u;
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_op + (size_t)r * D), "r"(my_slot),
                     "r"(kRowBytes)
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      // every store of this operand has landed in the staging slot: hand it to the workers
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
      __threadfence_block();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(ptx::smem_u32(&ctl->stage_ready[k & 3u][o]));
    }
  }
}

// ---- TMA producer + tile scheduler: streams the packed planes of every step's matrix ---
// PAIR (EXPERIMENT, off unless GQE_PAIR=1 is set; measured, not faster -- DESIGN.md 3.1): the CTA is one of
// a cluster of two that walk the SAME formula's tiles in lockstep (virtual tiles 2j, 2j+1: the host pads
// every segment to an even tile count, see gqe_capi.cu).  Each CTA fetches HALF of every weight stage and
// multicasts it into both CTAs' rings (cp.async.bulk ... .multicast::cluster), the MMA issuers release a
// stage in both CTAs (tcgen05.commit ... .multicast::cluster): the weight stream is read from L2 once per
// 256 rows instead of once per 128.  Tiles are dealt round-robin (tile + gridDim.x): the pair needs the same
// sequence in both CTAs.  Result on the benchmark mix: bit-identical scores, 0.146 ms against 0.132 -- the
// contractions are not bound by the weight stream (halving it changes their duration by < 2 %), and the
// round-robin deal loses 12 us to imbalance that the dynamic scheduler does not have.
// HALF (the STAGE instantiation): the ring holds FOUR HALF stages -- output rows [0, d/2) and [d/2, d) of a
// K-block plane, 16 KiB each, contiguous in the packed image -- in the first 64 KiB of the weight region, so
// that its last 32 KiB are free for the helper warp's row ring while the weight stream still runs three
// copies ahead of the tensor pipe (a ring of two full stages costs 17 % of the step: measured).
template <int D, int STRUCT, bool PAIR, bool HALF = false>
__device__ __forceinline__ void producer(const LaunchParams& p, uint8_t* smem, Ctl* ctl) {
  using C = Cfg<D>;
  constexpr int NST = HALF ? 4 : kStages;
  constexpr uint32_t kCopyBytes = HALF ? C::kStageBytes / 2 : C::kStageBytes;
  const bool deepsets = p.inter == GQE_INTER_DEEPSETS_MEAN || p.inter == GQE_INTER_DEEPSETS_MIN;
  uint32_t slot = 0, phase = 0;
  const uint32_t rank = PAIR ? ptx::cluster_ctarank() : 0u;
  // tile k of this CTA: the first is blockIdx.x, the rest come from the global counter.
  // Tile k+1 is published while tile k's weights are being streamed, so the workers can
  // prefetch its rows; an id >= n_tiles is the stop marker.
  auto publish = [&](uint32_t k, int64_t id) {
    ptx::mbar_wait(ptx::smem_u32(&ctl->sched_empty[k & 1]), ((k >> 1) & 1) ^ 1);
    *reinterpret_cast<volatile int64_t*>(&ctl->tile_id[k & 1]) = id;
    __threadfence_block();
    ptx::mbar_arrive(ptx::smem_u32(&ctl->sched_full[k & 1]));
  };
  int64_t tile = blockIdx.x;
  int si = tile < p.n_tiles ? seg_of_tile<STRUCT>(p, tile) : 0;
  publish(0, ring_pack(tile, si));
  // The packed weights come from gqe_pack, the kernel before this one in the stream; this kernel
  // is launched programmatically dependent on it, so that the set-up and the first gathers of
  // the workers (tables and indices only) overlap its tail.  Only this thread reads weights.
  ptx::griddep_wait();
  for (uint32_t k = 0; tile < p.n_tiles; ++k) {
    const int64_t next = PAIR ? tile + (int64_t)gridDim.x : (int64_t)gridDim.x + (int64_t)atomicAdd(p.tile_counter, 1u);
    const int si_next = next < p.n_tiles ? seg_of_tile<STRUCT>(p, next) : 0;
    publish(k + 1, next < p.n_tiles ? ring_pack(next, si_next) : p.n_tiles);
    const SegDev& s = p.seg[si];
    const Prog& pg = s.prog;
    for (int st = 0; st < pg.n; ++st) {
      const uint8_t* src = step_matrix(s, pg.mat[st]);
#pragma unroll 1
      for (int i = 0; i < (HALF ? 4 : 2) * C::kKB; ++i) {
        const uint32_t full = ptx::smem_u32(&ctl->full[slot]), empty = ptx::smem_u32(&ctl->empty[slot]);
        ptx::mbar_wait(empty, phase ^ 1);          // PAIR: released by BOTH CTAs' MMA issuers (the copy lands in both rings)
        ptx::mbar_arrive_expect_tx(full, kCopyBytes);
        if (HALF) {
          ptx::tma_bulk_g2s(ptx::smem_u32(smem + C::kOffB + slot * kCopyBytes), src + (size_t)i * kCopyBytes, kCopyBytes, full);
        } else if (PAIR) {
          constexpr uint32_t kHalf = C::kStageBytes / 2;
          const uint32_t off = rank * kHalf;
          ptx::tma_bulk_g2s_multicast(ptx::smem_u32(smem + C::kOffB + slot * C::kStageBytes) + off,
                                      src + (size_t)i * C::kStageBytes + off, kHalf, full, (uint16_t)3);
        } else {
          ptx::tma_bulk_g2s(ptx::smem_u32(smem + C::kOffB + slot * C::kStageBytes), src + (size_t)i * C::kStageBytes,
                            C::kStageBytes, full);
        }
        if (++slot == NST) { slot = 0; phase ^= 1; }
      }
    }
    tile = next;
    si = si_next;
  }
}

// ---- MMA issuer ------------------------------------------------------------------------
template <int D, int STRUCT, bool PAIR, bool HALF = false>
__device__ __forceinline__ void mma_issuer(const LaunchParams& p, uint8_t* smem, Ctl* ctl) {
  using C = Cfg<D>;
  constexpr int NST = HALF ? 4 : kStages;
  constexpr int NH = HALF ? 2 : 1;                         // column halves of the accumulator, one ring slot each
  constexpr uint32_t kSlotBytes = C::kStageBytes / NH;
  const bool deepsets = p.inter == GQE_INTER_DEEPSETS_MEAN || p.inter == GQE_INTER_DEEPSETS_MIN;
  constexpr uint32_t idesc = ptx::umma_idesc_bf16_f32(kRows, D / NH);
  const uint32_t tmem_base = ctl->tmem_base;
  const uint32_t a_hi = ptx::smem_u32(smem + C::kOffAhi), a_lo = ptx::smem_u32(smem + C::kOffAlo);
  const uint32_t b0 = ptx::smem_u32(smem + C::kOffB);
  const uint32_t bar_a_ready = ptx::smem_u32(&ctl->a_ready), bar_acc_full = ptx::smem_u32(&ctl->acc_full);
  uint32_t slot = 0, phase = 0, gs = 0;
  TileRing ring;
  for (;;) {
    const int64_t tile_v = ring.take(ctl);
    const int64_t tile = ring_tile(tile_v);
    if (tile >= p.n_tiles) break;
    const SegDev& s = p.seg[ring_seg(tile_v)];
    const int structure = STRUCT >= 0 ? STRUCT : s.structure;
    const Prog& pg = s.prog;
    // the accumulator / aggregate roles of the two TMEM regions alternate per tile (see worker())
    const uint32_t tmem_main = tmem_base + ((ring.k - 1) & 1u) * D;
    const uint32_t tmem_alt = tmem_base + (ring.k & 1u) * D;
    for (int st = 0; st < pg.n; ++st, ++gs) {
      const uint32_t tmem_acc = (pg.epi[st] & F_MMA_ALT) ? tmem_alt : tmem_main;
      ptx::mbar_wait(bar_a_ready, gs & 1);
      ptx::tc_fence_after_sync();
#pragma unroll 1
      for (int kb = 0; kb < C::kKB; ++kb) {
        // plane 0 of this K block: B_hi, used by A_hi and A_lo (HALF: output columns [0, d/2), then [d/2, d))
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const uint32_t full = ptx::smem_u32(&ctl->full[slot]), empty = ptx::smem_u32(&ctl->empty[slot]);
          ptx::mbar_wait(full, phase);
          ptx::tc_fence_after_sync();
          const uint32_t b = b0 + slot * kSlotBytes;
          const uint32_t acc = tmem_acc + h * (D / NH);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_ss(acc, ptx::umma_desc_sw128(a_hi + kb * C::kABlockBytes + 32 * k),
                              ptx::umma_desc_sw128(b + 32 * k), idesc, (kb | k) != 0);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_ss(acc, ptx::umma_desc_sw128(a_lo + kb * C::kABlockBytes + 32 * k),
                              ptx::umma_desc_sw128(b + 32 * k), idesc, 1u);
          if (PAIR) ptx::umma_commit_multicast(empty, (uint16_t)3); else ptx::umma_commit(empty);
          if (++slot == NST) { slot = 0; phase ^= 1; }
        }
        // plane 1: B_lo, used by A_hi
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const uint32_t full = ptx::smem_u32(&ctl->full[slot]), empty = ptx::smem_u32(&ctl->empty[slot]);
          ptx::mbar_wait(full, phase);
          ptx::tc_fence_after_sync();
          const uint32_t b = b0 + slot * kSlotBytes;
          const uint32_t acc = tmem_acc + h * (D / NH);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16_ss(acc, ptx::umma_desc_sw128(a_hi + kb * C::kABlockBytes + 32 * k),
                              ptx::umma_desc_sw128(b + 32 * k), idesc, 1u);
          if (PAIR) ptx::umma_commit_multicast(empty, (uint16_t)3); else ptx::umma_commit(empty);
          if (++slot == NST) { slot = 0; phase ^= 1; }
        }
      }
      ptx::umma_commit(bar_acc_full);
    }
  }
}

// ---------------------------------------------------------------------------------------
// STRUCT >= 0: the single-formula kernel of that query structure; STRUCT < 0: the grouped
// kernel, which looks its tile's structure up at run time (CTA-uniform).  Persistent:
// launched with min(n_tiles, SMs x CTAs/SM) CTAs.
template <int D, int STRUCT, bool PAIR = false, bool STAGE = false>
__global__ void __launch_bounds__(Cfg<D>::kThreads + (STAGE ? 32 * Cfg<D>::kHelperWarps : 0), Cfg<D>::kCtasPerSm)
    gqe_fused_tc(const __grid_constant__ LaunchParams p) {
  using C = Cfg<D>;
  extern __shared__ __align__(1024) uint8_t smem[];
  Ctl* ctl = reinterpret_cast<Ctl*>(smem + C::kOffCtl);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) cta_stamp(p, 0);

  if (wid == C::kWorkerWarps && lane == 0) {
    if ((ptx::smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B atoms need 1024-byte alignment
    for (int i = 0; i < kStages + 1; ++i) {
      ptx::mbar_init(ptx::smem_u32(&ctl->full[i]), 1);
      ptx::mbar_init(ptx::smem_u32(&ctl->empty[i]), PAIR ? 2 : 1);
    }
    ptx::mbar_init(ptx::smem_u32(&ctl->a_ready), C::kWorkerThreads);
    ptx::mbar_init(ptx::smem_u32(&ctl->acc_full), 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(ptx::smem_u32(&ctl->sched_full[i]), 1);
      // workers + MMA issuer (+ lane 0 of every helper warp)
      ptx::mbar_init(ptx::smem_u32(&ctl->sched_empty[i]), C::kWorkerThreads + 1 + ((STAGE && p.stage_on == 1) ? C::kHelperWarps : 0));
    }
    if (STAGE)
    {
      for (int i = 0; i < 4 * 5; ++i) ptx::mbar_init(ptx::smem_u32(&ctl->stage_ready[i / 5][i % 5]), C::kHelperWarps);
      for (int i = 0; i < 32; ++i) ptx::mbar_init(ptx::smem_u32(&ctl->row_bar[i]), 1);
    }
    ptx::fence_mbar_init();
  } else if (wid == C::kWorkerWarps + 1) {
    ptx::tmem_alloc(ptx::smem_u32(&ctl->tmem_base), C::kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before_sync();
  __syncthreads();
  if (PAIR) ptx::cluster_sync();   // the partner's barriers exist before anything is sent to them
  // the index arrays come from gqe_fetch_indices, the kernel in front of this one (a *_host call on pinned
  // buffers): this kernel was launched while that one ran; its writes are visible after the wait
  if (p.wait_prev) ptx::griddep_wait();
  ptx::tc_fence_after_sync();
  if (threadIdx.x == 0) cta_stamp(p, 1);

  if (wid < C::kWorkerWarps) {
    worker<D, STRUCT, STAGE>(p, smem, ctl);
  } else if (wid == C::kWorkerWarps) {
    if (lane == 0) producer<D, STRUCT, PAIR, STAGE>(p, smem, ctl);
    __syncwarp();
  } else if (wid == C::kWorkerWarps + 1) {
    if (lane == 0) mma_issuer<D, STRUCT, PAIR, STAGE>(p, smem, ctl);
    __syncwarp();
  } else if (STAGE) {
    if (p.stage_on == 1) helper<D, STRUCT>(p, smem, ctl, lane);
  }

  ptx::tc_fence_before_sync();
  __syncthreads();
  if (PAIR) ptx::cluster_sync();   // no CTA leaves while its partner may still signal its barriers
  if (wid == C::kWorkerWarps + 1) {
    ptx::tc_fence_after_sync();
    ptx::tmem_dealloc(ctl->tmem_base, C::kTmemCols);
  }
  if (threadIdx.x == 0) cta_stamp(p, 4);
}

// ---------------------------------------------------------------------------------------
// gqe_pack: fp32 [d,d] parameter -> two bf16 planes per 64-wide K block in the smem image
// the tensor-core kernel streams.  B[n][k] = chain_form ? M[k][n] : M[n][k]
//   chain_form  : act.mm(M)       (decoders.py:145)
//   !chain_form : M.mm(embeds)    (decoders.py:150,289,299)
template <int D>
__global__ void __launch_bounds__(256) gqe_pack(const __grid_constant__ PackParams p) {
  ptx::griddep_launch_dependents();   // the fused kernel may start its set-up and first gathers now
  const PackEntry& e = p.e[blockIdx.y];
  uint8_t* out = e.dst;
  const int item = blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk (8 k) of one n
  if (item >= D * D / 8) return;
  int n, kc;
  if (!e.chain_form) { n = item / (D / 8); kc = item % (D / 8); }
  else { kc = item / D; n = item % D; }
  const int ns = e.perm ? score_col_src(n) : n;   // accumulator column n holds output column ns
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int k = kc * 8 + i;
    x[i] = e.chain_form ? __ldg(e.src + (size_t)k * D + ns) : __ldg(e.src + (size_t)ns * D + k);
  }
  uint4 hi, lo;
  split2(x[0], x[1], hi.x, lo.x);
  split2(x[2], x[3], hi.y, lo.y);
  split2(x[4], x[5], hi.z, lo.z);
  split2(x[6], x[7], hi.w, lo.w);
  const int kb = kc >> 3, c = kc & 7;
  const uint32_t off = (uint32_t)((n >> 3) * 1024 + (n & 7) * 128 + ((c ^ (n & 7)) << 4));
  *reinterpret_cast<uint4*>(out + (size_t)(2 * kb) * Cfg<D>::kStageBytes + off) = hi;
  *reinterpret_cast<uint4*>(out + (size_t)(2 * kb + 1) * Cfg<D>::kStageBytes + off) = lo;
}

}  // namespace tc
}  // namespace gqe
