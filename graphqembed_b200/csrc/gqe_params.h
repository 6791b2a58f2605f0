// gqe_params.h -- launch parameter blocks shared by the host side (gqe_capi.cu)
// and the kernels.  Plain structs, no device code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gqe.h"

namespace gqe {

constexpr int kTileRows = 64;
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kRowsPerWarp = kTileRows / kWarps;  // 8
constexpr int kPanelK = 16;
constexpr int kRowStride = kTileRows + 4;  // 68 floats: 16B-aligned rows, 4-bank skew
#ifndef GQE_MAX_SEGS
#define GQE_MAX_SEGS 96
#endif
#ifndef GQE_MAX_MODES
#define GQE_MAX_MODES 16
#endif
constexpr int kMaxSegs = GQE_MAX_SEGS;     // formulas per launch
constexpr int kMaxModes = GQE_MAX_MODES;   // node types a launch can address (Bio: 5, configs[4]: 8)
constexpr float kCosEps = 1e-8f;  // nn.CosineSimilarity default eps (model.py:68)

// One node type as the kernels see it: how an index of the caller becomes a row of the mode's
// table, and how many rows that table has.  The index arrays of a call hold either table rows
// (index_kind == 0) or NODE IDS (index_kind == 1, the *_nodes entry points), which are mapped as
//     row = lut ? lut[node - base] : node - base          (gqe_bind_node_maps)
// i.e. the node_maps[mode][n] + 1 of reference netquery/bio/data_utils.py:20-21 as a dense
// per-mode table.  Every resolved row is checked against `rows`; a miss (unknown node, row out of
// range) is reported through LaunchParams::err and reads row 0 instead of out of bounds.
struct ModeDev {
  const int32_t* lut;   // DEVICE int32 [len]: table row of node (base + i), -1 = not a node of this mode
  int32_t base;
  uint32_t len;
  uint32_t rows;        // rows of the bound table
  uint32_t pad_;
};
enum { IDX_ERR_NONE = 0, IDX_ERR_UNKNOWN_NODE = 1, IDX_ERR_ROW_RANGE = 2 };

// ---- the per-structure program of the tensor-core kernel (gqe_tc.cuh) ---------------------------
// Built by the host once per formula of a launch (SegDev::prog); the kernel's three roles walk it.
constexpr int kMaxSteps = 8;
#ifndef __CUDACC__
#define __host__
#define __device__
#endif
enum { G_NONE = 7, G_TARGET = 3 };                       // gather source: anchor 0..2, target, none
enum { M_REL0 = 0, M_REL1 = 1, M_REL2 = 2, M_PRE = 3, M_POST = 4 };
// epilogue kinds (E_KIND masks them) and flags.  E_NONE: the accumulator is left in TMEM, raw,
// for the NEXT step's epilogue (first DeepSets branch).  F_MMA_ALT: this step's MMAs write the
// tile's second TMEM region, so that the raw first branch (F_AGG_RAW) or the running aggregate
// in the first region survives; the epilogue then combines the two and keeps the aggregate in
// the first region.
enum { E_TO_A = 0, E_AGG = 1, E_SCORE = 2, E_NONE = 3, E_KIND = 3, F_RELU = 4, F_FIRST = 8, F_LAST = 16, F_DEST_ACC = 32,
       F_AGG_RAW = 64, F_MMA_ALT = 128 };

struct Prog {
  int n;
  uint8_t mat[kMaxSteps];
  uint8_t gather[kMaxSteps];
  uint8_t epi[kMaxSteps];
};

// Operator order of reference netquery/model.py:70-109 (see include/gqe.h gqe_plan).
// `composed`: the host pre-multiplied every run of consecutive linear operators of this
// formula into one matrix (gqe_compose, fp32), so a run is ONE contraction here:
//   chains          act.mm(M1).mm(M2).mm(M3)      -> act.mm(M1 M2 M3)          rel[0]
//   DeepSets branch relu(pre.mm(R.mm(e)))         -> relu((pre R).mm(e))       rel[b]
//   3-inter_chain   pre.mm(R2a.mm(R2b.mm(e)))     -> (pre R2a R2b).mm(e)       rel[1]
//   3-chain_inter   R1.mm(post.mm(combined))      -> (R1 post).mm(combined)    post
// Same algebra, different fp32 rounding (~1e-7 relative), far inside the 1e-4 bound.
inline __host__ __device__ void build_program(Prog& pg, int structure, bool deepsets, bool composed) {
  int n = 0;
  auto push = [&](int mat, int gather, int epi) {
    pg.mat[n] = (uint8_t)mat;
    pg.gather[n] = (uint8_t)gather;
    pg.epi[n] = (uint8_t)epi;
    ++n;
  };
  if (structure <= GQE_CHAIN3) {
    const int hops = composed ? 1 : structure + 1;
    for (int h = 0; h < hops; ++h) push(h, h == 0 ? G_TARGET : G_NONE, h == hops - 1 ? E_SCORE : E_TO_A);
  } else {
    const int nb = structure == GQE_INTER3 ? 3 : 2;
    for (int b = 0; b < nb; ++b) {
      const int pos = (b == 0 ? F_FIRST : 0) | (b == nb - 1 ? F_LAST : 0);
      const int agg_simple = E_AGG | pos | ((b == nb - 1 && structure != GQE_CHAIN_INTER3) ? F_DEST_ACC : 0);
      if (composed) {
        // DeepSets: the first branch has no epilogue of its own -- its accumulator stays in TMEM
        // and the second branch's epilogue applies relu to both (one TMEM pass and one
        // workers <-> issuer round trip less per tile)
        if (deepsets) push(b, b, b == 0 ? E_NONE : (E_AGG | F_RELU | F_MMA_ALT | (b == 1 ? F_AGG_RAW : 0) | (b == nb - 1 ? F_LAST : 0)));
        else push(b, b, agg_simple);
        continue;
      }
      if (structure == GQE_INTER_CHAIN3 && b == 1) {
        push(M_REL1, b, E_TO_A);                               // reverse(r2b) first (model.py:85)
        push(M_REL2, G_NONE, deepsets ? E_TO_A : agg_simple);  // then reverse(r2a)
      } else {
        push(b, b, deepsets ? E_TO_A : agg_simple);
      }
      if (deepsets) push(M_PRE, G_NONE, E_AGG | F_RELU | pos);  // relu(pre.mm(e)) decoders.py:289-292
    }
    if (composed) {
      if (deepsets) push(M_POST, G_NONE, E_SCORE);                          // post, or R1 post for 3-chain_inter
      else if (structure == GQE_CHAIN_INTER3) push(M_REL2, G_NONE, E_SCORE);
    } else {
      if (deepsets) push(M_POST, G_NONE, structure == GQE_CHAIN_INTER3 ? E_TO_A : E_SCORE);  // decoders.py:299
      if (structure == GQE_CHAIN_INTER3) push(M_REL2, G_NONE, E_SCORE);                      // model.py:107
    }
  }
  pg.n = n;
}


// One formula's slice of a launch, fully resolved to device pointers.
struct SegDev {
  int32_t structure;
  int32_t n_anchor;
  uint32_t remote_mask;  // bit k: anchor table k lives in a PEER GPU's HBM; bit 3: the target table
  int32_t composed;      // tensor-core path: runs of linear operators were pre-multiplied (gqe_compose)
  int8_t tgt_mode;       // index into LaunchParams::mode
  int8_t anc_mode[GQE_MAX_ANCHORS];
  const float* tgt_table;
  const float* anc_table[GQE_MAX_ANCHORS];
  const float* rel[GQE_MAX_RELS];  // relation parameters in application order
  const float* pre;                // DeepSets pre/post of the intersection mode
  const float* post;
  int64_t q_begin, q_end;          // query range in the concatenated arrays
  int64_t tile_begin;              // first tile of this segment inside the launch
  Prog prog;                       // tensor-core path: the contractions of this formula, in reference order
};

struct LaunchParams {
  SegDev seg[kMaxSegs];
  int32_t n_segs;
  int32_t decoder;  // gqe_decoder
  int32_t inter;    // gqe_inter
  const int32_t* anchor_rows;
  int64_t anchor_stride;  // total queries (distance between anchor slots)
  const int32_t* target_rows;
  const int64_t* target_offsets;  // ragged layout (single segment only) or null
  int64_t n_pairs;
  int32_t T;  // regular layout: targets per query
  float* out_scores;
  // fused margin loss (T == 2): deterministic two-level reduction
  float* out_loss;
  float margin;
  double inv_q;
  double* partials;   // [gridDim.x]
  double* loss_acc;   // running sum across the launches of one call
  int32_t final_launch;  // last launch of the call: leave *loss_acc zeroed for the next call
  unsigned int* ticket;
  // many targets per query (the eval shape): intersection tiles write their query
  // embedding rows here (fp32 [n_queries_total, D]) instead of scoring; gqe_score_pairs
  // then scores every (query, target) pair against them
  float* q_out;
  // persistent tensor-core kernel: tiles of this launch and the dynamic tile counter
  int64_t n_tiles;
  unsigned int* tile_counter;
  // diagnostics (gqe_debug_set_phase_log): kPhaseSlots (tag << 56 | clock64) stamps per tile
  unsigned long long* phase_log;
  int64_t phase_cap;  // tiles the log has room for
  // index resolution (see ModeDev)
  int32_t index_kind;           // 0: the index arrays hold table rows, 1: node ids
  unsigned long long* err;      // DEVICE [2]: (kind << 32 | mode, offending value) of the first bad index
  unsigned long long* err_host; // mapped pinned copy written by the last CTA of a *_host call, or null
  int32_t pair;                 // tensor-core path: launched as clusters of two CTAs (tiles padded to pairs per segment)
  int32_t stage_on;             // tensor-core path (STAGE kernel): helper warps copy rows of peer GPUs' tables into `stage`
  float* stage;                 // [CTAs][4 tiles][5 operands][128 rows][d] fp32, local memory
  uint32_t stage_mask;          // operands the helper may stage (bit o as in tc::staged_ops); the others are gathered in place
  int32_t wait_prev;            // tensor-core path: the index arrays are written by the kernel in front of this one in
                                // the stream (gqe_fetch_indices): every thread waits for it before its first index load
  ModeDev mode[kMaxModes];
};
constexpr int kPhaseSlots = 32;

struct OpParams {
  int32_t op;       // see OP_* below
  int32_t decoder;
  int32_t inter;
  int32_t n_rels;
  int32_t mutate;   // TransE path score: write the translated embeds1 back
  int64_t n;
  const float* table;
  int64_t table_rows;        // OP_ENCODE: rows of `table` (bounds check)
  unsigned long long* err;   // see LaunchParams::err
  const int32_t* rows;
  const float* rel[GQE_MAX_RELS];
  const float* pre;
  const float* post;
  const float* in0;
  const float* in1;
  const float* in2;
  float* io0;       // embeds1 when mutated in place
  float* out;
};
// Tensor-core path: rows per tile and the weight-packing request (gqe_pack).
constexpr int kTcTileRows = 128;
struct PackEntry {
  const float* src;
  uint8_t* dst;        // packed image of this entry (a slot of the context's weight cache)
  int32_t chain_form;  // 1: B[n][k] = M[k][n] (act.mm(M)); 0: B[n][k] = M[n][k] (M.mm(embeds))
  int32_t perm;        // 1: output columns permuted inside 16-column blocks for the fragment-layout
                       //    scoring of the accumulator (tc::score_col_src)
};
constexpr int kMaxPack = 5 * kMaxSegs;
struct PackParams {
  PackEntry e[kMaxPack];
};
// gqe_compose: dst = a . b ([d,d] row-major, bf16x3 products accumulated in fp32), one product
// per blockIdx.z.
struct ComposeEntry {
  const float* a;
  const float* b;
  float* dst;
};
constexpr int kMaxCompose = 3 * kMaxSegs;  // a structure needs at most three products
struct ComposeParams {
  ComposeEntry e[kMaxCompose];
};
// gqe_fetch_indices: index arrays of a *_host call copied from MAPPED pinned host memory by the SMs
constexpr int kMaxFetch = 6;
struct FetchParams {
  const int32_t* src[kMaxFetch];  // device alias of the caller's pinned host array
  int32_t* dst[kMaxFetch];
  int64_t n[kMaxFetch];           // 32-bit words
  int32_t count;
};
// gqe_store_batch: the slices of a device-resident query store gathered into the index arrays of a call,
// with one negative drawn per query (gqe_margin_loss_store_device)
struct StoreSliceDev {
  const int32_t* anchors;   // [n_anchor][block_q] node ids
  const int32_t* targets;   // [block_q]
  const int64_t* neg_ptr;   // CSR offsets [block_q + 1], or null: every query draws from the pool negs[0 .. pool_n)
  const int32_t* negs;
  int64_t block_q, start, n, out_q0, pool_n;
  int32_t n_anchor;
  int32_t tgt_mode;         // for the error report
};
struct StoreBatchParams {
  StoreSliceDev s[kMaxSegs];
  int32_t count;
  int64_t nq_total;
  unsigned long long seed;
  int32_t* anchors_out;     // [GQE_MAX_ANCHORS][nq_total]
  int32_t* pairs_out;       // [nq_total][2]
  unsigned long long* err;
};
enum { IDX_ERR_NO_NEGATIVE = 3 };
// gqe_score_pairs: one formula segment's (query, target) pairs against stored query embeddings
struct PairSeg {
  const float* tgt_table;
  int64_t q_begin, q_end;
  ModeDev mode;                   // the target mode (index resolution)
};
struct PairParams {
  PairSeg seg[kMaxSegs];
  int32_t n_segs;
  int32_t T;                      // regular layout: targets per query (offsets == null)
  const float* q;                 // fp32 [n_queries_total, D] query embeddings
  const int32_t* target_rows;
  const int64_t* target_offsets;  // ragged layout (single segment) or null
  int64_t n_pairs;                // ragged: total pairs
  float* out_scores;
  int32_t index_kind;
  unsigned long long* err;
};

// ---- index resolution (device side) ----------------------------------------------------
// All 32-bit: node ids, rows and map lengths are < 2^31, so "inside the map" and "inside the
// table" are ONE unsigned compare each (a negative value wraps to >= 2^31).
#ifdef __CUDACC__
static __device__ __noinline__ void report_index(unsigned long long* err, int kind, int mode, int32_t value) {
  if (err && atomicCAS(err, 0ull, ((unsigned long long)kind << 32) | (unsigned int)mode) == 0ull)
    err[1] = (unsigned long long)(long long)value;
}
// raw index (a table row, or a node id when index_kind != 0) -> row candidate; issues the node-map
// load and does not branch on loaded data.  -1: the id is outside the map.
__device__ __forceinline__ int32_t index_lookup(const ModeDev& m, int32_t v, int index_kind) {
  if (!index_kind) return v;
  const uint32_t k = (uint32_t)v - (uint32_t)m.base;
  const bool inside = k < m.len;
  return inside ? (m.lut ? __ldg(m.lut + k) : (int32_t)k) : -1;
}
// candidate -> checked row: a miss is reported (with the caller's raw index `v`) and reads row 0
__device__ __forceinline__ int32_t index_check(const ModeDev& m, int mode, int32_t cand, int32_t v, int index_kind,
                                               unsigned long long* err) {
  if ((uint32_t)cand >= m.rows) {
    report_index(err, (index_kind && cand < 0) ? IDX_ERR_UNKNOWN_NODE : IDX_ERR_ROW_RANGE, mode, v);
    return 0;
  }
  return cand;
}
// both halves back to back (callers that have nothing to put between the dependent loads)
__device__ __forceinline__ int32_t resolve_index(const ModeDev& m, int mode, int32_t v, int index_kind,
                                                 unsigned long long* err) {
  return index_check(m, mode, index_lookup(m, v, index_kind), v, index_kind, err);
}
#endif

enum { OP_ENCODE = 0, OP_PROJECT = 1, OP_PATH_SCORE = 2, OP_INTERSECT = 3, OP_COSINE = 4, OP_MATMUL = 5 };

}  // namespace gqe
