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
#define GQE_MAX_SEGS 32
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
