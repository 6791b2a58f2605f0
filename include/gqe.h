/*
 * gqe.h -- C ABI of the B200-native conjunctive-query scorer.
 *
 * This is the drop-in boundary for ONE path of williamleif/graphqembed
 * ("netquery"): the batched forward scoring of conjunctive graph queries and
 * its margin loss.  The reference has no FFI of its own (it is pure Python on
 * top of torch); what a maintainer would bind are the operator methods below.
 * Each entry point names the reference interface it replaces (file:line under
 * the reference tree).  See INTEGRATION.md for the ctypes stubs.
 *
 * Conventions
 *   - plain C types only; every pointer marked DEVICE is a CUDA device
 *     pointer on the context's device, every pointer marked HOST is host
 *     memory (pinned for true async, pageable works).
 *   - all work is enqueued on the context's stream; *_device entry points do
 *     not synchronise, *_host entry points return after their result landed.
 *   - return value: GQE_OK (0) or a negative gqe_status; the message of the
 *     last failure on a context is kept in gqe_last_error().
 *   - "row" always means a row of a mode's embedding table, i.e. the value
 *     node_maps[mode][node] + 1 of reference netquery/bio/data_utils.py:20-21.
 *   - "relation id" is the position of the canonical triple
 *     (from_mode, name, to_mode) in the decoder's registration order,
 *     reference netquery/decoders.py:135-137.
 *   - embeddings exchanged by the operator-level calls are FEATURE-MAJOR
 *     [d, n] fp32 (reference layout: encoders.py:41, decoders.py:150).
 *   - index arrays hold table ROWS (the entry points without "_nodes") or NODE IDS (the
 *     "_nodes" twins, mapped to rows on the device through gqe_bind_node_maps).  Every index is
 *     bounds-checked inside the kernels: a bad one reads row 0 and is reported -- the *_host
 *     calls return GQE_ERR_INDEX, the asynchronous *_device calls leave it for gqe_index_error.
 *   - packed / pre-multiplied images of the relation and DeepSets matrices are CACHED across
 *     calls on the tensor-core path: after changing a bound matrix in place (an optimiser
 *     step) call gqe_invalidate_weights, or turn the cache off with gqe_set_weight_cache.
 *     Embedding tables are always read in place.
 *   - there is no CPU fallback anywhere behind this header.
 */
#ifndef GQE_H_
#define GQE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GQE_ABI_VERSION 6
#define GQE_MAX_ANCHORS 3
#define GQE_MAX_RELS 3

typedef enum gqe_status {
  GQE_OK = 0,
  GQE_ERR_INVALID = -1,      /* bad argument (null pointer, negative size, unknown enum) */
  GQE_ERR_UNBOUND = -2,      /* tables / relations / intersection not bound yet */
  GQE_ERR_UNSUPPORTED = -3,  /* dimension or combination this build has no kernel for */
  GQE_ERR_CUDA = -4,         /* a CUDA runtime call failed; see gqe_last_error */
  GQE_ERR_NOMEM = -5,
  GQE_ERR_INDEX = -6         /* a node id is not in the bound node map (the reference's KeyError,
                                bio/data_utils.py:21) or a row index is outside its table (nn.Embedding's
                                IndexError); the offending value is in gqe_last_error */
} gqe_status;

/* Query structures of reference netquery/model.py:70-109. */
typedef enum gqe_structure {
  GQE_CHAIN1 = 0,       /* "1-chain"       model.py:71-76  */
  GQE_CHAIN2 = 1,       /* "2-chain"                       */
  GQE_CHAIN3 = 2,       /* "3-chain"                       */
  GQE_INTER2 = 3,       /* "2-inter"       model.py:77-98  */
  GQE_INTER3 = 4,       /* "3-inter"                       */
  GQE_INTER_CHAIN3 = 5, /* "3-inter_chain" model.py:83-86  */
  GQE_CHAIN_INTER3 = 6  /* "3-chain_inter" model.py:99-109 */
} gqe_structure;

/* Metapath decoders selectable by reference netquery/utils.py:128-137. */
typedef enum gqe_decoder {
  GQE_DEC_BILINEAR = 0, /* BilinearMetapathDecoder      decoders.py:123-150, params [R,d,d] */
  GQE_DEC_TRANSE = 1,   /* TransEMetapathDecoder        decoders.py:181-208, params [R,d]   */
  GQE_DEC_DISTMULT = 2  /* BilinearDiagMetapathDecoder  decoders.py:211-236, params [R,d]   */
} gqe_decoder;

/* Intersection operators selectable by reference netquery/utils.py:139-150. */
typedef enum gqe_inter {
  GQE_INTER_DEEPSETS_MEAN = 0, /* SetIntersection(agg=mean)       decoders.py:270-300 */
  GQE_INTER_DEEPSETS_MIN = 1,  /* SetIntersection(agg=min)                            */
  GQE_INTER_SIMPLE_MEAN = 2,   /* SimpleSetIntersection(agg=mean) decoders.py:302-319 */
  GQE_INTER_SIMPLE_MIN = 3     /* SimpleSetIntersection(agg=min)                      */
} gqe_inter;

/* Arithmetic of the d x d contractions (Bilinear projections, DeepSets pre/post).
 * Everything else (gather, normalisation, cosine, loss) is fp32 in both modes. */
typedef enum gqe_precision {
  GQE_PREC_BF16X3 = 0, /* default: tcgen05 tensor cores, operands split x = hi + lo in bf16, three
                          products hi*hi + lo*hi + hi*lo accumulated in fp32 (relative error ~2^-17
                          per product; scores within 1e-4 of the fp32 reference).  Used for the
                          Bilinear decoder at d = 128 / 256 with the regular target layout; every
                          other combination runs the fp32 kernels. */
  GQE_PREC_FP32 = 1    /* exact fp32 FMA on the CUDA cores everywhere */
} gqe_precision;

/*
 * A lowered Formula (reference netquery/graph.py:11-36): which table feeds
 * which operand and which relation parameter is applied in which order.
 * rel[] is in APPLICATION order, which is exactly the order in which the
 * reference touches its parameter dict:
 *   chains          rel[0..n-1] = r1..rn, applied to the TARGET (decoders.py:143-145)
 *   2/3-inter       rel[k] = reverse(r_k), applied to anchor k (model.py:81,88,92)
 *   3-inter_chain   rel[0] = reverse(r1) on anchor 0;
 *                   rel[1] = reverse(r2b) then rel[2] = reverse(r2a) on anchor 1
 *                   (model.py:84-86 walks rels[1][::-1])
 *   3-chain_inter   rel[0] = reverse(r2a) on anchor 0, rel[1] = reverse(r2b) on
 *                   anchor 1, rel[2] = reverse(r1) applied after the
 *                   intersection (model.py:102-107)
 * inter_mode is the mode whose pre/post matrices are used: the target mode, or
 * for 3-chain_inter the intermediate node's mode (model.py:106); -1 for chains.
 */
typedef struct gqe_plan {
  int32_t structure;                    /* gqe_structure */
  int32_t target_mode;
  int32_t anchor_mode[GQE_MAX_ANCHORS]; /* unused slots: -1 */
  int32_t inter_mode;
  int32_t rel[GQE_MAX_RELS];            /* unused slots: -1 */
} gqe_plan;

/* One formula's slice of a grouped (multi-formula) batch: queries
 * [query_begin, query_end) of the concatenated index arrays. */
typedef struct gqe_segment {
  gqe_plan plan;
  int64_t query_begin;
  int64_t query_end;
} gqe_segment;

typedef struct gqe_ctx gqe_ctx;

/* ---- context ----------------------------------------------------------- */
int gqe_abi_version(void);
/* stream: a cudaStream_t (NULL = the legacy default stream of `device`). */
int gqe_create(int device, void* stream, gqe_ctx** out);
void gqe_destroy(gqe_ctx* ctx);
int gqe_set_stream(gqe_ctx* ctx, void* stream);
/* Select / query the arithmetic of the contractions (gqe_precision). */
int gqe_set_precision(gqe_ctx* ctx, int32_t precision);
int gqe_get_precision(const gqe_ctx* ctx);
/* Operator pre-composition on the tensor-core path.  Runs of consecutive linear operators of
 * a formula (chained relation matrices; DeepSets pre x relation; relation x post) are
 * multiplied together once (split-bf16 products accumulated in fp32, ~2^-17 relative -- the
 * same arithmetic as the contractions themselves, NOT exact fp32) and cached, so the fused
 * kernel runs one contraction per run: a 3-chain costs one contraction per (query, target) pair
 * instead of three.  Same algebra as reference netquery/decoders.py:143-150,289-299, different
 * rounding: scores stay within ~4e-6 of the exact-fp32 kernels (tests/test_gpu_parity.py).
 * AUTO (default) composes whenever the weight cache is on, else when at least 8 tiles
 * (1024 rows) of a formula share the product. */
typedef enum gqe_compose_mode { GQE_COMPOSE_OFF = 0, GQE_COMPOSE_AUTO = 1, GQE_COMPOSE_ALWAYS = 2 } gqe_compose_mode;
int gqe_set_compose(gqe_ctx* ctx, int32_t mode);
/* Weight cache of the tensor-core path.  The d x d operator matrices a formula uses are split
 * into bf16 planes, swizzled and (with pre-composition) multiplied together before the fused
 * kernel can stream them; the context keeps those images keyed by the source pointers, so only
 * the first call that needs a matrix pays for it (two small kernels) and every later call is the
 * fused kernel alone.  The cache cannot see in-place updates of the parameters:
 *   gqe_invalidate_weights  drop every image (call after optimizer.step(); re-binding does it too)
 *   gqe_set_weight_cache    on (default) / off = re-prepare on every call, parameters fully live
 *   gqe_weight_prep_count   matrices packed so far on this context (bookkeeping) */
int gqe_set_weight_cache(gqe_ctx* ctx, int32_t on);
int gqe_invalidate_weights(gqe_ctx* ctx);
int64_t gqe_weight_prep_count(const gqe_ctx* ctx);
/* Message of the last failure on ctx (ctx == NULL: last gqe_create failure). */
const char* gqe_last_error(const gqe_ctx* ctx);
/* Diagnostics: while `log` (DEVICE uint64 [n_records][32], zeroed by the caller) is set, thread 0
 * of every tile < n_records - 256 of the tensor-core kernel records (tag << 56 | SM clock) stamps
 * at its phase boundaries (tags: 1+16*structure start, 2 gather done, 9 A operand handed over and
 * prefetches issued, 8 previous chain tile scored, 3 contraction done, 4 epilogue done,
 * 5 transposed, 6 scored, 7 end; slot 31 = %smid), and, when n_records >= 512, CTA b writes
 * (%globaltimer ns, clock64) pairs into record n_records-1-b at kernel entry, set-up done, first
 * tile taken, tile loop left and exit.  tools/phase_report.py decodes both.  NULL / 0 turns it
 * off. */
int gqe_debug_set_phase_log(gqe_ctx* ctx, uint64_t* log, int64_t n_tiles);
/* Diagnostics: the output column held by accumulator column n of a contraction whose result the
 * tensor-core kernel scores in the TMEM fragment layout (the last hop of a chain): gqe_pack
 * permutes the packed weights by it so that the four accumulator columns a lane owns are four
 * contiguous output columns.  A bijection of every 16-column block (host function, no GPU). */
int gqe_debug_score_col_src(int n);
/* Number of kernels this context has launched so far (bench bookkeeping). */
int64_t gqe_launch_count(const gqe_ctx* ctx);

/* ---- parameter binding (no copies: the pointers alias the owner's storage,
 *      so in-place optimiser updates are seen by the next call) -------------
 * Replaces: the `features` closure + nn.Embedding tables of reference
 * netquery/bio/data_utils.py:16-21 (one [rows, d] fp32 table per mode). */
int gqe_bind_tables(gqe_ctx* ctx, int32_t n_modes, const float* const* tables /*HOST array of DEVICE ptrs*/,
                    const int64_t* rows /*HOST [n_modes]*/, int32_t d);
/* rows[m] is also the bound every row index of mode m is checked against inside the kernels. */
/* Node id -> table row maps, one per mode, in the order of gqe_bind_tables.  Replaces the
 * node_maps dict lookup of the `features` closure (reference netquery/bio/data_utils.py:20-21,
 * CUDA variant utils.py:17-24): the "_nodes" entry points take the reference's node ids and
 * every kernel computes
 *     row = lut[m] ? lut[m][node - base[m]] : node - base[m]
 * in its index prologue.  lut[m]: DEVICE int32 [len[m]], entry = node_maps[mode][node] + 1, or -1
 * for an id that is not a node of the mode (reported as GQE_ERR_INDEX: the reference's KeyError).
 * lut == NULL (or lut[m] == NULL) selects the affine form; identity ids (utils.py:18-20:
 * row = node + 1) are base = -1.  n_modes == 0 unbinds.  The arrays are not copied. */
int gqe_bind_node_maps(gqe_ctx* ctx, int32_t n_modes, const int32_t* const* lut /*HOST array of DEVICE ptrs*/,
                       const int64_t* base /*HOST [n_modes]*/, const int64_t* len /*HOST [n_modes]*/);
/* First index error a kernel of this context has seen since the last call (synchronises the
 * stream): GQE_OK and *kind = 0 when there is none, else GQE_ERR_INDEX with *kind = 1 (unknown
 * node) or 2 (row out of range), the mode and the offending value; the record is cleared. */
int gqe_index_error(gqe_ctx* ctx, int32_t* kind, int32_t* mode, int64_t* value);
/* A table pointer may be a PEER pointer obtained from gqe_ipc_open (another
 * GPU's shard of the node-type-sharded table): the fused kernels then read
 * those rows in place over NVLink.  tables[m] == NULL with rows[m] == 0 marks
 * a mode that is absent on this rank; a plan that references it fails with
 * GQE_ERR_UNBOUND. */
/* Replaces: the `mats` / `vecs` parameter dicts of reference
 * netquery/decoders.py:129-140,188-197,217-226.  params[r] is [d,d] row-major
 * (bilinear) or [d] (transe / distmult). */
int gqe_bind_relations(gqe_ctx* ctx, int32_t decoder /*gqe_decoder*/, int32_t n_rels,
                       const float* const* params /*HOST array of DEVICE ptrs*/, int32_t d);
/* Replaces: `pre_mats` / `post_mats` of reference netquery/decoders.py:275-286
 * (pre[m]: [d_exp, d], post[m]: [d, d_exp], row-major; d_exp must equal d, as
 * reference netquery/utils.py:141 constructs it).  For the SIMPLE kinds pass
 * NULL arrays. */
int gqe_bind_intersection(gqe_ctx* ctx, int32_t inter /*gqe_inter*/, int32_t n_modes,
                          const float* const* pre, const float* const* post /*HOST arrays of DEVICE ptrs*/,
                          int32_t d, int32_t d_expand);

/* ---- the fused hot path --------------------------------------------------
 * gqe_score_device replaces QueryEncoderDecoder.forward (reference
 * netquery/model.py:70-109) for one formula.
 *   anchor_rows   DEVICE int32 [n_anchors][n_queries] (anchor k of query q at k*n_queries+q)
 *   target_rows   DEVICE int32 [n_pairs]; the targets of query q are
 *                 target_rows[target_offsets[q] .. target_offsets[q+1]) or, when
 *                 target_offsets == NULL, the regular layout
 *                 target_rows[q*T .. (q+1)*T) with T = n_pairs / n_queries
 *   target_offsets DEVICE int64 [n_queries+1] (offsets[0] == 0,
 *                 offsets[n_queries] == n_pairs, non-decreasing) or NULL
 *   out_scores    DEVICE fp32 [n_pairs], same order as target_rows
 * The query side (anchors, projections, intersection) is evaluated ONCE per
 * query however many targets it is scored against (the reference re-evaluates
 * it per pair: model.py:122-123, utils.py:58-60,86-88). */
int gqe_score_device(gqe_ctx* ctx, const gqe_plan* plan, int64_t n_queries,
                     const int32_t* anchor_rows, int64_t n_pairs, const int32_t* target_rows,
                     const int64_t* target_offsets, float* out_scores);

/* gqe_margin_loss_device replaces QueryEncoderDecoder.margin_loss (reference
 * netquery/model.py:112-127) once the negatives are chosen: per query one
 * positive and one negative target row,
 *   loss = mean_q max(0, margin - (score(q,pos) - score(q,neg))).
 *   pair_rows   DEVICE int32 [n_queries][2] (positive, negative)
 *   out_loss    DEVICE fp32 [1]
 *   out_scores  DEVICE fp32 [n_queries][2] or NULL
 * The reduction order is fixed (deterministic result). */
int gqe_margin_loss_device(gqe_ctx* ctx, const gqe_plan* plan, int64_t n_queries,
                           const int32_t* anchor_rows, const int32_t* pair_rows,
                           float margin, float* out_loss, float* out_scores);

/* Grouped variant: many formulas in one call (the "full mix" workload); any number of segments, one
 * launch per 96 formulas (one loss accumulator across the launches).
 *   segments      HOST array; segment s owns queries [query_begin, query_end)
 *   anchor_rows   DEVICE int32 [GQE_MAX_ANCHORS][n_queries_total]
 *   target_rows   DEVICE int32 [n_queries_total][targets_per_query]
 *   out_scores    DEVICE fp32 [n_queries_total][targets_per_query] or NULL when out_loss given
 *   out_loss      DEVICE fp32 [1] or NULL; requires targets_per_query == 2; mean over ALL queries */
int gqe_score_grouped_device(gqe_ctx* ctx, const gqe_segment* segments, int32_t n_segments,
                             int64_t n_queries_total, const int32_t* anchor_rows,
                             const int32_t* target_rows, int32_t targets_per_query,
                             float* out_scores, float margin, float* out_loss);

/* Node-id variants: identical to the three calls above except that the index arrays hold the
 * reference's NODE IDS (int32) instead of table rows; the lookup of
 * netquery/bio/data_utils.py:20-21 happens inside the kernel (gqe_bind_node_maps). */
int gqe_score_nodes_device(gqe_ctx* ctx, const gqe_plan* plan, int64_t n_queries,
                           const int32_t* anchor_nodes, int64_t n_pairs, const int32_t* target_nodes,
                           const int64_t* target_offsets, float* out_scores);
int gqe_margin_loss_nodes_device(gqe_ctx* ctx, const gqe_plan* plan, int64_t n_queries,
                                 const int32_t* anchor_nodes, const int32_t* pair_nodes,
                                 float margin, float* out_loss, float* out_scores);
int gqe_score_grouped_nodes_device(gqe_ctx* ctx, const gqe_segment* segments, int32_t n_segments,
                                   int64_t n_queries_total, const int32_t* anchor_nodes,
                                   const int32_t* target_nodes, int32_t targets_per_query,
                                   float* out_scores, float margin, float* out_loss);

/* Margin loss straight from a DEVICE-RESIDENT query store: what replaces, per training / evaluation-loss step,
 * the batch slicing of netquery/train_helpers.py:95-107 and the per-query random.choice of the negative
 * target (netquery/model.py:113-120) -- no index array crosses PCIe.  A store block is one formula's queries as
 * flat arrays uploaded once (graphqembed_b200/store.py); slice i supplies the queries of segments[i]:
 *   anchors        DEVICE int32 [n_anchor][block_queries], slot-major node ids of the whole block
 *   targets        DEVICE int32 [block_queries] positive target node ids
 *   neg_ptr, negs  DEVICE CSR (int64 [block_queries + 1], int32 values): the stored (or hard) negatives of every
 *                  query; or neg_ptr == NULL and negs = a pool of pool_size candidates every query draws from
 *                  (1-chain: every node of the target mode, model.py:118-119)
 *   start          first query of the slice inside the block; its length is the segment's query range
 * One kernel gathers the slices into the call's index arrays and draws, per query, one negative uniformly
 * from its list (counter-based generator: query position and `seed` -> the same draw whatever the launch
 * shape; pass a new seed every step); the fused scoring kernel follows in the same stream.  A query without
 * negatives is reported as GQE_ERR_INDEX (gqe_index_error), like the reference's IndexError.
 *   out_loss    DEVICE fp32 [1];  out_scores DEVICE fp32 [n_queries_total][2] or NULL
 *   out_pairs   DEVICE int32 [n_queries_total][2] or NULL: the (positive, negative) node ids that were scored
 * Needs gqe_bind_node_maps (the store holds node ids). */
typedef struct gqe_store_slice {
  const int32_t* anchors;
  const int32_t* targets;
  const int64_t* neg_ptr;
  const int32_t* negs;
  int64_t block_queries;
  int64_t start;
  int64_t pool_size;
} gqe_store_slice;
int gqe_margin_loss_store_device(gqe_ctx* ctx, const gqe_segment* segments, int32_t n_segments,
                                 const gqe_store_slice* slices, uint64_t seed, float margin,
                                 float* out_loss, float* out_scores, int32_t* out_pairs);

/* Host-buffer variants: same semantics, HOST index arrays in, HOST results
 * out; the H2D / D2H copies and a stream synchronise happen inside the call.
 * This is the call a non-torch host (the reference's own CPU pipeline) makes. */
int gqe_score_host(gqe_ctx* ctx, const gqe_plan* plan, int64_t n_queries,
                   const int32_t* anchor_rows, int64_t n_pairs, const int32_t* target_rows,
                   const int64_t* target_offsets, float* out_scores);
int gqe_margin_loss_host(gqe_ctx* ctx, const gqe_plan* plan, int64_t n_queries,
                         const int32_t* anchor_rows, const int32_t* pair_rows,
                         float margin, float* out_loss, float* out_scores);
int gqe_score_grouped_host(gqe_ctx* ctx, const gqe_segment* segments, int32_t n_segments,
                           int64_t n_queries_total, const int32_t* anchor_rows,
                           const int32_t* target_rows, int32_t targets_per_query,
                           float* out_scores, float margin, float* out_loss);

/* ... and from HOST node-id arrays: what replaces model.py:75-92 + bio/data_utils.py:20-21 +
 * the scoring itself for a host that holds the reference's query data as flat int32 arrays. */
int gqe_score_nodes_host(gqe_ctx* ctx, const gqe_plan* plan, int64_t n_queries,
                         const int32_t* anchor_nodes, int64_t n_pairs, const int32_t* target_nodes,
                         const int64_t* target_offsets, float* out_scores);
int gqe_margin_loss_nodes_host(gqe_ctx* ctx, const gqe_plan* plan, int64_t n_queries,
                               const int32_t* anchor_nodes, const int32_t* pair_nodes,
                               float margin, float* out_loss, float* out_scores);
int gqe_score_grouped_nodes_host(gqe_ctx* ctx, const gqe_segment* segments, int32_t n_segments,
                                 int64_t n_queries_total, const int32_t* anchor_nodes,
                                 const int32_t* target_nodes, int32_t targets_per_query,
                                 float* out_scores, float margin, float* out_loss);

/* ---- operator-level entry points (the un-fused reference surface) -------
 * All embeddings are DEVICE fp32 feature-major [d, n]. */
/* DirectEncoder.forward (reference netquery/encoders.py:29-43): gather rows of
 * `mode`, L2-normalise each column (no epsilon). */
int gqe_encode_device(gqe_ctx* ctx, int32_t mode, int64_t n, const int32_t* rows /*DEVICE*/, float* out);
/* *MetapathDecoder.project (decoders.py:149-150,207-208,235-236). out may alias in. */
int gqe_project_device(gqe_ctx* ctx, int32_t rel, int64_t n, const float* in, float* out);
/* *MetapathDecoder.forward (decoders.py:142-147,200-205,228-233).  For TransE
 * the reference translates embeds1 in place (decoders.py:203); set
 * mutate_embeds1 != 0 to reproduce that side effect. */
int gqe_path_score_device(gqe_ctx* ctx, int32_t n_rels, const int32_t* rels /*HOST*/, int64_t n,
                          float* embeds1, const float* embeds2, int32_t mutate_embeds1, float* out /*[n]*/);
/* SetIntersection.forward / SimpleSetIntersection.forward (decoders.py:288-300,311-319).
 * embeds3 may be NULL (two operands). */
int gqe_intersect_device(gqe_ctx* ctx, int32_t mode, int64_t n, const float* embeds1,
                         const float* embeds2, const float* embeds3, float* out);
/* nn.CosineSimilarity(dim=0, eps=1e-8) as used at model.py:68,97,108. */
int gqe_cosine_device(gqe_ctx* ctx, int32_t d, int64_t n, const float* x, const float* y, float* out /*[n]*/);

/* ---- differentiable operator surface ----------------------------------------
 * The training step of the reference is loss.backward() through the operators above
 * (netquery/train_helpers.py:78).  These are the forward pieces autograd needs as separate
 * calls and the vector-Jacobian product of each operator, all on feature-major fp32 [d, n]
 * DEVICE tensors; graphqembed_b200/autograd.py chains them the way torch.autograd chains
 * the reference's ops.  Parameter gradients (gw, gv, gtable) are ACCUMULATED into their
 * destination with atomicAdd; zero it first. */
/* out = W in (transpose == 0: M.mm(embeds), decoders.py:150,289,299) or W^T in
 * (transpose != 0: act.mm(M), decoders.py:145); w: DEVICE [d,d] row-major. */
int gqe_matmul_device(gqe_ctx* ctx, const float* w, int32_t transpose, int32_t d, int64_t n, const float* in, float* out);
/* gw += gy x^T (transpose == 0) or x gy^T (transpose != 0): gradient of the matrix above. */
int gqe_matmul_wgrad_device(gqe_ctx* ctx, int32_t transpose, int32_t d, int64_t n, const float* gy, const float* x,
                            float* gw);
/* gv[i] += sum_c gy[i][c] * (x ? x[i][c] : 1): gradient of a TransE / BilinearDiag relation
 * vector (decoders.py:203,208 / 231,236). */
int gqe_rowsum_device(gqe_ctx* ctx, int32_t d, int64_t n, const float* gy, const float* x, float* gv);
/* out = agg_k act(e_k): act = relu (DeepSets, decoders.py:289-296) or identity
 * (SimpleSetIntersection, decoders.py:311-316); agg = mean (use_min == 0) or min; e3 may be NULL. */
int gqe_aggregate_device(gqe_ctx* ctx, int32_t d, int64_t n, const float* e1, const float* e2, const float* e3,
                         int32_t relu, int32_t use_min, float* out);
int gqe_aggregate_bwd_device(gqe_ctx* ctx, int32_t d, int64_t n, const float* e1, const float* e2, const float* e3,
                             int32_t relu, int32_t use_min, const float* gout, float* g1, float* g2, float* g3);
/* out[c] = sum_k x[k][c] y[k][c]: the un-normalised BilinearDiag chain score (decoders.py:232). */
int gqe_dot_device(gqe_ctx* ctx, int32_t d, int64_t n, const float* x, const float* y, float* out);
/* VJP of gqe_cosine_device (raw_dot == 0) or gqe_dot_device (raw_dot != 0); gx or gy may be NULL. */
int gqe_cosine_bwd_device(gqe_ctx* ctx, int32_t d, int64_t n, const float* x, const float* y, const float* gout,
                          int32_t raw_dot, float* gx, float* gy);
/* VJP of gqe_encode_device: gtable[rows[c], :] += (g_c - x_hat (x_hat . g_c)) / |t|, gtable: DEVICE
 * [table rows, d] dense gradient of the mode's table (what nn.Embedding's backward produces). */
int gqe_encode_bwd_device(gqe_ctx* ctx, int32_t mode, int64_t n, const int32_t* rows, const float* gout, float* gtable);

/* ---- sparse training step -------------------------------------------------------
 * The reference trains with nn.Embedding tables and dense torch.optim.Adam
 * (netquery/bio/train.py:59-62, train_helpers.py:78-79): a dense [rows, d] gradient and a full-table
 * update per touched mode and step.  These two calls are the sparse equivalent.
 * gqe_encode_bwd_rows_device: VJP of gqe_encode_device PER GATHERED ROW,
 *   grad_rows[c, :] = (g_c - x_hat (x_hat . g_c)) / |t|  for column c of gout (DEVICE [d, n]) --
 *   with `rows` the table gradient as (index, value) pairs: no dense tensor, no atomics. */
int gqe_encode_bwd_rows_device(gqe_ctx* ctx, int32_t mode, int64_t n, const int32_t* rows, const float* gout,
                               float* grad_rows /*DEVICE [n, d]*/);
/* gqe_adam_rows_device: Adam on selected rows of one table with the trajectory of DENSE Adam.
 * exp_avg / exp_avg_sq: DEVICE [table_rows, d] moments; last_step: DEVICE int32 [table_rows], the
 * optimiser step of the TABLE up to which each row is up to date (all zero initially).
 *   grad_rows == NULL: catch the listed rows (rows == NULL: every row) up to `step`, i.e. apply the
 *     zero-gradient Adam steps last_step[r]+1 .. step that dense Adam applied to them (moments
 *     decay, the row keeps moving); duplicates in `rows` are allowed.  Call it for the rows a batch
 *     is about to read, and for all rows before evaluating / saving the table.
 *   grad_rows != NULL: `rows` (DEVICE int64 [n]) are UNIQUE and grad_rows (DEVICE [n, d]) their summed
 *     gradients: each row is caught up to step - 1, then takes Adam step `step` (numbered from 1;
 *     bias corrections 1 - beta^step as torch.optim.Adam, no weight decay, no amsgrad). */
int gqe_adam_rows_device(gqe_ctx* ctx, float* table, float* exp_avg, float* exp_avg_sq, int32_t* last_step,
                         int64_t table_rows, int32_t d, int64_t n, const int64_t* rows, const float* grad_rows,
                         int32_t step, float lr, float beta1, float beta2, float eps);

/* ---- native training step -----------------------------------------------------------
 * The body of the reference's training loop for one formula batch (netquery/train_helpers.py:76-79,
 * netquery/bio/train.py:59-62):
 *     optimizer.zero_grad(); loss = enc_dec.margin_loss(formula, queries); loss.backward(); optimizer.step()
 * with torch.optim.Adam (no weight decay, no amsgrad), as ONE call: the forward pass of
 * netquery/model.py:70-127 in exact fp32, its backward pass and the Adam update of every parameter
 * that received a gradient are launched back to back from native code.  The bound tables and operator
 * parameters are updated IN PLACE (they alias the owner's storage).  Optimiser state -- moments, step
 * counters, one [rows, d] gradient-accumulation buffer per trained table -- lives in the context.
 *   tables: only the rows the batch gathered are touched; a row is first caught up with the
 *     zero-gradient steps dense Adam applied to it since it was last touched (gqe_adam_rows_device),
 *     so the trajectory is dense torch.optim.Adam's.  Call gqe_train_flush before reading the tables
 *     outside the training step (evaluation, torch.save).
 *   operators: dense Adam on the relation matrices / vectors and pre / post matrices the formula used
 *     (torch.optim.Adam skips parameters without a gradient; so does this).
 * anchors: int32 [n_anchors, n_queries]; pairs: int32 [n_queries, 2] = (positive, negative) target --
 * table rows, or node ids for the *_nodes calls (gqe_bind_node_maps).  out_loss: the batch's margin
 * loss BEFORE the update, DEVICE (or, *_host: HOST) float.  The *_host calls take HOST index buffers,
 * synchronise and report a bad index as GQE_ERR_INDEX (the step has then been applied with row 0 in its
 * place); the *_device calls are asynchronous (poll gqe_index_error). */
typedef struct gqe_adam {
  float lr, beta1, beta2, eps;
} gqe_adam;
int gqe_train_step_device(gqe_ctx* ctx, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_rows,
                          const int32_t* pair_rows, float margin, const gqe_adam* hyper, float* out_loss);
int gqe_train_step_nodes_device(gqe_ctx* ctx, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_nodes,
                                const int32_t* pair_nodes, float margin, const gqe_adam* hyper, float* out_loss);
int gqe_train_step_host(gqe_ctx* ctx, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_rows,
                        const int32_t* pair_rows, float margin, const gqe_adam* hyper, float* out_loss);
int gqe_train_step_nodes_host(gqe_ctx* ctx, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_nodes,
                              const int32_t* pair_nodes, float margin, const gqe_adam* hyper, float* out_loss);
/* The reference's loop sums the weighted losses of several batches before ONE backward() / optimizer.step()
 * (netquery/train_helpers.py:63-79: the edge batch, plus path_weight / inter_weight times the other query
 * types, intersections also with hard negatives).  gqe_train_backward_*: forward + backward of one batch, the
 * gradients SCALED BY `weight` accumulated on top of those of earlier calls, no update (out_loss: the batch's own,
 * unweighted loss); gqe_train_apply: the Adam step on every parameter and table row that received a gradient
 * since the last apply.  gqe_train_step_* == gqe_train_backward_*(weight 1) + gqe_train_apply. */
int gqe_train_backward_device(gqe_ctx* ctx, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_rows,
                              const int32_t* pair_rows, float margin, float weight, const gqe_adam* hyper, float* out_loss);
int gqe_train_backward_nodes_device(gqe_ctx* ctx, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_nodes,
                                    const int32_t* pair_nodes, float margin, float weight, const gqe_adam* hyper,
                                    float* out_loss);
int gqe_train_backward_host(gqe_ctx* ctx, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_rows,
                            const int32_t* pair_rows, float margin, float weight, const gqe_adam* hyper, float* out_loss);
int gqe_train_backward_nodes_host(gqe_ctx* ctx, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_nodes,
                                  const int32_t* pair_nodes, float margin, float weight, const gqe_adam* hyper,
                                  float* out_loss);
int gqe_train_apply(gqe_ctx* ctx, const gqe_adam* hyper);
/* every row of every trained table up to date (the zero-gradient steps it still owes) */
int gqe_train_flush(gqe_ctx* ctx);
/* forget the optimiser state (moments, step counters); the parameters keep their values */
int gqe_train_reset(gqe_ctx* ctx);
/* Adam steps the table of `mode` has taken so far */
int64_t gqe_train_steps(const gqe_ctx* ctx, int32_t mode);

/* ---- GraphSAGE-style encoder (the --depth > 0 path; reference netquery/encoders.py:47-129,
 *      netquery/aggregators.py:17-68, built by netquery/utils.py:93-126) ---------------------
 * MeanAggregator.forward after its neighbour sampling: out[i, :] = mean of src[cols[j], :] over
 * j in [ptr[i], ptr[i+1]).  The reference multiplies a dense [batch, unique neighbours] mask with
 * the gathered feature matrix (aggregators.py:55-67); ptr / cols is that mask as the CSR it is.
 * src: DEVICE fp32 [n_src, d] row-major (an embedding table, or the output of a lower encoder for
 * the unique neighbours); ptr: DEVICE int64 [n + 1]; cols: DEVICE int32; out: DEVICE fp32 [n, d]. */
int gqe_segment_mean_device(gqe_ctx* ctx, const float* src, int64_t n_src, int32_t d, int64_t n,
                            const int64_t* ptr, const int32_t* cols, float* out);
/* The compress step of Encoder.forward: out[m, n] = W[m, k] . X[k, n], followed by ReLU when
 * relu != 0 (compress_params[mode].mm(combined); F.relu -- encoders.py:118-123).  Exact fp32. */
int gqe_linear_device(gqe_ctx* ctx, const float* w, int32_t m, int32_t k, int64_t n, const float* x,
                      int32_t relu, float* out);

/* ---- node-type-sharded tables across the GPUs of one box -------------------
 * (no counterpart in the reference, which is single-process: this is the
 * multi-GPU form of the `features` lookup of netquery/bio/data_utils.py:20-21.)
 * One process per GPU.  The owner of a shard exports its table with
 * gqe_ipc_export; the 64-byte handle and the offset travel to the other ranks
 * by whatever the host uses (torch.distributed here); each of them maps the
 * shard with gqe_ipc_open and binds the returned peer pointer in
 * gqe_bind_tables.  Peer access over NVLink is enabled by the open. */
#define GQE_IPC_HANDLE_BYTES 64
/* dev_ptr: any pointer inside a cudaMalloc'ed allocation of this process (not
 * a VMM / expandable-segments mapping).  handle_out: HOST [64]; offset_out:
 * byte offset of dev_ptr inside the exported allocation. */
int gqe_ipc_export(gqe_ctx* ctx, const void* dev_ptr, uint8_t* handle_out, int64_t* offset_out);
/* handle: HOST [64] from a DIFFERENT process on the same box.  The mapping is
 * reference counted per process (several tables may live in one allocation). */
int gqe_ipc_open(gqe_ctx* ctx, const uint8_t* handle, int64_t offset, void** peer_ptr_out);
int gqe_ipc_close(gqe_ctx* ctx, void* peer_ptr);
/* Staged alternative (owner side of an NCCL exchange): out[i, :] =
 * table[mode][rows[i], :], raw rows, DEVICE int32 rows [n] -> DEVICE fp32 [n, d]. */
int gqe_gather_rows_device(gqe_ctx* ctx, int32_t mode, int64_t n, const int32_t* rows, float* out);

#ifdef __cplusplus
}
#endif
#endif /* GQE_H_ */
