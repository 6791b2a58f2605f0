// gqe_capi.cu -- host side of the C ABI declared in include/gqe.h.
//
// Owns the context (bound parameter pointers, stream, scratch buffers),
// validates and resolves lowered formulas to device pointers, sizes grids and
// launches the fused kernels.  No torch types, no CPU compute path.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gqe.h"
#include "gqe_ctx.h"
#include "gqe_launch.h"

using namespace gqe;

static std::string g_create_error;

int gqe_fail(gqe_ctx* c, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (c) c->err = buf; else g_create_error = buf;
  return code;
}
#define fail gqe_fail

// The runtime keeps ONE "last error"; a launcher that reads it would otherwise report an
// error some earlier, unrelated call left behind.  Entry points drop (and, with
// GQE_DEBUG_ERRORS set, name) such leftovers before they launch anything.
static void drop_stale_error(const char* where) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess && getenv("GQE_DEBUG_ERRORS"))
    fprintf(stderr, "[gqe] stale CUDA error seen at %s: %s\n", where, cudaGetErrorString(e));
}

static bool is_ipc_pointer(const void* p);  // defined with the IPC entry points below

static bool dim_supported(int d) { return d == 32 || d == 64 || d == 128 || d == 256; }

static int n_anchors_of(int structure) {
  switch (structure) {
    case GQE_CHAIN1: case GQE_CHAIN2: case GQE_CHAIN3: return 1;
    case GQE_INTER2: case GQE_INTER_CHAIN3: case GQE_CHAIN_INTER3: return 2;
    case GQE_INTER3: return 3;
    default: return -1;
  }
}
static int n_rels_of(int structure) {
  switch (structure) {
    case GQE_CHAIN1: return 1;
    case GQE_CHAIN2: case GQE_INTER2: return 2;
    case GQE_CHAIN3: case GQE_INTER3: case GQE_INTER_CHAIN3: case GQE_CHAIN_INTER3: return 3;
    default: return -1;
  }
}

extern "C" int gqe_abi_version(void) { return GQE_ABI_VERSION; }

extern "C" int gqe_create(int device, void* stream, gqe_ctx** out) {
  if (!out) return fail(nullptr, GQE_ERR_INVALID, "gqe_create: out is null");
  *out = nullptr;
  int n_dev = 0;
  cudaError_t e = cudaGetDeviceCount(&n_dev);
  if (e != cudaSuccess || n_dev == 0)
    return fail(nullptr, GQE_ERR_CUDA, "gqe_create: no CUDA device (%s); this library has no CPU path",
                cudaGetErrorString(e));
  if (device < 0 || device >= n_dev) return fail(nullptr, GQE_ERR_INVALID, "gqe_create: device %d out of range", device);
  cudaDeviceProp prop;
  GQE_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(nullptr, GQE_ERR_UNSUPPORTED, "gqe_create: device %d is sm_%d%d; this build targets sm_100a only",
                device, prop.major, prop.minor);
  GQE_CUDA(nullptr, cudaSetDevice(device));
  gqe_ctx* c = new gqe_ctx();
  c->device = device;
  c->stream = (cudaStream_t)stream;
  if (cudaMalloc(&c->loss_acc, sizeof(double)) != cudaSuccess ||
      cudaMalloc(&c->d_err, 2 * sizeof(unsigned long long)) != cudaSuccess ||
      cudaHostAlloc((void**)&c->h_err, 2 * sizeof(unsigned long long), cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer((void**)&c->h_err_dev, c->h_err, 0) != cudaSuccess ||
      cudaMalloc(&c->ticket, sizeof(unsigned int)) != cudaSuccess ||
      cudaMalloc(&c->tile_counter, sizeof(unsigned int)) != cudaSuccess) {
    delete c;
    return fail(nullptr, GQE_ERR_NOMEM, "gqe_create: cudaMalloc failed");
  }
  cudaMemset(c->loss_acc, 0, sizeof(double));
  cudaMemset(c->d_err, 0, 2 * sizeof(unsigned long long));
  c->h_err[0] = c->h_err[1] = 0ull;
  cudaMemset(c->ticket, 0, sizeof(unsigned int));
  cudaMemset(c->tile_counter, 0, sizeof(unsigned int));
  drop_stale_error("gqe_create (end)");
  *out = c;
  return GQE_OK;
}

extern "C" void gqe_destroy(gqe_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  gqe_train_destroy(c);
  cudaFree(c->partials);
  cudaFree(c->packed);
  cudaFree(c->qbuf);
  cudaFree(c->stage_buf);
  cudaFree(c->compose_buf);
  cudaFree(c->loss_acc);
  cudaFree(c->d_err);
  if (c->h_err) cudaFreeHost(c->h_err);
  cudaFree(c->ticket);
  cudaFree(c->tile_counter);
  for (void* p : c->stage) cudaFree(p);
  if (c->h_loss) cudaFreeHost(c->h_loss);
  drop_stale_error("gqe_destroy (end)");
  delete c;
}

extern "C" int gqe_set_stream(gqe_ctx* c, void* stream) {
  if (!c) return GQE_ERR_INVALID;
  c->stream = (cudaStream_t)stream;
  return GQE_OK;
}

extern "C" int gqe_set_precision(gqe_ctx* c, int32_t precision) {
  if (!c) return GQE_ERR_INVALID;
  if (precision != GQE_PREC_BF16X3 && precision != GQE_PREC_FP32)
    return fail(c, GQE_ERR_INVALID, "gqe_set_precision: unknown precision %d", precision);
  c->precision = precision;
  return GQE_OK;
}
extern "C" int gqe_get_precision(const gqe_ctx* c) { return c ? c->precision : GQE_ERR_INVALID; }

extern "C" int gqe_set_compose(gqe_ctx* c, int32_t mode) {
  if (!c) return GQE_ERR_INVALID;
  if (mode < GQE_COMPOSE_OFF || mode > GQE_COMPOSE_ALWAYS)
    return fail(c, GQE_ERR_INVALID, "gqe_set_compose: unknown mode %d", mode);
  c->compose = mode;
  return GQE_OK;
}

extern "C" int gqe_set_weight_cache(gqe_ctx* c, int32_t on) {
  if (!c) return GQE_ERR_INVALID;
  c->wcache_on = on != 0;
  c->wcache.clear();
  return GQE_OK;
}
extern "C" int gqe_invalidate_weights(gqe_ctx* c) {
  if (!c) return GQE_ERR_INVALID;
  c->wcache.clear();   // slots are rewritten in stream order: kernels in flight keep reading the old images
  return GQE_OK;
}
extern "C" int64_t gqe_weight_prep_count(const gqe_ctx* c) { return c ? c->weight_preps : 0; }

extern "C" int gqe_bind_node_maps(gqe_ctx* c, int32_t n_modes, const int32_t* const* lut, const int64_t* base,
                                  const int64_t* len) {
  if (!c) return GQE_ERR_INVALID;
  if (n_modes == 0) { c->node_maps.clear(); return GQE_OK; }
  if (n_modes < 0 || !base || !len) return fail(c, GQE_ERR_INVALID, "gqe_bind_node_maps: bad arguments");
  if (n_modes > kMaxModes) return fail(c, GQE_ERR_UNSUPPORTED, "gqe_bind_node_maps: more than %d node types", kMaxModes);
  std::vector<ModeDev> maps(n_modes);
  for (int m = 0; m < n_modes; ++m) {
    if (len[m] < 0 || len[m] > 0x7fffffffLL || base[m] < -0x80000000LL || base[m] > 0x7fffffffLL)
      return fail(c, GQE_ERR_INVALID, "gqe_bind_node_maps: mode %d: node ids and map lengths must fit 32 bits", m);
    maps[m].lut = lut ? lut[m] : nullptr;   // null: affine map row = node - base (identity ids: base = -1)
    maps[m].base = (int32_t)base[m];
    maps[m].len = (uint32_t)len[m];
    maps[m].rows = 0;
    maps[m].pad_ = 0;
  }
  c->node_maps.swap(maps);
  return GQE_OK;
}

extern "C" int gqe_debug_set_phase_log(gqe_ctx* c, uint64_t* log, int64_t n_tiles) {
  if (!c) return GQE_ERR_INVALID;
  if ((log == nullptr) != (n_tiles == 0) || n_tiles < 0) return fail(c, GQE_ERR_INVALID, "gqe_debug_set_phase_log: bad arguments");
  c->phase_log = reinterpret_cast<unsigned long long*>(log);
  c->phase_cap = n_tiles;
  return GQE_OK;
}

extern "C" const char* gqe_last_error(const gqe_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }
extern "C" int64_t gqe_launch_count(const gqe_ctx* c) { return c ? c->launches : 0; }
extern "C" int gqe_debug_score_col_src(int n) { return n < 0 ? -1 : score_col_src_host(n); }

// ---- binding ---------------------------------------------------------------
extern "C" int gqe_bind_tables(gqe_ctx* c, int32_t n_modes, const float* const* tables, const int64_t* rows,
                               int32_t d) {
  if (!c) return GQE_ERR_INVALID;
  if (n_modes <= 0 || !tables || !rows) return fail(c, GQE_ERR_INVALID, "gqe_bind_tables: bad arguments");
  if (!dim_supported(d))
    return fail(c, GQE_ERR_UNSUPPORTED, "gqe_bind_tables: embedding dimension %d not supported (32/64/128/256)", d);
  for (int m = 0; m < n_modes; ++m) {
    // (NULL, 0 rows) = a mode that is absent on this rank (sharded tables)
    if (!tables[m] && rows[m] == 0) continue;
    if (!tables[m] || rows[m] <= 0) return fail(c, GQE_ERR_INVALID, "gqe_bind_tables: mode %d has no table", m);
  }
  GQE_CUDA(c, cudaSetDevice(c->device));
  std::vector<char> remote(n_modes, 0);
  for (int m = 0; m < n_modes; ++m) {
    if (!tables[m]) continue;
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, tables[m]);
    if (e != cudaSuccess || (attr.type != cudaMemoryTypeDevice && attr.type != cudaMemoryTypeManaged)) {
      cudaGetLastError();
      return fail(c, GQE_ERR_INVALID, "gqe_bind_tables: table %d is not a device pointer", m);
    }
    // a peer shard: mapped by gqe_ipc_open (CUDA reports IPC mappings under the importing
    // device), or any other pointer that resolves to a different device
    remote[m] = attr.device != c->device || is_ipc_pointer(tables[m]);
  }
  c->tables.assign(tables, tables + n_modes);
  c->table_rows.assign(rows, rows + n_modes);
  c->table_remote.swap(remote);
  c->d = d;
  drop_stale_error("gqe_bind_tables (end)");
  return GQE_OK;
}

extern "C" int gqe_bind_relations(gqe_ctx* c, int32_t decoder, int32_t n_rels, const float* const* params,
                                  int32_t d) {
  if (!c) return GQE_ERR_INVALID;
  if (decoder < GQE_DEC_BILINEAR || decoder > GQE_DEC_DISTMULT)
    return fail(c, GQE_ERR_INVALID, "gqe_bind_relations: unknown decoder %d", decoder);
  if (n_rels <= 0 || !params) return fail(c, GQE_ERR_INVALID, "gqe_bind_relations: bad arguments");
  if (!dim_supported(d)) return fail(c, GQE_ERR_UNSUPPORTED, "gqe_bind_relations: dimension %d not supported", d);
  for (int r = 0; r < n_rels; ++r)
    if (!params[r]) return fail(c, GQE_ERR_INVALID, "gqe_bind_relations: relation %d has no parameter", r);
  c->rels.assign(params, params + n_rels);
  c->wcache.clear();
  c->decoder = decoder;
  c->rel_d = d;
  return GQE_OK;
}

extern "C" int gqe_bind_intersection(gqe_ctx* c, int32_t inter, int32_t n_modes, const float* const* pre,
                                     const float* const* post, int32_t d, int32_t d_expand) {
  if (!c) return GQE_ERR_INVALID;
  if (inter < GQE_INTER_DEEPSETS_MEAN || inter > GQE_INTER_SIMPLE_MIN)
    return fail(c, GQE_ERR_INVALID, "gqe_bind_intersection: unknown kind %d", inter);
  const bool deepsets = inter <= GQE_INTER_DEEPSETS_MIN;
  c->pre.clear();
  c->post.clear();
  c->wcache.clear();
  if (deepsets) {
    if (n_modes <= 0 || !pre || !post) return fail(c, GQE_ERR_INVALID, "gqe_bind_intersection: pre/post required");
    if (d_expand != d)
      return fail(c, GQE_ERR_UNSUPPORTED, "gqe_bind_intersection: expand dim %d != embed dim %d not supported", d_expand, d);
    if (!dim_supported(d)) return fail(c, GQE_ERR_UNSUPPORTED, "gqe_bind_intersection: dimension %d not supported", d);
    for (int m = 0; m < n_modes; ++m)
      if (!pre[m] || !post[m]) return fail(c, GQE_ERR_INVALID, "gqe_bind_intersection: mode %d has no matrices", m);
    c->pre.assign(pre, pre + n_modes);
    c->post.assign(post, post + n_modes);
  }
  c->inter = inter;
  c->inter_d = d;
  return GQE_OK;
}

// ---- plan resolution ---------------------------------------------------------
static int resolve(gqe_ctx* c, const gqe_plan& pl, SegDev* s) {
  const int na = n_anchors_of(pl.structure), nr = n_rels_of(pl.structure);
  if (na < 0) return fail(c, GQE_ERR_INVALID, "unknown query structure %d", pl.structure);
  if (c->tables.empty()) return fail(c, GQE_ERR_UNBOUND, "embedding tables are not bound");
  if (c->rels.empty()) return fail(c, GQE_ERR_UNBOUND, "relation parameters are not bound");
  if (c->rel_d != c->d) return fail(c, GQE_ERR_UNSUPPORTED, "relation dim %d != table dim %d", c->rel_d, c->d);
  const int nm = (int)c->tables.size();
  if (pl.target_mode < 0 || pl.target_mode >= nm) return fail(c, GQE_ERR_INVALID, "target mode %d out of range", pl.target_mode);
  std::memset(s, 0, sizeof *s);
  s->structure = pl.structure;
  s->n_anchor = na;
  s->tgt_table = c->tables[pl.target_mode];
  s->tgt_mode = (int8_t)pl.target_mode;
  if (!s->tgt_table) return fail(c, GQE_ERR_UNBOUND, "target mode %d has no table on this rank", pl.target_mode);
  if (c->table_remote[pl.target_mode]) s->remote_mask |= 8u;
  for (int k = 0; k < na; ++k) {
    if (pl.anchor_mode[k] < 0 || pl.anchor_mode[k] >= nm)
      return fail(c, GQE_ERR_INVALID, "anchor %d mode %d out of range", k, pl.anchor_mode[k]);
    s->anc_table[k] = c->tables[pl.anchor_mode[k]];
    s->anc_mode[k] = (int8_t)pl.anchor_mode[k];
    if (!s->anc_table[k])
      return fail(c, GQE_ERR_UNBOUND, "anchor %d mode %d has no table on this rank", k, pl.anchor_mode[k]);
    if (c->table_remote[pl.anchor_mode[k]]) s->remote_mask |= 1u << k;
  }
  for (int k = 0; k < nr; ++k) {
    if (pl.rel[k] < 0 || pl.rel[k] >= (int)c->rels.size())
      return fail(c, GQE_ERR_INVALID, "relation id %d out of range", pl.rel[k]);
    s->rel[k] = c->rels[pl.rel[k]];
  }
  if (pl.structure >= GQE_INTER2) {
    if (c->inter < 0) return fail(c, GQE_ERR_UNBOUND, "intersection operator is not bound");
    if (c->inter <= GQE_INTER_DEEPSETS_MIN) {
      if (c->inter_d != c->d) return fail(c, GQE_ERR_UNSUPPORTED, "intersection dim %d != table dim %d", c->inter_d, c->d);
      if (pl.inter_mode < 0 || pl.inter_mode >= (int)c->pre.size())
        return fail(c, GQE_ERR_INVALID, "intersection mode %d out of range", pl.inter_mode);
      s->pre = c->pre[pl.inter_mode];
      s->post = c->post[pl.inter_mode];
    }
  }
  return GQE_OK;
}

static int ensure_partials(gqe_ctx* c, int64_t n) {
  if (n <= c->partials_cap) return GQE_OK;
  // the old buffer may still be read by a kernel in flight on the stream
  GQE_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaFree(c->partials);
  c->partials = nullptr;
  c->partials_cap = 0;
  const int64_t cap = std::max<int64_t>(n, 4096);
  GQE_CUDA(c, cudaMalloc(&c->partials, cap * sizeof(double)));
  c->partials_cap = cap;
  return GQE_OK;
}

// ---- index errors ----------------------------------------------------------------
// Kernels report the first bad index (unknown node id, row outside its table) in c->d_err and
// read row 0 instead.  *_host calls fetch the word with their result and fail with
// GQE_ERR_INDEX; *_device calls are asynchronous, the caller polls gqe_index_error().
static int index_error_from(gqe_ctx* c, unsigned long long w0, unsigned long long w1) {
  const int kind = (int)(w0 >> 32), mode = (int)(w0 & 0xffffffffu);
  const long long value = (long long)w1;
  if (kind == IDX_ERR_UNKNOWN_NODE)
    return fail(c, GQE_ERR_INDEX, "unknown node: id %lld is not in the node map of mode %d", value, mode);
  if (kind == IDX_ERR_NO_NEGATIVE)
    return fail(c, GQE_ERR_INDEX, "query %lld of the call has no negative sample to draw from", value);
  return fail(c, GQE_ERR_INDEX, "row index out of range: %lld is outside the table of mode %d", value, mode);
}

extern "C" int gqe_index_error(gqe_ctx* c, int32_t* kind, int32_t* mode, int64_t* value) {
  if (!c) return GQE_ERR_INVALID;
  GQE_CUDA(c, cudaSetDevice(c->device));
  unsigned long long w[2] = {0ull, 0ull};
  GQE_CUDA(c, cudaMemcpyAsync(w, c->d_err, sizeof w, cudaMemcpyDeviceToHost, c->stream));
  GQE_CUDA(c, cudaStreamSynchronize(c->stream));
  if (kind) *kind = (int32_t)(w[0] >> 32);
  if (mode) *mode = (int32_t)(w[0] & 0xffffffffu);
  if (value) *value = (int64_t)w[1];
  if (w[0] == 0ull) return GQE_OK;
  GQE_CUDA(c, cudaMemsetAsync(c->d_err, 0, sizeof w, c->stream));
  return index_error_from(c, w[0], w[1]);
}

// ---- packed-weight cache ------------------------------------------------------------
constexpr int kWeightSlots = 1024;  // packed images the cache can hold (256 MiB at d = 256, 64 MiB at d = 128)

static int ensure_weight_buffers(gqe_ctx* c) {
  const size_t need = (size_t)kWeightSlots * tc_packed_bytes(c->d);
  if (c->packed_cap < need) {
    GQE_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->packed);
    c->packed = nullptr;
    c->packed_cap = 0;
    c->wcache.clear();
    GQE_CUDA(c, cudaMalloc(&c->packed, need));
    c->packed_cap = need;
  }
  const size_t need_c = (size_t)kMaxCompose * c->d * c->d * sizeof(float);
  if (c->compose != GQE_COMPOSE_OFF && c->compose_cap < need_c) {
    GQE_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaFree(c->compose_buf);
    c->compose_buf = nullptr;
    c->compose_cap = 0;
    GQE_CUDA(c, cudaMalloc(&c->compose_buf, need_c));
    c->compose_cap = need_c;
  }
  return GQE_OK;
}

// The one launcher behind every fused entry point.  index_kind: 0 = the index arrays hold table
// rows, 1 = node ids (mapped through gqe_bind_node_maps inside the kernels).  err_host: mapped
// pinned word the last CTA copies the index-error word to (the *_host calls), or null.
static const bool kStageRemote = [] {
  const char* e = getenv("GQE_STAGE");
  return e ? atoi(e) != 0 : true;
}();
static const uint32_t kStageMask = [] {   // GQE_STAGE_MASK=0x..: operands the helper warp stages (default: all five)
  const char* e = getenv("GQE_STAGE_MASK");
  return e ? (uint32_t)strtoul(e, nullptr, 0) : 0x1Fu;
}();
static const int kForceStage = [] {
  const char* e = getenv("GQE_FORCE_STAGE");
  return e ? atoi(e) : 0;
}();
static int run_fused(gqe_ctx* c, const gqe_segment* segs, int32_t n_segs, int64_t nq_total,
                     const int32_t* anchor_rows, int64_t n_pairs, const int32_t* target_rows,
                     const int64_t* target_offsets, int32_t T, float* out_scores, float margin, float* out_loss,
                     int index_kind, unsigned long long* err_host = nullptr) {
  if (!c) return GQE_ERR_INVALID;
  if (!segs || n_segs <= 0) return fail(c, GQE_ERR_INVALID, "no segments");
  if (nq_total < 0 || n_pairs < 0) return fail(c, GQE_ERR_INVALID, "negative size");
  if (!out_scores && !out_loss && n_pairs > 0) return fail(c, GQE_ERR_INVALID, "no output requested");
  if (target_offsets && n_segs != 1) return fail(c, GQE_ERR_INVALID, "ragged targets need a single formula");
  if (target_offsets && out_loss) return fail(c, GQE_ERR_INVALID, "margin loss needs the regular (pos,neg) layout");
  if (out_loss && T != 2) return fail(c, GQE_ERR_INVALID, "margin loss needs exactly 2 targets per query");
  drop_stale_error("fused entry");
  if (!target_offsets && T <= 0 && nq_total > 0) return fail(c, GQE_ERR_INVALID, "targets per query must be positive");
  if (index_kind && c->node_maps.empty()) return fail(c, GQE_ERR_UNBOUND, "node maps are not bound (gqe_bind_node_maps)");
  if (index_kind && c->node_maps.size() != c->tables.size())
    return fail(c, GQE_ERR_INVALID, "node maps cover %d modes, tables %d", (int)c->node_maps.size(), (int)c->tables.size());
  if ((int)c->tables.size() > kMaxModes)
    return fail(c, GQE_ERR_UNSUPPORTED, "more than %d node types are not supported by the fused kernels", kMaxModes);
  GQE_CUDA(c, cudaSetDevice(c->device));

  if (out_loss) {
    // the last launch of a call leaves the accumulator zeroed (LaunchParams::final_launch)
    if (c->loss_dirty) GQE_CUDA(c, cudaMemsetAsync(c->loss_acc, 0, sizeof(double), c->stream));
    c->loss_dirty = false;
    // mean over an empty batch is NaN in the reference (torch mean of an empty tensor)
    if (nq_total == 0) GQE_CUDA(c, cudaMemsetAsync(out_loss, 0xFF, sizeof(float), c->stream));
  }
  if (nq_total == 0) return GQE_OK;
  if (!anchor_rows || !target_rows) return fail(c, GQE_ERR_INVALID, "index arrays are null");

  LaunchParams lp;
  std::memset(&lp, 0, sizeof lp);
  lp.decoder = c->decoder;
  lp.inter = c->inter < 0 ? 0 : c->inter;
  lp.anchor_rows = anchor_rows;
  lp.anchor_stride = nq_total;
  lp.target_rows = target_rows;
  lp.target_offsets = target_offsets;
  lp.n_pairs = n_pairs;
  lp.T = T;
  lp.out_scores = out_scores;
  lp.out_loss = out_loss;
  lp.margin = margin;
  lp.inv_q = 1.0 / (double)nq_total;
  lp.loss_acc = c->loss_acc;
  lp.ticket = c->ticket;
  lp.tile_counter = c->tile_counter;
  lp.phase_log = c->phase_log;
  lp.phase_cap = c->phase_cap;
  lp.index_kind = index_kind;
  lp.err = c->d_err;
  for (size_t m = 0; m < c->tables.size(); ++m) {
    ModeDev& md = lp.mode[m];
    if (index_kind) md = c->node_maps[m];
    else { md.lut = nullptr; md.base = 0; md.len = 0; }
    md.rows = (uint32_t)std::min<int64_t>(c->table_rows[m], 0x7fffffffLL);
  }

  // Bilinear d x d contractions go to the tensor cores (tcgen05, bf16x3 split) unless the
  // context asks for exact fp32; the ragged target layout stays on the fp32 kernels.
  const bool ragged_chain = target_offsets != nullptr && segs[0].plan.structure <= GQE_CHAIN3;
  const bool use_tc = c->precision == GQE_PREC_BF16X3 && c->decoder == GQE_DEC_BILINEAR && tc_dim_supported(c->d) &&
                      !ragged_chain;
  // More than (positive, negative) per query -- the evaluation shape (utils.py:70-91): the
  // intersection structures build each query embedding once, leave it in HBM and a second,
  // purely HBM-bound kernel scores every (query, target) pair against it.
  const bool pairs_mode = (target_offsets != nullptr || T > 2) && out_scores != nullptr;
  if (pairs_mode) {
    const size_t need = (size_t)nq_total * c->d * sizeof(float);
    if (c->qbuf_cap < need) {
      GQE_CUDA(c, cudaStreamSynchronize(c->stream));
      cudaFree(c->qbuf);
      c->qbuf = nullptr;
      c->qbuf_cap = 0;
      GQE_CUDA(c, cudaMalloc(&c->qbuf, need + need / 4));
      c->qbuf_cap = need + need / 4;
    }
    lp.q_out = c->qbuf;
  }
  // TransE / DistMult with chains or the element-wise intersections have no contraction at all:
  // they run the streaming warp-per-query kernel (gqe_vec.cu), pure HBM traffic
  const bool use_vec = !use_tc && c->decoder != GQE_DEC_BILINEAR && !pairs_mode && !target_offsets && T <= 2 &&
                       (c->inter >= GQE_INTER_SIMPLE_MEAN || c->inter < 0 || [&] {
                         for (int32_t k = 0; k < n_segs; ++k)
                           if (segs[k].plan.structure >= GQE_INTER2 && segs[k].query_end > segs[k].query_begin) return false;
                         return true;
                       }());
  const int64_t tile_rows = use_tc ? kTcTileRows : kTileRows;
  if (use_tc) {
    if (int rc = ensure_weight_buffers(c)) return rc;
    if (!c->wcache_on) c->wcache.clear();   // parameters are treated as live: re-prepared every call
  }

  int32_t last_nonempty = -1;
  for (int32_t k = 0; k < n_segs; ++k)
    if (segs[k].query_end > segs[k].query_begin) last_nonempty = k;
  int32_t i = 0;
  while (i < n_segs) {
    int n = 0;
    int64_t tiles = 0;
    PackParams pp;
    int n_pack = 0;
    // fp32 products of runs of linear operators, level 0 = both operands are parameters, level 1 =
    // one operand is a level-0 product of this call (three-factor runs): two stream-ordered launches
    ComposeParams cw[2];
    int n_cw[2] = {0, 0};
    // cache full (worst case of this launch: 5 images per formula): start over
    if (use_tc && (int)c->wcache.size() + 5 * std::min<int>(n_segs - i, kMaxSegs) > kWeightSlots) c->wcache.clear();
    auto product = [&](const float* a, const float* b, int level) -> const float* {
      for (int l = 0; l < 2; ++l)
        for (int k = 0; k < n_cw[l]; ++k)
          if (cw[l].e[k].a == a && cw[l].e[k].b == b) return cw[l].e[k].dst;
      ComposeEntry& e = cw[level].e[n_cw[level]];
      e.a = a;
      e.b = b;
      e.dst = c->compose_buf + (size_t)(n_cw[0] + n_cw[1]) * c->d * c->d;
      ++n_cw[level];
      return e.dst;
    };
    // Packed image of a matrix or of a product of up to three (f0 f1 f2; right_assoc: f0 (f1 f2))
    // in the orientation its use needs.  `perm`: the matrix of a contraction whose accumulator is
    // scored in the TMEM fragment layout (the last hop of a chain) has its output columns
    // permuted at packing time.  Served from the context's cache when present.
    auto packed_of = [&](const float* f0, const float* f1, const float* f2, int right_assoc, int chain_form,
                         int perm) -> const float* {
      gqe_ctx::WKey key;
      key.f[0] = f0; key.f[1] = f1; key.f[2] = f2;
      key.n = f2 ? 3 : (f1 ? 2 : 1);
      key.right_assoc = key.n == 3 ? right_assoc : 0;
      key.chain_form = chain_form;
      key.perm = perm;
      size_t slot = 0;
      for (; slot < c->wcache.size(); ++slot)
        if (c->wcache[slot] == key) break;
      uint8_t* dst = c->packed + slot * tc_packed_bytes(c->d);
      if (slot == c->wcache.size()) {
        c->wcache.push_back(key);
        const float* src = f0;
        if (key.n == 2) src = product(f0, f1, 0);
        else if (key.n == 3) src = right_assoc ? product(f0, product(f1, f2, 0), 1) : product(product(f0, f1, 0), f2, 1);
        pp.e[n_pack].src = src;
        pp.e[n_pack].dst = dst;
        pp.e[n_pack].chain_form = chain_form;
        pp.e[n_pack].perm = perm;
        ++n_pack;
      }
      return reinterpret_cast<const float*>(dst);
    };
    int first_structure = -1;
    bool uniform = true;
    while (i < n_segs && n < kMaxSegs) {
      const gqe_segment& g = segs[i++];
      if (g.query_begin < 0 || g.query_end < g.query_begin || g.query_end > nq_total)
        return fail(c, GQE_ERR_INVALID, "segment query range [%lld,%lld) outside [0,%lld)", (long long)g.query_begin,
                    (long long)g.query_end, (long long)nq_total);
      if (g.query_end == g.query_begin) continue;
      SegDev* s = &lp.seg[n];
      int rc = resolve(c, g.plan, s);
      if (rc != GQE_OK) return rc;
      s->q_begin = g.query_begin;
      s->q_end = g.query_end;
      s->tile_begin = tiles;
      const int64_t nq = g.query_end - g.query_begin;
      const int64_t rows = g.plan.structure <= GQE_CHAIN3 ? (target_offsets ? n_pairs : nq * T) : nq;
      tiles += (rows + tile_rows - 1) / tile_rows;
      if (use_tc) {
        const int st = g.plan.structure;
        const int chain_form = st <= GQE_CHAIN3 ? 1 : 0;
        // composing pays off once a few tiles share the product (d^3 FMAs per product); with the
        // cache on, the product is reused by later calls as well
        const bool compose = c->compose == GQE_COMPOSE_ALWAYS ||
                             (c->compose == GQE_COMPOSE_AUTO && (rows >= 8 * kTcTileRows || c->wcache_on));
        const bool ds = s->pre != nullptr;
        auto one = [&](const float* m, int cf, int perm = 0) { return packed_of(m, nullptr, nullptr, 0, cf, perm); };
        if (!compose || st == GQE_CHAIN1 || (st >= GQE_INTER2 && !ds && st != GQE_INTER_CHAIN3)) {
          // nothing to merge: one contraction per operator, as written in the reference; the
          // accumulator of a chain's last hop is scored in the TMEM fragment layout (perm)
          for (int k = 0; k < n_rels_of(st); ++k) {
            const bool last = k == n_rels_of(st) - 1;
            s->rel[k] = one(s->rel[k], chain_form, last && st <= GQE_CHAIN3);
          }
          if (s->pre) s->pre = one(s->pre, 0);
          if (s->post) s->post = one(s->post, 0);
        } else {
          s->composed = 1;
          const float *r0 = s->rel[0], *r1 = s->rel[1], *r2 = s->rel[2], *pre = s->pre, *post = s->post;
          s->rel[0] = s->rel[1] = s->rel[2] = s->pre = s->post = nullptr;
          if (st <= GQE_CHAIN3) {                       // act.mm(M1).mm(M2)[.mm(M3)]  (decoders.py:143-145)
            s->rel[0] = packed_of(r0, r1, st == GQE_CHAIN3 ? r2 : nullptr, 0, 1, 1);
          } else if (st == GQE_INTER2 || st == GQE_INTER3) {   // relu(pre.mm(R_b.mm(e)))  (decoders.py:289-292)
            s->rel[0] = packed_of(pre, r0, nullptr, 0, 0, 0);
            s->rel[1] = packed_of(pre, r1, nullptr, 0, 0, 0);
            if (st == GQE_INTER3) s->rel[2] = packed_of(pre, r2, nullptr, 0, 0, 0);
            s->post = one(post, 0);
          } else if (st == GQE_INTER_CHAIN3) {          // branch 1: R2a.mm(R2b.mm(e))  (model.py:84-86)
            if (ds) {
              s->rel[0] = packed_of(pre, r0, nullptr, 0, 0, 0);
              s->rel[1] = packed_of(pre, r2, r1, 1, 0, 0);
              s->post = one(post, 0);
            } else {
              s->rel[0] = one(r0, 0);
              s->rel[1] = packed_of(r2, r1, nullptr, 0, 0, 0);
            }
          } else {                                      // 3-chain_inter (DeepSets): R1.mm(post.mm(.))  (model.py:106-107)
            s->rel[0] = packed_of(pre, r0, nullptr, 0, 0, 0);
            s->rel[1] = packed_of(pre, r1, nullptr, 0, 0, 0);
            s->post = packed_of(r2, post, nullptr, 0, 0, 0);
          }
        }
      }
      if (use_tc)
        build_program(s->prog, g.plan.structure, c->inter == GQE_INTER_DEEPSETS_MEAN || c->inter == GQE_INTER_DEEPSETS_MIN,
                      s->composed != 0);
      if (first_structure < 0) first_structure = g.plan.structure;
      else if (first_structure != g.plan.structure) uniform = false;
      ++n;
    }
    if (n == 0 || tiles == 0) continue;
    if (tiles > 0x7fffffffLL) return fail(c, GQE_ERR_UNSUPPORTED, "batch too large for one launch (%lld tiles)", (long long)tiles);
    // heaviest structures first: CTAs are dispatched in tile order, so the cheap
    // tiles (1-chain) fill the tail of the last wave instead of the 7-contraction ones
    {
      auto cost = [](const SegDev& sd) {
        switch (sd.structure) {
          case GQE_CHAIN1: return 1; case GQE_CHAIN2: return 2; case GQE_CHAIN3: return 3;
          case GQE_INTER2: return 5; case GQE_INTER3: return 7; default: return 6;
        }
      };
      std::stable_sort(lp.seg, lp.seg + n, [&](const SegDev& a, const SegDev& b) { return cost(a) > cost(b); });
      // CTA pairs (tc::producer<PAIR>): every segment is padded to an even number of tiles, so that
      // tiles (2j, 2j+1) always belong to one formula; a padding tile has no rows
      lp.pair = use_tc ? tc_use_pair(c->d, tiles) : 0;
      int64_t t = 0;
      for (int k = 0; k < n; ++k) {
        const int64_t nq = lp.seg[k].q_end - lp.seg[k].q_begin;
        const int64_t rows = lp.seg[k].structure <= GQE_CHAIN3 ? (target_offsets ? n_pairs : nq * T) : nq;
        lp.seg[k].tile_begin = t;
        const int64_t nt = (rows + tile_rows - 1) / tile_rows;
        t += lp.pair ? ((nt + 1) & ~(int64_t)1) : nt;
      }
      tiles = t;
    }
    lp.n_segs = n;
    lp.n_tiles = tiles;
    lp.final_launch = i > last_nonempty ? 1 : 0;
    lp.err_host = lp.final_launch ? err_host : nullptr;
    if (out_loss) {
      c->loss_dirty = true;   // until the final launch of the call has been issued
      int rc = ensure_partials(c, tiles);
      if (rc != GQE_OK) return rc;
      lp.partials = c->partials;
    }
    // single formula -> that structure's own kernel; otherwise the grouped kernel
    const int structure = (n == 1 && uniform) ? first_structure : -1;
    if (use_tc) {
      for (int l = 0; l < 2; ++l)
        if (n_cw[l] > 0) {
          GQE_CUDA(c, launch_compose(c->d, cw[l], n_cw[l], c->stream));
          c->launches += 1;
        }
      if (n_pack > 0) {
        GQE_CUDA(c, launch_pack(c->d, pp, n_pack, c->stream));
        c->launches += 1;
        c->weight_preps += n_pack;
      }
      // Node-type-sharded tables (some operand rows live in a peer GPU's HBM), grouped d = 256 kernel: the
      // instantiation whose helper warps fetch those rows into a local staging area a tile or two ahead
      // (GQE_STAGE=0: gather them in place as round 1 did; GQE_FORCE_STAGE=1: treat every table as remote --
      // a single-GPU test of the mechanism)
      lp.stage_on = 0;
      if (kStageRemote && c->d == 256 && structure < 0 && !lp.pair) {
        bool any = kForceStage >= 2;      // 2: the STAGE instantiation with nothing to stage (its fixed cost)
        for (int k = 0; k < n; ++k) {
          if (kForceStage == 1) lp.seg[k].remote_mask = 0xFu;
          any = any || lp.seg[k].remote_mask != 0;
        }
        if (any) {
          int sms = 148;
          cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device);
          const size_t need = (size_t)sms * tc_stage_bytes_per_cta(c->d);
          if (c->stage_buf_cap < need) {
            GQE_CUDA(c, cudaStreamSynchronize(c->stream));
            cudaFree(c->stage_buf);
            c->stage_buf = nullptr;
            c->stage_buf_cap = 0;
            GQE_CUDA(c, cudaMalloc(&c->stage_buf, need));
            c->stage_buf_cap = need;
          }
          lp.stage = c->stage_buf;
          lp.stage_mask = kStageMask;
          lp.stage_on = kForceStage == 3 ? 2 : 1;   // 3: the STAGE instantiation with its helper warp idle (diagnostics)
        }
      }
      // indices fetched by gqe_fetch_indices right in front of this launch: the kernel waits for it itself
      // (with weight preparation in between, the ordinary stream order has already done that)
      lp.wait_prev = (c->idx_by_kernel && n_pack == 0 && n_cw[0] == 0 && n_cw[1] == 0) ? 1 : 0;
      c->idx_by_kernel = false;
      GQE_CUDA(c, launch_fused_tc(c->d, structure, lp, tiles, c->stream));
      c->launches += 1;
      if (lp.err_host) c->err_posted = true;
    } else if (use_vec) {
      GQE_CUDA(c, launch_fused_vec(c->d, lp, c->stream));
      c->launches += 1;
    } else {
      GQE_CUDA(c, launch_fused_simt(c->d, structure, lp, tiles, c->stream));
      c->launches += 1;
    }
    if (out_loss && lp.final_launch) c->loss_dirty = false;
    if (pairs_mode) {
      PairParams qp;
      std::memset(&qp, 0, sizeof qp);
      int64_t n_p = 0;
      for (int k = 0; k < n; ++k) {
        if (lp.seg[k].structure <= GQE_CHAIN3) continue;  // chain tiles scored their own pairs
        PairSeg& ps = qp.seg[qp.n_segs++];
        ps.tgt_table = lp.seg[k].tgt_table;
        ps.q_begin = lp.seg[k].q_begin;
        ps.q_end = lp.seg[k].q_end;
        ps.mode = lp.mode[lp.seg[k].tgt_mode];
        n_p += target_offsets ? n_pairs : (ps.q_end - ps.q_begin) * T;
      }
      if (qp.n_segs > 0) {
        qp.T = T;
        qp.q = c->qbuf;
        qp.target_rows = target_rows;
        qp.target_offsets = target_offsets;
        qp.n_pairs = n_pairs;
        qp.out_scores = out_scores;
        qp.index_kind = index_kind;
        qp.err = c->d_err;
        GQE_CUDA(c, launch_score_pairs(c->d, qp, n_p, c->stream));
        c->launches += 1;
      }
    }
  }
  return GQE_OK;
}

static int regular_T(gqe_ctx* c, int64_t nq, int64_t n_pairs, const int64_t* offsets, int32_t* T) {
  *T = 0;
  if (offsets) return GQE_OK;
  if (nq == 0) return n_pairs == 0 ? GQE_OK : fail(c, GQE_ERR_INVALID, "pairs without queries");
  if (n_pairs % nq != 0 || n_pairs / nq <= 0 || n_pairs / nq > 0x7fffffff)
    return fail(c, GQE_ERR_INVALID, "regular layout needs n_pairs (%lld) to be a positive multiple of n_queries (%lld)",
                (long long)n_pairs, (long long)nq);
  *T = (int32_t)(n_pairs / nq);
  return GQE_OK;
}

static int score_device_impl(gqe_ctx* c, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_rows,
                             int64_t n_pairs, const int32_t* target_rows, const int64_t* target_offsets,
                             float* out_scores, int index_kind) {
  if (!c) return GQE_ERR_INVALID;
  if (!plan || (!out_scores && n_pairs > 0)) return fail(c, GQE_ERR_INVALID, "gqe_score_device: null argument");
  int32_t T;
  int rc = regular_T(c, n_queries, n_pairs, target_offsets, &T);
  if (rc != GQE_OK) return rc;
  gqe_segment seg{*plan, 0, n_queries};
  return run_fused(c, &seg, 1, n_queries, anchor_rows, n_pairs, target_rows, target_offsets, T, out_scores, 0.f, nullptr,
                   index_kind);
}
extern "C" int gqe_score_device(gqe_ctx* c, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_rows,
                                int64_t n_pairs, const int32_t* target_rows, const int64_t* target_offsets,
                                float* out_scores) {
  return score_device_impl(c, plan, n_queries, anchor_rows, n_pairs, target_rows, target_offsets, out_scores, 0);
}
extern "C" int gqe_score_nodes_device(gqe_ctx* c, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_nodes,
                                      int64_t n_pairs, const int32_t* target_nodes, const int64_t* target_offsets,
                                      float* out_scores) {
  return score_device_impl(c, plan, n_queries, anchor_nodes, n_pairs, target_nodes, target_offsets, out_scores, 1);
}

static int margin_loss_device_impl(gqe_ctx* c, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_rows,
                                   const int32_t* pair_rows, float margin, float* out_loss, float* out_scores,
                                   int index_kind) {
  if (!c) return GQE_ERR_INVALID;
  if (!plan || !out_loss) return fail(c, GQE_ERR_INVALID, "gqe_margin_loss_device: null argument");
  gqe_segment seg{*plan, 0, n_queries};
  return run_fused(c, &seg, 1, n_queries, anchor_rows, 2 * n_queries, pair_rows, nullptr, 2, out_scores, margin, out_loss,
                   index_kind);
}
extern "C" int gqe_margin_loss_device(gqe_ctx* c, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_rows,
                                      const int32_t* pair_rows, float margin, float* out_loss, float* out_scores) {
  return margin_loss_device_impl(c, plan, n_queries, anchor_rows, pair_rows, margin, out_loss, out_scores, 0);
}
extern "C" int gqe_margin_loss_nodes_device(gqe_ctx* c, const gqe_plan* plan, int64_t n_queries,
                                            const int32_t* anchor_nodes, const int32_t* pair_nodes, float margin,
                                            float* out_loss, float* out_scores) {
  return margin_loss_device_impl(c, plan, n_queries, anchor_nodes, pair_nodes, margin, out_loss, out_scores, 1);
}

extern "C" int gqe_score_grouped_device(gqe_ctx* c, const gqe_segment* segments, int32_t n_segments,
                                        int64_t n_queries_total, const int32_t* anchor_rows, const int32_t* target_rows,
                                        int32_t targets_per_query, float* out_scores, float margin, float* out_loss) {
  if (!c) return GQE_ERR_INVALID;
  return run_fused(c, segments, n_segments, n_queries_total, anchor_rows, n_queries_total * targets_per_query,
                   target_rows, nullptr, targets_per_query, out_scores, margin, out_loss, 0);
}
extern "C" int gqe_score_grouped_nodes_device(gqe_ctx* c, const gqe_segment* segments, int32_t n_segments,
                                              int64_t n_queries_total, const int32_t* anchor_nodes,
                                              const int32_t* target_nodes, int32_t targets_per_query, float* out_scores,
                                              float margin, float* out_loss) {
  if (!c) return GQE_ERR_INVALID;
  return run_fused(c, segments, n_segments, n_queries_total, anchor_nodes, n_queries_total * targets_per_query,
                   target_nodes, nullptr, targets_per_query, out_scores, margin, out_loss, 1);
}

// ---- device-resident query store ---------------------------------------------------
enum { ST_ANCHOR = 0, ST_TARGET = 1, ST_OFFSETS = 2, ST_SCORES = 3, ST_LOSS = 4 };
extern "C" int gqe_margin_loss_store_device(gqe_ctx* c, const gqe_segment* segs, int32_t n_segs,
                                            const gqe_store_slice* slices, uint64_t seed, float margin, float* out_loss,
                                            float* out_scores, int32_t* out_pairs) {
  if (!c) return GQE_ERR_INVALID;
  if (!segs || !slices || n_segs <= 0) return fail(c, GQE_ERR_INVALID, "no segments");
  if (!out_loss) return fail(c, GQE_ERR_INVALID, "gqe_margin_loss_store_device: out_loss is null");
  if (c->node_maps.empty()) return fail(c, GQE_ERR_UNBOUND, "node maps are not bound (gqe_bind_node_maps)");
  GQE_CUDA(c, cudaSetDevice(c->device));
  int64_t nq = 0;
  for (int32_t i = 0; i < n_segs; ++i) {
    if (segs[i].query_begin < 0 || segs[i].query_end < segs[i].query_begin)
      return fail(c, GQE_ERR_INVALID, "segment %d: bad query range", i);
    nq = std::max<int64_t>(nq, segs[i].query_end);
  }
  if (nq == 0) {
    GQE_CUDA(c, cudaMemsetAsync(out_loss, 0xff, sizeof(float), c->stream));   // mean of nothing: NaN, like torch
    return GQE_OK;
  }
  int rc;
  if ((rc = gqe_stage_reserve(c, ST_ANCHOR, sizeof(int32_t) * (size_t)GQE_MAX_ANCHORS * nq)) != GQE_OK) return rc;
  if ((rc = gqe_stage_reserve(c, ST_TARGET, sizeof(int32_t) * (size_t)2 * nq)) != GQE_OK) return rc;
  for (int32_t i0 = 0; i0 < n_segs; i0 += kMaxSegs) {
    StoreBatchParams sp;
    sp.count = 0;
    sp.nq_total = nq;
    sp.seed = seed;
    sp.anchors_out = (int32_t*)c->stage[ST_ANCHOR];
    sp.pairs_out = (int32_t*)c->stage[ST_TARGET];
    sp.err = c->d_err;
    int64_t max_n = 0;
    for (int32_t i = i0; i < std::min<int32_t>(n_segs, i0 + kMaxSegs); ++i) {
      const int64_t n = segs[i].query_end - segs[i].query_begin;
      if (n == 0) continue;
      const gqe_store_slice& g = slices[i];
      const int na = n_anchors_of(segs[i].plan.structure);
      if (na <= 0) return fail(c, GQE_ERR_INVALID, "segment %d: unknown query structure", i);
      if (!g.anchors || !g.targets || !g.negs || g.start < 0 || g.start + n > g.block_queries)
        return fail(c, GQE_ERR_INVALID, "segment %d: store slice [%lld,%lld) outside its block of %lld queries", i,
                    (long long)g.start, (long long)(g.start + n), (long long)g.block_queries);
      StoreSliceDev& d = sp.s[sp.count++];
      d.anchors = g.anchors; d.targets = g.targets; d.neg_ptr = g.neg_ptr; d.negs = g.negs;
      d.block_q = g.block_queries; d.start = g.start; d.n = n; d.out_q0 = segs[i].query_begin; d.pool_n = g.pool_size;
      d.n_anchor = na;
      d.tgt_mode = segs[i].plan.target_mode;
      max_n = std::max(max_n, n);
    }
    if (sp.count > 0) {
      GQE_CUDA(c, launch_store_batch(sp, max_n, c->stream));
      c->launches += 1;
    }
  }
  if (out_pairs)
    GQE_CUDA(c, cudaMemcpyAsync(out_pairs, c->stage[ST_TARGET], sizeof(int32_t) * (size_t)2 * nq, cudaMemcpyDeviceToDevice, c->stream));
  else
    c->idx_by_kernel = true;   // the scoring kernel, launched programmatically dependent, waits for the batch kernel itself
  rc = run_fused(c, segs, n_segs, nq, (const int32_t*)c->stage[ST_ANCHOR], 2 * nq, (const int32_t*)c->stage[ST_TARGET], nullptr,
                 2, out_scores, margin, out_loss, 1);
  c->idx_by_kernel = false;
  return rc;
}

// ---- host-buffer variants ------------------------------------------------------
int gqe_stage_reserve(gqe_ctx* c, int slot, size_t bytes) {
  if (bytes <= c->stage_cap[slot]) return GQE_OK;
  GQE_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaFree(c->stage[slot]);
  c->stage[slot] = nullptr;
  c->stage_cap[slot] = 0;
  const size_t cap = std::max(bytes + bytes / 4, (size_t)1 << 16);
  GQE_CUDA(c, cudaMalloc(&c->stage[slot], cap));
  c->stage_cap[slot] = cap;
  return GQE_OK;
}
static const size_t kInPlaceIndexBytes = [] {   // GQE_INPLACE_BYTES in the environment overrides (0 = always copy)
  const char* e = getenv("GQE_INPLACE_BYTES");
  return e ? (size_t)atoll(e) : (size_t)128 << 10;
}();

static const bool kFetchKernel = [] {
  const char* e = getenv("GQE_FETCH_KERNEL");
  return e ? atoi(e) != 0 : true;
}();

static int max_anchors(const gqe_segment* segs, int n) {
  int m = 0;
  for (int i = 0; i < n; ++i) m = std::max(m, n_anchors_of(segs[i].plan.structure));
  return m;
}

static int run_fused_host(gqe_ctx* c, const gqe_segment* segs, int32_t n_segs, int64_t nq, const int32_t* anchor_rows,
                          int64_t n_pairs, const int32_t* target_rows, const int64_t* target_offsets, int32_t T,
                          float* out_scores, float margin, float* out_loss, int index_kind) {
  if (!c) return GQE_ERR_INVALID;
  if (!segs || n_segs <= 0) return fail(c, GQE_ERR_INVALID, "no segments");
  GQE_CUDA(c, cudaSetDevice(c->device));
  const int na = max_anchors(segs, n_segs);
  if (na <= 0) return fail(c, GQE_ERR_INVALID, "unknown query structure");
  int rc;
  if ((rc = gqe_stage_reserve(c, ST_ANCHOR, sizeof(int32_t) * (size_t)na * nq)) != GQE_OK) return rc;
  if ((rc = gqe_stage_reserve(c, ST_TARGET, sizeof(int32_t) * (size_t)n_pairs)) != GQE_OK) return rc;
  if (target_offsets && (rc = gqe_stage_reserve(c, ST_OFFSETS, sizeof(int64_t) * (size_t)(nq + 1))) != GQE_OK) return rc;
  if (out_scores && (rc = gqe_stage_reserve(c, ST_SCORES, sizeof(float) * (size_t)n_pairs)) != GQE_OK) return rc;
  if (out_loss && (rc = gqe_stage_reserve(c, ST_LOSS, sizeof(float))) != GQE_OK) return rc;
  // The loss goes straight to a mapped pinned word (one posted write by the last CTA) instead of
  // a device word + a D2H copy: one stream operation less on the critical path of the call.
  float* loss_dst = out_loss ? (float*)c->stage[ST_LOSS] : nullptr;
  if (out_loss && nq > 0) {
    if (!c->h_loss) {
      GQE_CUDA(c, cudaHostAlloc((void**)&c->h_loss, sizeof(float), cudaHostAllocMapped));
      GQE_CUDA(c, cudaHostGetDevicePointer((void**)&c->h_loss_dev, c->h_loss, 0));
    }
    loss_dst = c->h_loss_dev;
  }
  // Small calls on pinned (mapped) host buffers: the kernels read the indices in place over PCIe.  A
  // copy operation costs ~8 us of fixed latency, two or three of them are most of a 512-query call
  // (configs[0]); the in-place reads add one ~2 us round trip to the first tile instead.  Large calls keep
  // the copies (in place they stall every tile's first gather: profiles/r02_experiments.md).
  const int32_t *zc_anchor = nullptr, *zc_target = nullptr;
  if (nq > 0 && anchor_rows && target_rows && !target_offsets &&
      sizeof(int32_t) * ((size_t)na * nq + (size_t)n_pairs) <= kInPlaceIndexBytes) {
    cudaPointerAttributes pa, pt;
    if (cudaPointerGetAttributes(&pa, anchor_rows) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer &&
        cudaPointerGetAttributes(&pt, target_rows) == cudaSuccess && pt.type == cudaMemoryTypeHost && pt.devicePointer) {
      zc_anchor = (const int32_t*)pa.devicePointer;
      zc_target = (const int32_t*)pt.devicePointer;
    }
    (void)cudaGetLastError();    // (a pageable pointer is an error to the query on some drivers)
  }
  // Larger calls on pinned (mapped) buffers: ONE kernel copies every range over PCIe (gqe_fetch_indices)
  // instead of one copy-engine operation per range (GQE_FETCH_KERNEL=0: always the copy engine).
  bool fetched = false;
  c->idx_by_kernel = false;
  if (nq > 0 && !zc_anchor && kFetchKernel && anchor_rows && target_rows) {
    cudaPointerAttributes pa, pt, po;
    const bool ok = cudaPointerGetAttributes(&pa, anchor_rows) == cudaSuccess && pa.type == cudaMemoryTypeHost && pa.devicePointer &&
                    cudaPointerGetAttributes(&pt, target_rows) == cudaSuccess && pt.type == cudaMemoryTypeHost && pt.devicePointer &&
                    (!target_offsets || (cudaPointerGetAttributes(&po, target_offsets) == cudaSuccess &&
                                         po.type == cudaMemoryTypeHost && po.devicePointer));
    (void)cudaGetLastError();
    if (ok) {
      FetchParams fp;
      fp.count = 0;
      auto add = [&](const void* src_dev, void* dst, int64_t words) {
        if (words <= 0) return;
        fp.src[fp.count] = (const int32_t*)src_dev;
        fp.dst[fp.count] = (int32_t*)dst;
        fp.n[fp.count] = words;
        ++fp.count;
      };
      for (int k = 0; k < na; ++k) {     // per slot, the query range of the segments that read it (see below)
        int64_t lo = nq, hi = 0;
        for (int32_t i = 0; i < n_segs; ++i)
          if (n_anchors_of(segs[i].plan.structure) > k && segs[i].query_end > segs[i].query_begin) {
            lo = std::min(lo, std::max<int64_t>(segs[i].query_begin, 0));
            hi = std::max(hi, std::min<int64_t>(segs[i].query_end, nq));
          }
        add((const int32_t*)pa.devicePointer + (size_t)k * nq + lo, (int32_t*)c->stage[ST_ANCHOR] + (size_t)k * nq + lo, hi - lo);
      }
      add(pt.devicePointer, c->stage[ST_TARGET], n_pairs);
      if (target_offsets) add(po.devicePointer, c->stage[ST_OFFSETS], 2 * (nq + 1));
      GQE_CUDA(c, launch_fetch_indices(fp, c->stream));
      c->launches += 1;
      c->idx_by_kernel = true;
      fetched = true;
    }
  }
  if (nq > 0 && !zc_anchor && !fetched) {
    if (!anchor_rows || !target_rows) return fail(c, GQE_ERR_INVALID, "index arrays are null");
    // in stream order, in front of the kernels: with the packed weights cached there is nothing
    // to overlap the copies with, and a second stream would only add event traffic to the call
    // anchor slot k is only read for the queries of segments with more than k anchors: copy, per
    // slot, the query range those segments span (chains use one slot of the three-slot layout)
    for (int k = 0; k < na; ++k) {
      int64_t lo = nq, hi = 0;
      for (int32_t i = 0; i < n_segs; ++i)
        if (n_anchors_of(segs[i].plan.structure) > k && segs[i].query_end > segs[i].query_begin) {
          lo = std::min(lo, std::max<int64_t>(segs[i].query_begin, 0));
          hi = std::max(hi, std::min<int64_t>(segs[i].query_end, nq));
        }
      if (hi > lo)
        GQE_CUDA(c, cudaMemcpyAsync((int32_t*)c->stage[ST_ANCHOR] + (size_t)k * nq + lo, anchor_rows + (size_t)k * nq + lo,
                                    sizeof(int32_t) * (size_t)(hi - lo), cudaMemcpyHostToDevice, c->stream));
    }
    GQE_CUDA(c, cudaMemcpyAsync(c->stage[ST_TARGET], target_rows, sizeof(int32_t) * (size_t)n_pairs,
                                cudaMemcpyHostToDevice, c->stream));
    if (target_offsets)
      GQE_CUDA(c, cudaMemcpyAsync(c->stage[ST_OFFSETS], target_offsets, sizeof(int64_t) * (size_t)(nq + 1),
                                  cudaMemcpyHostToDevice, c->stream));
  }
  c->h_err[0] = 0ull;
  c->err_posted = false;
  rc = run_fused(c, segs, n_segs, nq, zc_anchor ? zc_anchor : (const int32_t*)c->stage[ST_ANCHOR], n_pairs,
                 zc_target ? zc_target : (const int32_t*)c->stage[ST_TARGET],
                 target_offsets ? (const int64_t*)c->stage[ST_OFFSETS] : nullptr, T,
                 out_scores ? (float*)c->stage[ST_SCORES] : nullptr, margin, loss_dst, index_kind, c->h_err_dev);
  c->idx_by_kernel = false;
  if (rc != GQE_OK) return rc;
  if (out_scores && n_pairs > 0)
    GQE_CUDA(c, cudaMemcpyAsync(out_scores, c->stage[ST_SCORES], sizeof(float) * (size_t)n_pairs, cudaMemcpyDeviceToHost,
                                c->stream));
  const bool mapped_loss = out_loss && loss_dst == c->h_loss_dev;
  if (out_loss && !mapped_loss)
    GQE_CUDA(c, cudaMemcpyAsync(out_loss, c->stage[ST_LOSS], sizeof(float), cudaMemcpyDeviceToHost, c->stream));
  // the index-error word: posted by the last CTA of the tensor-core kernel, fetched otherwise
  if (!c->err_posted && nq > 0)
    GQE_CUDA(c, cudaMemcpyAsync(c->h_err, c->d_err, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  GQE_CUDA(c, cudaStreamSynchronize(c->stream));
  if (mapped_loss) *out_loss = *(volatile float*)c->h_loss;
  const unsigned long long w0 = *(volatile unsigned long long*)&c->h_err[0];
  if (w0 != 0ull) {
    const unsigned long long w1 = *(volatile unsigned long long*)&c->h_err[1];
    cudaMemsetAsync(c->d_err, 0, 2 * sizeof(unsigned long long), c->stream);
    return index_error_from(c, w0, w1);
  }
  return GQE_OK;
}

#define GQE_HOST_ENTRY_POINTS(SUFFIX, KIND)                                                                           \
  extern "C" int gqe_score##SUFFIX##_host(gqe_ctx* c, const gqe_plan* plan, int64_t n_queries, const int32_t* anchors, \
                                          int64_t n_pairs, const int32_t* targets, const int64_t* target_offsets,     \
                                          float* out_scores) {                                                        \
    if (!c) return GQE_ERR_INVALID;                                                                                   \
    if (!plan || (!out_scores && n_pairs > 0)) return fail(c, GQE_ERR_INVALID, "gqe_score_host: null argument");      \
    int32_t T;                                                                                                        \
    int rc = regular_T(c, n_queries, n_pairs, target_offsets, &T);                                                    \
    if (rc != GQE_OK) return rc;                                                                                      \
    gqe_segment seg{*plan, 0, n_queries};                                                                             \
    return run_fused_host(c, &seg, 1, n_queries, anchors, n_pairs, targets, target_offsets, T, out_scores, 0.f,      \
                          nullptr, KIND);                                                                             \
  }                                                                                                                   \
  extern "C" int gqe_margin_loss##SUFFIX##_host(gqe_ctx* c, const gqe_plan* plan, int64_t n_queries,                 \
                                                const int32_t* anchors, const int32_t* pairs, float margin,          \
                                                float* out_loss, float* out_scores) {                                \
    if (!c) return GQE_ERR_INVALID;                                                                                   \
    if (!plan || !out_loss) return fail(c, GQE_ERR_INVALID, "gqe_margin_loss_host: null argument");                   \
    gqe_segment seg{*plan, 0, n_queries};                                                                             \
    return run_fused_host(c, &seg, 1, n_queries, anchors, 2 * n_queries, pairs, nullptr, 2, out_scores, margin,      \
                          out_loss, KIND);                                                                            \
  }                                                                                                                   \
  extern "C" int gqe_score_grouped##SUFFIX##_host(gqe_ctx* c, const gqe_segment* segments, int32_t n_segments,       \
                                                  int64_t n_queries_total, const int32_t* anchors,                   \
                                                  const int32_t* targets, int32_t targets_per_query,                 \
                                                  float* out_scores, float margin, float* out_loss) {                \
    if (!c) return GQE_ERR_INVALID;                                                                                   \
    if (!segments || n_segments <= 0) return fail(c, GQE_ERR_INVALID, "no segments");                                 \
    /* grouped anchors are always laid out with GQE_MAX_ANCHORS slots */                                              \
    return run_fused_host(c, segments, n_segments, n_queries_total, anchors, n_queries_total * targets_per_query,     \
                          targets, nullptr, targets_per_query, out_scores, margin, out_loss, KIND);                   \
  }
GQE_HOST_ENTRY_POINTS(, 0)
GQE_HOST_ENTRY_POINTS(_nodes, 1)
#undef GQE_HOST_ENTRY_POINTS

// ---- operator-level entry points -------------------------------------------------
static int launch_op(gqe_ctx* c, int d, const OpParams& op) {
  if (op.n == 0) return GQE_OK;
  if (op.n < 0) return fail(c, GQE_ERR_INVALID, "negative column count");
  GQE_CUDA(c, cudaSetDevice(c->device));
  if (!dim_supported(d)) return fail(c, GQE_ERR_UNSUPPORTED, "dimension %d not supported (32/64/128/256)", d);
  drop_stale_error("operator entry");
  cudaError_t e = launch_op_simt(d, op, c->stream);
  if (e != cudaSuccess) return fail(c, GQE_ERR_CUDA, "operator launch failed: %s", cudaGetErrorString(e));
  c->launches += 1;
  return GQE_OK;
}

extern "C" int gqe_encode_device(gqe_ctx* c, int32_t mode, int64_t n, const int32_t* rows, float* out) {
  if (!c) return GQE_ERR_INVALID;
  if (c->tables.empty()) return fail(c, GQE_ERR_UNBOUND, "embedding tables are not bound");
  if (mode < 0 || mode >= (int)c->tables.size()) return fail(c, GQE_ERR_INVALID, "mode %d out of range", mode);
  if (!c->tables[mode]) return fail(c, GQE_ERR_UNBOUND, "mode %d has no table on this rank", mode);
  if (n > 0 && (!rows || !out)) return fail(c, GQE_ERR_INVALID, "gqe_encode_device: null argument");
  OpParams op;
  std::memset(&op, 0, sizeof op);
  op.op = OP_ENCODE;
  op.n = n;
  op.table = c->tables[mode];
  op.table_rows = c->table_rows[mode];
  op.err = c->d_err;
  op.rows = rows;
  op.out = out;
  return launch_op(c, c->d, op);
}

extern "C" int gqe_project_device(gqe_ctx* c, int32_t rel, int64_t n, const float* in, float* out) {
  if (!c) return GQE_ERR_INVALID;
  if (c->rels.empty()) return fail(c, GQE_ERR_UNBOUND, "relation parameters are not bound");
  if (rel < 0 || rel >= (int)c->rels.size()) return fail(c, GQE_ERR_INVALID, "relation id %d out of range", rel);
  if (n > 0 && (!in || !out)) return fail(c, GQE_ERR_INVALID, "gqe_project_device: null argument");
  OpParams op;
  std::memset(&op, 0, sizeof op);
  op.op = OP_PROJECT;
  op.decoder = c->decoder;
  op.n = n;
  op.rel[0] = c->rels[rel];
  op.in0 = in;
  op.out = out;
  return launch_op(c, c->rel_d, op);
}

extern "C" int gqe_path_score_device(gqe_ctx* c, int32_t n_rels, const int32_t* rels, int64_t n, float* embeds1,
                                     const float* embeds2, int32_t mutate_embeds1, float* out) {
  if (!c) return GQE_ERR_INVALID;
  if (c->rels.empty()) return fail(c, GQE_ERR_UNBOUND, "relation parameters are not bound");
  if (n_rels < 1 || n_rels > GQE_MAX_RELS || !rels) return fail(c, GQE_ERR_INVALID, "metapath must have 1..3 relations");
  if (n > 0 && (!embeds1 || !embeds2 || !out)) return fail(c, GQE_ERR_INVALID, "gqe_path_score_device: null argument");
  OpParams op;
  std::memset(&op, 0, sizeof op);
  op.op = OP_PATH_SCORE;
  op.decoder = c->decoder;
  op.n_rels = n_rels;
  op.mutate = mutate_embeds1;
  op.n = n;
  for (int k = 0; k < n_rels; ++k) {
    if (rels[k] < 0 || rels[k] >= (int)c->rels.size()) return fail(c, GQE_ERR_INVALID, "relation id %d out of range", rels[k]);
    op.rel[k] = c->rels[rels[k]];
  }
  op.io0 = embeds1;
  op.in1 = embeds2;
  op.out = out;
  return launch_op(c, c->rel_d, op);
}

extern "C" int gqe_intersect_device(gqe_ctx* c, int32_t mode, int64_t n, const float* e1, const float* e2,
                                    const float* e3, float* out) {
  if (!c) return GQE_ERR_INVALID;
  if (c->inter < 0) return fail(c, GQE_ERR_UNBOUND, "intersection operator is not bound");
  if (n > 0 && (!e1 || !e2 || !out)) return fail(c, GQE_ERR_INVALID, "gqe_intersect_device: null argument");
  OpParams op;
  std::memset(&op, 0, sizeof op);
  op.op = OP_INTERSECT;
  op.inter = c->inter;
  op.n = n;
  if (c->inter <= GQE_INTER_DEEPSETS_MIN) {
    if (mode < 0 || mode >= (int)c->pre.size()) return fail(c, GQE_ERR_INVALID, "intersection mode %d out of range", mode);
    op.pre = c->pre[mode];
    op.post = c->post[mode];
  }
  op.in0 = e1;
  op.in1 = e2;
  op.in2 = e3;
  op.out = out;
  return launch_op(c, c->inter_d, op);
}

extern "C" int gqe_cosine_device(gqe_ctx* c, int32_t d, int64_t n, const float* x, const float* y, float* out) {
  if (!c) return GQE_ERR_INVALID;
  if (n > 0 && (!x || !y || !out)) return fail(c, GQE_ERR_INVALID, "gqe_cosine_device: null argument");
  OpParams op;
  std::memset(&op, 0, sizeof op);
  op.op = OP_COSINE;
  op.n = n;
  op.in0 = x;
  op.in1 = y;
  op.out = out;
  return launch_op(c, d, op);
}

// ---- differentiable operator surface: forward pieces + vector-Jacobian products -----------
static int bwd_common(gqe_ctx* c, int d, int64_t n, const char* who) {
  if (!c) return GQE_ERR_INVALID;
  if (n < 0) return fail(c, GQE_ERR_INVALID, "%s: negative column count", who);
  if (!dim_supported(d)) return fail(c, GQE_ERR_UNSUPPORTED, "%s: dimension %d not supported (32/64/128/256)", who, d);
  GQE_CUDA(c, cudaSetDevice(c->device));
  drop_stale_error(who);
  return GQE_OK;
}
#define GQE_BWD_LAUNCH(c, call)                                                                 \
  do {                                                                                          \
    cudaError_t e_ = (call);                                                                    \
    if (e_ != cudaSuccess) return fail((c), GQE_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
    (c)->launches += 1;                                                                         \
  } while (0)

extern "C" int gqe_matmul_device(gqe_ctx* c, const float* w, int32_t transpose, int32_t d, int64_t n, const float* in,
                                 float* out) {
  int rc = bwd_common(c, d, n, "gqe_matmul_device");
  if (rc != GQE_OK || n == 0) return rc;
  if (!w || !in || !out) return fail(c, GQE_ERR_INVALID, "gqe_matmul_device: null argument");
  OpParams op;
  std::memset(&op, 0, sizeof op);
  op.op = OP_MATMUL;
  op.mutate = transpose;
  op.n = n;
  op.rel[0] = w;
  op.in0 = in;
  op.out = out;
  return launch_op(c, d, op);
}
extern "C" int gqe_matmul_wgrad_device(gqe_ctx* c, int32_t transpose, int32_t d, int64_t n, const float* gy,
                                       const float* x, float* gw) {
  int rc = bwd_common(c, d, n, "gqe_matmul_wgrad_device");
  if (rc != GQE_OK || n == 0) return rc;
  if (!gy || !x || !gw) return fail(c, GQE_ERR_INVALID, "gqe_matmul_wgrad_device: null argument");
  GQE_BWD_LAUNCH(c, launch_matmul_wgrad(d, n, gy, x, transpose ? 1 : 0, gw, c->stream));
  return GQE_OK;
}
extern "C" int gqe_rowsum_device(gqe_ctx* c, int32_t d, int64_t n, const float* gy, const float* x, float* gv) {
  int rc = bwd_common(c, d, n, "gqe_rowsum_device");
  if (rc != GQE_OK || n == 0) return rc;
  if (!gy || !gv) return fail(c, GQE_ERR_INVALID, "gqe_rowsum_device: null argument");
  GQE_BWD_LAUNCH(c, launch_rowsum(d, n, gy, x, gv, c->stream));
  return GQE_OK;
}
extern "C" int gqe_aggregate_device(gqe_ctx* c, int32_t d, int64_t n, const float* e1, const float* e2, const float* e3,
                                    int32_t relu, int32_t use_min, float* out) {
  int rc = bwd_common(c, d, n, "gqe_aggregate_device");
  if (rc != GQE_OK || n == 0) return rc;
  if (!e1 || !e2 || !out) return fail(c, GQE_ERR_INVALID, "gqe_aggregate_device: null argument");
  GQE_BWD_LAUNCH(c, launch_aggregate((int64_t)d * n, e1, e2, e3, relu, use_min, out, c->stream));
  return GQE_OK;
}
extern "C" int gqe_aggregate_bwd_device(gqe_ctx* c, int32_t d, int64_t n, const float* e1, const float* e2,
                                        const float* e3, int32_t relu, int32_t use_min, const float* gout, float* g1,
                                        float* g2, float* g3) {
  int rc = bwd_common(c, d, n, "gqe_aggregate_bwd_device");
  if (rc != GQE_OK || n == 0) return rc;
  if (!e1 || !e2 || !gout || !g1 || !g2 || (e3 && !g3)) return fail(c, GQE_ERR_INVALID, "gqe_aggregate_bwd_device: null argument");
  GQE_BWD_LAUNCH(c, launch_aggregate_bwd((int64_t)d * n, e1, e2, e3, relu, use_min, gout, g1, g2, g3, c->stream));
  return GQE_OK;
}
extern "C" int gqe_dot_device(gqe_ctx* c, int32_t d, int64_t n, const float* x, const float* y, float* out) {
  int rc = bwd_common(c, d, n, "gqe_dot_device");
  if (rc != GQE_OK || n == 0) return rc;
  if (!x || !y || !out) return fail(c, GQE_ERR_INVALID, "gqe_dot_device: null argument");
  GQE_BWD_LAUNCH(c, launch_dot(d, n, x, y, out, c->stream));
  return GQE_OK;
}
extern "C" int gqe_cosine_bwd_device(gqe_ctx* c, int32_t d, int64_t n, const float* x, const float* y, const float* gout,
                                     int32_t raw_dot, float* gx, float* gy) {
  int rc = bwd_common(c, d, n, "gqe_cosine_bwd_device");
  if (rc != GQE_OK || n == 0) return rc;
  if (!x || !y || !gout || (!gx && !gy)) return fail(c, GQE_ERR_INVALID, "gqe_cosine_bwd_device: null argument");
  GQE_BWD_LAUNCH(c, launch_cosine_bwd(d, n, x, y, gout, raw_dot, gx, gy, c->stream));
  return GQE_OK;
}
extern "C" int gqe_encode_bwd_device(gqe_ctx* c, int32_t mode, int64_t n, const int32_t* rows, const float* gout,
                                     float* gtable) {
  if (!c) return GQE_ERR_INVALID;
  if (c->tables.empty()) return fail(c, GQE_ERR_UNBOUND, "embedding tables are not bound");
  if (mode < 0 || mode >= (int)c->tables.size()) return fail(c, GQE_ERR_INVALID, "mode %d out of range", mode);
  if (!c->tables[mode]) return fail(c, GQE_ERR_UNBOUND, "mode %d has no table on this rank", mode);
  int rc = bwd_common(c, c->d, n, "gqe_encode_bwd_device");
  if (rc != GQE_OK || n == 0) return rc;
  if (!rows || !gout || !gtable) return fail(c, GQE_ERR_INVALID, "gqe_encode_bwd_device: null argument");
  GQE_BWD_LAUNCH(c, launch_encode_bwd(c->d, n, c->tables[mode], rows, gout, gtable, c->table_rows[mode], c->d_err, c->stream));
  return GQE_OK;
}

extern "C" int gqe_encode_bwd_rows_device(gqe_ctx* c, int32_t mode, int64_t n, const int32_t* rows, const float* gout,
                                          float* grad_rows) {
  if (!c) return GQE_ERR_INVALID;
  if (c->tables.empty()) return fail(c, GQE_ERR_UNBOUND, "embedding tables are not bound");
  if (mode < 0 || mode >= (int)c->tables.size()) return fail(c, GQE_ERR_INVALID, "mode %d out of range", mode);
  if (!c->tables[mode]) return fail(c, GQE_ERR_UNBOUND, "mode %d has no table on this rank", mode);
  int rc = bwd_common(c, c->d, n, "gqe_encode_bwd_rows_device");
  if (rc != GQE_OK || n == 0) return rc;
  if (!rows || !gout || !grad_rows) return fail(c, GQE_ERR_INVALID, "gqe_encode_bwd_rows_device: null argument");
  GQE_BWD_LAUNCH(c, launch_encode_bwd_rows(c->d, n, c->tables[mode], rows, gout, grad_rows, c->table_rows[mode], c->d_err,
                                           c->stream));
  return GQE_OK;
}

extern "C" int gqe_adam_rows_device(gqe_ctx* c, float* table, float* exp_avg, float* exp_avg_sq, int32_t* last_step,
                                    int64_t table_rows, int32_t d, int64_t n, const int64_t* rows, const float* grad_rows,
                                    int32_t step, float lr, float beta1, float beta2, float eps) {
  if (!c) return GQE_ERR_INVALID;
  if (n < 0 || table_rows <= 0 || d <= 0 || step < 0) return fail(c, GQE_ERR_INVALID, "gqe_adam_rows_device: bad size");
  if (d > 256) return fail(c, GQE_ERR_UNSUPPORTED, "gqe_adam_rows_device: d = %d > 256", d);
  if (n == 0) return GQE_OK;
  if (!table || !exp_avg || !exp_avg_sq || !last_step) return fail(c, GQE_ERR_INVALID, "gqe_adam_rows_device: null argument");
  if (grad_rows && !rows) return fail(c, GQE_ERR_INVALID, "gqe_adam_rows_device: gradient rows need their row indices");
  if (grad_rows && step < 1) return fail(c, GQE_ERR_INVALID, "gqe_adam_rows_device: Adam steps are numbered from 1");
  GQE_CUDA(c, cudaSetDevice(c->device));
  drop_stale_error("gqe_adam_rows_device");
  GQE_BWD_LAUNCH(c, launch_adam_rows(table, exp_avg, exp_avg_sq, last_step, table_rows, d, n, rows, grad_rows, step, lr,
                                     beta1, beta2, eps, c->stream));
  return GQE_OK;
}

// ---- GraphSAGE-style encoder steps ---------------------------------------------------
extern "C" int gqe_segment_mean_device(gqe_ctx* c, const float* src, int64_t n_src, int32_t d, int64_t n,
                                       const int64_t* ptr, const int32_t* cols, float* out) {
  if (!c) return GQE_ERR_INVALID;
  if (n < 0 || n_src < 0 || d <= 0 || d % 4 != 0) return fail(c, GQE_ERR_INVALID, "gqe_segment_mean_device: bad size");
  if (n == 0) return GQE_OK;
  if (!src || !ptr || !cols || !out) return fail(c, GQE_ERR_INVALID, "gqe_segment_mean_device: null argument");
  GQE_CUDA(c, cudaSetDevice(c->device));
  drop_stale_error("gqe_segment_mean_device");
  GQE_BWD_LAUNCH(c, launch_segment_mean(src, n_src, d, n, ptr, cols, out, c->d_err, c->stream));
  return GQE_OK;
}

extern "C" int gqe_linear_device(gqe_ctx* c, const float* w, int32_t m, int32_t k, int64_t n, const float* x,
                                 int32_t relu, float* out) {
  if (!c) return GQE_ERR_INVALID;
  if (m <= 0 || k <= 0 || n < 0) return fail(c, GQE_ERR_INVALID, "gqe_linear_device: bad size");
  if (n == 0) return GQE_OK;
  if (!w || !x || !out) return fail(c, GQE_ERR_INVALID, "gqe_linear_device: null argument");
  GQE_CUDA(c, cudaSetDevice(c->device));
  drop_stale_error("gqe_linear_device");
  GQE_BWD_LAUNCH(c, launch_linear(w, m, k, n, x, relu, out, c->stream));
  return GQE_OK;
}

// ---- node-type-sharded tables: CUDA IPC mapping of peer shards ----------------------
#include <map>
#include <mutex>

namespace {
struct IpcKey {
  unsigned char b[GQE_IPC_HANDLE_BYTES];
  bool operator<(const IpcKey& o) const { return std::memcmp(b, o.b, sizeof b) < 0; }
};
struct IpcMapping {
  void* base;
  int refs;
  int device;
};
std::mutex g_ipc_mu;
std::map<IpcKey, IpcMapping> g_ipc_open;       // handle -> mapping (a handle can be opened once per process)
std::map<void*, IpcKey> g_ipc_by_ptr;          // returned peer pointer -> handle

// cuMemGetAddressRange through the runtime's driver entry point (no -lcuda at link time)
typedef int (*cuMemGetAddressRange_t)(unsigned long long* pbase, size_t* psize, unsigned long long dptr);
cuMemGetAddressRange_t address_range_fn() {
  static cuMemGetAddressRange_t fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<cuMemGetAddressRange_t>(p);
  }
  return fn;
}
}  // namespace

static_assert(sizeof(cudaIpcMemHandle_t) == GQE_IPC_HANDLE_BYTES, "IPC handle size");

static bool is_ipc_pointer(const void* p) {
  std::lock_guard<std::mutex> lock(g_ipc_mu);
  return g_ipc_by_ptr.count(const_cast<void*>(p)) != 0;
}

extern "C" int gqe_ipc_export(gqe_ctx* c, const void* dev_ptr, uint8_t* handle_out, int64_t* offset_out) {
  if (!c) return GQE_ERR_INVALID;
  if (!dev_ptr || !handle_out || !offset_out) return fail(c, GQE_ERR_INVALID, "gqe_ipc_export: null argument");
  GQE_CUDA(c, cudaSetDevice(c->device));
  cuMemGetAddressRange_t range = address_range_fn();
  if (!range) return fail(c, GQE_ERR_CUDA, "gqe_ipc_export: cuMemGetAddressRange is not available");
  unsigned long long base = 0;
  size_t size = 0;
  const int rc = range(&base, &size, (unsigned long long)(uintptr_t)dev_ptr);
  if (rc != 0) return fail(c, GQE_ERR_CUDA, "gqe_ipc_export: cuMemGetAddressRange failed (CUresult %d)", rc);
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, reinterpret_cast<void*>((uintptr_t)base));
  if (e != cudaSuccess)
    return fail(c, GQE_ERR_CUDA,
                "gqe_ipc_export: cudaIpcGetMemHandle failed: %s (the table must live in a cudaMalloc allocation, "
                "not an expandable-segments / VMM mapping)", cudaGetErrorString(e));
  std::memcpy(handle_out, &h, GQE_IPC_HANDLE_BYTES);
  *offset_out = (int64_t)((unsigned long long)(uintptr_t)dev_ptr - base);
  return GQE_OK;
}

extern "C" int gqe_ipc_open(gqe_ctx* c, const uint8_t* handle, int64_t offset, void** peer_ptr_out) {
  if (!c) return GQE_ERR_INVALID;
  if (!handle || !peer_ptr_out || offset < 0) return fail(c, GQE_ERR_INVALID, "gqe_ipc_open: bad argument");
  *peer_ptr_out = nullptr;
  GQE_CUDA(c, cudaSetDevice(c->device));
  IpcKey key;
  std::memcpy(key.b, handle, sizeof key.b);
  std::lock_guard<std::mutex> lock(g_ipc_mu);
  auto it = g_ipc_open.find(key);
  if (it == g_ipc_open.end()) {
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    void* base = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
      return fail(c, GQE_ERR_CUDA, "gqe_ipc_open: cudaIpcOpenMemHandle failed: %s (peer shard not reachable)",
                  cudaGetErrorString(e));
    it = g_ipc_open.insert({key, IpcMapping{base, 0, c->device}}).first;
  } else if (it->second.device != c->device) {
    return fail(c, GQE_ERR_UNSUPPORTED, "gqe_ipc_open: handle already mapped on device %d", it->second.device);
  }
  it->second.refs += 1;
  void* p = static_cast<char*>(it->second.base) + offset;
  g_ipc_by_ptr[p] = key;
  *peer_ptr_out = p;
  return GQE_OK;
}

extern "C" int gqe_ipc_close(gqe_ctx* c, void* peer_ptr) {
  if (!c) return GQE_ERR_INVALID;
  std::lock_guard<std::mutex> lock(g_ipc_mu);
  auto pit = g_ipc_by_ptr.find(peer_ptr);
  if (pit == g_ipc_by_ptr.end()) return fail(c, GQE_ERR_INVALID, "gqe_ipc_close: not a pointer from gqe_ipc_open");
  auto it = g_ipc_open.find(pit->second);
  if (it != g_ipc_open.end() && --it->second.refs <= 0) {
    cudaSetDevice(it->second.device);
    cudaDeviceSynchronize();  // no kernel may still be reading the peer shard
    cudaIpcCloseMemHandle(it->second.base);
    g_ipc_open.erase(it);
    for (auto q = g_ipc_by_ptr.begin(); q != g_ipc_by_ptr.end();)
      q = (std::memcmp(q->second.b, pit->second.b, sizeof pit->second.b) == 0 && q != pit) ? g_ipc_by_ptr.erase(q) : std::next(q);
  }
  g_ipc_by_ptr.erase(pit);
  return GQE_OK;
}

extern "C" int gqe_gather_rows_device(gqe_ctx* c, int32_t mode, int64_t n, const int32_t* rows, float* out) {
  if (!c) return GQE_ERR_INVALID;
  if (c->tables.empty()) return fail(c, GQE_ERR_UNBOUND, "embedding tables are not bound");
  if (mode < 0 || mode >= (int)c->tables.size()) return fail(c, GQE_ERR_INVALID, "mode %d out of range", mode);
  if (!c->tables[mode]) return fail(c, GQE_ERR_UNBOUND, "mode %d has no table on this rank", mode);
  if (n < 0) return fail(c, GQE_ERR_INVALID, "negative row count");
  if (n == 0) return GQE_OK;
  if (!rows || !out) return fail(c, GQE_ERR_INVALID, "gqe_gather_rows_device: null argument");
  GQE_CUDA(c, cudaSetDevice(c->device));
  GQE_CUDA(c, launch_gather_rows(c->tables[mode], rows, n, c->d, out, c->table_rows[mode], c->d_err, c->stream));
  c->launches += 1;
  return GQE_OK;
}
