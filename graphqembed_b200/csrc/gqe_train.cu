// gqe_train.cu -- the native training step behind gqe_train_step_*.
//
// The body of the reference's training loop (netquery/train_helpers.py:76-79,
// netquery/bio/train.py:59-62),
//     optimizer.zero_grad(); loss = enc_dec.margin_loss(formula, queries); loss.backward(); optimizer.step()
// for ONE formula batch and torch.optim.Adam, as one call: node ids in, loss out, every kernel of
// the forward pass, the backward pass and the optimiser launched from here with no interpreter
// between them (through Python + autograd the same ~20 small kernels cost 1.3 ms at the
// reference's batch of 512; launched back to back they are launch-bound at a few us each).
//
//   forward   exact fp32 operator kernels, chained as model.py:70-127 chains the reference's ops:
//             the query side is built once and scored against the positive and the negative targets
//   backward  the vector-Jacobian kernels of gqe_bwd.cu in reverse order (a small tape below does
//             what autograd does for graphqembed_b200/autograd.py)
//   update    Adam with the reference's trajectory.  Operator matrices / vectors: dense Adam on the
//             parameters that received a gradient (torch.optim.Adam skips parameters without one).
//             Embedding tables: the row gradients are accumulated into a zero [rows, d] buffer
//             (k_encode_bwd, as nn.Embedding's backward does) and k_adam_rows<GSUM> updates exactly
//             the rows the batch touched -- first catching each one up with the zero-gradient steps
//             dense Adam applied to it since it was last touched (gqe_opt.cu) -- and zeroes those
//             rows of the buffer again.  Rows are also caught up before the forward pass reads them.
//
// The optimiser state (moments, step counters, the accumulation buffers) belongs to the context:
// gqe_train_flush brings every row up to date (before evaluation / saving the tables),
// gqe_train_reset forgets the state.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <vector>

#include "gqe_ctx.h"

using namespace gqe;

struct gqe_train_state {
  struct Dense {
    float *m = nullptr, *v = nullptr, *g = nullptr;
    int64_t count = 0, step = 0;
    bool touched = false;
  };
  struct Table {
    float *m = nullptr, *v = nullptr, *gsum = nullptr;
    int32_t* last = nullptr;
    int64_t rows = 0, step = 0;
    bool touched = false;
  };
  std::map<const float*, Dense> dense;   // keyed by the bound parameter pointer
  std::vector<Table> tables;             // per mode
  // workspace of one step: [d, n] activations and gradients, bump-allocated
  std::vector<std::pair<float*, size_t>> blocks;   // (base, capacity in floats)
  size_t used = 0;                                 // floats used of blocks.back()
  int32_t* idx = nullptr;                          // lowered table rows of the batches since the last apply
  size_t idx_cap = 0, idx_used = 0;
  struct Pending { int mode; size_t offset; int64_t count; };
  std::vector<Pending> pending;                    // row lists whose accumulated gradients wait for the Adam step
  gqe_adam last_hyper{1e-3f, 0.9f, 0.999f, 1e-8f};
};

namespace {

constexpr int kLists = GQE_MAX_ANCHORS + 2;   // anchors 0..2, positive targets, negative targets

struct LowerParams {
  const int32_t* src[kLists];
  int32_t stride[kLists], offset[kLists], mode[kLists];
  int32_t* dst[kLists];
  ModeDev md[kLists];
  int64_t n;
  int32_t index_kind;
  unsigned long long* err;
};

// caller indices (table rows, or node ids) -> checked table rows, all index lists of the batch at once
__global__ void __launch_bounds__(256) k_lower(const __grid_constant__ LowerParams p) {
  const int l = blockIdx.y;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n || !p.src[l]) return;
  const int32_t v = __ldg(p.src[l] + i * p.stride[l] + p.offset[l]);
  p.dst[l][i] = resolve_index(p.md[l], p.mode[l], v, p.index_kind, p.err);
}

// model.py:124-126: loss = mean(clamp(margin - (pos - neg), 0)); d loss / d pos = -[active] / n, d loss / d neg = +[active] / n.
// One block: the mean is summed in a fixed order (deterministic).
// `weight` scales the gradient (the reference sums several batches' losses with weights before backward():
// train_helpers.py:63-75); the loss written is the batch's own, unweighted.
__global__ void __launch_bounds__(1024) k_hinge(const float* __restrict__ pos, const float* __restrict__ neg, int64_t n,
                                                float margin, float weight, float* __restrict__ gpos,
                                                float* __restrict__ gneg, float* __restrict__ out_loss) {
  __shared__ double red[32];
  const float inv = weight / (float)n;
  double local = 0.0;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float h = margin - (pos[i] - neg[i]);
    const bool on = h > 0.f;     // torch.clamp(min=0): gradient 1 where the input is >= min ... 0 at a tie is measure-zero
    local += on ? (double)h : 0.0;
    gpos[i] = on ? -inv : 0.f;
    gneg[i] = on ? inv : 0.f;
    if (h != h) {                // NaN propagates into the loss like the reference
      local += (double)h;
      gpos[i] = h; gneg[i] = h;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    *out_loss = (float)(s / (double)n);
  }
}

// out[k, c] = in[k, c mod n] for c < 2n: one query side against the positive and the negative targets
__global__ void __launch_bounds__(256) k_dup(const float* __restrict__ in, float* __restrict__ out, int d, int64_t n) {
  const int64_t total = (int64_t)d * n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = i / n, c = i - k * n;
    const float x = in[i];
    out[k * 2 * n + c] = x;
    out[k * 2 * n + n + c] = x;
  }
}
// its backward: gin[k, c] (+)= gout[k, c] + gout[k, n + c]
__global__ void __launch_bounds__(256) k_fold(const float* __restrict__ gout, float* __restrict__ gin, int d, int64_t n,
                                              int accumulate) {
  const int64_t total = (int64_t)d * n;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = i / n, c = i - k * n;
    const float g = gout[k * 2 * n + c] + gout[k * 2 * n + n + c];
    gin[i] = accumulate ? gin[i] + g : g;
  }
}

__global__ void __launch_bounds__(256) k_add(float* __restrict__ dst, const float* __restrict__ src, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] += src[i];
}

#define TR_CUDA(call)                                                                                  \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess && rc == GQE_OK)                                                             \
      rc = gqe_fail(c, GQE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)

// The tape: forward ops run immediately and are recorded; backward() walks them in reverse.
struct Tape {
  gqe_ctx* c;
  gqe_train_state* ts;
  int d;
  int rc = GQE_OK;
  int64_t launches = 0;

  struct Var { float* v; float* g; int64_t n; };   // value [d, n] (scores: [n]), gradient (lazily allocated)
  enum Kind { ENCODE, MATMUL, VEC, AGG, COS, DUP };
  struct Node {
    Kind k;
    int in[3], out;
    const float* w;        // MATMUL: matrix, VEC: vector
    int f1, f2;            // MATMUL: transpose; VEC: mul; AGG: relu, use_min; COS: raw dot
    int mode;              // ENCODE
    const int32_t* rows;   // ENCODE
  };
  std::vector<Var> vars;
  std::vector<Node> nodes;

  float* take(size_t floats) {
    floats = (floats + 63) & ~(size_t)63;
    if (ts->blocks.empty() || ts->used + floats > ts->blocks.back().second) {
      size_t cap = ts->blocks.empty() ? (size_t)1 << 22 : ts->blocks.back().second * 2;
      cap = std::max(cap, floats * 4);
      float* base = nullptr;
      cudaError_t e = cudaMalloc(&base, cap * sizeof(float));
      if (e != cudaSuccess) {
        if (rc == GQE_OK) rc = gqe_fail(c, GQE_ERR_CUDA, "training workspace: cudaMalloc(%zu bytes) failed: %s", cap * sizeof(float), cudaGetErrorString(e));
        return nullptr;
      }
      ts->blocks.emplace_back(base, cap);
      ts->used = 0;
    }
    float* p = ts->blocks.back().first + ts->used;
    ts->used += floats;
    return p;
  }
  int new_var(int64_t n, bool vector_only = false) {
    Var v;
    v.n = n;
    v.v = take(vector_only ? (size_t)n : (size_t)d * n);
    v.g = nullptr;
    vars.push_back(v);
    return (int)vars.size() - 1;
  }
  void op_launch(OpParams& op) {
    if (rc != GQE_OK) return;
    TR_CUDA(launch_op_simt(d, op, c->stream));
    ++launches;
  }

  // ---- forward ----------------------------------------------------------------------------
  int encode(int mode, const int32_t* rows, int64_t n) {
    const int out = new_var(n);
    if (rc != GQE_OK) return out;
    OpParams op;
    std::memset(&op, 0, sizeof op);
    op.op = OP_ENCODE; op.n = n; op.table = c->tables[mode]; op.table_rows = c->table_rows[mode]; op.err = c->d_err;
    op.rows = rows; op.out = vars[out].v;
    op_launch(op);
    nodes.push_back(Node{ENCODE, {-1, -1, -1}, out, nullptr, 0, 0, mode, rows});
    return out;
  }
  int matmul(const float* w, int x, int transpose) {
    const int out = new_var(vars[x].n);
    if (rc != GQE_OK) return out;
    OpParams op;
    std::memset(&op, 0, sizeof op);
    op.op = OP_MATMUL; op.mutate = transpose; op.n = vars[x].n; op.rel[0] = w; op.in0 = vars[x].v; op.out = vars[out].v;
    op_launch(op);
    nodes.push_back(Node{MATMUL, {x, -1, -1}, out, w, transpose, 0, 0, nullptr});
    return out;
  }
  int vec(const float* v, int x, int mul) {
    const int out = new_var(vars[x].n);
    if (rc != GQE_OK) return out;
    OpParams op;
    std::memset(&op, 0, sizeof op);
    op.op = OP_PROJECT; op.decoder = mul ? GQE_DEC_DISTMULT : GQE_DEC_TRANSE; op.n = vars[x].n; op.rel[0] = v;
    op.in0 = vars[x].v; op.out = vars[out].v;
    op_launch(op);
    nodes.push_back(Node{VEC, {x, -1, -1}, out, v, mul, 0, 0, nullptr});
    return out;
  }
  int agg(int e1, int e2, int e3, int relu, int use_min) {
    const int out = new_var(vars[e1].n);
    if (rc != GQE_OK) return out;
    TR_CUDA(launch_aggregate((int64_t)d * vars[e1].n, vars[e1].v, vars[e2].v, e3 >= 0 ? vars[e3].v : nullptr, relu, use_min,
                             vars[out].v, c->stream));
    ++launches;
    nodes.push_back(Node{AGG, {e1, e2, e3}, out, nullptr, relu, use_min, 0, nullptr});
    return out;
  }
  int dup(int x) {     // [d, n] -> [d, 2n], both halves equal
    const int64_t n = vars[x].n;
    const int out = new_var(2 * n);
    if (rc != GQE_OK) return out;
    k_dup<<<(unsigned)std::min<int64_t>(((int64_t)d * n + 255) / 256, 1184), 256, 0, c->stream>>>(vars[x].v, vars[out].v, d, n);
    TR_CUDA(cudaGetLastError());
    ++launches;
    nodes.push_back(Node{DUP, {x, -1, -1}, out, nullptr, 0, 0, 0, nullptr});
    return out;
  }
  int cosine(int x, int y, int raw) {
    const int out = new_var(vars[x].n, true);
    if (rc != GQE_OK) return out;
    if (raw) {
      TR_CUDA(launch_dot(d, vars[x].n, vars[x].v, vars[y].v, vars[out].v, c->stream));
      ++launches;
    } else {
      OpParams op;
      std::memset(&op, 0, sizeof op);
      op.op = OP_COSINE; op.n = vars[x].n; op.in0 = vars[x].v; op.in1 = vars[y].v; op.out = vars[out].v;
      op_launch(op);
    }
    nodes.push_back(Node{COS, {x, y, -1}, out, nullptr, raw, 0, 0, nullptr});
    return out;
  }

  // ---- backward ---------------------------------------------------------------------------
  // where a kernel may WRITE the gradient of `var`: its gradient buffer when it has none yet, a
  // temporary otherwise (accumulate() then adds it in)
  float* grad_dst(int var, bool& fresh) {
    Var& v = vars[var];
    fresh = v.g == nullptr;
    float* p = take((size_t)d * v.n);
    if (fresh) v.g = p;
    return p;
  }
  void accumulate(int var, float* tmp, bool fresh) {
    if (fresh || rc != GQE_OK) return;
    const int64_t total = (int64_t)d * vars[var].n;
    k_add<<<(unsigned)std::min<int64_t>((total + 255) / 256, 1184), 256, 0, c->stream>>>(vars[var].g, tmp, total);
    TR_CUDA(cudaGetLastError());
    ++launches;
  }
  // gradient buffer of a dense parameter, zeroed on its first use in this step
  float* param_grad(const float* w, int64_t count) {
    gqe_train_state::Dense& ds = ts->dense[w];
    if (!ds.m) {
      TR_CUDA(cudaMalloc(&ds.m, 3 * count * sizeof(float)));
      if (rc != GQE_OK) return nullptr;
      ds.v = ds.m + count;
      ds.g = ds.v + count;
      ds.count = count;
      TR_CUDA(cudaMemsetAsync(ds.m, 0, 2 * count * sizeof(float), c->stream));
    }
    if (!ds.touched) {
      TR_CUDA(cudaMemsetAsync(ds.g, 0, count * sizeof(float), c->stream));
      ds.touched = true;
    }
    return ds.g;
  }

  void backward() {
    for (int i = (int)nodes.size() - 1; i >= 0 && rc == GQE_OK; --i) {
      const Node& nd = nodes[i];
      float* gout = vars[nd.out].g;
      if (!gout) continue;    // nothing flowed into this output
      const int64_t n = vars[nd.out].n;
      switch (nd.k) {
        case COS: {
          bool fx, fy;
          float* gx = grad_dst(nd.in[0], fx);
          float* gy = grad_dst(nd.in[1], fy);
          if (rc != GQE_OK) break;
          TR_CUDA(launch_cosine_bwd(d, n, vars[nd.in[0]].v, vars[nd.in[1]].v, gout, nd.f1, gx, gy, c->stream));
          ++launches;
          accumulate(nd.in[0], gx, fx);
          accumulate(nd.in[1], gy, fy);
          break;
        }
        case MATMUL: {
          float* gw = param_grad(nd.w, (int64_t)d * d);
          if (rc != GQE_OK) break;
          TR_CUDA(launch_matmul_wgrad(d, n, gout, vars[nd.in[0]].v, nd.f1 ? 1 : 0, gw, c->stream));
          ++launches;
          bool fx;
          float* gx = grad_dst(nd.in[0], fx);
          if (rc != GQE_OK) break;
          OpParams op;
          std::memset(&op, 0, sizeof op);
          op.op = OP_MATMUL; op.mutate = !nd.f1; op.n = n; op.rel[0] = nd.w; op.in0 = gout; op.out = gx;
          op_launch(op);
          accumulate(nd.in[0], gx, fx);
          break;
        }
        case VEC: {
          float* gv = param_grad(nd.w, d);
          if (rc != GQE_OK) break;
          TR_CUDA(launch_rowsum(d, n, gout, nd.f1 ? vars[nd.in[0]].v : nullptr, gv, c->stream));
          ++launches;
          Var& x = vars[nd.in[0]];
          if (!nd.f1 && !x.g) {
            x.g = gout;                 // y = x + v: the gradient passes through unchanged
          } else {
            bool fx;
            float* gx = grad_dst(nd.in[0], fx);
            if (rc != GQE_OK) break;
            if (nd.f1) {                // y = x * v: g * v
              OpParams op;
              std::memset(&op, 0, sizeof op);
              op.op = OP_PROJECT; op.decoder = GQE_DEC_DISTMULT; op.n = n; op.rel[0] = nd.w; op.in0 = gout; op.out = gx;
              op_launch(op);
              accumulate(nd.in[0], gx, fx);
            } else {
              accumulate(nd.in[0], gout, false);
            }
          }
          break;
        }
        case AGG: {
          bool f1, f2, f3 = true;
          float* g1 = grad_dst(nd.in[0], f1);
          float* g2 = grad_dst(nd.in[1], f2);
          float* g3 = nd.in[2] >= 0 ? grad_dst(nd.in[2], f3) : nullptr;
          if (rc != GQE_OK) break;
          TR_CUDA(launch_aggregate_bwd((int64_t)d * n, vars[nd.in[0]].v, vars[nd.in[1]].v, nd.in[2] >= 0 ? vars[nd.in[2]].v : nullptr,
                                       nd.f1, nd.f2, gout, g1, g2, g3, c->stream));
          ++launches;
          accumulate(nd.in[0], g1, f1);
          accumulate(nd.in[1], g2, f2);
          if (nd.in[2] >= 0) accumulate(nd.in[2], g3, f3);
          break;
        }
        case DUP: {
          Var& x = vars[nd.in[0]];
          const bool fresh = x.g == nullptr;
          if (fresh) x.g = take((size_t)d * x.n);
          if (rc != GQE_OK) break;
          k_fold<<<(unsigned)std::min<int64_t>(((int64_t)d * x.n + 255) / 256, 1184), 256, 0, c->stream>>>(gout, x.g, d, x.n,
                                                                                                     fresh ? 0 : 1);
          TR_CUDA(cudaGetLastError());
          ++launches;
          break;
        }
        case ENCODE: {
          gqe_train_state::Table& t = ts->tables[nd.mode];
          TR_CUDA(launch_encode_bwd(d, n, c->tables[nd.mode], nd.rows, gout, t.gsum, c->table_rows[nd.mode], c->d_err, c->stream));
          ++launches;
          t.touched = true;
          break;
        }
      }
    }
  }
};

int ensure_table_state(gqe_ctx* c, gqe_train_state* ts, int mode) {
  if ((int)ts->tables.size() < (int)c->tables.size()) ts->tables.resize(c->tables.size());
  gqe_train_state::Table& t = ts->tables[mode];
  const int64_t rows = c->table_rows[mode];
  if (t.m && t.rows == rows) return GQE_OK;
  if (t.m) return gqe_fail(c, GQE_ERR_INVALID, "the table of mode %d changed size under the optimiser state: call gqe_train_reset", mode);
  if (c->table_remote[mode]) return gqe_fail(c, GQE_ERR_UNSUPPORTED, "training needs the table of mode %d on this GPU", mode);
  const size_t cells = (size_t)rows * c->d;
  GQE_CUDA(c, cudaMalloc(&t.m, 3 * cells * sizeof(float)));
  t.v = t.m + cells;
  t.gsum = t.v + cells;
  GQE_CUDA(c, cudaMalloc(&t.last, (size_t)rows * sizeof(int32_t)));
  GQE_CUDA(c, cudaMemsetAsync(t.m, 0, 3 * cells * sizeof(float), c->stream));
  GQE_CUDA(c, cudaMemsetAsync(t.last, 0, (size_t)rows * sizeof(int32_t), c->stream));
  t.rows = rows;
  return GQE_OK;
}

int n_anchors_of(int structure) {
  switch (structure) {
    case GQE_CHAIN1: case GQE_CHAIN2: case GQE_CHAIN3: return 1;
    case GQE_INTER2: case GQE_INTER_CHAIN3: case GQE_CHAIN_INTER3: return 2;
    case GQE_INTER3: return 3;
    default: return -1;
  }
}
int n_rels_of(int structure) {
  switch (structure) {
    case GQE_CHAIN1: return 1;
    case GQE_CHAIN2: case GQE_INTER2: return 2;
    case GQE_CHAIN3: case GQE_INTER3: case GQE_INTER_CHAIN3: case GQE_CHAIN_INTER3: return 3;
    default: return -1;
  }
}

// forward + backward of one batch: gradients (scaled by `weight`) are ACCUMULATED in the state
int train_backward(gqe_ctx* c, const gqe_plan* plan, int64_t n, const int32_t* anchors, const int32_t* pairs, float margin,
                   float weight, const gqe_adam* hyper, float* out_loss, int index_kind) {
  if (!c) return GQE_ERR_INVALID;
  if (!plan || !hyper || !out_loss) return gqe_fail(c, GQE_ERR_INVALID, "gqe_train_step: null argument");
  if (n <= 0) return gqe_fail(c, GQE_ERR_INVALID, "gqe_train_step: empty batch");
  if (n > 0x7fffffffLL / 4) return gqe_fail(c, GQE_ERR_UNSUPPORTED, "gqe_train_step: batch too large");
  if (!anchors || !pairs) return gqe_fail(c, GQE_ERR_INVALID, "gqe_train_step: index arrays are null");
  const int st = plan->structure;
  const int na = n_anchors_of(st), nr = n_rels_of(st);
  if (na < 0) return gqe_fail(c, GQE_ERR_INVALID, "unknown query structure %d", st);
  if (c->tables.empty()) return gqe_fail(c, GQE_ERR_UNBOUND, "embedding tables are not bound");
  if (c->rels.empty()) return gqe_fail(c, GQE_ERR_UNBOUND, "relation parameters are not bound");
  if (c->rel_d != c->d) return gqe_fail(c, GQE_ERR_UNSUPPORTED, "relation dim %d != table dim %d", c->rel_d, c->d);
  if (c->d != 32 && c->d != 64 && c->d != 128 && c->d != 256)
    return gqe_fail(c, GQE_ERR_UNSUPPORTED, "dimension %d not supported (32/64/128/256)", c->d);
  if (index_kind && c->node_maps.size() != c->tables.size())
    return gqe_fail(c, GQE_ERR_UNBOUND, "node maps are not bound (gqe_bind_node_maps)");
  const int nm = (int)c->tables.size();
  int modes[kLists];
  for (int k = 0; k < GQE_MAX_ANCHORS; ++k) modes[k] = k < na ? plan->anchor_mode[k] : -1;
  modes[3] = modes[4] = plan->target_mode;
  for (int l = 0; l < kLists; ++l)
    if (modes[l] >= nm || (modes[l] < 0 && (l >= 3 || l < na)) || (modes[l] >= 0 && !c->tables[modes[l]]))
      return gqe_fail(c, GQE_ERR_INVALID, "mode %d of the plan has no bound table", modes[l]);
  const float* rel[GQE_MAX_RELS] = {nullptr, nullptr, nullptr};
  for (int k = 0; k < nr; ++k) {
    if (plan->rel[k] < 0 || plan->rel[k] >= (int)c->rels.size())
      return gqe_fail(c, GQE_ERR_INVALID, "relation id %d out of range", plan->rel[k]);
    rel[k] = c->rels[plan->rel[k]];
  }
  const bool chain = st <= GQE_CHAIN3;
  const bool bilinear = c->decoder == GQE_DEC_BILINEAR;
  const int mul = c->decoder == GQE_DEC_DISTMULT ? 1 : 0;
  const float *pre = nullptr, *post = nullptr;
  bool deepsets = false, use_min = false;
  if (!chain) {
    if (c->inter < 0) return gqe_fail(c, GQE_ERR_UNBOUND, "intersection operator is not bound");
    deepsets = c->inter <= GQE_INTER_DEEPSETS_MIN;
    use_min = c->inter == GQE_INTER_DEEPSETS_MIN || c->inter == GQE_INTER_SIMPLE_MIN;
    if (deepsets) {
      if (c->inter_d != c->d) return gqe_fail(c, GQE_ERR_UNSUPPORTED, "intersection dim %d != table dim %d", c->inter_d, c->d);
      if (plan->inter_mode < 0 || plan->inter_mode >= (int)c->pre.size())
        return gqe_fail(c, GQE_ERR_INVALID, "intersection mode %d out of range", plan->inter_mode);
      pre = c->pre[plan->inter_mode];
      post = c->post[plan->inter_mode];
    }
  }
  GQE_CUDA(c, cudaSetDevice(c->device));
  (void)cudaGetLastError();
  if (!c->train) c->train = new gqe_train_state();
  gqe_train_state* ts = c->train;
  ts->last_hyper = *hyper;

  // workspace: one block (several only while it is still growing)
  if (ts->blocks.size() > 1) {
    GQE_CUDA(c, cudaStreamSynchronize(c->stream));
    size_t total = 0;
    for (auto& b : ts->blocks) { total += b.second; cudaFree(b.first); }
    ts->blocks.clear();
    float* base = nullptr;
    GQE_CUDA(c, cudaMalloc(&base, total * sizeof(float)));
    ts->blocks.emplace_back(base, total);
  }
  ts->used = 0;
  if (ts->idx_cap < ts->idx_used + (size_t)kLists * n) {   // (the lists of earlier batches of this step are kept)
    GQE_CUDA(c, cudaStreamSynchronize(c->stream));
    const size_t cap = 2 * (ts->idx_used + (size_t)kLists * n);
    int32_t* grown = nullptr;
    GQE_CUDA(c, cudaMalloc(&grown, cap * sizeof(int32_t)));
    if (ts->idx_used) GQE_CUDA(c, cudaMemcpy(grown, ts->idx, ts->idx_used * sizeof(int32_t), cudaMemcpyDeviceToDevice));
    cudaFree(ts->idx);
    ts->idx = grown;
    ts->idx_cap = cap;
  }
  const size_t idx0 = ts->idx_used;
  ts->idx_used += (size_t)kLists * n;
  for (int l = 0; l < kLists; ++l)
    if (modes[l] >= 0) {
      int rc = ensure_table_state(c, ts, modes[l]);
      if (rc != GQE_OK) return rc;
    }

  // ---- indices -> checked table rows (one launch for all lists) ------------------------------
  int32_t* rows[kLists];
  LowerParams lw;
  std::memset(&lw, 0, sizeof lw);
  for (int l = 0; l < kLists; ++l) {
    rows[l] = ts->idx + idx0 + (size_t)l * n;
    if (modes[l] < 0) continue;
    const bool tgt = l >= 3;
    lw.src[l] = tgt ? pairs : anchors + (size_t)l * n;
    lw.stride[l] = tgt ? 2 : 1;
    lw.offset[l] = l == 4 ? 1 : 0;
    lw.mode[l] = modes[l];
    lw.dst[l] = rows[l];
    ModeDev& md = lw.md[l];
    if (index_kind) md = c->node_maps[modes[l]];
    else { md.lut = nullptr; md.base = 0; md.len = 0; }
    md.rows = (uint32_t)std::min<int64_t>(c->table_rows[modes[l]], 0x7fffffffLL);
  }
  lw.n = n;
  lw.index_kind = index_kind;
  lw.err = c->d_err;
  k_lower<<<dim3((unsigned)((n + 255) / 256), kLists), 256, 0, c->stream>>>(lw);
  GQE_CUDA(c, cudaGetLastError());
  c->launches += 1;

  // ---- rows about to be read: catch them up with the zero-gradient steps they missed ------------
  const gqe_adam& h = *hyper;
  for (int l = 0; l < kLists; ++l) {
    if (modes[l] < 0) continue;
    gqe_train_state::Table& t = ts->tables[modes[l]];
    if (t.step <= 0 || l == 4) continue;               // (list 4 = the negatives, adjacent to list 3: one call for both)
    GQE_CUDA(c, launch_adam_rows_accum(const_cast<float*>(c->tables[modes[l]]), t.m, t.v, t.last, t.rows, c->d, l == 3 ? 2 * n : n, rows[l], nullptr,
                                       (int)t.step, h.lr, h.beta1, h.beta2, h.eps, c->stream));
    c->launches += 1;
  }

  // ---- forward (model.py:70-109), the query side once, positive and negative targets ------------
  Tape tp;
  tp.c = c; tp.ts = ts; tp.d = c->d;
  auto project = [&](int x, const float* r) { return bilinear ? tp.matmul(r, x, 0) : tp.vec(r, x, mul); };   // decoders.py:150 / 208 / 236
  // the positive and the negative targets form ONE [d, 2n] block (columns [0, n) positive, [n, 2n)
  // negative: their row lists are adjacent in ts->idx), scored against the duplicated query side
  int score;
  if (chain) {
    const int a2 = tp.dup(tp.encode(modes[0], rows[0], n));
    int act = tp.encode(modes[3], rows[3], 2 * n);
    for (int k = 0; k < nr; ++k) act = bilinear ? tp.matmul(rel[k], act, 1) : tp.vec(rel[k], act, mul);   // decoders.py:143-145
    score = tp.cosine(act, a2, mul);                                                                      // raw dot for DistMult
  } else {
    int e[3] = {-1, -1, -1};
    e[0] = project(tp.encode(modes[0], rows[0], n), rel[0]);
    e[1] = project(tp.encode(modes[1], rows[1], n), rel[1]);
    if (st == GQE_INTER_CHAIN3) e[1] = project(e[1], rel[2]);
    if (st == GQE_INTER3) e[2] = project(tp.encode(modes[2], rows[2], n), rel[2]);
    int q;
    if (deepsets) {
      int hdn[3] = {-1, -1, -1};
      for (int b = 0; b < na; ++b) hdn[b] = tp.matmul(pre, e[b], 0);        // decoders.py:289-292
      q = tp.matmul(post, tp.agg(hdn[0], hdn[1], hdn[2], 1, use_min), 0);    // decoders.py:293-299
    } else {
      q = tp.agg(e[0], e[1], e[2], 0, use_min);                              // decoders.py:311-319
    }
    if (st == GQE_CHAIN_INTER3) q = project(q, rel[2]);                      // model.py:107
    score = tp.cosine(tp.encode(modes[3], rows[3], 2 * n), tp.dup(q), 0);
  }
  if (tp.rc != GQE_OK) return tp.rc;

  // ---- loss and its gradient ------------------------------------------------------------------
  tp.vars[score].g = tp.take((size_t)2 * n);
  if (tp.rc != GQE_OK) return tp.rc;
  k_hinge<<<1, 1024, 0, c->stream>>>(tp.vars[score].v, tp.vars[score].v + n, n, margin, weight, tp.vars[score].g,
                                    tp.vars[score].g + n, out_loss);
  GQE_CUDA(c, cudaGetLastError());
  c->launches += 1;

  // ---- backward ---------------------------------------------------------------------------------
  // (score vars hold [n] gradients; every other gradient is [d, n])
  tp.backward();
  c->launches += tp.launches;
  if (tp.rc != GQE_OK) return tp.rc;

  // the rows whose gradients now sit in the accumulation buffers
  for (int l = 0; l < 4; ++l)
    if (modes[l] >= 0) ts->pending.push_back(gqe_train_state::Pending{modes[l], idx0 + (size_t)l * n, l == 3 ? 2 * n : n});
  return GQE_OK;
}

// Adam on everything that received a gradient since the last apply
int train_apply(gqe_ctx* c, const gqe_adam* hyper) {
  if (!c) return GQE_ERR_INVALID;
  if (!hyper) return gqe_fail(c, GQE_ERR_INVALID, "gqe_train_apply: null argument");
  if (!c->train) return GQE_OK;
  GQE_CUDA(c, cudaSetDevice(c->device));
  gqe_train_state* ts = c->train;
  ts->last_hyper = *hyper;
  const gqe_adam& h = *hyper;
  for (int m = 0; m < (int)ts->tables.size(); ++m) {
    gqe_train_state::Table& t = ts->tables[m];
    if (!t.touched) continue;
    t.touched = false;
    t.step += 1;
    for (const auto& pr : ts->pending) {
      if (pr.mode != m) continue;
      GQE_CUDA(c, launch_adam_rows_accum(const_cast<float*>(c->tables[m]), t.m, t.v, t.last, t.rows, c->d, pr.count,
                                         ts->idx + pr.offset, t.gsum, (int)t.step, h.lr, h.beta1, h.beta2, h.eps, c->stream));
      c->launches += 1;
    }
  }
  ts->pending.clear();
  ts->idx_used = 0;
  for (auto& kv : ts->dense) {
    gqe_train_state::Dense& ds = kv.second;
    if (!ds.touched) continue;
    ds.touched = false;
    ds.step += 1;
    GQE_CUDA(c, launch_adam_dense(const_cast<float*>(kv.first), ds.m, ds.v, ds.g, ds.count, (int)ds.step, h.lr, h.beta1, h.beta2,
                                  h.eps, c->stream));
    c->launches += 1;
  }
  // the operator matrices changed: packed / pre-multiplied images of the tensor-core path are stale
  c->wcache.clear();
  return GQE_OK;
}

int train_step(gqe_ctx* c, const gqe_plan* plan, int64_t n, const int32_t* anchors, const int32_t* pairs, float margin,
               const gqe_adam* hyper, float* out_loss, int index_kind) {
  if (c && c->train && !c->train->pending.empty())
    return gqe_fail(c, GQE_ERR_INVALID, "gqe_train_step: gradients of gqe_train_backward calls are pending; call gqe_train_apply first");
  int rc = train_backward(c, plan, n, anchors, pairs, margin, 1.f, hyper, out_loss, index_kind);
  return rc != GQE_OK ? rc : train_apply(c, hyper);
}

}  // namespace

void gqe_train_destroy(gqe_ctx* c) {
  if (!c || !c->train) return;
  gqe_train_state* ts = c->train;
  for (auto& kv : ts->dense) cudaFree(kv.second.m);
  for (auto& t : ts->tables) { cudaFree(t.m); cudaFree(t.last); }
  for (auto& b : ts->blocks) cudaFree(b.first);
  cudaFree(ts->idx);
  delete ts;
  c->train = nullptr;
}

extern "C" int gqe_train_step_device(gqe_ctx* c, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_rows,
                                     const int32_t* pair_rows, float margin, const gqe_adam* hyper, float* out_loss) {
  return train_step(c, plan, n_queries, anchor_rows, pair_rows, margin, hyper, out_loss, 0);
}
extern "C" int gqe_train_step_nodes_device(gqe_ctx* c, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_nodes,
                                           const int32_t* pair_nodes, float margin, const gqe_adam* hyper, float* out_loss) {
  return train_step(c, plan, n_queries, anchor_nodes, pair_nodes, margin, hyper, out_loss, 1);
}

// host index buffers in, host loss out: one H2D copy per array in front, the loss through the mapped
// pinned word, the index-error word fetched with it
static int train_step_host(gqe_ctx* c, const gqe_plan* plan, int64_t n, const int32_t* anchors, const int32_t* pairs,
                           float margin, const gqe_adam* hyper, float* out_loss, int index_kind, bool apply = true,
                           float weight = 1.f) {
  if (!c) return GQE_ERR_INVALID;
  if (!plan || !hyper || !out_loss || !anchors || !pairs) return gqe_fail(c, GQE_ERR_INVALID, "gqe_train_step_host: null argument");
  if (n <= 0) return gqe_fail(c, GQE_ERR_INVALID, "gqe_train_step_host: empty batch");
  const int na = n_anchors_of(plan->structure);
  if (na < 0) return gqe_fail(c, GQE_ERR_INVALID, "unknown query structure %d", plan->structure);
  GQE_CUDA(c, cudaSetDevice(c->device));
  int rc;
  if ((rc = gqe_stage_reserve(c, 0, sizeof(int32_t) * (size_t)na * n)) != GQE_OK) return rc;
  if ((rc = gqe_stage_reserve(c, 1, sizeof(int32_t) * (size_t)2 * n)) != GQE_OK) return rc;
  if (!c->h_loss) {
    GQE_CUDA(c, cudaHostAlloc((void**)&c->h_loss, sizeof(float), cudaHostAllocMapped));
    GQE_CUDA(c, cudaHostGetDevicePointer((void**)&c->h_loss_dev, c->h_loss, 0));
  }
  GQE_CUDA(c, cudaMemcpyAsync(c->stage[0], anchors, sizeof(int32_t) * (size_t)na * n, cudaMemcpyHostToDevice, c->stream));
  GQE_CUDA(c, cudaMemcpyAsync(c->stage[1], pairs, sizeof(int32_t) * (size_t)2 * n, cudaMemcpyHostToDevice, c->stream));
  rc = apply ? train_step(c, plan, n, (const int32_t*)c->stage[0], (const int32_t*)c->stage[1], margin, hyper, c->h_loss_dev, index_kind)
             : train_backward(c, plan, n, (const int32_t*)c->stage[0], (const int32_t*)c->stage[1], margin, weight, hyper,
                              c->h_loss_dev, index_kind);
  if (rc != GQE_OK) return rc;
  GQE_CUDA(c, cudaMemcpyAsync(c->h_err, c->d_err, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
  GQE_CUDA(c, cudaStreamSynchronize(c->stream));
  *out_loss = *(volatile float*)c->h_loss;
  if (*(volatile unsigned long long*)&c->h_err[0] != 0ull) {
    int32_t kind, mode;
    int64_t value;
    return gqe_index_error(c, &kind, &mode, &value);
  }
  return GQE_OK;
}
extern "C" int gqe_train_step_host(gqe_ctx* c, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_rows,
                                   const int32_t* pair_rows, float margin, const gqe_adam* hyper, float* out_loss) {
  return train_step_host(c, plan, n_queries, anchor_rows, pair_rows, margin, hyper, out_loss, 0);
}
extern "C" int gqe_train_step_nodes_host(gqe_ctx* c, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_nodes,
                                         const int32_t* pair_nodes, float margin, const gqe_adam* hyper, float* out_loss) {
  return train_step_host(c, plan, n_queries, anchor_nodes, pair_nodes, margin, hyper, out_loss, 1);
}

// gradient accumulation over several batches, then one Adam step (the reference's loop sums weighted losses
// of several query types before backward(): train_helpers.py:63-79)
extern "C" int gqe_train_backward_device(gqe_ctx* c, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_rows,
                                         const int32_t* pair_rows, float margin, float weight, const gqe_adam* hyper,
                                         float* out_loss) {
  return train_backward(c, plan, n_queries, anchor_rows, pair_rows, margin, weight, hyper, out_loss, 0);
}
extern "C" int gqe_train_backward_nodes_device(gqe_ctx* c, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_nodes,
                                               const int32_t* pair_nodes, float margin, float weight, const gqe_adam* hyper,
                                               float* out_loss) {
  return train_backward(c, plan, n_queries, anchor_nodes, pair_nodes, margin, weight, hyper, out_loss, 1);
}
extern "C" int gqe_train_backward_host(gqe_ctx* c, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_rows,
                                       const int32_t* pair_rows, float margin, float weight, const gqe_adam* hyper,
                                       float* out_loss) {
  return train_step_host(c, plan, n_queries, anchor_rows, pair_rows, margin, hyper, out_loss, 0, false, weight);
}
extern "C" int gqe_train_backward_nodes_host(gqe_ctx* c, const gqe_plan* plan, int64_t n_queries, const int32_t* anchor_nodes,
                                             const int32_t* pair_nodes, float margin, float weight, const gqe_adam* hyper,
                                             float* out_loss) {
  return train_step_host(c, plan, n_queries, anchor_nodes, pair_nodes, margin, hyper, out_loss, 1, false, weight);
}
extern "C" int gqe_train_apply(gqe_ctx* c, const gqe_adam* hyper) { return train_apply(c, hyper); }

extern "C" int gqe_train_flush(gqe_ctx* c) {
  if (!c) return GQE_ERR_INVALID;
  if (!c->train) return GQE_OK;
  GQE_CUDA(c, cudaSetDevice(c->device));
  gqe_train_state* ts = c->train;
  const gqe_adam h = ts->last_hyper;
  for (int m = 0; m < (int)ts->tables.size(); ++m) {
    gqe_train_state::Table& t = ts->tables[m];
    if (!t.m || t.step <= 0 || m >= (int)c->tables.size() || !c->tables[m]) continue;
    GQE_CUDA(c, launch_adam_rows_accum(const_cast<float*>(c->tables[m]), t.m, t.v, t.last, t.rows, c->d, t.rows, nullptr, nullptr,
                                       (int)t.step, h.lr, h.beta1, h.beta2, h.eps, c->stream));
    c->launches += 1;
  }
  return GQE_OK;
}

extern "C" int gqe_train_reset(gqe_ctx* c) {
  if (!c) return GQE_ERR_INVALID;
  if (!c->train) return GQE_OK;
  GQE_CUDA(c, cudaSetDevice(c->device));
  GQE_CUDA(c, cudaStreamSynchronize(c->stream));
  gqe_train_destroy(c);
  return GQE_OK;
}

extern "C" int64_t gqe_train_steps(const gqe_ctx* c, int32_t mode) {
  if (!c || !c->train || mode < 0 || mode >= (int)c->train->tables.size()) return 0;
  return c->train->tables[mode].step;
}
