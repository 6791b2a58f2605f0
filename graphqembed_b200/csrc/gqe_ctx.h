// gqe_ctx.h -- the context behind the C ABI (include/gqe.h), shared by the host-side translation
// units (gqe_capi.cu: binding, scoring entry points; gqe_train.cu: the native training step).
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/gqe.h"
#include "gqe_launch.h"

// state of the native training step (gqe_train.cu): Adam moments per parameter, owned by the context
struct gqe_train_state;

struct gqe_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  int64_t launches = 0;

  int d = 0;
  std::vector<const float*> tables;
  std::vector<int64_t> table_rows;
  std::vector<char> table_remote;  // the table lives in a peer GPU's HBM (mapped with gqe_ipc_open)

  int decoder = -1;
  int rel_d = 0;
  std::vector<const float*> rels;

  int inter = -1;
  int inter_d = 0;
  std::vector<const float*> pre, post;

  // arithmetic of the d x d contractions (gqe_precision)
  int precision = GQE_PREC_BF16X3;
  // tensor-core path: packed (bf16 hi/lo, pre-swizzled) images of the operator matrices.  The
  // buffer is a CACHE across calls: slot i holds the image described by wcache[i] (a parameter
  // matrix or a product of up to three, one orientation, one column order) until
  // gqe_invalidate_weights / a re-bind drops it, so a call whose matrices are all cached
  // launches the fused kernel alone.
  uint8_t* packed = nullptr;
  size_t packed_cap = 0;
  struct WKey {
    const float* f[3];
    int n, right_assoc, chain_form, perm;
    bool operator==(const WKey& o) const {
      return n == o.n && f[0] == o.f[0] && f[1] == o.f[1] && f[2] == o.f[2] && right_assoc == o.right_assoc &&
             chain_form == o.chain_form && perm == o.perm;
    }
  };
  std::vector<WKey> wcache;
  bool wcache_on = true;
  int64_t weight_preps = 0;    // matrices packed so far (bench bookkeeping)

  // node id -> table row maps (gqe_bind_node_maps); rows of the bound tables are in table_rows
  std::vector<gqe::ModeDev> node_maps;
  // first bad index seen by a kernel: DEVICE [2] and its mapped pinned twin for the *_host calls
  unsigned long long* d_err = nullptr;
  unsigned long long* h_err = nullptr;
  unsigned long long* h_err_dev = nullptr;

  // staging area of the fused kernel's helper warps (rows of peer GPUs' tables, node-type-sharded tables)
  float* stage_buf = nullptr;
  size_t stage_buf_cap = 0;

  // a *_host call copied its index arrays with gqe_fetch_indices: the next tensor-core launch waits for that
  // kernel itself (LaunchParams::wait_prev)
  bool idx_by_kernel = false;

  // diagnostics: per-tile phase stamps of the tensor-core kernel
  unsigned long long* phase_log = nullptr;
  int64_t phase_cap = 0;

  // operator pre-composition on the tensor-core path (gqe_compose_mode) and its fp32 scratch
  int compose = GQE_COMPOSE_AUTO;
  float* compose_buf = nullptr;
  size_t compose_cap = 0;

  // query embeddings of the many-targets-per-query path (fp32 [n_queries, d], grows on demand)
  float* qbuf = nullptr;
  size_t qbuf_cap = 0;

  // margin-loss reduction scratch
  double* partials = nullptr;
  int64_t partials_cap = 0;
  double* loss_acc = nullptr;
  bool loss_dirty = false;     // *loss_acc may be non-zero (a call failed between its launches)
  float* h_loss = nullptr;     // mapped pinned word the kernels write the loss of a *_host call to
  float* h_loss_dev = nullptr;
  unsigned int* ticket = nullptr;
  unsigned int* tile_counter = nullptr;

  bool err_posted = false;     // the last kernel of the call copies the index-error word to h_err itself

  // native training step (gqe_train.cu), created on first use
  gqe_train_state* train = nullptr;

  // staging for the *_host entry points (grow on demand)
  void* stage[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t stage_cap[6] = {0, 0, 0, 0, 0, 0};
};


// error return of an entry point: records the message (gqe_last_error) and hands the code back
int gqe_fail(gqe_ctx* c, int code, const char* fmt, ...);
void gqe_train_destroy(gqe_ctx* c);   // gqe_train.cu: frees the optimiser state (gqe_destroy)
// fetch (and clear) the device index-error word of a *_device call; GQE_OK or GQE_ERR_INDEX with the message set
int gqe_stage_reserve(gqe_ctx* c, int slot, size_t bytes);

#define GQE_CUDA(c, call)                                                                        \
  do {                                                                                           \
    cudaError_t e_ = (call);                                                                     \
    if (e_ != cudaSuccess)                                                                       \
      return gqe_fail((c), GQE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
  } while (0)
