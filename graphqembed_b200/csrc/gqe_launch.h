// gqe_launch.h -- host-callable launchers of the kernels, one translation unit
// per embedding dimension behind them.
#pragma once
#include "gqe_params.h"

namespace gqe {

#define GQE_DECLARE_DIM(D)                                                                                      \
  cudaError_t launch_fused_simt_d##D(int structure, const LaunchParams& lp, int64_t grid, cudaStream_t st);     \
  cudaError_t launch_op_simt_d##D(const OpParams& op, cudaStream_t st);
GQE_DECLARE_DIM(32)
GQE_DECLARE_DIM(64)
GQE_DECLARE_DIM(128)
GQE_DECLARE_DIM(256)
#undef GQE_DECLARE_DIM

// tensor-core (tcgen05) kernels: Bilinear decoder, d = 128 / 256
#define GQE_DECLARE_TC_DIM(D)                                                                                \
  cudaError_t launch_fused_tc_d##D(int structure, const LaunchParams& lp, int64_t grid, cudaStream_t st);    \
  cudaError_t launch_pack_d##D(const PackParams& pp, int n_entries, cudaStream_t st);
GQE_DECLARE_TC_DIM(128)
GQE_DECLARE_TC_DIM(256)
#undef GQE_DECLARE_TC_DIM

// raw row gather (gqe_rows.cu): out[i] = table[rows[i]], any d % 4 == 0
cudaError_t launch_gather_rows(const float* table, const int32_t* rows, int64_t n, int d, float* out, int64_t table_rows,
                               unsigned long long* err, cudaStream_t st);

// index arrays of a host call: mapped pinned host memory -> device staging, all ranges in ONE launch (gqe_rows.cu)
cudaError_t launch_fetch_indices(const FetchParams& fp, cudaStream_t st);

// batches out of a device-resident query store (gqe_rows.cu)
cudaError_t launch_store_batch(const StoreBatchParams& sp, int64_t max_n, cudaStream_t st);

// (query, target) pair scoring against stored query embeddings (gqe_pairs.cu)
cudaError_t launch_score_pairs(int d, const PairParams& pp, int64_t n_pairs_total, cudaStream_t st);

// operator-level backward kernels (gqe_bwd.cu), any supported d
cudaError_t launch_matmul_wgrad(int d, int64_t n, const float* gy, const float* x, int form, float* gw, cudaStream_t st);
cudaError_t launch_rowsum(int d, int64_t n, const float* gy, const float* x, float* gv, cudaStream_t st);
cudaError_t launch_aggregate(int64_t total, const float* e1, const float* e2, const float* e3, int relu, int use_min,
                             float* out, cudaStream_t st);
cudaError_t launch_aggregate_bwd(int64_t total, const float* e1, const float* e2, const float* e3, int relu, int use_min,
                                 const float* gout, float* g1, float* g2, float* g3, cudaStream_t st);
cudaError_t launch_cosine_bwd(int d, int64_t n, const float* x, const float* y, const float* gout, int raw_dot, float* gx,
                              float* gy, cudaStream_t st);
cudaError_t launch_dot(int d, int64_t n, const float* x, const float* y, float* out, cudaStream_t st);
cudaError_t launch_encode_bwd(int d, int64_t n, const float* table, const int32_t* rows, const float* gout, float* gtable,
                              int64_t table_rows, unsigned long long* err, cudaStream_t st);

// sparse training step (gqe_opt.cu): per-row encoder gradients and the row-wise Adam
cudaError_t launch_encode_bwd_rows(int d, int64_t n, const float* table, const int32_t* rows, const float* gout, float* out,
                                   int64_t table_rows, unsigned long long* err, cudaStream_t st);
cudaError_t launch_adam_rows(float* table, float* m, float* v, int32_t* last, int64_t table_rows, int d, int64_t n,
                             const int64_t* rows, const float* grads, int step, float lr, float beta1, float beta2,
                             float eps, cudaStream_t st);

// native training step (gqe_train.cu): row-wise Adam over a dense accumulated-gradient buffer (rows may
// repeat; gsum == nullptr: catch the rows up to `step`), and Adam on a small dense parameter
cudaError_t launch_adam_rows_accum(float* table, float* m, float* v, int32_t* last, int64_t table_rows, int d, int64_t n,
                                   const int32_t* rows, float* gsum, int step, float lr, float beta1, float beta2, float eps,
                                   cudaStream_t st);
cudaError_t launch_adam_dense(float* p, float* m, float* v, const float* g, int64_t n, int step, float lr, float beta1,
                              float beta2, float eps, cudaStream_t st);

// GraphSAGE-style encoder steps (gqe_sage.cu)
cudaError_t launch_segment_mean(const float* src, int64_t n_src, int d, int64_t n, const int64_t* ptr, const int32_t* cols,
                                float* out, unsigned long long* err, cudaStream_t st);
cudaError_t launch_linear(const float* w, int m, int k, int64_t n, const float* x, int relu, float* out, cudaStream_t st);

// streaming kernel of the contraction-free decoders (gqe_vec.cu): TransE / DistMult chains and
// element-wise intersections, any supported d, regular layout with T <= 2
cudaError_t launch_fused_vec(int d, const LaunchParams& lp, cudaStream_t st);

// products of operator matrices (gqe_compose.cu), d = 128 / 256
cudaError_t launch_compose(int d, const ComposeParams& cp, int n_entries, cudaStream_t st);
int score_col_src_host(int n);   // tc::score_col_src (gqe_tc.cuh) for the host side

inline bool tc_dim_supported(int d) { return d == 128 || d == 256; }
// staging area of the STAGE instantiation of the fused kernel (d = 256, grouped kernel): bytes per CTA
inline size_t tc_stage_bytes_per_cta(int d) { return (size_t)4 * 5 * 128 * d * sizeof(float); }
// CTA pairs sharing the weight stream (gqe_tc.cuh, PAIR): an experiment, on only with GQE_PAIR=1 in the environment
int tc_use_pair(int d, int64_t tiles);
inline size_t tc_packed_bytes(int d) { return (size_t)4 * d * d; }
inline cudaError_t launch_fused_tc(int d, int structure, const LaunchParams& lp, int64_t grid, cudaStream_t st) {
  switch (d) {
    case 128: return launch_fused_tc_d128(structure, lp, grid, st);
    case 256: return launch_fused_tc_d256(structure, lp, grid, st);
    default: return cudaErrorInvalidValue;
  }
}
inline cudaError_t launch_pack(int d, const PackParams& pp, int n_entries, cudaStream_t st) {
  switch (d) {
    case 128: return launch_pack_d128(pp, n_entries, st);
    case 256: return launch_pack_d256(pp, n_entries, st);
    default: return cudaErrorInvalidValue;
  }
}

// structure < 0 selects the grouped kernel
inline cudaError_t launch_fused_simt(int d, int structure, const LaunchParams& lp, int64_t grid, cudaStream_t st) {
  switch (d) {
    case 32: return launch_fused_simt_d32(structure, lp, grid, st);
    case 64: return launch_fused_simt_d64(structure, lp, grid, st);
    case 128: return launch_fused_simt_d128(structure, lp, grid, st);
    case 256: return launch_fused_simt_d256(structure, lp, grid, st);
    default: return cudaErrorInvalidValue;
  }
}
inline cudaError_t launch_op_simt(int d, const OpParams& op, cudaStream_t st) {
  switch (d) {
    case 32: return launch_op_simt_d32(op, st);
    case 64: return launch_op_simt_d64(op, st);
    case 128: return launch_op_simt_d128(op, st);
    case 256: return launch_op_simt_d256(op, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace gqe
