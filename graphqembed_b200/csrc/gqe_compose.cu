// gqe_compose.cu -- products of the (tiny) operator matrices of a formula.
//
// Consecutive linear operators of a query structure (relation projections, DeepSets pre /
// post; reference netquery/decoders.py:145,150,289,299) are multiplied together ONCE per
// call so that the fused tensor-core kernel runs one contraction per run instead of one per
// operator.  A product is d x d x d (d = 128 / 256): far too small for the tcgen05 pipeline of
// the fused kernel to pay off, and latency-bound on the CUDA cores (17 us per 64x64 tile with
// FFMA), so it runs on the warp-level tensor-core path (wmma, bf16 operands, fp32 accumulate)
// with the same hi/lo split as the fused kernel: c = a_hi b_hi + a_lo b_hi + a_hi b_lo,
// ~2^-17 relative (NOT exact fp32).  One CTA = one 64x64 tile of one product; the whole K extent
// of both operands is fetched in one round of loads; a launch computes up to 96 products.  A
// three-factor run is two stream-ordered launches (the inner product first).  The results are
// packed into the context's weight cache, so this kernel only runs when a parameter changed.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <mma.h>
#include <stdint.h>

#include "gqe_launch.h"

namespace gqe {

// tile: kComposeRows rows (of a) x 64 columns (of b)
template <int D, int kComposeRows>
struct ComposeSmem {
  __nv_bfloat16 a_hi[kComposeRows][D + 8], a_lo[kComposeRows][D + 8];   // a[i0 .. i0+128)[0 .. D)
  __nv_bfloat16 b_hi[D][64 + 8], b_lo[D][64 + 8];                       // b[0 .. D)[j0 .. j0+64)
};

__device__ __forceinline__ void split_store4(__nv_bfloat16* hi, __nv_bfloat16* lo, const float4 v) {
  const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
  const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - __low2float(h0), v.y - __high2float(h0));
  const __nv_bfloat162 l1 = __floats2bfloat162_rn(v.z - __low2float(h1), v.w - __high2float(h1));
  reinterpret_cast<__nv_bfloat162*>(hi)[0] = h0;
  reinterpret_cast<__nv_bfloat162*>(hi)[1] = h1;
  reinterpret_cast<__nv_bfloat162*>(lo)[0] = l0;
  reinterpret_cast<__nv_bfloat162*>(lo)[1] = l1;
}

constexpr int kComposeThreads = 512;

template <int D, int kComposeRows>
__global__ void __launch_bounds__(kComposeThreads, 1) gqe_compose(const __grid_constant__ ComposeParams p) {
  using namespace nvcuda;
  extern __shared__ __align__(32) unsigned char smem_raw[];
  ComposeSmem<D, kComposeRows>& sm = *reinterpret_cast<ComposeSmem<D, kComposeRows>*>(smem_raw);
  constexpr int NH = kComposeRows / 64;
  const ComposeEntry& e = p.e[blockIdx.z];
  const int i0 = blockIdx.y * kComposeRows, j0 = blockIdx.x * 64;
  // Every load of an operand tile is issued before its first shared store (N float4 per thread
  // in registers), so the CTA pays ~one memory latency per operand.
  constexpr int NA = kComposeRows * D / 4 / kComposeThreads;   // float4 per thread of a (16 at d = 256)
  constexpr int N = 64 * D / 4 / kComposeThreads;              // float4 per thread of b (8 at d = 256)
  float4 va[NA], vb[N];
#pragma unroll
  for (int u = 0; u < NA; ++u) {
    const int idx = threadIdx.x + u * kComposeThreads;
    va[u] = __ldcg(reinterpret_cast<const float4*>(e.a + (size_t)(i0 + idx / (D / 4)) * D) + idx % (D / 4));
  }
#pragma unroll
  for (int u = 0; u < N; ++u) {
    const int idx = threadIdx.x + u * kComposeThreads;
    vb[u] = __ldcg(reinterpret_cast<const float4*>(e.b + (size_t)(idx / 16) * D + j0) + idx % 16);
  }
#pragma unroll
  for (int u = 0; u < NA; ++u) {
    const int idx = threadIdx.x + u * kComposeThreads;
    split_store4(&sm.a_hi[idx / (D / 4)][4 * (idx % (D / 4))], &sm.a_lo[idx / (D / 4)][4 * (idx % (D / 4))], va[u]);
  }
#pragma unroll
  for (int u = 0; u < N; ++u) {
    const int idx = threadIdx.x + u * kComposeThreads;
    split_store4(&sm.b_hi[idx / 16][4 * (idx % 16)], &sm.b_lo[idx / 16][4 * (idx % 16)], vb[u]);
  }
  __syncthreads();
  // warp w (16 warps): rows 16 (w / 4) .. +16 and 64 + 16 (w / 4) .. +16, columns 16 (w % 4) .. +16
  const int w = threadIdx.x >> 5;
  const int r0 = 16 * (w >> 2), c0 = 16 * (w & 3);
  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[NH];
#pragma unroll
  for (int h = 0; h < NH; ++h) wmma::fill_fragment(acc[h], 0.f);
#pragma unroll 4
  for (int k = 0; k < D; k += 16) {
    wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> bh, bl;
    wmma::load_matrix_sync(bh, &sm.b_hi[k][c0], 64 + 8);
    wmma::load_matrix_sync(bl, &sm.b_lo[k][c0], 64 + 8);
#pragma unroll
    for (int h = 0; h < NH; ++h) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::row_major> ah, al;
      wmma::load_matrix_sync(ah, &sm.a_hi[r0 + 64 * h][k], D + 8);
      wmma::load_matrix_sync(al, &sm.a_lo[r0 + 64 * h][k], D + 8);
      wmma::mma_sync(acc[h], ah, bh, acc[h]);
      wmma::mma_sync(acc[h], al, bh, acc[h]);
      wmma::mma_sync(acc[h], ah, bl, acc[h]);
    }
  }
#pragma unroll
  for (int h = 0; h < NH; ++h)
    wmma::store_matrix_sync(e.dst + (size_t)(i0 + r0 + 64 * h) * D + j0 + c0, acc[h], D, wmma::mem_row_major);
}

template <int D, int ROWS>
static cudaError_t launch_compose_t(const ComposeParams& cp, int n_entries, cudaStream_t st) {
  static bool configured[64] = {false};
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 63;
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(gqe_compose<D, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(ComposeSmem<D, ROWS>));
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  const dim3 grid(D / 64, D / ROWS, (unsigned)n_entries);
  gqe_compose<D, ROWS><<<grid, kComposeThreads, sizeof(ComposeSmem<D, ROWS>), st>>>(cp);
  return cudaGetLastError();
}

// 64-row tiles: 160 CTAs for the 10 products of the benchmark mix.  128-row tiles (80 CTAs, one
// wave) were measured 8 us SLOWER per call: the time of a CTA is its load -> split -> smem -> MMA
// chain, which grows with the tile.
cudaError_t launch_compose(int d, const ComposeParams& cp, int n_entries, cudaStream_t st) {
  if (n_entries <= 0) return cudaSuccess;
  switch (d) {
    case 128: return launch_compose_t<128, 64>(cp, n_entries, st);
    case 256: return launch_compose_t<256, 64>(cp, n_entries, st);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace gqe
