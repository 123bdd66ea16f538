// Engine dispatch: every dense contraction of the hot path goes through gemm_nt / gemm_nn / gemm_tn.
//   engine 0: exact-fp32 FFMA kernel (gemm_simt.cuh) -- always used for small / odd shapes;
//   engine 1: tcgen05 3xBF16-split tensor-core kernels (gemm_tc.cuh) when the caller supplies the pre-split weight
//             image built at fold time (gemm_nt / gemm_nn) or for the weight-gradient contraction (gemm_tn).
// Which chains may use the tensor engine is a bit mask (NUDF_TC_MASK, see tc_mask()): the forward value chain needs
// fp32-grade accuracy (udf feeds exp(-25000 u) and sigmoid(400 u)), the other chains tolerate the 3xBF16 split.
#pragma once
#include <stdlib.h>
#include "gemm_tc.cuh"
#include "gemm_pl.cuh"
#include "epi_planes.cuh"

namespace nudf {

// TC_FWD: UDF value chain (fused exact fp16-slice kernel); TC_REV/TAN/BWD: UDF gradient, tangent and backward chains; TC_WGRAD:
// weight gradients; TC_COLOR / TC_NERF: BACKWARD data GEMMs of the ReLU networks (2 planes); TC_RELU_FWD: their forward passes,
// with 3 planes / 6 products (gemm_w<3>, per-K-slice accumulators summed in fp32): a 4e-6 perturbation of a pre-activation (the
// 2-plane split) flips ~60x more ReLU gates than the reference's own fp32 rounding does, which shows up as O(1/batch) jumps in
// the parameter gradients and fails test_color_network_vs_reference_and_grads; the 3-plane product (3.5e-7 per layer) passes.
enum TcChain { TC_FWD = 1, TC_REV = 2, TC_TAN = 4, TC_BWD = 8, TC_WGRAD = 16, TC_COLOR = 32, TC_NERF = 64, TC_RELU_FWD = 128 };

int get_engine();
int tc_mask();
static inline bool tc_on(int chain) { return get_engine() == 1 && (tc_mask() & chain) != 0; }

template <class Epi>
static inline int gemm_nt(const float* A, int64_t lda, const float* W, int64_t ldw, int64_t M, int N, int64_t K,
                          const Epi& epi, cudaStream_t st, const uint16_t* img = nullptr, int chain = 0, int planes = 2) {
  if (img != nullptr && tc_on(chain) && K >= 32 && N >= 16)
    return planes == 3 ? tc::gemm_w<3>(A, lda, M, N, (int)K, img, epi, st) : tc::gemm_w<2>(A, lda, M, N, (int)K, img, epi, st);
  return gemm_simt<true, true, Epi>(A, lda, W, ldw, M, N, K, epi, st, 1);
}
template <class Epi>
static inline int gemm_nn(const float* A, int64_t lda, const float* W, int64_t ldw, int64_t M, int N, int64_t K,
                          const Epi& epi, cudaStream_t st, const uint16_t* img = nullptr, int chain = 0) {
  if (img != nullptr && tc_on(chain) && K >= 32 && N >= 16)
    return tc::gemm_w<2>(A, lda, M, N, (int)K, img, epi, st);
  return gemm_simt<true, false, Epi>(A, lda, W, ldw, M, N, K, epi, st, 1);
}
// C[M x N] += A[K x M]^T B[K x N]   (contraction over points)
int colsum(const float* X, int64_t ldx, const float* w, float wscale, int64_t P, int N, float* out, cudaStream_t st, bool t128 = false);

// ---- plane-fed chains (gemm_pl.cuh) -------------------------------------------------------------------------------
// Optional mode (nudf_set_chain_planes(1) / NUDF_PLANES=1): the reverse-sweep and tangent chains of the UDF network carry
// D[l] / Adot[l] between kernels as split-bf16 plane tensors -- operands fetched by cp.async.bulk, weight gradients on the
// MN-major plane kernel.  Parity-tested; off by default in round 1 because the plane-side epilogues (8-byte plane loads /
// stores under a 112-register budget) are slower than the fp32 ones by about what the operand path and the weight
// gradients gain (C2 step 7.57 ms vs 7.52 ms); DESIGN.md 5.
int chain_planes_flag();
static inline bool chain_planes_on() { return chain_planes_flag() == 1 && tc_on(TC_REV) && tc_on(TC_TAN) && tc_on(TC_WGRAD); }
static inline int64_t plane_rows(int64_t P) { return round_up(P, 128); }     // gemm_wrp reads whole 128-row tiles
template <class Epi>
static inline int gemm_tn_planes(const tc::Planes& X, int M, const tc::Planes& Y, int N, int64_t P, const Epi& epi, cudaStream_t st) {
  const int tiles = (int)(cdiv(M, 128) * cdiv(N, 256));
  int splits = tc::sm_count() / tiles;
  if (splits < 1) splits = 1;
  return tc::gemm_tn_pl(X, M, Y, N, P, epi, st, splits);
}
// dst planes[:, col0 + c] = src[:, c] * scale  for c < ncols (skip-concatenated columns of the tangent chain)
static __global__ void copy_cols_planes_kernel(const float* __restrict__ src, int64_t lds, tc::Planes dst, int col0, int ncols, int64_t P,
                                               float scale) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t row = idx / ncols;
  int c = (int)(idx - row * ncols);
  if (row >= P) return;
  tc::pl_store1(dst, row, col0 + c, src[row * lds + c] * scale);
}

// colsum_a (optional): colsum_a[m] += sum_k A[k, m] -- the bias gradient that goes with a weight gradient; fused into the
// tensor-engine kernel's operand staging, a separate reduction kernel on the FFMA path.
// t128: both operands are stored in the T128 layout (common.cuh) -- tensor-engine kernel only.
template <class Epi>
static inline int gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, int M, int N, int64_t K,
                          const Epi& epi, cudaStream_t st, int split_k, int chain = TC_WGRAD, float* colsum_a = nullptr,
                          bool t128 = false) {
  if (t128) {
    const tc::TnPair pr{A, lda, B, ldb, colsum_a};
    return tc::gemm_tn2(&pr, 1, M, N, K, epi, st);
  }
  if (tc_on(chain) && M >= 32 && N >= 32 && K >= 1024 && (lda & 3) == 0 && (ldb & 3) == 0 && (reinterpret_cast<uintptr_t>(A) & 15u) == 0 && (reinterpret_cast<uintptr_t>(B) & 15u) == 0) {
    const tc::TnPair pr{A, lda, B, ldb, colsum_a};        // row-major operands through the cp.async-fed kernel
    return tc::gemm_tn2(&pr, 1, M, N, K, epi, st, true);
  }
  if (tc_on(chain) && M >= 32 && N >= 32 && K >= 128) {
    // split the points so that (M tiles x N tiles x splits) fills the SMs once, with at least 8 slices of 64 points per CTA
    const int tiles = (int)(cdiv(M, 128) * cdiv(N, 256));
    int splits = tc::sm_count() / tiles;
    const int max_splits = (int)cdiv(K, 512);
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    return tc::gemm_tn(A, lda, B, ldb, M, N, K, epi, st, splits, colsum_a, t128);
  }
  if (int rc = gemm_simt<false, false, Epi>(A, lda, B, ldb, M, N, K, epi, st, split_k)) return rc;
  if (colsum_a != nullptr) return colsum(A, lda, nullptr, 1.f, K, M, colsum_a, st);
  return 0;
}

// T128 operands (fused UDF chains): up to two (A, B) pairs accumulate into the same C in one launch (tc::gemm_tn2)
template <class Epi>
static inline int gemm_tn_pairs(const tc::TnPair* pairs, int n_pairs, int M, int N, int64_t K, const Epi& epi, cudaStream_t st) {
  return tc::gemm_tn2(pairs, n_pairs, M, N, K, epi, st);
}

}  // namespace nudf
