// Engine dispatch: every dense contraction of the hot path goes through gemm_nt / gemm_nn / gemm_tn.
//   engine 0: exact-fp32 FFMA kernel (gemm_simt.cuh) -- always used for small / odd shapes;
//   engine 1: tcgen05 3xBF16-split tensor-core kernel (gemm_tc.cuh) for the wide layers, when the caller
//             supplies the pre-split weight image (TcW) produced at fold time.
#pragma once
#include "gemm_simt.cuh"

namespace nudf {

// Pre-split bf16 (hi, lo) weight images in UMMA shared-memory order (see gemm_tc.cuh); null pointers = absent.
struct TcW {
  const uint16_t* nt_img;  // operand for X * W^T   (K = in  contiguous)
  const uint16_t* nn_img;  // operand for dY * W    (K = out contiguous, i.e. W^T image)
  int n_pad_nt, k_pad_nt, n_pad_nn, k_pad_nn;
};

int get_engine();

template <class Epi>
static inline int gemm_nt(const float* A, int64_t lda, const float* W, int64_t ldw, int64_t M, int N, int64_t K,
                          const Epi& epi, cudaStream_t st, const TcW* tw = nullptr) {
  (void)tw;
  return gemm_simt<true, true, Epi>(A, lda, W, ldw, M, N, K, epi, st, 1);
}
template <class Epi>
static inline int gemm_nn(const float* A, int64_t lda, const float* W, int64_t ldw, int64_t M, int N, int64_t K,
                          const Epi& epi, cudaStream_t st, const TcW* tw = nullptr) {
  (void)tw;
  return gemm_simt<true, false, Epi>(A, lda, W, ldw, M, N, K, epi, st, 1);
}
// C[M x N] += A[K x M]^T B[K x N]   (contraction over points)
template <class Epi>
static inline int gemm_tn(const float* A, int64_t lda, const float* B, int64_t ldb, int M, int N, int64_t K,
                          const Epi& epi, cudaStream_t st, int split_k) {
  return gemm_simt<false, false, Epi>(A, lda, B, ldb, M, N, K, epi, st, split_k);
}

int colsum(const float* X, int64_t ldx, const float* w, float wscale, int64_t P, int N, float* out, cudaStream_t st);

}  // namespace nudf
