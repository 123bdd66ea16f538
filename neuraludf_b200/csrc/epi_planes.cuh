// Epilogue functors of the plane-fed layer chains (gemm_pl.cuh): same maths as EpiRev / EpiTan (gemm_simt.cuh) with the
// chain-carried tensors D[l] and Adot[l] stored ONLY as split-bf16 plane tensors -- written here, fetched by the next chain
// kernel and by the plane-fed weight gradient with cp.async.bulk.
#pragma once
#include "gemm_simt.cuh"
#include "planes.cuh"

namespace nudf {

// Reverse sweep: acc = G = d udf / d A[l]  ->  D[l-1] = G * s * sigma(100 z[l-1]) (planes; columns >= n_main zero-filled up
// to the plane width: they are the K padding of the next GEMM); skip-concatenated columns are routed to Gpe (fp32).
struct EpiRevP {
  int n_main; float post_scale;
  const float* Anext; int64_t lda; float a_unscale;
  tc::Planes dpl;
  float* Gpe; int64_t ldg;
  struct Aux { float a[4]; };
  __device__ __forceinline__ void load(int64_t row, int col, int nv, Aux& x) const {
    int n = n_main - col;
    n = n < nv ? n : nv;
    if (n > 0) ld4(Anext, lda, row, col, n, x.a);
    else { x.a[0] = x.a[1] = x.a[2] = x.a[3] = 0.f; }
  }
  __device__ __forceinline__ void apply(int64_t row, int col, const float acc[4], int nv, const Aux& x) const {
    float d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col + j;
      const float g = acc[j] * post_scale;
      d[j] = (j < nv && c < n_main) ? g * sig_from_softplus(x.a[j] * a_unscale) : 0.f;
      if (j < nv && c >= n_main && Gpe != nullptr) Gpe[row * ldg + (c - n_main)] = g;
    }
    if (col < dpl.cb * 64) tc::pl_store4(dpl, row, col, d);
  }
  NUDF_EPI_CALL
};

// Tangent chain: acc = Zdot_l.  Q_l = Zdot * D_l * 100 (1 - S_l) (fp32, consumed by the backward chain's epilogue);
// Adot_{l+1} = S_l * Zdot * post_scale (planes; optionally also fp32 for the last layer's weighted column sum).
struct EpiTanP {
  const float* Anext; int64_t lda; float a_unscale;
  tc::Planes D;
  float* Q; int64_t ldq;
  tc::Planes npl; float post_scale;
  float* AdotNext; int64_t ldn;          // may be null
  struct Aux { float a[4], d[4]; };
  __device__ __forceinline__ void load(int64_t row, int col, int nv, Aux& x) const {
    ld4(Anext, lda, row, col, nv, x.a);
    tc::pl_load4(D, row, col, x.d);
  }
  __device__ __forceinline__ void apply(int64_t row, int col, const float acc[4], int nv, const Aux& x) const {
    float q[4], n[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float s = sig_from_softplus(x.a[j] * a_unscale);
      q[j] = acc[j] * x.d[j] * (100.0f * (1.0f - s));
      n[j] = j < nv ? s * acc[j] * post_scale : 0.f;
    }
    st4(Q, ldq, row, col, nv, q);
    tc::pl_store(npl, row, col, nv, n);
    if (AdotNext != nullptr) st4(AdotNext, ldn, row, col, nv, n);
  }
  NUDF_EPI_CALL
};

template <> struct epi_family<EpiRevP> { static constexpr int value = FAM_TC_REV; };
template <> struct epi_family<EpiTanP> { static constexpr int value = FAM_TC_TAN; };

}  // namespace nudf
