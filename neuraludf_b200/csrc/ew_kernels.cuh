// Small element-wise kernels shared by the network files (each translation unit gets its own static copy).
#pragma once
#include "common.cuh"

namespace nudf {

static inline unsigned ew_blocks(int64_t n, int t) { return (unsigned)((n + t - 1) / t); }

// dst[:, col0 + c] = src[:, c] * scale   for c < ncols
static __global__ void ew_copy_cols_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd,
                                           int col0, int ncols, int64_t P, float scale) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int64_t row = idx / ncols;
  int c = (int)(idx - row * ncols);
  if (row >= P) return;
  dst[row * ldd + col0 + c] = src[row * lds + c] * scale;
}

// dst[row, col0 + :] = PE(src[row / spr, :d])  (models/embedder.py:11-36); spr = samples sharing one source row (>=1)
static __global__ void ew_pe_kernel(const float* __restrict__ src, int d, int L, int spr, int64_t P, float* __restrict__ dst,
                                    int64_t ldd, int col0, float* __restrict__ dst2, int64_t ldd2, int col02) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P) return;
  const float* x = src + (i / spr) * d;
  float* e = dst + i * ldd + col0;
  float* e2 = dst2 ? dst2 + i * ldd2 + col02 : nullptr;
  for (int c = 0; c < d; ++c) { e[c] = x[c]; if (e2) e2[c] = x[c]; }
  float f = 1.0f;
  for (int k = 0; k < L; ++k) {
    for (int c = 0; c < d; ++c) {
      float s, co;
      sincosf(x[c] * f, &s, &co);
      e[d * (1 + 2 * k) + c] = s; e[d * (2 + 2 * k) + c] = co;
      if (e2) { e2[d * (1 + 2 * k) + c] = s; e2[d * (2 + 2 * k) + c] = co; }
    }
    f *= 2.0f;
  }
}

}  // namespace nudf
