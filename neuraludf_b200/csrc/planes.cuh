// Split-bf16 "plane" tensors: storage format shared by the epilogues that write them (gemm_simt.cuh functors) and the
// tcgen05 kernels that fetch them with cp.async.bulk (gemm_pl.cuh, which documents the layout and its two operand roles).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace nudf {
namespace tc {

// byte offset of element (row, k) inside a [rows x 64] bf16 K-major SWIZZLE_128B tile (tile base 1024-aligned)
__host__ __device__ inline uint32_t sw128(uint32_t row, uint32_t k) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((((k >> 3) ^ (row & 7u)) & 7u) << 4) + ((k & 7u) << 1);
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
// split 4 consecutive values into NP bf16 planes (packed pairs)
template <int NP>
__device__ __forceinline__ void split4(const float x[4], uint2 planes[NP]) {
  float r[4] = {x[0], x[1], x[2], x[3]};
#pragma unroll
  for (int p = 0; p < NP; ++p) {
    float h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __bfloat162float(__float2bfloat16_rn(r[j]));
      r[j] -= h[j];
    }
    planes[p] = make_uint2(pack_bf16(h[0], h[1]), pack_bf16(h[2], h[3]));
  }
}

constexpr int PL_BLOCK = 64;                       // rows and cols of a block
constexpr int PL_PLANE_ELEMS = PL_BLOCK * PL_BLOCK;   // uint16 elements of one plane of one block
constexpr uint32_t PL_PLANE_BYTES = PL_PLANE_ELEMS * 2;

struct Planes {
  uint16_t* p;    // 1024-byte aligned
  int cb;         // number of 64-column blocks
};

__host__ __device__ inline int64_t planes_elems(int64_t rows, int cols) {
  return ((rows + 63) / 64) * (int64_t)((cols + 63) / 64) * 2 * PL_PLANE_ELEMS;
}
__host__ __device__ inline uint16_t* pl_block(const Planes& t, int64_t mb, int cb, int plane) {
  return t.p + ((mb * t.cb + cb) * 2 + plane) * (int64_t)PL_PLANE_ELEMS;
}
// 4 consecutive columns (col % 4 == 0) of one row
__device__ __forceinline__ void pl_store4(const Planes& t, int64_t row, int col, const float v[4]) {
  uint2 pl[2];
  split4<2>(v, pl);
  uint8_t* b = reinterpret_cast<uint8_t*>(pl_block(t, row >> 6, col >> 6, 0)) + sw128((uint32_t)(row & 63), (uint32_t)(col & 63));
  *reinterpret_cast<uint2*>(b) = pl[0];
  *reinterpret_cast<uint2*>(b + PL_PLANE_BYTES) = pl[1];
}
__device__ __forceinline__ void pl_load4(const Planes& t, int64_t row, int col, float v[4]) {
  const uint8_t* b = reinterpret_cast<const uint8_t*>(pl_block(t, row >> 6, col >> 6, 0)) + sw128((uint32_t)(row & 63), (uint32_t)(col & 63));
  const uint2 h = *reinterpret_cast<const uint2*>(b);
  const uint2 l = *reinterpret_cast<const uint2*>(b + PL_PLANE_BYTES);
  v[0] = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
  v[1] = __uint_as_float(h.x & 0xFFFF0000u) + __uint_as_float(l.x & 0xFFFF0000u);
  v[2] = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
  v[3] = __uint_as_float(h.y & 0xFFFF0000u) + __uint_as_float(l.y & 0xFFFF0000u);
}

// one element
__device__ __forceinline__ void pl_store1(const Planes& t, int64_t row, int col, float v) {
  const float h = __bfloat162float(__float2bfloat16_rn(v));
  uint8_t* b = reinterpret_cast<uint8_t*>(pl_block(t, row >> 6, col >> 6, 0)) + sw128((uint32_t)(row & 63), (uint32_t)(col & 63));
  *reinterpret_cast<uint16_t*>(b) = __bfloat16_as_ushort(__float2bfloat16_rn(v));
  *reinterpret_cast<uint16_t*>(b + PL_PLANE_BYTES) = __bfloat16_as_ushort(__float2bfloat16_rn(v - h));
}
// nv of 4 consecutive columns (col % 4 == 0) valid: the others are left untouched
__device__ __forceinline__ void pl_store(const Planes& t, int64_t row, int col, int nv, const float v[4]) {
  if (nv >= 4) { pl_store4(t, row, col, v); return; }
  for (int j = 0; j < nv; ++j) pl_store1(t, row, col + j, v[j]);
}

}  // namespace tc
}  // namespace nudf
