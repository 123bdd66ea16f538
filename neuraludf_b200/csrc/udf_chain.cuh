// Fused UDF value chain for sm_100a: ONE persistent kernel walks every layer of UDFNetwork.forward (reference
// models/fields.py:192-211) for a 128-point tile; the activations never leave the SM between layers.
//
//   pts --PE--> A_0 --[W_0]--> softplus --> A_1 --[W_1]--> ... --[W_last]--> y          (one CTA per SM, tiles of 128 points)
//
// * every layer is a tcgen05.mma contraction (M = 128 points, N <= 128 per output tile, K = 16 per instruction, fp32
//   accumulators in TMEM); the epilogue warps read the accumulator with tcgen05.ld, apply bias + softplus(beta = 100), and
//   write the NEXT layer's A operand straight into shared memory in the UMMA K-major SWIZZLE_128B layout -- no HBM round trip
//   and no LSU operand path.  Weight slices stream from L2 through a cp.async.bulk (TMA engine) ring.
// * fp32-grade accuracy on the tensor engine (the udf head feeds exp(-25000 u): SURVEY.md section 0, fact 3) comes from an
//   EXACT-MAIN fp16 slice scheme instead of the truncating 3xBF16 split:
//       activation row r :  a = 2^(ea_r - 11) (a0 + a1),  a0 = rint(a 2^(11 - ea_r)) in [-2048, 2048] (12-bit integer, exact in
//                           fp16), a1 = fp16(remainder) in [-1/2, 1/2];   2^ea_r > max_k |a_rk|  (per-row power of two)
//       weight row n     :  w = 2^(ew_n - 13) (w0 + w1 + w2),  w0 = 1024 rint(w 2^(3 - ew_n)) (4-bit integer x 2^10),
//                           w1 = fp16(1024 remainder), w2 = fp16(second remainder)
//   Main accumulator  M = sum_k a0 w0: every product is an integer multiple of 2^10 below 2^24 and the 256-term sum stays below
//   2^22 units, so the tensor core's truncating fp32 accumulation is EXACT.  Correction accumulator
//   C = sum_k a0 w1 + a0 w2 + a1 w0 + a1 w1 is ~2^-3 of M, so its truncation error is ~2^-27 of the result.  z = 2^(ea_r +
//   ew_n - 24) (M + C) + b in fp32.  Dropped terms: a1 w2 (2^-24).  5 tensor products per K step; 2 fp16 planes of A (128 KB of
//   shared memory for a [128 x 256] tile) + a 2-slot ring of 3-plane weight slices (2 x 48 KB).
// * roles: warps 0-7 epilogue (warp w: TMEM lane quadrant w & 3 = tile rows 32 (w & 3).., output-column half w >> 2 = N tile
//   w >> 2), warp 8 MMA issuer + TMEM owner, warp 9 weight-ring loader.  The two N tiles of a layer have their own
//   accumulator pairs (M0 C0 M1 C1 = 512 TMEM columns), so the epilogue of tile 0 overlaps the MMAs of tile 1.
//   Per-row scales need the row maximum over all 256 columns: pass 1 (activation, partial row max, value written back to
//   TMEM over M), exchange through shared memory, pass 2 (slicing into the A planes).
#pragma once
#include <cuda_fp16.h>
#include <stdlib.h>

#include "gemm_tc.cuh"

namespace nudf {
namespace chain {

using namespace tc;

constexpr int CH_NT = 128;                                  // output columns per N tile
constexpr int CH_MAX_SLICES = 4;                            // K <= 256
constexpr uint32_t CH_APLANE = 128u * 128u;                 // bytes of one plane of one 64-wide K slice of A (128 rows x 128 B)
constexpr uint32_t CH_A_BYTES = CH_MAX_SLICES * 2u * CH_APLANE;     // 128 KB
constexpr int CH_WPL = 3;                                   // weight planes
constexpr uint32_t CH_WSLOT = CH_WPL * CH_NT * 128u;        // 48 KB: one K slice of one N tile, 3 planes
constexpr int CH_NSLOT = 2;
constexpr int CH_EPI_WARPS = 16;
constexpr int CH_EPI_THREADS = CH_EPI_WARPS * 32;           // 512
constexpr int CH_MMA_WARP = 16, CH_LOAD_WARP = 17;
constexpr int CH_THREADS = 576;
constexpr int CH_MAX_STAGES = 112;
constexpr int CH_MAX_PE = 64;                               // positional-encoding width (K of layer 0) <= one K slice

struct ChainLayer {
  int K, N;                  // logical contraction / output widths
  int n_kslices, n_tiles;
  int stage0;                // first weight stage of this layer in ChainParams::S
  int last;                  // 1: last layer (plain output, no activation)
  int pe_next;               // PE columns appended to the next layer's input (skip layer), else 0
  float post_scale;          // 1/sqrt(2) when the next layer is the skip layer
  const float* bias;         // [128 n_tiles], zero for padded columns
  const float* wscale;       // DEVICE scalar 2^(E_l - 13): weights of this layer are 2^(E_l - 13) (w0 + w1 + w2)
  float* out;                // hidden: A[l+1] [P, ld_out] or null; last: Y [P, ld_out] or null
  int64_t ld_out;
};
struct ChainStage {
  uint32_t src;              // uint16-element offset of plane 0 of this (layer, N tile, K slice) in the chain image
  uint16_t rows;             // rows to fetch / N of the MMA (multiple of 16)
  uint16_t plane_rows;       // rows of the full tile in the image (plane stride = plane_rows * 64 elements)
};
struct ChainParams {
  int n_layers, n_stages;
  ChainLayer L[NUDF_MAX_LAYERS];
  ChainStage S[CH_MAX_STAGES];
  const uint16_t* img;
  const float* pts; int64_t P; float scale; int n_freq, d_pe;
  float* e0; int pe_ld;                  // PE(x) [P, pe_ld] or null
  float* udf_out; float inv_scale;       // value-only mode: udf_out[P] = |y_0| / scale
  long long* trace;                      // profiling aid (NUDF_CHAIN_TRACE=1): clock64() stamps of CTA 0, first point tile
};
// trace layout: [role 0 = epilogue warp 0, 1 = epilogue warp 12, 2 = MMA issuer][layer][8]
constexpr int CH_TRACE_WORDS = 4 * NUDF_MAX_LAYERS * 8;
#define CH_TR(role, l, k, v) do { if (tr_on) p.trace[((role) * NUDF_MAX_LAYERS + (l)) * 8 + (k)] = (v); } while (0)

__host__ __device__ inline int ch_tile_rows(int N, int t) { int r = N - CH_NT * t; return pad16(r < CH_NT ? r : CH_NT); }
__host__ __device__ inline int ch_n_tiles(int N) { return (N + CH_NT - 1) / CH_NT; }
// uint16 elements of the chain image of one layer: [tile][k slice][plane][rows_t x 64]
__host__ __device__ inline int64_t ch_layer_elems(int N, int K) {
  int64_t e = 0;
  for (int t = 0; t < ch_n_tiles(N); ++t) e += (int64_t)ch_tile_rows(N, t) * pad64(K) * CH_WPL;
  return e;
}
__host__ __device__ inline int64_t ch_tile_off(int N, int K, int t) {
  int64_t e = 0;
  for (int i = 0; i < t; ++i) e += (int64_t)ch_tile_rows(N, i) * pad64(K) * CH_WPL;
  return e;
}

// ---- weight image ------------------------------------------------------------------------------------------------------
// meta[0] = 2^(3 - E), meta[1] = 2^(E - 13) with 2^E > max |W| over the whole layer (one power of two per LAYER: the integer
// slice w0 then has up to 4 bits for the largest weights and fewer for small rows, whose precision lives in the floating
// fp16 remainders w1, w2 -- 22 more bits relative to each element)
static __global__ void chain_layer_scale_kernel(const float* __restrict__ W, int64_t ldw, int N, int K, float* __restrict__ meta) {
  float mx = 0.f;
  for (int64_t i = threadIdx.x; i < (int64_t)N * K; i += blockDim.x) mx = fmaxf(mx, fabsf(W[(i / K) * ldw + (i % K)]));
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
    int e = 0;
    if (mx > 1e-30f && mx < 1e30f) e = (int)((__float_as_uint(mx) >> 23) & 0xffu) - 126;
    meta[0] = __uint_as_float((uint32_t)(3 - e + 127) << 23);
    meta[1] = __uint_as_float((uint32_t)(e - 13 + 127) << 23);
  }
}
// one block per padded output row; W row-major [N, ldw]
static __global__ void chain_prep_kernel(const float* __restrict__ W, int64_t ldw, const float* __restrict__ bias, int N, int K,
                                         const float* __restrict__ meta, uint16_t* __restrict__ img, float* __restrict__ bias_tab) {
  const int col = blockIdx.x;                    // padded output column: 128 t + local row
  const int t = col / CH_NT, nl = col - t * CH_NT;
  const int rows_t = ch_tile_rows(N, t);
  const bool valid = col < N;
  if (threadIdx.x == 0) bias_tab[col] = (valid && bias) ? bias[col] : 0.f;
  if (nl >= rows_t) return;                      // beyond the padded tile: only the bias table entry exists
  const int Kp = pad64(K);
  const float up = meta[0];                      // 2^(3 - E): scaled weights are in (-8, 8)
  uint16_t* base = img + ch_tile_off(N, K, t);
  for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
    float w = (valid && k < K) ? W[(int64_t)col * ldw + k] * up : 0.f;
    const float w0 = rintf(w);
    const float r1 = (w - w0) * 1024.0f;
    const __half h1 = __float2half_rn(r1);
    const __half h2 = __float2half_rn(r1 - __half2float(h1));
    const __half h0 = __float2half_rn(w0 * 1024.0f);
    const int s = k >> 6;
    const uint32_t off = sw128((uint32_t)nl, (uint32_t)(k & 63)) >> 1;
    uint16_t* b = base + (int64_t)s * CH_WPL * rows_t * 64 + off;
    b[0] = __half_as_ushort(h0);
    b[(int64_t)rows_t * 64] = __half_as_ushort(h1);
    b[(int64_t)2 * rows_t * 64] = __half_as_ushort(h2);
  }
}

// ---- device helpers ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t make_idesc_f16(uint32_t n) {     // kind::f16: D = F32, A = B = F16, K-major, M = 128
  return (1u << 4) | ((n >> 3) << 17) | ((128u >> 4) << 24);
}
// TMEM -> registers, 16 consecutive columns of this thread's lane (row); load and wait::ld in ONE asm statement so that no
// use of the destination registers can be scheduled between them
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// two loads (main and correction accumulator) in flight together, one wait
__device__ __forceinline__ void tmem_ld16x2(uint32_t ta, uint32_t tb, float v[16], float w[16]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  uint32_t* q = reinterpret_cast<uint32_t*>(w);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%32];\n"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%33];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]),
        "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]), "=r"(q[8]), "=r"(q[9]), "=r"(q[10]), "=r"(q[11]), "=r"(q[12]), "=r"(q[13]),
        "=r"(q[14]), "=r"(q[15])
      : "r"(ta), "r"(tb)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float v[16]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 512;" ::: "memory"); }

// softplus(beta = 100) with the MUFU base-2 primitives: max(z, 0) + ln2/100 * log2(1 + 2^(-|100 z| log2 e)); identical to
// common.cuh's softplus100 up to the rounding of the two constant folds (|diff| < 1e-9)
__device__ __forceinline__ float softplus100_fast(float z) {
  float t, l;
  const float az = fabsf(z) * -144.26950408889634f;          // -100 log2(e) |z|
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(az));
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + t));
  const float sp = fmaf(l, 0.0069314718055994531f, fmaxf(z, 0.0f));     // ln(2) / 100
  return 100.0f * z > 20.0f ? z : sp;
}
// biased exponent field of a non-negative float: the per-row scale only needs max over the row of this
__device__ __forceinline__ uint32_t expo_bits(float m) { return (__float_as_uint(m) >> 23) & 0xffu; }
// per-row power-of-two scale from the exponent field E of the row maximum: 2^e > max with e = E - 126
__device__ __forceinline__ void row_scale(uint32_t E, float& sa, float& inv) {
  int e = 0;
  if (E > 27u && E < 227u) e = (int)E - 126;
  sa = __uint_as_float((uint32_t)(e - 11 + 127) << 23);
  inv = __uint_as_float((uint32_t)(11 - e + 127) << 23);
}
// 8 consecutive activations -> 16 bytes of plane 0 (integer part) and plane 1 (remainder), fp16
__device__ __forceinline__ void slice8(const float* a, float inv, uint4& p0, uint4& p1) {
  uint32_t o0[4], o1[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const float t0 = a[2 * q] * inv, t1 = a[2 * q + 1] * inv;
    const float i0 = rintf(t0), i1 = rintf(t1);
    const __half2 h0 = __floats2half2_rn(i0, i1);
    const __half2 h1 = __floats2half2_rn(t0 - i0, t1 - i1);
    o0[q] = *reinterpret_cast<const uint32_t*>(&h0);
    o1[q] = *reinterpret_cast<const uint32_t*>(&h1);
  }
  p0 = make_uint4(o0[0], o0[1], o0[2], o0[3]);
  p1 = make_uint4(o1[0], o1[1], o1[2], o1[3]);
}

struct ChainCtl {
  uint64_t w_full[CH_NSLOT], w_empty[CH_NSLOT];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t a_ready[CH_MAX_SLICES];
  uint32_t tmem_addr;
  uint32_t pad_;
  uint8_t rowexp[2][4][128];    // [layer parity][column quarter][row]: exponent field of the partial row maximum
};
// 128 KB of A planes + 2 x 48 KB weight slots + control block = 225.1 KB of the 227 KB a CTA may have: no room for an alignment
// slack, so the dynamic shared-memory window itself is declared 1024-byte aligned (SWIZZLE_128B operands need it) and the
// kernel traps if the runtime did not honour that.
constexpr size_t CH_SMEM_BYTES = (size_t)CH_A_BYTES + (size_t)CH_NSLOT * CH_WSLOT + sizeof(ChainCtl);
static_assert(CH_SMEM_BYTES <= 227 * 1024, "fused chain: shared-memory budget exceeded");

__global__ void __launch_bounds__(CH_THREADS, 1) udf_chain_kernel(const __grid_constant__ ChainParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0u) __trap();
  uint8_t* a_smem = smem;                                   // [k slice][plane][128 rows x 128 B]
  uint8_t* w_smem = smem + CH_A_BYTES;                      // [slot][plane][rows x 128 B]
  ChainCtl* ctl = reinterpret_cast<ChainCtl*>(smem + CH_A_BYTES + CH_NSLOT * CH_WSLOT);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_ptiles = (p.P + 127) / 128;

  if (tid == 0) {
    for (int s = 0; s < CH_NSLOT; ++s) { mbar_init(&ctl->w_full[s], 1); mbar_init(&ctl->w_empty[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&ctl->acc_full[s], 1); mbar_init(&ctl->acc_empty[s], CH_EPI_THREADS); }
    for (int s = 0; s < CH_MAX_SLICES; ++s) mbar_init(&ctl->a_ready[s], CH_EPI_THREADS / 2);
    fence_barrier_init();
  }
  if (warp == CH_MMA_WARP) tmem_alloc(&ctl->tmem_addr, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = ctl->tmem_addr;

  if (warp < CH_EPI_WARPS) {
    // =========================================== epilogue warps ===========================================
    // warp w: TMEM lane quadrant w & 3 (tile rows 32 (w & 3) ..), column quarter cq = w >> 2: columns [32 cq, 32 cq + 32) of
    // EVERY N tile, so all 16 warps work on tile 0 while the tensor core is still busy with tile 1.
    const int quad = warp & 3, cq = warp >> 2;
    const int r_in = quad * 32 + lane;                       // row of the tile = TMEM lane
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    uint32_t af_cnt[2] = {0u, 0u};                           // completed waits on acc_full[slot]
    uint32_t lp = 0;                                         // layer parity of the row-exponent exchange buffers
    for (int64_t pt = blockIdx.x; pt < n_ptiles; pt += gridDim.x) {
      const bool tr_on = p.trace != nullptr && blockIdx.x == 0 && pt == blockIdx.x && lane == 0 && (warp == 0 || warp == 12);
      const int tr_role = warp == 0 ? 0 : 1;
      const int64_t row = pt * 128 + r_in;
      const bool row_ok = row < p.P;
      float x[3] = {0.f, 0.f, 0.f};
      if (row_ok) { x[0] = p.pts[row * 3 + 0] * p.scale; x[1] = p.pts[row * 3 + 1] * p.scale; x[2] = p.pts[row * 3 + 2] * p.scale; }
      // PE(x) of this thread's row (models/embedder.py:22-36): [x | sin(2^q x) | cos(2^q x)]_q; kept for the skip layer
      float pe[CH_MAX_PE];
#pragma unroll
      for (int j = 0; j < CH_MAX_PE; ++j) pe[j] = 0.f;
      pe[0] = x[0]; pe[1] = x[1]; pe[2] = x[2];
      {
        float f = 1.0f;
#pragma unroll 1
        for (int q = 0; q < p.n_freq; ++q) {
#pragma unroll 1
          for (int c = 0; c < 3; ++c) {
            float sn, cs;
            sincosf(x[c] * f, &sn, &cs);
            pe[3 + 6 * q + c] = sn;
            pe[3 + 6 * q + 3 + c] = cs;
          }
          f *= 2.0f;
        }
      }
      // ---- layer-0 operand: PE(x) -> planes of K slice 0 (column quarters 0 and 1 hold its 64 columns) ----
      float sa, inv;
      {
        float mx = 0.f;
        if (cq < 2) {
#pragma unroll 1
          for (int j = 0; j < 32; ++j) mx = fmaxf(mx, fabsf(pe[32 * cq + j]));
        }
        ctl->rowexp[lp][cq][r_in] = (uint8_t)expo_bits(mx);
        epi_bar_sync();
        {
          const uint32_t e01 = max((uint32_t)ctl->rowexp[lp][0][r_in], (uint32_t)ctl->rowexp[lp][1][r_in]);
          const uint32_t e23 = max((uint32_t)ctl->rowexp[lp][2][r_in], (uint32_t)ctl->rowexp[lp][3][r_in]);
          row_scale(max(e01, e23), sa, inv);
        }
        lp ^= 1;
        if (cq < 2) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 p0, p1;
            slice8(pe + 32 * cq + 8 * g, inv, p0, p1);
            const uint32_t off = sw128((uint32_t)r_in, (uint32_t)(32 * cq + 8 * g));
            *reinterpret_cast<uint4*>(a_smem + off) = p0;
            *reinterpret_cast<uint4*>(a_smem + CH_APLANE + off) = p1;
          }
          fence_proxy_async();
          mbar_arrive(&ctl->a_ready[0]);
          if (p.e0 != nullptr && row_ok) {
            float* e = p.e0 + row * p.pe_ld;
#pragma unroll 1
            for (int j = 32 * cq; j < 32 * cq + 32; ++j)
              if (j < p.pe_ld) e[j] = pe[j];
          }
        }
      }
      for (int l = 0; l < p.n_layers; ++l) {
        const ChainLayer& L = p.L[l];
        CH_TR(tr_role, l, 0, clock64());
        const float sl = sa * __ldg(L.wscale);              // z = sl (M + C) + b
        if (!L.last) {
          // ---------------- hidden layer: produce A_{l+1} ----------------
          const int n_next = L.N + L.pe_next;                          // width of the next layer's input
          float rmax = 0.f;
#pragma unroll 1
          for (int t = 0; t < 2; ++t) {
            if (CH_NT * t >= n_next) break;
            const bool has_acc = t < L.n_tiles;
            int rows_t = 0, n_valid = 0;
            float bl = 0.f;
            if (has_acc) {
              rows_t = ch_tile_rows(L.N, t);
              n_valid = L.N - CH_NT * t; n_valid = n_valid < CH_NT ? n_valid : CH_NT;
              bl = __ldg(L.bias + CH_NT * t + 32 * cq + lane);      // lane j holds the bias of this warp's column j
              mbar_wait(&ctl->acc_full[t], af_cnt[t] & 1u);
              ++af_cnt[t];
              tcgen05_fence_after();
            }
            if (t == 0) CH_TR(tr_role, l, 1, clock64());
            const uint32_t t_m = t_lane + (uint32_t)t * 256u;
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              const int c0 = 32 * cq + 16 * sub;                     // column inside the tile
              float v[16];
              if (c0 < rows_t) {
                float cc[16];
                tmem_ld16x2(t_m + (uint32_t)c0, t_m + 128u + (uint32_t)c0, v, cc);
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] += cc[j];
              }
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const int lc = c0 + j, col = CH_NT * t + lc;
                const float b = __shfl_sync(0xffffffffu, bl, 16 * sub + j);
                float a = 0.f;
                if (lc < n_valid) a = softplus100_fast(fmaf(v[j], sl, b)) * L.post_scale;
                else if (col >= L.N && col < n_next) a = pe[col - L.N] * L.post_scale;
                v[j] = a;
                rmax = fmaxf(rmax, fabsf(a));
              }
              tmem_st16(t_m + (uint32_t)c0, v);
              if (L.out != nullptr && row_ok) {
                float* o = L.out + row * L.ld_out + CH_NT * t + c0;
                const int nv = n_next - (CH_NT * t + c0);
                if (nv >= 16 && (L.ld_out & 3) == 0) {
#pragma unroll
                  for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(o + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                } else {
#pragma unroll
                  for (int j = 0; j < 16; ++j)
                    if (j < nv) o[j] = v[j];
                }
              }
            }
          }
          tmem_wait_st();
          ctl->rowexp[lp][cq][r_in] = (uint8_t)expo_bits(rmax);
          CH_TR(tr_role, l, 2, clock64());
          epi_bar_sync();                                    // all MMAs of this layer are complete, all row maxima are in
          CH_TR(tr_role, l, 3, clock64());
          {
            const uint32_t e01 = max((uint32_t)ctl->rowexp[lp][0][r_in], (uint32_t)ctl->rowexp[lp][1][r_in]);
            const uint32_t e23 = max((uint32_t)ctl->rowexp[lp][2][r_in], (uint32_t)ctl->rowexp[lp][3][r_in]);
            row_scale(max(e01, e23), sa, inv);
          }
          lp ^= 1;
          const int nks_next = pad64(n_next) / 64;
#pragma unroll 1
          for (int t = 0; t < 2; ++t) {
            if (CH_NT * t >= n_next) break;
            const int s = 2 * t + (cq >> 1);                 // K slice of the next layer this warp's columns belong to
            const uint32_t t_m = t_lane + (uint32_t)t * 256u;
            if (s < nks_next) {
              uint8_t* sl_base = a_smem + (size_t)s * 2 * CH_APLANE;
#pragma unroll
              for (int sub = 0; sub < 2; ++sub) {
                float v[16];
                tmem_ld16(t_m + (uint32_t)(32 * cq + 16 * sub), v);
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                  uint4 p0, p1;
                  slice8(v + 8 * g, inv, p0, p1);
                  const uint32_t off = sw128((uint32_t)r_in, (uint32_t)((cq & 1) * 32 + 16 * sub + 8 * g));
                  *reinterpret_cast<uint4*>(sl_base + off) = p0;
                  *reinterpret_cast<uint4*>(sl_base + CH_APLANE + off) = p1;
                }
              }
              fence_proxy_async();
              mbar_arrive(&ctl->a_ready[s]);
            }
            if (t < L.n_tiles) {
              tcgen05_fence_before();
              mbar_arrive(&ctl->acc_empty[t]);
            }
          }
          CH_TR(tr_role, l, 4, clock64());
        } else {
          // ---------------- last layer: plain output ----------------
#pragma unroll 1
          for (int t = 0; t < L.n_tiles; ++t) {
            const int slot = t & 1;
            const int rows_t = ch_tile_rows(L.N, t);
            const float bl = __ldg(L.bias + CH_NT * t + 32 * cq + lane);
            mbar_wait(&ctl->acc_full[slot], af_cnt[slot] & 1u);
            ++af_cnt[slot];
            tcgen05_fence_after();
            const uint32_t t_m = t_lane + (uint32_t)slot * 256u;
#pragma unroll
            for (int sub = 0; sub < 2; ++sub) {
              const int c0 = 32 * cq + 16 * sub;
              if (c0 < rows_t) {
                float v[16], cc[16];
                tmem_ld16x2(t_m + (uint32_t)c0, t_m + 128u + (uint32_t)c0, v, cc);
                const int col0 = CH_NT * t + c0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  const float b = __shfl_sync(0xffffffffu, bl, 16 * sub + j);
                  v[j] = (col0 + j < L.N) ? fmaf(v[j] + cc[j], sl, b) : 0.f;
                }
                if (row_ok) {
                  if (L.out != nullptr) {
                    float* o = L.out + row * L.ld_out + col0;
                    const int nv = L.N - col0;
                    if (nv >= 16 && (L.ld_out & 3) == 0) {
#pragma unroll
                      for (int q = 0; q < 4; ++q) *reinterpret_cast<float4*>(o + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                    } else {
#pragma unroll
                      for (int j = 0; j < 16; ++j)
                        if (j < nv) o[j] = v[j];
                    }
                  }
                  if (p.udf_out != nullptr && col0 == 0) p.udf_out[row] = fabsf(v[0]) * p.inv_scale;
                }
              } else {
                // keep the warp-collective shuffles of the other branch matched: nothing to do
              }
            }
            tcgen05_fence_before();
            mbar_arrive(&ctl->acc_empty[slot]);
          }
          CH_TR(tr_role, l, 4, clock64());
        }
      }
      epi_bar_sync();     // every MMA of this point tile has completed (all warps waited for the last N tile): the A planes
                          // may be overwritten by the next point tile's layer-0 operand
    }
  } else if (warp == CH_MMA_WARP) {
    // =========================================== MMA issuer ===========================================
    if (lane == 0) {
      uint32_t wcnt = 0, ae_cnt[2] = {0u, 0u}, ar_cnt[CH_MAX_SLICES] = {0u, 0u, 0u, 0u};
      const uint32_t a_addr = smem_u32(a_smem), w_addr = smem_u32(w_smem);
      for (int64_t pt = blockIdx.x; pt < n_ptiles; pt += gridDim.x) {
        const bool tr_on = p.trace != nullptr && blockIdx.x == 0 && pt == blockIdx.x;
        for (int l = 0; l < p.n_layers; ++l) {
          const ChainLayer& L = p.L[l];
          int st = L.stage0;
          long long w_acc = 0, w_a = 0, w_w = 0, t0;
          CH_TR(2, l, 0, clock64());
          for (int t = 0; t < L.n_tiles; ++t) {
            const uint32_t slot = (uint32_t)(t & 1);
            t0 = clock64();
            mbar_wait(&ctl->acc_empty[slot], (ae_cnt[slot] & 1u) ^ 1u);       // the epilogue has drained this accumulator pair
            w_acc += clock64() - t0;
            ++ae_cnt[slot];
            tcgen05_fence_after();
            const uint32_t acc_m = tmem_base + slot * 256u, acc_c = acc_m + 128u;
            for (int s = 0; s < L.n_kslices; ++s, ++st, ++wcnt) {
              t0 = clock64();
              if (t == 0) { mbar_wait(&ctl->a_ready[s], ar_cnt[s] & 1u); ++ar_cnt[s]; }
              w_a += clock64() - t0;
              const uint32_t ws = wcnt % CH_NSLOT, wu = wcnt / CH_NSLOT;
              t0 = clock64();
              mbar_wait(&ctl->w_full[ws], wu & 1u);
              w_w += clock64() - t0;
              tcgen05_fence_after();
              const uint32_t rows = p.S[st].rows;
              const uint32_t idesc = make_idesc_f16(rows);
              const uint32_t a0 = a_addr + (uint32_t)s * 2u * CH_APLANE, a1 = a0 + CH_APLANE;
              const uint32_t w0 = w_addr + ws * CH_WSLOT, w1 = w0 + rows * 128u, w2 = w1 + rows * 128u;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const uint32_t first = (s == 0 && j == 0) ? 0u : 1u;
                const uint64_t da0 = make_desc(a0 + j * 32), da1 = make_desc(a1 + j * 32);
                const uint64_t db0 = make_desc(w0 + j * 32), db1 = make_desc(w1 + j * 32), db2 = make_desc(w2 + j * 32);
                mma_bf16(acc_m, da0, db0, idesc, first);            // exact main term
                mma_bf16(acc_c, da1, db1, idesc, first);            // corrections, smallest first
                mma_bf16(acc_c, da0, db2, idesc, 1u);
                mma_bf16(acc_c, da1, db0, idesc, 1u);
                mma_bf16(acc_c, da0, db1, idesc, 1u);
              }
              mma_commit(&ctl->w_empty[ws]);
            }
            mma_commit(&ctl->acc_full[slot]);
          }
          CH_TR(2, l, 1, clock64());
          CH_TR(2, l, 2, w_acc);
          CH_TR(2, l, 3, w_a);
          CH_TR(2, l, 4, w_w);
        }
      }
    }
    __syncwarp();
  } else {
    // =========================================== weight-ring loader ===========================================
    if (lane == 0) {
      uint32_t wcnt = 0;
      for (int64_t pt = blockIdx.x; pt < n_ptiles; pt += gridDim.x) {
        for (int st = 0; st < p.n_stages; ++st, ++wcnt) {
          const uint32_t ws = wcnt % CH_NSLOT, wu = wcnt / CH_NSLOT;
          mbar_wait(&ctl->w_empty[ws], (wu & 1u) ^ 1u);
          const ChainStage S = p.S[st];
          const uint32_t bytes = (uint32_t)S.rows * 128u;
          mbar_arrive_expect_tx(&ctl->w_full[ws], CH_WPL * bytes);
          uint8_t* dst = w_smem + ws * CH_WSLOT;
#pragma unroll
          for (int pl = 0; pl < CH_WPL; ++pl)
            bulk_g2s(dst + pl * bytes, p.img + S.src + (int64_t)pl * S.plane_rows * 64, bytes, &ctl->w_full[ws]);
        }
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == CH_MMA_WARP) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// host: launch on `st`
static inline int launch_chain(const ChainParams& p, cudaStream_t st) {
  if (p.P <= 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    NUDF_CUDA_OK(cudaFuncSetAttribute(udf_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_SMEM_BYTES));
    attr_set = true;
  }
  int64_t grid = (p.P + 127) / 128;
  if (grid > sm_count()) grid = sm_count();
  static int trace_mode = -1;
  if (trace_mode < 0) { const char* e = getenv("NUDF_CHAIN_TRACE"); trace_mode = (e && atoi(e) != 0) ? 1 : 0; }
  if (trace_mode == 1 && p.P >= 128 * 148) {              // profiling aid: synchronous, prints CTA 0's pipeline stamps
    static long long* dbuf = nullptr;
    if (!dbuf) NUDF_CUDA_OK(cudaMalloc(&dbuf, sizeof(long long) * CH_TRACE_WORDS));
    NUDF_CUDA_OK(cudaMemsetAsync(dbuf, 0, sizeof(long long) * CH_TRACE_WORDS, st));
    ChainParams q = p;
    q.trace = dbuf;
    udf_chain_kernel<<<(unsigned)grid, CH_THREADS, CH_SMEM_BYTES, st>>>(q);
    NUDF_LAUNCH_OK();
    static long long h[CH_TRACE_WORDS];
    NUDF_CUDA_OK(cudaMemcpyAsync(h, dbuf, sizeof(h), cudaMemcpyDeviceToHost, st));
    NUDF_CUDA_OK(cudaStreamSynchronize(st));
    const long long z = h[(0 * NUDF_MAX_LAYERS + 0) * 8 + 0];
    fprintf(stderr, "[chain trace] P=%lld layers=%d (cycles rel. to epilogue layer-0 start)\n", (long long)p.P, p.n_layers);
    for (int l = 0; l < p.n_layers; ++l) {
      const long long* e0 = h + (0 * NUDF_MAX_LAYERS + l) * 8;
      const long long* e1 = h + (1 * NUDF_MAX_LAYERS + l) * 8;
      const long long* m = h + (2 * NUDF_MAX_LAYERS + l) * 8;
      fprintf(stderr, "  L%d epi0: start %7lld acc %7lld p1 %7lld bar %7lld p2 %7lld | epi1: acc %7lld p1 %7lld p2 %7lld | mma: start %7lld end %7lld "
                      "wait acc_empty %6lld a_ready %6lld w_full %6lld\n",
              l, e0[0] - z, e0[1] - z, e0[2] - z, e0[3] - z, e0[4] - z, e1[1] ? e1[1] - z : 0, e1[2] ? e1[2] - z : 0, e1[4] ? e1[4] - z : 0,
              m[0] - z, m[1] - z, m[2], m[3], m[4]);
    }
    trace_mode = 2;                                         // once per process
    return 0;
  }
  LaunchTimer lt_(FAM_UDF_FWD_CHAIN, st);
  udf_chain_kernel<<<(unsigned)grid, CH_THREADS, CH_SMEM_BYTES, st>>>(p);
  NUDF_LAUNCH_OK();
  return 0;
}

}  // namespace chain
}  // namespace nudf
