// Fused layer chains of UDFNetwork for sm_100a: ONE persistent kernel walks a whole sequence of dense layers for a 128-point
// tile; the chain-carried activations never leave the SM between layers.  Four chains of the network run this way
// (reference models/fields.py:192-231 and their autograd adjoints; maths: tests/proto/udf_pipeline.py):
//
//   F  value chain          pts --PE--> A_0 --[W_0]--> softplus --> A_1 ... --[W_last]--> y            (:192-211)
//   R  reverse sweep        seed (sgn/scale) W_last[0,:] --> D_l = G_{l+1} s(100 z_l),  G_l = D_l W_l   (exact grad_x udf, :219-231)
//   T  tangent chain        Adot_0 = J_e gbar,  Zdot_l = Adot_l W_l^T,  Q_l = Zdot D_l 100 (1 - S_l),  Adot_{l+1} = S_l Zdot_l
//   B  backward chain       Zbar_{l-1} = (Zbar_l W_l) S_{l-1} + Q_{l-1}
//
// F and R are chained in one launch per forward call (the activations R needs were written a few microseconds earlier by
// the same CTA and come back from L2), T and B in one launch per backward call.  Every chain-local tensor that another
// kernel consumes later (saved activations, D, Q / Zbar, Adot for the weight-gradient contractions) is written once, in
// fp32, by the thread that owns the row.
//
// * every layer is a tcgen05.mma contraction (M = 128 points, N <= 128 per output tile, K = 16 per instruction, fp32
//   accumulators in TMEM); the epilogue warps read the accumulator with tcgen05.ld, apply the element-wise part of the layer
//   and write the NEXT layer's A operand straight into shared memory in the UMMA K-major SWIZZLE_128B layout -- no HBM round
//   trip and no LSU operand path for the chain itself.  Weight slices stream from L2 through a cp.async.bulk (TMA) ring.
// * fp32-grade accuracy on the tensor engine (the udf head feeds exp(-25000 u): SURVEY.md section 0, fact 3) comes from an
//   EXACT-MAIN fp16 slice scheme instead of the truncating 3xBF16 split:
//       operand row r  :  a = 2^(ea_r - 11) (a0 + a1),  a0 = rint(a 2^(11 - ea_r)) in [-2048, 2048] (12-bit integer, exact in
//                         fp16), a1 = fp16(remainder) in [-1/2, 1/2];   2^ea_r > max_k |a_rk|  (per-row power of two)
//       weights, layer :  w = 2^(E - 13) (w0 + w1 + w2),  w0 = 1024 rint(w 2^(3 - E)) (<= 4-bit integer x 2^10),
//                         w1 = fp16(1024 remainder), w2 = fp16(second remainder);  2^E > max |w| of the layer
//   Main accumulator  M = sum_k a0 w0: every product is an integer multiple of 2^10 below 2^24 and the 256-term sum stays below
//   2^22 units, so the tensor core's truncating fp32 accumulation is EXACT.  Correction accumulator
//   C = sum_k a0 w1 + a0 w2 + a1 w0 + a1 w1 is ~2^-3 of M, so its truncation error is ~2^-27 of the result.
//   z = 2^(ea_r + E - 24) (M + C) (+ b) in fp32.  Dropped: a1 w2 (2^-24).  5 tensor products per K step; 2 fp16 planes of A
//   (128 KB of shared memory for a [128 x 256] tile) + a 2-slot ring of 3-plane weight slices (2 x 48 KB).
//   Measured: udf head error 3.7e-7 vs 1.2e-6 for the reference's own fp32 run (tests/test_gpu_chain.py).
// * roles: warps 0-15 epilogue (warp w: TMEM lane quadrant w & 3 = tile rows 32 (w & 3).., column quarter w >> 2 = columns
//   [32 cq, 32 cq + 32) of EVERY N tile), warp 16 MMA issuer + TMEM owner, warp 17 weight-ring loader.  The two N tiles of a
//   layer have their own accumulator pairs (M0 C0 M1 C1 = 512 TMEM columns), so the epilogue of tile 0 overlaps the MMAs of
//   tile 1.  Per-row scales need the row maximum over all columns: pass 1 (element-wise part, partial row max, value written
//   back to TMEM over M), exchange of the exponent bytes through shared memory, pass 2 (slicing into the A planes).
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
constexpr int CH_MAX_STEPS = 24;
constexpr int CH_MAX_GRID = 148;                           // CTAs per launch (<= SM count); sizes the per-CTA encoding stash
#ifndef CH_RING2
#define CH_RING2 2
#endif
constexpr int CH_MAX_PE = 64;                               // positional-encoding width (K of layer 0) <= one K slice

// what the epilogue does with the accumulator of a step (or, for steps without a GEMM, how the next operand is made)
enum StepKind {
  ST_PE = 0,         // no GEMM: next = PE(x)                          out0 = E0
  ST_FWD = 1,        // a = softplus100(z + b) post                    out0 = A[l+1]        next = a (| PE columns at the skip layer)
  ST_FWD_LAST = 2,   // y = z + b                                      out0 = Y, udf_out    (records sign(y_0) for a following R chain)
  ST_REV_SEED = 3,   // no GEMM: g = (sgn/scale) W_last[0, c] post     as ST_REV
  ST_REV = 4,        // g = z post; c < n_main: d = g s(in0)           out0 = D[l-1]        next = d;   c >= n_main: out1 (Gpe) = g
  ST_REV_FINAL = 5,  // ge = z + in1 (Gpe)                             out0 = Ge
  ST_EDOT = 6,       // no GEMM: next = scale J_e(x) gbar              out0 = Edot
  ST_TAN = 7,        // q = z in1 (D) 100 (1 - S), n = S z post        out0 = Q[l], out1 = Adot[l+1]   next = n (| Edot columns at the skip layer)
  ST_LOAD = 8,       // no GEMM: next = in0 (zf)
  ST_BWD = 9,        // ab = z (+ rowv vec0[c]); c < n_main: zb = ab post s(in0) + in1 (Q)    out0 = Zbar[l-1]    next = zb
};

struct ChainStep {
  int kind;
  int K, N;                  // GEMM: logical contraction / output widths (N = 0: no GEMM in this step)
  int n_kslices, n_tiles;    // of the GEMM
  int rows_override;         // > 0: fetch / multiply only this many weight rows of tile 0 (value-only last layer)
  int img_N;                 // N the weight image was laid out for (differs from N only in the value-only last layer)
  int n_wpl;                 // weight planes of the GEMM: 3 = exact fp16 slice scheme (F chain), 2 = split-bf16 (R, T, B chains)
  int n_main;                // REV / BWD: output columns that continue the chain
  int n_next;                // width of the operand this step produces for the next GEMM (0: none)
  int sync_before;           // 1: the step starts with a barrier of the epilogue warps (reads state other threads wrote)
  float post_scale, a_unscale;
  uint32_t img_off;          // uint16-element offset of this GEMM's weight image in ChainParams::img
  const float* bias;         // [128 n_tiles] or null
  const float* wscale;       // DEVICE scalar 2^(E - 13) of the weight image
  const float* in0; const float* in1;          // auxiliary inputs  [P, ld]
  float* out0; float* out1;                    // outputs           [P, ld]
  const float* vec0;         // per-column vector (REV_SEED: W_last[0,:], BWD first step: W_last[0,:])
  const float* rowv;         // per-row vector (BWD first step: z0 [P]) or null
  int ld_in0, ld_in1, ld_out0, ld_out1;
};
struct ChainParams {
  int n_steps;
  ChainStep S[CH_MAX_STEPS];
  const uint16_t* img;
  const float* pts; int64_t P; float scale; int n_freq, d_pe;
  const float* gbar;                     // T chain: upstream gradient of grad_x udf [P,3]
  const float* pe_src; int pe_ld;        // per-CTA stash (128 rows per CTA, T128, pe_ld columns) of what the first step made (PE / Edot),
  int pe_cta;                            // re-read for the skip layer's appended columns; always CTA-local rows (pe_cta = 1)
  int pe_sh;                             // stash column = encoding column + pe_sh, pe_sh = (width of the skip layer's main part) % 8:
                                         // the appended columns of the skip layer then sit at octet-aligned stash columns
  int t128;                              // 1: every [P, ld] tensor of the steps is stored in the T128 layout (common.cuh)
  float* udf_out; float inv_scale;       // value-only mode: udf_out[P] = |y_0| / scale
  long long* trace;                      // profiling aid (NUDF_CHAIN_TRACE=n): clock64() stamps of CTA 0, first point tile
};
// trace layout: [role 0 = epilogue warp 0, 1 = epilogue warp 12, 2 = MMA issuer][step][8]
constexpr int CH_TRACE_WORDS = 3 * CH_MAX_STEPS * 8;
#define CH_TR(role, l, k, v) do { if (tr_on) p.trace[((role) * CH_MAX_STEPS + (l)) * 8 + (k)] = (v); } while (0)

__host__ __device__ inline int ch_tile_rows(int N, int t) { int r = N - CH_NT * t; return pad16(r < CH_NT ? r : CH_NT); }
__host__ __device__ inline int ch_n_tiles(int N) { return (N + CH_NT - 1) / CH_NT; }
// uint16 elements of the chain image of one GEMM: [tile][k slice][plane][rows_t x 64]
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
static __global__ void chain_layer_scale_kernel(const float* __restrict__ W, int64_t ldw, int rows, int cols, float* __restrict__ meta) {
  float mx = 0.f;
  for (int64_t i = threadIdx.x; i < (int64_t)rows * cols; i += blockDim.x) mx = fmaxf(mx, fabsf(W[(i / cols) * ldw + (i % cols)]));
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
// one block per padded operand row n of B(n, k), n < N, k < K.  transposed == 0: B(n,k) = W[n*ldw + k] (X W^T);
// transposed == 1: B(n,k) = W[k*ldw + n] (dY W).  bias_tab (optional): [128 n_tiles] copy of bias, zero padded.
static __device__ __forceinline__ void chain_prep_row(const float* __restrict__ W, int64_t ldw, const float* __restrict__ bias, int N, int K,
                                                      int transposed, const float* __restrict__ meta, uint16_t* __restrict__ img,
                                                      float* __restrict__ bias_tab, int split_bf16, int col) {
  // col = padded output column: 128 t + local row
  const int t = col / CH_NT, nl = col - t * CH_NT;
  const int rows_t = ch_tile_rows(N, t);
  const bool valid = col < N;
  if (threadIdx.x == 0 && bias_tab != nullptr) bias_tab[col] = (valid && bias) ? bias[col] : 0.f;
  if (nl >= rows_t) return;                      // beyond the padded tile: only the bias table entry exists
  const int Kp = pad64(K);
  const float up = split_bf16 ? 1.0f : meta[0];  // 2^(3 - E): scaled weights are in (-8, 8)
  uint16_t* base = img + ch_tile_off(N, K, t);
  for (int k = threadIdx.x; k < Kp; k += blockDim.x) {
    float w = 0.f;
    if (valid && k < K) w = (transposed ? W[(int64_t)k * ldw + col] : W[(int64_t)col * ldw + k]) * up;
    if (split_bf16) {                            // planes 0, 1 = bf16(w), bf16(w - plane 0); plane 2 unused
      const __nv_bfloat16 b0 = __float2bfloat16_rn(w);
      const __nv_bfloat16 b1 = __float2bfloat16_rn(w - __bfloat162float(b0));
      const uint32_t off = sw128((uint32_t)nl, (uint32_t)(k & 63)) >> 1;
      uint16_t* b = base + (int64_t)(k >> 6) * CH_WPL * rows_t * 64 + off;
      b[0] = __bfloat16_as_ushort(b0);
      b[(int64_t)rows_t * 64] = __bfloat16_as_ushort(b1);
      continue;
    }
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

// all layers of a network in one launch each: grid.y = job
struct ScaleJob { const float* W; float* meta; int ldw, rows, cols; };
struct ScaleJobs { int n; ScaleJob j[16]; };
struct PrepJob { const float* W; const float* bias; const float* meta; uint16_t* img; float* bias_tab; int ldw, N, K, transposed, split; };
struct PrepJobs { int n; PrepJob j[48]; };
static __global__ void chain_scale_jobs_kernel(const __grid_constant__ ScaleJobs jobs) {
  const ScaleJob& J = jobs.j[blockIdx.x];
  float mx = 0.f;
  for (int64_t i = threadIdx.x; i < (int64_t)J.rows * J.cols; i += blockDim.x) mx = fmaxf(mx, fabsf(J.W[(i / J.cols) * J.ldw + (i % J.cols)]));
  __shared__ float red[32];
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) mx = fmaxf(mx, red[i]);
    int e = 0;
    if (mx > 1e-30f && mx < 1e30f) e = (int)((__float_as_uint(mx) >> 23) & 0xffu) - 126;
    J.meta[0] = __uint_as_float((uint32_t)(3 - e + 127) << 23);
    J.meta[1] = __uint_as_float((uint32_t)(e - 13 + 127) << 23);
  }
}
static __device__ __forceinline__ void chain_prep_row(const float* __restrict__ W, int64_t ldw, const float* __restrict__ bias, int N, int K,
                                                      int transposed, const float* __restrict__ meta, uint16_t* __restrict__ img,
                                                      float* __restrict__ bias_tab, int split_bf16, int col);
static __global__ void chain_prep_jobs_kernel(const __grid_constant__ PrepJobs jobs) {
  const PrepJob& J = jobs.j[blockIdx.y];
  const int col = blockIdx.x;
  if (col >= CH_NT * ch_n_tiles(J.N)) return;
  chain_prep_row(J.W, J.ldw, J.bias, J.N, J.K, J.transposed, J.meta, J.img, J.bias_tab, J.split, col);
}

static inline int run_prep_jobs(const ScaleJobs& sj, const PrepJobs& pj, cudaStream_t st) {
  if (sj.n > 0) {
    chain_scale_jobs_kernel<<<sj.n, 256, 0, st>>>(sj);
    NUDF_LAUNCH_OK();
  }
  if (pj.n > 0) {
    int mx = 0;
    for (int i = 0; i < pj.n; ++i) { const int c = CH_NT * ch_n_tiles(pj.j[i].N); mx = c > mx ? c : mx; }
    chain_prep_jobs_kernel<<<dim3((unsigned)mx, (unsigned)pj.n), 64, 0, st>>>(pj);
    NUDF_LAUNCH_OK();
  }
  return 0;
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
// elect.sync: one lane of a converged warp.  The MMA warp runs its issue loop with warp-uniform control flow and issues inside
// `if (elect_one())`: ptxas then keeps the descriptor arithmetic on the uniform datapath (UIADD3 / UMOV between consecutive
// UTCHMMAs) instead of the R2UR round trips a `lane == 0` loop needs (~90 cycles per 64-cycle MMA in the first version).
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "selp.u32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred;
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
// 8 consecutive values -> 16 bytes of plane 0 (integer part) and plane 1 (remainder), fp16
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
// 16 consecutive floats of one row of a [P, ld] tensor (columns c .. c+15, c % 16 == 0, `nv` of them valid); 16-byte accesses.
// Plain (coherent) loads: tensors written earlier by this very kernel are read here (F -> R, T -> B hand-offs).
// t128: T128 layout (common.cuh) -- lanes = consecutive rows => every access instruction of a warp is contiguous.
__device__ __forceinline__ void ld_row16(const float* base, int64_t ld, int64_t row, int c, int nv, bool t128, float v[16]) {
  if (t128) {
    const float* q = base + t128_off(row, c, ld);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
      if (4 * i < nv) t = *reinterpret_cast<const float4*>(q + (int64_t)i * 512);
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
    if (nv < 16) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j >= nv) v[j] = 0.f;
    }
    return;
  }
  const float* q = base + row * ld + c;
  if (nv >= 16 && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(q) & 15u) == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 t = reinterpret_cast<const float4*>(q)[i];
      v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = j < nv ? q[j] : 0.f;
  }
}
// store; in the T128 layout whole 4-column groups are written (values beyond nv must already be zero: they land in the
// tensor's own padding columns, nv <= ld - c)
__device__ __forceinline__ void st_row16(float* base, int64_t ld, int64_t row, int c, int nv, bool t128, const float v[16]) {
  if (t128) {
    float* q = base + t128_off(row, c, ld);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (4 * i < nv) *reinterpret_cast<float4*>(q + (int64_t)i * 512) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    return;
  }
  float* q = base + row * ld + c;
  if (nv >= 16 && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(q) & 15u) == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) reinterpret_cast<float4*>(q)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
#pragma unroll
    for (int j = 0; j < 16; ++j)
      if (j < nv) q[j] = v[j];
  }
}

struct ChainCtl {
  uint64_t w_full[CH_NSLOT], w_empty[CH_NSLOT];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t a_ready[CH_MAX_SLICES];
  uint32_t tmem_addr;
  uint32_t pad_;
  uint8_t rowexp[2][4][128];    // [step parity][column quarter][row]: exponent field of the partial row maximum
  float rowsgn[128];            // sign(y_0) / scale of the last F layer, for the seed of a following R chain
};
// 128 KB of A planes + 2 x 48 KB weight slots + control block = 225.6 KB of the 227 KB a CTA may have: no room for an alignment
// slack, so the dynamic shared-memory window itself is declared 1024-byte aligned (SWIZZLE_128B operands need it) and the
// kernel traps if the runtime did not honour that.
constexpr size_t CH_SMEM_BYTES = (size_t)CH_A_BYTES + (size_t)CH_NSLOT * CH_WSLOT + sizeof(ChainCtl);
static_assert(CH_SMEM_BYTES <= 227 * 1024, "fused chain: shared-memory budget exceeded");

// ---- epilogue -------------------------------------------------------------------------------------------------------
// The epilogue walks a step's output in OCTETS (8 consecutive columns of the thread's row): per octet two tcgen05.ld.x8 (main and
// correction accumulator), the element-wise part of the step, 16-byte global stores, one tcgen05.st.x8 that parks the value
// which continues the chain.  Interior octets (all 8 columns inside the step's main column range, T128 tensors) take a
// predicate-free fast path whose auxiliary global loads (saved activations, D, Q) are software-pipelined two octets ahead in a
// 3-deep register ring -- the first two are issued before the accumulator-ready wait; boundary octets (skip-layer columns,
// partial tiles) and row-major tensors go through one generic out-of-line routine.
__device__ __forceinline__ void tmem_ld8x2(uint32_t ta, uint32_t tb, float v[8], float w[8]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  uint32_t* q = reinterpret_cast<uint32_t*>(w);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%16];\n"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8, %9, %10, %11, %12, %13, %14, %15}, [%17];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(q[0]), "=r"(q[1]),
        "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
      : "r"(ta), "r"(tb)
      : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t ta, float v[8]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "r"(ta)
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t ta, const float v[8]) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(ta), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ float ex2_fast(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// sigma(100 z) recovered from the stored softplus output a = post softplus100(z):  1 - exp(-100 a / post) = 1 - 2^(ak a) with
// ak = -100 log2(e) / post.  In the linear regime (100 z > 20) 2^(ak a) < 2.1e-9 and the result rounds to exactly 1, which is
// what sig_from_softplus (common.cuh) returns there.
__device__ __forceinline__ float sig_ak(float a, float ak) { return 1.0f - ex2_fast(a * ak); }

struct EpiT {                         // per-thread constants of one point tile
  int r_in, cq;
  uint32_t t_lane;
  int64_t row;                        // global row (point)
  int64_t srow;                       // row inside the PE / Edot tensor (CTA-local stash in value-only launches)
  bool row_ok, t128;
};
__device__ __forceinline__ int64_t row_off(int64_t ld, int64_t row, bool t128) {
  return t128 ? ((row >> 7) * (ld >> 2) * 512 + (row & 127) * 4) : row * ld;
}
// T128 tensors, c % 8 == 0: the octet of this thread's row is two float4, 512 floats apart; rp = base + row_off
__device__ __forceinline__ void ldg8(const float* rp, int c, float v[8]) {
  const float4 a = *reinterpret_cast<const float4*>(rp + c * 128);
  const float4 b = *reinterpret_cast<const float4*>(rp + c * 128 + 512);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void stg8(float* rp, int c, const float v[8]) {
  *reinterpret_cast<float4*>(rp + c * 128) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(rp + c * 128 + 512) = make_float4(v[4], v[5], v[6], v[7]);
}
// 8 consecutive values -> 16 bytes of plane 0 / plane 1 into the A planes of K slice `sl_base`
__device__ __forceinline__ void slice_store8(const float v[8], float inv, uint8_t* sl_base, int r_in, int kcol) {
  uint4 p0, p1;
  slice8(v, inv, p0, p1);
  const uint32_t off = sw128((uint32_t)r_in, (uint32_t)kcol);
  *reinterpret_cast<uint4*>(sl_base + off) = p0;
  *reinterpret_cast<uint4*>(sl_base + CH_APLANE + off) = p1;
}

struct StepR {                        // the fields of a step the fast paths use, read once per step
  int N, n_tiles, rows_override, n_main, n_next, n_cols, n_out_tiles;
  int ld_in0, ld_in1, ld_out0, ld_out1;
  int pe_base;                        // FWD / TAN skip layer: stash column of logical column c is c - pe_base (N rounded down to 8)
  int gpe_base;                       // REV skip layer: likewise for the shifted Gpe tensor (n_main rounded down to 8)
  float ps, ak, rz;
  bool has_rowv;
  const float* bias; const float* vec0;
  const float* in0r; const float* in1r; float* out0r; float* out1r;      // row pointers (base + row_off) or null
  const float* pesr;                  // this thread's row of the encoding stash (T128) or null
};
__device__ __forceinline__ StepR load_step(const ChainStep& G, const ChainParams& p, const EpiT& e) {
  StepR S;
  S.N = G.N; S.n_tiles = G.n_tiles; S.rows_override = G.rows_override; S.n_main = G.n_main; S.n_next = G.n_next;
  int n_cols = G.n_next > G.N ? G.n_next : G.N;
  if (G.kind == ST_TAN && G.n_main > n_cols) n_cols = G.n_main;
  S.n_cols = n_cols;
  S.n_out_tiles = (n_cols + CH_NT - 1) / CH_NT;
  S.ld_in0 = G.ld_in0; S.ld_in1 = G.ld_in1; S.ld_out0 = G.ld_out0; S.ld_out1 = G.ld_out1;
  S.pe_base = G.N & ~7;
  S.gpe_base = G.n_main & ~7;
  S.ps = G.post_scale;
  S.ak = -144.26950408889634f * G.a_unscale;
  S.has_rowv = G.rowv != nullptr;
  S.rz = (G.rowv != nullptr && e.row_ok) ? G.rowv[e.row] : 0.f;
  S.bias = G.bias; S.vec0 = G.vec0;
  S.in0r = G.in0 ? G.in0 + row_off(G.ld_in0, e.row, e.t128) : nullptr;
  S.in1r = G.in1 ? G.in1 + row_off(G.ld_in1, e.row, e.t128) : nullptr;
  S.out0r = G.out0 ? G.out0 + row_off(G.ld_out0, e.row, e.t128) : nullptr;
  S.out1r = G.out1 ? G.out1 + row_off(G.ld_out1, e.row, e.t128) : nullptr;
  S.pesr = p.pe_src ? p.pe_src + row_off(p.pe_ld, e.srow, true) : nullptr;
  return S;
}

// Generic octet for ROW-MAJOR tensors (t128 == 0: the F chain of a partially fused configuration -- only PE / FWD / FWD_LAST
// steps exist there): z[8] in = accumulator values, out = the value that continues the chain; performs the step's stores.
__device__ __noinline__ void octet_slow(const ChainStep* S, const ChainParams* p, int64_t row, int64_t srow, bool row_ok, int col0, float* z) {
  const int kind = S->kind;
  const float ps = S->post_scale;
  const int N = S->N, n_next = S->n_next;
#pragma unroll 1
  for (int j = 0; j < 8; ++j) {
    const int col = col0 + j;
    const float zz = z[j];
    float nx = 0.f;
    if (kind == ST_FWD) {
      if (col < N) nx = softplus100_fast(zz + __ldg(S->bias + col)) * ps;
      else if (col < n_next) nx = (row_ok ? p->pe_src[t128_off(srow, col - (N & ~7), p->pe_ld)] : 0.f) * ps;      // stash: T128, shifted
      if (row_ok && S->out0 != nullptr && col < n_next) S->out0[row * S->ld_out0 + col] = nx;
    } else if (kind == ST_FWD_LAST) {
      if (col < N) {
        nx = zz + __ldg(S->bias + col);
        if (row_ok && S->out0 != nullptr) S->out0[row * S->ld_out0 + col] = nx;
      }
    }
    z[j] = nx;
  }
}

// One octet of a step on T128 tensors.  a / b: prefetched auxiliary octets (FWD: a = bias).  Columns beyond the step's ranges
// come out as exact zeros (they are the tensors' padding and the next operand's K padding).  Interior octets (`full`) take the
// predicate-free branch; boundary octets (skip-layer columns, last partial octet) the masked one.
template <int KIND>
__device__ __forceinline__ void octet_fast(const StepR& S, const EpiT& e, int col0, float z[8], const float a[8], float b[8]) {
  if (KIND == ST_FWD) {                    // a = the bias octet
    const int n1 = S.N - col0;             // columns j < n1: softplus; n1 <= j < n2: appended encoding (skip layer); else 0
    if (n1 >= 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] = softplus100_fast(z[j] + a[j]) * S.ps;
    } else {
      const int n2 = S.n_next - col0;
      float pe[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) pe[j] = 0.f;
      if (n2 > n1 && n2 > 0 && e.row_ok && S.pesr != nullptr) ldg8(S.pesr, col0 - S.pe_base, pe);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float sp = softplus100_fast(z[j] + a[j]);
        z[j] = (j < n1 ? sp : (j < n2 ? pe[j] : 0.f)) * S.ps;
      }
    }
    if (S.out0r != nullptr && e.row_ok && col0 < S.ld_out0) stg8(S.out0r, col0, z);
  } else if (KIND == ST_FWD_LAST) {        // a = the bias octet
    const int n1 = S.N - col0;
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = (j < n1) ? z[j] + a[j] : 0.f;
    if (S.out0r != nullptr && e.row_ok && col0 < S.ld_out0) stg8(S.out0r, col0, z);
  } else if (KIND == ST_REV || KIND == ST_REV_SEED) {
    const int n1 = S.n_main - col0;        // j < n1: D = G s(A); n1 <= j (< N): gradient w.r.t. the appended PE columns -> shifted Gpe
    if (n1 >= 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] = z[j] * S.ps * sig_ak(a[j], S.ak);
    } else {
      const int n2 = S.N - col0;
      float gp[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float g = z[j] * S.ps;
        gp[j] = (j >= n1 && j < n2) ? g : 0.f;
        z[j] = (j < n1) ? g * sig_ak(a[j], S.ak) : 0.f;
      }
      if (S.out1r != nullptr && e.row_ok) stg8(S.out1r, col0 - S.gpe_base, gp);
    }
    if (S.out0r != nullptr && e.row_ok && col0 < S.ld_out0) stg8(S.out0r, col0, z);
  } else if (KIND == ST_REV_FINAL) {
    const int n1 = S.N - col0;
#pragma unroll
    for (int j = 0; j < 8; ++j) z[j] = (j < n1) ? z[j] : 0.f;
    if (S.out0r != nullptr && e.row_ok && col0 < S.ld_out0) stg8(S.out0r, col0, z);
  } else if (KIND == ST_TAN) {
    const int n1 = S.N - col0;             // j < n1: q = zdot d 100 (1 - s), adot = s zdot; n1 <= j < n2: appended Edot columns
    if (n1 >= 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float s = sig_ak(a[j], S.ak);
        b[j] = z[j] * b[j] * (100.0f * (1.0f - s));
        z[j] = s * z[j] * S.ps;
      }
    } else {
      const int n2 = S.n_main - col0;
      float pe[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) pe[j] = 0.f;
      if (n2 > n1 && n2 > 0 && e.row_ok && S.pesr != nullptr) ldg8(S.pesr, col0 - S.pe_base, pe);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float s = sig_ak(a[j], S.ak);
        const float q = z[j] * b[j] * (100.0f * (1.0f - s));
        const float n = s * z[j];
        b[j] = (j < n1) ? q : 0.f;
        z[j] = (j < n1 ? n : (j < n2 ? pe[j] : 0.f)) * S.ps;
      }
    }
    if (e.row_ok) {
      if (S.out0r != nullptr && col0 < S.ld_out0) stg8(S.out0r, col0, b);
      if (S.out1r != nullptr && col0 < S.ld_out1) stg8(S.out1r, col0, z);
    }
  } else if (KIND == ST_BWD) {
    const int n1 = S.n_main - col0;
    if (S.has_rowv) {
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] = fmaf(S.rz, (j < n1) ? __ldg(S.vec0 + col0 + j) : 0.f, z[j]);
    }
    if (n1 >= 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] = z[j] * S.ps * sig_ak(a[j], S.ak) + b[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] = (j < n1) ? z[j] * S.ps * sig_ak(a[j], S.ak) + b[j] : 0.f;
    }
    if (S.out0r != nullptr && e.row_ok && col0 < S.ld_out0) stg8(S.out0r, col0, z);
  }
}

// row maximum exchange + per-row scale of the operand the step produced; returns after the barrier
__device__ __forceinline__ void exchange_scale(ChainCtl* ctl, const EpiT& e, float rmax, uint32_t& lp, float& sa, float& inv) {
  ctl->rowexp[lp][e.cq][e.r_in] = (uint8_t)expo_bits(rmax);
  epi_bar_sync();            // all MMAs of this step are complete (every warp waited for every tile), all maxima are in
  const uint32_t e01 = max((uint32_t)ctl->rowexp[lp][0][e.r_in], (uint32_t)ctl->rowexp[lp][1][e.r_in]);
  const uint32_t e23 = max((uint32_t)ctl->rowexp[lp][2][e.r_in], (uint32_t)ctl->rowexp[lp][3][e.r_in]);
  row_scale(max(e01, e23), sa, inv);
  lp ^= 1;
}

#define CH_TRS(k) do { if (tr_on) p.trace[(tr_role * CH_MAX_STEPS + l) * 8 + (k)] = clock64(); } while (0)

// accumulator loads without the wait, and the wait with the destination registers as in/out operands: nothing that uses
// them can be scheduled above it
__device__ __forceinline__ void tmem_ld8_nowait(uint32_t ta, float v[8]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(ta)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_fence8(float v[8]) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile("tcgen05.wait::ld.sync.aligned;\n"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :
               : "memory");
}
// 8 consecutive values -> split-bf16 planes (hi = bf16(x), lo = bf16(x - hi)) of K slice `sl_base`
__device__ __forceinline__ void split_store8(const float v[8], uint8_t* sl_base, int r_in, int kcol) {
  uint32_t h[4], lo[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    h[q] = pack_bf16(v[2 * q], v[2 * q + 1]);
    lo[q] = pack_bf16(v[2 * q] - __uint_as_float(h[q] << 16), v[2 * q + 1] - __uint_as_float(h[q] & 0xFFFF0000u));
  }
  const uint32_t off = sw128((uint32_t)r_in, (uint32_t)kcol);
  *reinterpret_cast<uint4*>(sl_base + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(sl_base + CH_APLANE + off) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}
// the step's own GEMM ran in the exact fp16 scheme (two accumulators, row-scaled operand) -- F chain; the others split-bf16
#ifndef CH_ALL_EXACT
#define CH_ALL_EXACT 0
#endif
__host__ __device__ constexpr bool kind_exact(int kind) { return CH_ALL_EXACT || kind == ST_FWD || kind == ST_FWD_LAST; }

// ---- steps with a GEMM whose output is at most two N tiles wide (FWD, REV, TAN, BWD) ----
template <int KIND>
__device__ __forceinline__ void epi_step_gemm(const ChainStep& G, const ChainParams& p, ChainCtl* ctl, uint8_t* a_smem, const EpiT& e,
                                              uint32_t (&af_cnt)[2], uint32_t& lp, float& sa, float& inv, bool tr_on, int tr_role, int l) {
  constexpr bool EXACT = kind_exact(KIND);
  constexpr int NAUX = (KIND == ST_REV || KIND == ST_FWD) ? 1 : 2;          // FWD: the bias octet travels through the ring
  const StepR S = load_step(G, p, e);
  const float sl = EXACT ? sa * __ldg(G.wscale) : 1.0f;          // exact scheme: z = sl (M + C)
  constexpr int RD = (NAUX == 2) ? CH_RING2 : 3;                 // ring depth; prefetch distance RD - 1 octets
  float ax[RD][8], bx[RD][8];
  auto col_of = [&](int g) { return CH_NT * (g >> 2) + 32 * e.cq + 8 * (g & 3); };
  auto prefetch = [&](int g, float (&a)[8], float (&b)[8]) {
#pragma unroll
    for (int j = 0; j < 8; ++j) { a[j] = 0.f; b[j] = 0.f; }
    const int c = col_of(g);
    if (e.t128 && c < S.n_cols) {
      if (KIND == ST_FWD) {
        if (c < CH_NT * S.n_tiles) {
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(S.bias + c)), b1 = __ldg(reinterpret_cast<const float4*>(S.bias + c + 4));
          a[0] = b0.x; a[1] = b0.y; a[2] = b0.z; a[3] = b0.w; a[4] = b1.x; a[5] = b1.y; a[6] = b1.z; a[7] = b1.w;
        }
      } else if (e.row_ok) {
        if (c < S.ld_in0) ldg8(S.in0r, c, a);
        if (NAUX == 2 && c < S.ld_in1) ldg8(S.in1r, c, b);
      }
    }
  };
  float rmax = 0.f;
#pragma unroll
  for (int g = 0; g < RD - 1; ++g) prefetch(g, ax[g], bx[g]);
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const int t = g >> 2, o = g & 3;
    if (t < S.n_out_tiles) {
      const int slot = t & 1;
      if (o == 0 && t < S.n_tiles) {
        mbar_wait(&ctl->acc_full[slot], af_cnt[slot] & 1u);
        ++af_cnt[slot];
        tcgen05_fence_after();
        if (t == 0) CH_TRS(1);
      }
      if (g + RD - 1 < 8 && ((g + RD - 1) >> 2) < S.n_out_tiles) prefetch(g + RD - 1, ax[(g + RD - 1) % RD], bx[(g + RD - 1) % RD]);
      const int c0 = 32 * e.cq + 8 * o, col0 = CH_NT * t + c0;
      const uint32_t t_m = e.t_lane + (uint32_t)slot * 256u + (uint32_t)c0;
      const int rows_t = (t < S.n_tiles) ? (S.rows_override > 0 ? S.rows_override : ch_tile_rows(S.N, t)) : 0;
      float z[8];
      if (c0 < rows_t) {
        if (EXACT) {
          float cc[8];
          tmem_ld8x2(t_m, t_m + 128u, z, cc);
#pragma unroll
          for (int j = 0; j < 8; ++j) z[j] = (z[j] + cc[j]) * sl;
        } else {
          tmem_ld8(t_m, z);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = 0.f;
      }
      if (col0 < S.n_cols) {
        if (e.t128) {
          octet_fast<KIND>(S, e, col0, z, ax[g % RD], bx[g % RD]);
        } else {                               // row-major tensors (partially fused configurations: F chain only)
          float zs[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) zs[j] = z[j];
          octet_slow(&G, &p, e.row, e.srow, e.row_ok, col0, zs);
#pragma unroll
          for (int j = 0; j < 8; ++j) z[j] = zs[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = 0.f;
      }
      if (S.n_next > 0) {
        if (EXACT) {
#pragma unroll
          for (int j = 0; j < 8; ++j) rmax = fmaxf(rmax, fabsf(z[j]));
        }
        tmem_st8(t_m, z);                                            // parked in TMEM (over the main accumulator)
      }
      if (o == 3 && t < S.n_tiles && S.n_next == 0) {                // nothing continues: the accumulators are free again
        tcgen05_fence_before();
        mbar_arrive(&ctl->acc_empty[slot]);
      }
    }
  }
  CH_TRS(2);
  if (S.n_next > 0) {
    tmem_wait_st();
    // exact scheme: the row scale needs every warp's partial maximum (barrier).  Split-bf16: no scale; this warp has waited
    // for every tile's accumulator, so all MMAs that read the current A planes are complete -- it may overwrite its columns.
    if (EXACT) exchange_scale(ctl, e, rmax, lp, sa, inv);
    CH_TRS(3);
    // ---------------- pass 2: slice the new operand into the A planes ----------------
    const int nks_next = pad64(S.n_next) / 64;
    const int n_nx_tiles = (S.n_next + CH_NT - 1) / CH_NT;
#pragma unroll 1
    for (int t = 0; t < S.n_out_tiles; ++t) {
      const int slot = t & 1;
      const int s = 2 * t + (e.cq >> 1);                 // K slice of the next GEMM this warp's columns belong to
      if (t < n_nx_tiles && s < nks_next) {
        uint8_t* sl_base = a_smem + (size_t)s * 2 * CH_APLANE;
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          float v[8];
          tmem_ld8(e.t_lane + (uint32_t)slot * 256u + (uint32_t)(32 * e.cq + 8 * o), v);
          if (EXACT) slice_store8(v, inv, sl_base, e.r_in, (e.cq & 1) * 32 + 8 * o);
          else split_store8(v, sl_base, e.r_in, (e.cq & 1) * 32 + 8 * o);
        }
        fence_proxy_async();
        mbar_arrive(&ctl->a_ready[s]);
      }
      if (t < S.n_tiles) {
        tcgen05_fence_before();
        mbar_arrive(&ctl->acc_empty[slot]);
      }
    }
  }
}

// ---- last layer of F (up to 3 N tiles, nothing continues) and last step of R (39 columns) ----
template <int KIND>
__device__ __forceinline__ void epi_step_tail(const ChainStep& G, const ChainParams& p, ChainCtl* ctl, const EpiT& e, uint32_t (&af_cnt)[2],
                                              float sa, bool tr_on, int tr_role, int l) {
  constexpr bool EXACT = kind_exact(KIND);
  const StepR S = load_step(G, p, e);
  const float sl = EXACT ? sa * __ldg(G.wscale) : 1.0f;
#pragma unroll 1
  for (int t = 0; t < S.n_out_tiles; ++t) {
    const int slot = t & 1;
    const bool has_acc = t < S.n_tiles;
    if (has_acc) {
      mbar_wait(&ctl->acc_full[slot], af_cnt[slot] & 1u);
      ++af_cnt[slot];
      tcgen05_fence_after();
      if (t == 0) CH_TRS(1);
    }
    const int rows_t = has_acc ? (S.rows_override > 0 ? S.rows_override : ch_tile_rows(S.N, t)) : 0;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int c0 = 32 * e.cq + 8 * o, col0 = CH_NT * t + c0;
      if (col0 < S.n_cols && c0 < rows_t) {
        const uint32_t t_m = e.t_lane + (uint32_t)slot * 256u + (uint32_t)c0;
        float z[8], bb[8];
        if (EXACT) {
          float cc[8];
          tmem_ld8x2(t_m, t_m + 128u, z, cc);
#pragma unroll
          for (int j = 0; j < 8; ++j) z[j] = (z[j] + cc[j]) * sl;
        } else {
          tmem_ld8(t_m, z);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) bb[j] = 0.f;
        if (e.t128) {
          if (KIND == ST_FWD_LAST) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(S.bias + col0)), b1 = __ldg(reinterpret_cast<const float4*>(S.bias + col0 + 4));
            bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
          }
          octet_fast<KIND>(S, e, col0, z, bb, bb);
        } else {
          float zs[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) zs[j] = z[j];
          octet_slow(&G, &p, e.row, e.srow, e.row_ok, col0, zs);
#pragma unroll
          for (int j = 0; j < 8; ++j) z[j] = zs[j];
        }
        if (KIND == ST_FWD_LAST && col0 == 0) {
          const float y0 = z[0];
          ctl->rowsgn[e.r_in] = (y0 > 0.f ? 1.f : (y0 < 0.f ? -1.f : 0.f)) * p.inv_scale;
          if (p.udf_out != nullptr && e.row_ok) p.udf_out[e.row] = fabsf(y0) * p.inv_scale;
        }
      }
    }
    if (has_acc) {
      tcgen05_fence_before();
      mbar_arrive(&ctl->acc_empty[slot]);
    }
  }
  CH_TRS(2);
}

// ---- steps without a GEMM that make a split-bf16 operand from global data (T128 only): LOAD (zf), REV_SEED ----
template <int KIND>
__device__ __forceinline__ void epi_step_nogemm(const ChainStep& G, const ChainParams& p, ChainCtl* ctl, uint8_t* a_smem, const EpiT& e,
                                                bool tr_on, int tr_role, int l) {
  const StepR S = load_step(G, p, e);
  const float sgn_scaled = (KIND == ST_REV_SEED) ? ctl->rowsgn[e.r_in] : 0.f;
  const int nks_next = pad64(S.n_next) / 64;
#pragma unroll 1
  for (int t = 0; t < S.n_out_tiles; ++t) {
    const int s = 2 * t + (e.cq >> 1);
    uint8_t* sl_base = a_smem + (size_t)s * 2 * CH_APLANE;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int col0 = CH_NT * t + 32 * e.cq + 8 * o;
      float z[8], a[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) { z[j] = 0.f; a[j] = 0.f; }
      if (col0 < S.n_cols) {
        if (KIND == ST_LOAD) {
          if (e.row_ok && col0 < S.ld_in0) ldg8(S.in0r, col0, z);
#pragma unroll
          for (int j = 0; j < 8; ++j) z[j] = (col0 + j < S.n_next) ? z[j] : 0.f;
        } else {
          if (e.row_ok && col0 < S.ld_in0) ldg8(S.in0r, col0, a);
#pragma unroll
          for (int j = 0; j < 8; ++j) z[j] = (col0 + j < S.N) ? sgn_scaled * __ldg(S.vec0 + col0 + j) : 0.f;     // S.N = width of the last layer's input
          octet_fast<ST_REV_SEED>(S, e, col0, z, a, a);
        }
      }
      if (s < nks_next) split_store8(z, sl_base, e.r_in, (e.cq & 1) * 32 + 8 * o);
    }
    if (s < nks_next) {
      fence_proxy_async();
      mbar_arrive(&ctl->a_ready[s]);
    }
  }
  CH_TRS(2);
}

// ---- first step: PE(x) (models/embedder.py:22-36) for the F chain, Edot = scale J_e(x) gbar for the T chain ----
// out[c], c < 3 + 6 n_freq: c < 3: x_c; then per frequency q: sin(2^q x_k) (k = 0..2), cos(2^q x_k); the JVP variant returns
// the directional derivative along v.  Out of line and looped on purpose: sincosf is large, the step runs once per point tile.
template <bool JVP>
__device__ __noinline__ void pe_row(float x0, float x1, float x2, float v0, float v1, float v2, int n_freq, float* out) {
  const float x[3] = {x0, x1, x2}, v[3] = {v0, v1, v2};
  float f = 1.0f;
#pragma unroll 1
  for (int k = 0; k < 3; ++k) out[k] = JVP ? v[k] : x[k];
#pragma unroll 1
  for (int q = 0; q < n_freq; ++q) {
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
      float sn, cs;
      sincosf(x[k] * f, &sn, &cs);
      out[3 + 6 * q + k] = JVP ? f * cs * v[k] : sn;
      out[3 + 6 * q + 3 + k] = JVP ? -f * sn * v[k] : cs;
    }
    f *= 2.0f;
  }
}
template <bool JVP>
__device__ __forceinline__ void epi_step_pe(const ChainStep& G, const ChainParams& p, ChainCtl* ctl, uint8_t* a_smem, const EpiT& e,
                                            uint32_t& lp, float& sa, float& inv, bool tr_on, int tr_role, int l) {
  float vals[4][8];
  float rmax = 0.f;
  const bool mine = e.cq <= 1 && 32 * e.cq < G.n_next + 8;       // the encoding (+ stash shift) is at most 64 columns wide: K slice 0 only
  if (mine) {
    float x[3] = {0.f, 0.f, 0.f}, v[3] = {0.f, 0.f, 0.f};
    if (e.row_ok) {
      x[0] = p.pts[e.row * 3 + 0] * p.scale; x[1] = p.pts[e.row * 3 + 1] * p.scale; x[2] = p.pts[e.row * 3 + 2] * p.scale;
      if (JVP) { v[0] = p.gbar[e.row * 3 + 0] * p.scale; v[1] = p.gbar[e.row * 3 + 1] * p.scale; v[2] = p.gbar[e.row * 3 + 2] * p.scale; }
    }
    float row[CH_MAX_PE + 8];                                     // row[8 + c] = encoding column c; row[0..7] = 0 (stash shift)
#pragma unroll 1
    for (int c = 0; c < CH_MAX_PE + 8; ++c) row[c] = 0.f;
    pe_row<JVP>(x[0], x[1], x[2], v[0], v[1], v[2], p.n_freq, row + 8);
    float* outr = (G.out0 != nullptr && e.t128) ? G.out0 + row_off(G.ld_out0, e.row, true) : nullptr;
    float* stash = (p.pe_src != nullptr) ? const_cast<float*>(p.pe_src) + row_off(p.pe_ld, e.srow, true) : nullptr;
#pragma unroll
    for (int o = 0; o < 4; ++o) {
      const int col0 = 32 * e.cq + 8 * o;
      float sh[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        vals[o][j] = row[8 + col0 + j];
        sh[j] = row[8 + col0 + j - p.pe_sh];                      // stash column c holds encoding column c - pe_sh
        rmax = fmaxf(rmax, fabsf(vals[o][j]));
      }
      if (e.row_ok) {
        if (outr != nullptr && col0 < G.ld_out0) stg8(outr, col0, vals[o]);          // columns >= d_pe are zeros: the tensor's padding
        if (G.out0 != nullptr && !e.t128) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (col0 + j < G.ld_out0) G.out0[e.row * G.ld_out0 + col0 + j] = vals[o][j];
        }
        if (stash != nullptr && col0 < p.pe_ld) stg8(stash, col0, sh);
      }
    }
  }
  CH_TRS(2);
  if (!JVP) exchange_scale(ctl, e, rmax, lp, sa, inv);          // the F chain's operand is row-scaled fp16; the T chain's split-bf16
  CH_TRS(3);
  if (e.cq <= 1) {
    if (mine) {
#pragma unroll
      for (int o = 0; o < 4; ++o) {
        if (JVP) split_store8(vals[o], a_smem, e.r_in, 32 * e.cq + 8 * o);
        else slice_store8(vals[o], inv, a_smem, e.r_in, 32 * e.cq + 8 * o);
      }
    }
    fence_proxy_async();
    mbar_arrive(&ctl->a_ready[0]);
  }
}

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
    EpiT e;
    e.cq = warp >> 2;
    e.r_in = (warp & 3) * 32 + lane;                         // row of the tile = TMEM lane
    e.t_lane = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    e.t128 = p.t128 != 0;
    uint32_t af_cnt[2] = {0u, 0u};                           // completed waits on acc_full[slot]
    uint32_t lp = 0;                                         // parity of the row-exponent exchange buffers
    for (int64_t pt = blockIdx.x; pt < n_ptiles; pt += gridDim.x) {
      const bool tr_on = p.trace != nullptr && blockIdx.x == 0 && pt == blockIdx.x && lane == 0 && (warp == 0 || warp == 12);
      const int tr_role = warp == 0 ? 0 : 1;
      e.row = pt * 128 + e.r_in;
      e.row_ok = e.row < p.P;
      e.srow = p.pe_cta ? (int64_t)blockIdx.x * 128 + e.r_in : e.row;
      float sa = 1.0f, inv = 1.0f;                           // scale of the operand currently in shared memory
      for (int l = 0; l < p.n_steps; ++l) {
        const ChainStep& G = p.S[l];
        CH_TRS(0);
        if (G.sync_before) epi_bar_sync();
        switch (G.kind) {
          case ST_PE: epi_step_pe<false>(G, p, ctl, a_smem, e, lp, sa, inv, tr_on, tr_role, l); break;
          case ST_EDOT: epi_step_pe<true>(G, p, ctl, a_smem, e, lp, sa, inv, tr_on, tr_role, l); break;
          case ST_FWD: epi_step_gemm<ST_FWD>(G, p, ctl, a_smem, e, af_cnt, lp, sa, inv, tr_on, tr_role, l); break;
          case ST_REV: epi_step_gemm<ST_REV>(G, p, ctl, a_smem, e, af_cnt, lp, sa, inv, tr_on, tr_role, l); break;
          case ST_TAN: epi_step_gemm<ST_TAN>(G, p, ctl, a_smem, e, af_cnt, lp, sa, inv, tr_on, tr_role, l); break;
          case ST_BWD: epi_step_gemm<ST_BWD>(G, p, ctl, a_smem, e, af_cnt, lp, sa, inv, tr_on, tr_role, l); break;
          case ST_FWD_LAST: epi_step_tail<ST_FWD_LAST>(G, p, ctl, e, af_cnt, sa, tr_on, tr_role, l); break;
          case ST_REV_FINAL: epi_step_tail<ST_REV_FINAL>(G, p, ctl, e, af_cnt, sa, tr_on, tr_role, l); break;
          case ST_LOAD: epi_step_nogemm<ST_LOAD>(G, p, ctl, a_smem, e, tr_on, tr_role, l); break;
          case ST_REV_SEED: epi_step_nogemm<ST_REV_SEED>(G, p, ctl, a_smem, e, tr_on, tr_role, l); break;
          default: break;
        }
        CH_TRS(4);
      }
      epi_bar_sync();     // every MMA of this point tile has completed (all warps waited for the last N tile): the A planes
                          // may be overwritten by the next point tile's first operand
    }
  } else if (warp == CH_MMA_WARP) {
    // =========================================== MMA issuer ===========================================
    {
      // the whole warp walks the loop; one elected lane issues (see elect_one)
      uint32_t wcnt = 0, ae_cnt[2] = {0u, 0u}, ar_cnt[CH_MAX_SLICES] = {0u, 0u, 0u, 0u};
      const uint32_t a_addr = smem_u32(a_smem), w_addr = smem_u32(w_smem);
      for (int64_t pt = blockIdx.x; pt < n_ptiles; pt += gridDim.x) {
        const bool tr_on = p.trace != nullptr && blockIdx.x == 0 && pt == blockIdx.x && lane == 0;
        for (int l = 0; l < p.n_steps; ++l) {
          const ChainStep& S = p.S[l];
          long long w_acc = 0, w_a = 0, w_w = 0, t0;
          CH_TR(2, l, 0, clock64());
          for (int t = 0; t < S.n_tiles; ++t) {
            const uint32_t slot = (uint32_t)(t & 1);
            t0 = clock64();
            mbar_wait(&ctl->acc_empty[slot], (ae_cnt[slot] & 1u) ^ 1u);       // the epilogue has drained this accumulator pair
            w_acc += clock64() - t0;
            ++ae_cnt[slot];
            tcgen05_fence_after();
            const uint32_t acc_m = tmem_base + slot * 256u, acc_c = acc_m + 128u;
            const uint32_t rows = (uint32_t)(S.rows_override > 0 ? S.rows_override : ch_tile_rows(S.N, t));
            const bool exact = S.n_wpl == 3;
            const uint32_t idesc = exact ? make_idesc_f16(rows) : make_idesc(rows);
            for (int s = 0; s < S.n_kslices; ++s, ++wcnt) {
              t0 = clock64();
              if (t == 0) { mbar_wait(&ctl->a_ready[s], ar_cnt[s] & 1u); ++ar_cnt[s]; }
              w_a += clock64() - t0;
              const uint32_t ws = wcnt % CH_NSLOT, wu = wcnt / CH_NSLOT;
              t0 = clock64();
              mbar_wait(&ctl->w_full[ws], wu & 1u);
              w_w += clock64() - t0;
              tcgen05_fence_after();
              // descriptors of the K step j = 0; step j adds 32 bytes = 2 units of the 16-byte address field
              const uint64_t da0 = make_desc(a_addr + (uint32_t)s * 2u * CH_APLANE), da1 = make_desc(a_addr + (uint32_t)s * 2u * CH_APLANE + CH_APLANE);
              const uint32_t wb = w_addr + ws * CH_WSLOT;
              const uint64_t db0 = make_desc(wb), db1 = make_desc(wb + rows * 128u), db2 = make_desc(wb + 2u * rows * 128u);
              if (elect_one()) {
                if (exact) {
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const uint32_t first = (s == 0 && j == 0) ? 0u : 1u;
                    const uint64_t o = (uint64_t)(2 * j);
                    mma_bf16(acc_m, da0 + o, db0 + o, idesc, first);            // exact main term
                    mma_bf16(acc_c, da1 + o, db1 + o, idesc, first);            // corrections, smallest first
                    mma_bf16(acc_c, da0 + o, db2 + o, idesc, 1u);
                    mma_bf16(acc_c, da1 + o, db0 + o, idesc, 1u);
                    mma_bf16(acc_c, da0 + o, db1 + o, idesc, 1u);
                  }
                } else {                                                         // split-bf16: lo hi + hi lo + hi hi, one accumulator
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    const uint32_t first = (s == 0 && j == 0) ? 0u : 1u;
                    const uint64_t o = (uint64_t)(2 * j);
                    mma_bf16(acc_m, da1 + o, db0 + o, idesc, first);
                    mma_bf16(acc_m, da0 + o, db1 + o, idesc, 1u);
                    mma_bf16(acc_m, da0 + o, db0 + o, idesc, 1u);
                  }
                }
                mma_commit(&ctl->w_empty[ws]);
              }
              __syncwarp();
            }
            if (elect_one()) mma_commit(&ctl->acc_full[slot]);
            __syncwarp();
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
        for (int l = 0; l < p.n_steps; ++l) {
          const ChainStep& S = p.S[l];
          for (int t = 0; t < S.n_tiles; ++t) {
            const int rows_full = ch_tile_rows(S.img_N, t);
            const uint32_t rows = (uint32_t)(S.rows_override > 0 ? S.rows_override : rows_full);
            const uint32_t bytes = rows * 128u;
            const uint16_t* tile = p.img + S.img_off + ch_tile_off(S.img_N, S.K, t);
            for (int s = 0; s < S.n_kslices; ++s, ++wcnt) {
              const uint32_t ws = wcnt % CH_NSLOT, wu = wcnt / CH_NSLOT;
              mbar_wait(&ctl->w_empty[ws], (wu & 1u) ^ 1u);
              mbar_arrive_expect_tx(&ctl->w_full[ws], (uint32_t)S.n_wpl * bytes);
              uint8_t* dst = w_smem + ws * CH_WSLOT;
              const uint16_t* src = tile + (int64_t)s * CH_WPL * rows_full * 64;
#pragma unroll
              for (int pl = 0; pl < CH_WPL; ++pl)
                if (pl < S.n_wpl) bulk_g2s(dst + pl * bytes, src + (int64_t)pl * rows_full * 64, bytes, &ctl->w_full[ws]);
            }
          }
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
static inline int launch_chain(const ChainParams& p, int family, cudaStream_t st) {
  if (p.P <= 0) return 0;
  static bool attr_set = false;
  if (!attr_set) {
    NUDF_CUDA_OK(cudaFuncSetAttribute(udf_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)CH_SMEM_BYTES));
    attr_set = true;
  }
  int64_t grid = (p.P + 127) / 128;
  if (grid > sm_count()) grid = sm_count();
  if (grid > CH_MAX_GRID) grid = CH_MAX_GRID;                // the per-CTA encoding stash is sized for CH_MAX_GRID tiles
  static int trace_mode = -1;
  if (trace_mode < 0) { const char* e = getenv("NUDF_CHAIN_TRACE"); trace_mode = (e && atoi(e) > 0) ? atoi(e) : 0; }
  if (trace_mode > 0 && p.P >= 128 * 148) {              // profiling aid: synchronous, prints CTA 0's pipeline stamps
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
    const long long z = h[0];
    fprintf(stderr, "[chain trace] P=%lld steps=%d (cycles rel. to the first step's start; epi0 = warp 0, epi1 = warp 12)\n", (long long)p.P, p.n_steps);
    for (int l = 0; l < p.n_steps; ++l) {
      const long long* e0 = h + (0 * CH_MAX_STEPS + l) * 8;
      const long long* e1 = h + (1 * CH_MAX_STEPS + l) * 8;
      const long long* m = h + (2 * CH_MAX_STEPS + l) * 8;
      fprintf(stderr, "  S%02d kind %d epi0: start %7lld acc %7lld p1 %7lld bar %7lld p2 %7lld | epi1: acc %7lld p1 %7lld p2 %7lld | mma: start %7lld end %7lld "
                      "wait acc_empty %6lld a_ready %6lld w_full %6lld\n",
              l, p.S[l].kind, e0[0] - z, e0[1] ? e0[1] - z : 0, e0[2] - z, e0[3] ? e0[3] - z : 0, e0[4] - z, e1[1] ? e1[1] - z : 0,
              e1[2] ? e1[2] - z : 0, e1[4] ? e1[4] - z : 0, m[0] ? m[0] - z : 0, m[1] ? m[1] - z : 0, m[2], m[3], m[4]);
    }
    --trace_mode;                                           // NUDF_CHAIN_TRACE = number of launches to trace
    return 0;
  }
  LaunchTimer lt_(family, st);
  udf_chain_kernel<<<(unsigned)grid, CH_THREADS, CH_SMEM_BYTES, st>>>(p);
  NUDF_LAUNCH_OK();
  return 0;
}

}  // namespace chain
}  // namespace nudf
