// Ray generation on the device: the per-step host-side tensor algebra of the reference's data loader
//   dataset/dataset.py:228-294  gen_random_rays_patches_at  (pixel -> camera ray -> world ray, colour / mask gather, ndc uv)
//   dataset/dataset.py:151-164  gen_rays_at                 (full-image ray grid for validation)
//   dataset/dataset.py:329-335  near_far_from_sphere
// as two fused kernels: one thread per ray, coalesced [N,10] / [N,3] writes, images gathered straight from the resident
// [H,W,3] tensors.  The random pixel indices stay in torch (RNG parity with the reference).  HBM-bound: 40 B (+24 B of
// gathers) per ray in, 56 B out.
#include "../../include/nudf.h"
#include "common.cuh"

namespace nudf {

__device__ __forceinline__ void pixel_ray(const float* __restrict__ Ki, const float* __restrict__ pose, float x, float y,
                                          float o[3], float d[3]) {
  // p = K^-1 [x, y, 1]; v = p / |p|; d = R v; o = t      (dataset.py:283-287)
  float p[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) p[r] = Ki[r * 3 + 0] * x + Ki[r * 3 + 1] * y + Ki[r * 3 + 2];
  const float inv = 1.0f / sqrtf(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  p[0] *= inv; p[1] *= inv; p[2] *= inv;
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    d[r] = pose[r * 4 + 0] * p[0] + pose[r * 4 + 1] * p[1] + pose[r * 4 + 2] * p[2];
    o[r] = pose[r * 4 + 3];
  }
}
__device__ __forceinline__ void sphere_near_far(const float o[3], const float d[3], float* near, float* far) {
  const float a = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  const float b = 2.0f * (o[0] * d[0] + o[1] * d[1] + o[2] * d[2]);
  const float mid = 0.5f * (-b) / a;
  *near = mid - 1.0f;
  *far = mid + 1.0f;
}

__global__ void gen_rays_kernel(const float* __restrict__ Ki, const float* __restrict__ pose, const int64_t* __restrict__ px,
                                const int64_t* __restrict__ py, int n, const float* __restrict__ image,
                                const float* __restrict__ mask, int H, int W, float* __restrict__ rays, float* __restrict__ uv,
                                float* __restrict__ near, float* __restrict__ far) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t x = px[i], y = py[i];
  float o[3], d[3];
  pixel_ray(Ki, pose, (float)x, (float)y, o, d);
  float* r = rays + (int64_t)i * 10;
  r[0] = o[0]; r[1] = o[1]; r[2] = o[2]; r[3] = d[0]; r[4] = d[1]; r[5] = d[2];
  const int64_t pix = (y * W + x) * 3;
  r[6] = image[pix + 0]; r[7] = image[pix + 1]; r[8] = image[pix + 2];
  r[9] = mask != nullptr ? (mask[pix] > 0.f ? 1.f : 0.f) : 1.f;
  if (uv != nullptr) {
    uv[i * 2 + 0] = 2.0f * (float)x / (float)(W - 1) - 1.0f;
    uv[i * 2 + 1] = 2.0f * (float)y / (float)(H - 1) - 1.0f;
  }
  if (near != nullptr) sphere_near_far(o, d, near + i, far + i);
}

// rays of the [Hl, Wl] pixel grid x_j = linspace(0, W-1, Wl)[j], y_i = linspace(0, H-1, Hl)[i]; outputs [Hl, Wl, 3]
__global__ void gen_rays_grid_kernel(const float* __restrict__ Ki, const float* __restrict__ pose, int W, int H, int Wl, int Hl,
                                     float* __restrict__ rays_o, float* __restrict__ rays_d, float* __restrict__ near,
                                     float* __restrict__ far) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Wl * Hl) return;
  const int i = idx / Wl, j = idx - i * Wl;
  // torch.linspace(0, W-1, Wl): start + step * j computed symmetrically from both ends (torch's kernel): use the same form
  const float sx = Wl > 1 ? (float)(W - 1) / (float)(Wl - 1) : 0.f;
  const float sy = Hl > 1 ? (float)(H - 1) / (float)(Hl - 1) : 0.f;
  const float x = j < Wl / 2 ? sx * (float)j : (float)(W - 1) - sx * (float)(Wl - 1 - j);
  const float y = i < Hl / 2 ? sy * (float)i : (float)(H - 1) - sy * (float)(Hl - 1 - i);
  float o[3], d[3];
  pixel_ray(Ki, pose, x, y, o, d);
#pragma unroll
  for (int c = 0; c < 3; ++c) { rays_o[(int64_t)idx * 3 + c] = o[c]; rays_d[(int64_t)idx * 3 + c] = d[c]; }
  if (near != nullptr) sphere_near_far(o, d, near + idx, far + idx);
}

}  // namespace nudf

using namespace nudf;

extern "C" {

int nudf_gen_rays(const float* intrinsics_inv, const float* pose, const int64_t* px, const int64_t* py, int32_t n,
                  const float* image, const float* mask, int32_t H, int32_t W, float* rays, float* ndc_uv, float* near,
                  float* far, void* stream) {
  if (n <= 0) return 0;
  NUDF_REQUIRE(intrinsics_inv && pose && px && py && image && rays, "null pointer");
  NUDF_REQUIRE(H > 1 && W > 1, "image size");
  NUDF_REQUIRE((near == nullptr) == (far == nullptr), "near and far go together");
  LaunchTimer lt_(FAM_RAY, (cudaStream_t)stream);
  gen_rays_kernel<<<(unsigned)cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(intrinsics_inv, pose, px, py, n, image, mask, H, W, rays,
                                                                           ndc_uv, near, far);
  NUDF_LAUNCH_OK();
  return 0;
}

int nudf_gen_rays_grid(const float* intrinsics_inv, const float* pose, int32_t W, int32_t H, int32_t Wl, int32_t Hl, float* rays_o,
                       float* rays_d, float* near, float* far, void* stream) {
  if (Wl <= 0 || Hl <= 0) return 0;
  NUDF_REQUIRE(intrinsics_inv && pose && rays_o && rays_d, "null pointer");
  NUDF_REQUIRE((near == nullptr) == (far == nullptr), "near and far go together");
  LaunchTimer lt_(FAM_RAY, (cudaStream_t)stream);
  gen_rays_grid_kernel<<<(unsigned)cdiv((int64_t)Wl * Hl, 256), 256, 0, (cudaStream_t)stream>>>(intrinsics_inv, pose, W, H, Wl, Hl, rays_o,
                                                                                                rays_d, near, far);
  NUDF_LAUNCH_OK();
  return 0;
}

}  // extern "C"
