// Per-view, per-pixel math of the pixel / patch blending stage (fine-tuning stage of NeuralUDF), shared by the CUDA kernels
// (blend.cu) and the host harness of the CPU tests (tests/host/blend_host.cpp).
// Semantics followed (reference): models/projector_utils.py:8-85 (projection, 'zeros' padding, validity mask),
// models/patch_projector.py:131-166 (homography warp of the patch pixels, inside-image mask, clamp to [-10, 10]),
// torch.nn.functional.grid_sample(mode='bilinear', padding_mode='zeros', align_corners=True), models/fields.py:498-537.
#pragma once
#include "common.cuh"

namespace nudf {

// bilinear sample of a [3, H, W] image at unnormalised coordinates (ix, iy); taps outside the image contribute zero
NUDF_HD void bilinear3(const float* img, int H, int W, float ix, float iy, float out[3]) {
  const float x0f = floorf(ix), y0f = floorf(iy);
  const float wx1 = ix - x0f, wy1 = iy - y0f, wx0 = (x0f + 1.0f) - ix, wy0 = (y0f + 1.0f) - iy;
  out[0] = out[1] = out[2] = 0.f;
  // the comparisons are done in float: coordinates of masked-out pixels can be far outside the int range
  const bool xin0 = x0f >= 0.f && x0f <= (float)(W - 1), xin1 = x0f + 1.0f >= 0.f && x0f + 1.0f <= (float)(W - 1);
  const bool yin0 = y0f >= 0.f && y0f <= (float)(H - 1), yin1 = y0f + 1.0f >= 0.f && y0f + 1.0f <= (float)(H - 1);
  if (!((xin0 || xin1) && (yin0 || yin1))) return;
  const int x0 = (int)x0f, y0 = (int)y0f;
  const int64_t plane = (int64_t)H * W;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float* p = img + c * plane;
    float v = 0.f;
    if (yin0 && xin0) v += p[(int64_t)y0 * W + x0] * (wx0 * wy0);
    if (yin0 && xin1) v += p[(int64_t)y0 * W + x0 + 1] * (wx1 * wy0);
    if (yin1 && xin0) v += p[(int64_t)(y0 + 1) * W + x0] * (wx0 * wy1);
    if (yin1 && xin1) v += p[(int64_t)(y0 + 1) * W + x0 + 1] * (wx1 * wy1);
    out[c] = v;
  }
}

// Projection of a world point into a view: proj = K[:3,:3] @ w2c[:3,:] (row-major 3x4).  Returns the validity mask
// (|x_norm| < 1 and |y_norm| < 1) and the unnormalised sampling coordinates (only meaningful when valid).
NUDF_HD bool pixel_project(const float* proj, const float* p, int H, int W, float* ix, float* iy) {
  const float X = proj[0] * p[0] + proj[1] * p[1] + proj[2] * p[2] + proj[3];
  const float Y = proj[4] * p[0] + proj[5] * p[1] + proj[6] * p[2] + proj[7];
  float Z = proj[8] * p[0] + proj[9] * p[1] + proj[10] * p[2] + proj[11];
  Z = Z < 1e-3f ? 1e-3f : Z;
  float xn = 2.0f * (X / Z) / (float)(W - 1) - 1.0f;
  float yn = 2.0f * (Y / Z) / (float)(H - 1) - 1.0f;
  if (xn > 1.0f || xn < -1.0f) xn = 2.0f;
  if (yn > 1.0f || yn < -1.0f) yn = 2.0f;
  *ix = ((xn + 1.0f) / 2.0f) * (float)(W - 1);
  *iy = ((yn + 1.0f) / 2.0f) * (float)(H - 1);
  return fabsf(xn) < 1.0f && fabsf(yn) < 1.0f;
}

// Homography warp of one patch pixel (u, v) of the query image into a source view (hom row-major 3x3).  Returns the
// inside-image mask (depth > 0 and at least h_patch pixels away from the border) and the sampling coordinates.
NUDF_HD bool patch_warp_pixel(const float* hom, float u, float v, int H, int W, int h_patch, float* ix, float* iy) {
  const float wx = hom[0] * u + hom[1] * v + hom[2];
  const float wy = hom[3] * u + hom[4] * v + hom[5];
  const float wz = hom[6] * u + hom[7] * v + hom[8];
  const float zc = wz < 1e-8f ? 1e-8f : wz;
  const float gx = wx / zc, gy = wy / zc;
  const bool m = wz > 0.f && gx < (float)(W - h_patch) && gy < (float)(H - h_patch) && gx >= (float)h_patch && gy >= (float)h_patch;
  float xn = 2.0f * gx / (float)(W - 1) - 1.0f;
  float yn = 2.0f * gy / (float)(H - 1) - 1.0f;
  xn = xn < -10.f ? -10.f : (xn > 10.f ? 10.f : xn);
  yn = yn < -10.f ? -10.f : (yn > 10.f ? 10.f : yn);
  *ix = ((xn + 1.0f) / 2.0f) * (float)(W - 1);
  *iy = ((yn + 1.0f) / 2.0f) * (float)(H - 1);
  return m;
}

}  // namespace nudf
