// Host (g++) harness around neuraludf_b200/csrc/blendmath.cuh: a sequential per-point forward / backward that composes the
// SAME per-view, per-pixel functions the CUDA kernel (csrc/blend.cu) calls, with the same masked-softmax algebra.  Lets the
// CPU-only dev box check the semantics (projection, masks, bilinear taps, fusion, logits gradient) against the op-by-op
// torch form.  Test infrastructure.
#include <math.h>
#include <stdint.h>
#include <vector>

#include "../../neuraludf_b200/csrc/blendmath.cuh"

using namespace nudf;

struct Cfg { int32_t n_rays, n_samples, n_views, height, width, h_patch; };

static void point(const Cfg& c, int64_t p, const float* pts, const float* proj, const float* hom, const float* px,
                  const float* imgs, const float* logits, int64_t ld, const float* g_pix, const float* g_pat, float* c_pix,
                  float* c_pat, float* m_pat, float* g_logits) {
  const int V = c.n_views, H = c.height, W = c.width;
  const int64_t P = (int64_t)c.n_rays * c.n_samples, istr = (int64_t)3 * H * W;
  std::vector<float> sm(V), dsm(V, 0.f);
  float mx = -INFINITY, se = 0.f;
  for (int v = 0; v < V; ++v) mx = fmaxf(mx, logits[p * ld + v]);
  for (int v = 0; v < V; ++v) { sm[v] = expf(logits[p * ld + v] - mx); se += sm[v]; }
  for (int v = 0; v < V; ++v) sm[v] /= se;
  // pixel
  {
    std::vector<float> a(V, 0.f), col(3 * V, 0.f);
    float A = 0.f;
    for (int v = 0; v < V; ++v) {
      float ix, iy;
      if (pixel_project(proj + v * 12, pts + p * 3, H, W, &ix, &iy)) { bilinear3(imgs + v * istr, H, W, ix, iy, &col[3 * v]); a[v] = sm[v]; }
      A += a[v];
    }
    float cp[3] = {0.f, 0.f, 0.f};
    for (int v = 0; v < V; ++v) for (int k = 0; k < 3; ++k) cp[k] += a[v] / (A + 1e-8f) * col[3 * v + k];
    if (c_pix) for (int k = 0; k < 3; ++k) c_pix[p * 3 + k] = cp[k];
    if (g_pix) {
      const float* g = g_pix + p * 3;
      const float gc = g[0] * cp[0] + g[1] * cp[1] + g[2] * cp[2];
      for (int v = 0; v < V; ++v)
        if (a[v] != 0.f) dsm[v] += (g[0] * col[3 * v] + g[1] * col[3 * v + 1] + g[2] * col[3 * v + 2] - gc) / (A + 1e-8f);
    }
  }
  if (hom) {
    const int h = c.h_patch, side = 2 * h + 1, npx = side * side;
    const int64_t n = p / c.n_samples;
    std::vector<float> acc(3 * npx, 0.f), colv(3 * npx), tv(V, 0.f);
    std::vector<char> valid(V, 0);
    float Apat = 0.f;
    for (int v = 0; v < V; ++v) {
      const float* hm = hom + ((int64_t)v * P + p) * 9;
      bool all_in = true;
      std::vector<float> ix(npx), iy(npx);
      for (int q = 0; q < npx; ++q) {
        const float u = px[n * 2] + (float)(q % side - h), w = px[n * 2 + 1] + (float)(q / side - h);
        all_in = patch_warp_pixel(hm, u, w, H, W, h, &ix[q], &iy[q]) && all_in;
      }
      if (!all_in) continue;
      valid[v] = 1;
      Apat += sm[v];
      float t = 0.f;
      for (int q = 0; q < npx; ++q) {
        bilinear3(imgs + v * istr, H, W, ix[q], iy[q], &colv[3 * q]);
        for (int k = 0; k < 3; ++k) {
          acc[3 * q + k] += sm[v] * colv[3 * q + k];
          if (g_pat) t += g_pat[(p * npx + q) * 3 + k] * colv[3 * q + k];
        }
      }
      tv[v] = t;
    }
    const float inv = 1.0f / (Apat + 1e-8f);
    if (c_pat) for (int i = 0; i < 3 * npx; ++i) c_pat[p * npx * 3 + i] = acc[i] * inv;
    if (m_pat) m_pat[p] = Apat > 0.f ? 1.f : 0.f;
    if (g_pat) {
      float gc = 0.f;
      for (int i = 0; i < 3 * npx; ++i) gc += g_pat[p * npx * 3 + i] * acc[i] * inv;
      for (int v = 0; v < V; ++v)
        if (valid[v]) dsm[v] += (tv[v] - gc) * inv;
    }
  }
  if (g_logits) {
    float s = 0.f;
    for (int v = 0; v < V; ++v) s += sm[v] * dsm[v];
    for (int v = 0; v < V; ++v) g_logits[p * V + v] = sm[v] * (dsm[v] - s);
  }
}

extern "C" {
void blend_host_forward(const Cfg* c, const float* pts, const float* proj, const float* hom, const float* px, const float* imgs,
                        const float* logits, int64_t ld, float* c_pix, float* c_pat, float* m_pat) {
  for (int64_t p = 0; p < (int64_t)c->n_rays * c->n_samples; ++p)
    point(*c, p, pts, proj, hom, px, imgs, logits, ld, nullptr, nullptr, c_pix, c_pat, m_pat, nullptr);
}
void blend_host_backward(const Cfg* c, const float* pts, const float* proj, const float* hom, const float* px, const float* imgs,
                         const float* logits, int64_t ld, const float* g_pix, const float* g_pat, float* g_logits) {
  for (int64_t p = 0; p < (int64_t)c->n_rays * c->n_samples; ++p)
    point(*c, p, pts, proj, hom, px, imgs, logits, ld, g_pix, g_pat, nullptr, nullptr, nullptr, g_logits);
}
}
