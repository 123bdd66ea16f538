// Host (g++) harness around neuraludf_b200/csrc/raymath.cuh: a sequential per-ray forward / backward that composes
// the SAME per-sample functions the CUDA kernels call, in the same order as ray_kernels.cu.  Lets the CPU-only dev
// box check the formulas and hand-derived derivatives against the oracle's autograd.  Test infrastructure.
#include <math.h>
#include <stdint.h>
#include <vector>

#include "../../neuraludf_b200/csrc/raymath.cuh"

using namespace nudf;

struct Cfg {
  int S, O;
  float inv_s, beta, gamma, r;
  int has_r;
  float fs, ssf;
  int use_norm;
  int has_bg_rgb;
  float bg_rgb[3];
};

static void forward_state(const Cfg& c, const float* d, const float* dists, const float* udf, const float* g,
                          const float* bg_alpha, std::vector<float>& tc, std::vector<float>& q, std::vector<float>& t,
                          std::vector<float>& P, std::vector<float>& ap, std::vector<float>& am, std::vector<float>& alpha,
                          std::vector<float>& T) {
  int S = c.S, SO = c.S + c.O;
  tc.resize(S); q.resize(S); t.resize(S); P.resize(S); ap.resize(S); am.resize(S); alpha.resize(SO); T.resize(SO);
  for (int i = 0; i < S; ++i) tc[i] = grad_quantities(g + 3 * i, d, c.use_norm).tc;
  float run = 1.0f;
  for (int i = 0; i < S; ++i) {
    float raw, aocc;
    occ_forward(udf[i], dists[i], c.beta, c.gamma, &raw, &aocc);
    float vm = (i + 1 < S) ? (tc[i + 1] < 0.01f ? 1.f : 0.f) : 1.f;
    q[i] = 1.0f - aocc + c.fs * vm;
    t[i] = clampf_(q[i], 0.f, 1.f) + 1e-7f;
    P[i] = run;
    run *= t[i];
  }
  for (int i = 0; i < SO; ++i) {
    if (i < S) {
      float ic = iter_cos_forward(tc[i], c.has_r, c.r);
      ap[i] = neus_alpha_forward(udf[i], ic, dists[i], c.inv_s);
      am[i] = neus_alpha_forward(-udf[i], ic, dists[i], c.inv_s);
      float vis = clampf_(P[i], 0.f, 1.f);
      alpha[i] = ap[i] * vis + am[i] * (1.f - vis);
    } else {
      alpha[i] = bg_alpha[i];
    }
  }
  run = 1.0f;
  for (int i = 0; i < SO; ++i) { T[i] = run; run *= (1.0f - alpha[i] + 1e-7f); }
}

extern "C" {

// out: color_base[3], color[3], depth, ws_fg, ws_all, sums[5]; weights[SO]
void ray_forward_host(const Cfg* c, const float* d, const float* pts, const float* mid, const float* dists,
                      const float* udf, const float* g, const float* scb, const float* sc, const float* bg_alpha,
                      const float* bg_color, float* out, float* weights) {
  std::vector<float> tc, q, t, P, ap, am, alpha, T;
  forward_state(*c, d, dists, udf, g, bg_alpha, tc, q, t, P, ap, am, alpha, T);
  int S = c->S, SO = c->S + c->O;
  for (int k = 0; k < 14; ++k) out[k] = 0.f;
  for (int i = 0; i < SO; ++i) {
    float w = alpha[i] * T[i];
    weights[i] = w;
    out[8] += w;
    if (i < S) {
      out[7] += w;
      GradQ gq = grad_quantities(g + 3 * i, d, c->use_norm);
      for (int k = 0; k < 3; ++k) { out[k] += w * scb[3 * i + k]; out[3 + k] += w * sc[3 * i + k]; }
      out[6] += w * mid[i];
      float pn = sqrtf(pts[3 * i] * pts[3 * i] + pts[3 * i + 1] * pts[3 * i + 1] + pts[3 * i + 2] * pts[3 * i + 2]);
      float relax = pn < 1.2f, near = udf[i] < 0.05f;
      float ge = (gq.gmag - 1.f) * (gq.gmag - 1.f);
      out[9] += relax * ge; out[10] += relax; out[11] += near * ge; out[12] += near;
      out[13] += expf(-c->ssf * udf[i]);
    } else {
      for (int k = 0; k < 3; ++k) { out[k] += w * bg_color[3 * i + k]; out[3 + k] += w * bg_color[3 * i + k]; }
    }
  }
  if (c->has_bg_rgb)
    for (int k = 0; k < 3; ++k) out[3 + k] += c->bg_rgb[k] * (1.0f - out[8]);
}

// bar: color_base[3], color[3], depth, ws_fg, ws_all ; coef: ge_coef, ge_ns_coef, sparse_coef (already normalised)
// outputs: udf_bar[S], g_bar[3S], scb_bar[3S], sc_bar[3S], bg_alpha_bar[SO], bg_color_bar[3SO], scalar_bar[3]
void ray_backward_host(const Cfg* c, const float* d, const float* pts, const float* mid, const float* dists,
                       const float* udf, const float* g, const float* scb, const float* sc, const float* bg_alpha,
                       const float* bg_color, const float* bar, const float* coef, float* udf_bar, float* g_bar,
                       float* scb_bar, float* sc_bar, float* bg_alpha_bar, float* bg_color_bar, float* scalar_bar) {
  std::vector<float> tc, q, t, P, ap, am, alpha, T;
  forward_state(*c, d, dists, udf, g, bg_alpha, tc, q, t, P, ap, am, alpha, T);
  int S = c->S, SO = c->S + c->O;
  const float* cbb = bar; const float* ccb = bar + 3;
  float depth_b = bar[6], wsfg_b = bar[7], wsall_b = bar[8];
  if (c->has_bg_rgb) wsall_b -= ccb[0] * c->bg_rgb[0] + ccb[1] * c->bg_rgb[1] + ccb[2] * c->bg_rgb[2];
  std::vector<float> wbar(SO), abar(SO);
  for (int i = 0; i < SO; ++i) {
    float w = alpha[i] * T[i];
    float wb = wsall_b;
    if (i < S) {
      for (int k = 0; k < 3; ++k) {
        wb += cbb[k] * scb[3 * i + k] + ccb[k] * sc[3 * i + k];
        scb_bar[3 * i + k] = w * cbb[k]; sc_bar[3 * i + k] = w * ccb[k];
      }
      wb += depth_b * mid[i] + wsfg_b;
    } else {
      for (int k = 0; k < 3; ++k) {
        wb += (cbb[k] + ccb[k]) * bg_color[3 * i + k];
        bg_color_bar[3 * i + k] = w * (cbb[k] + ccb[k]);
      }
    }
    wbar[i] = wb;
  }
  float B = 0.f;  // B_{i+1}
  for (int i = SO - 1; i >= 0; --i) {
    abar[i] = T[i] * (wbar[i] - B);
    B = wbar[i] * alpha[i] + (1.0f - alpha[i] + 1e-7f) * B;
  }
  for (int i = 0; i < SO; ++i) bg_alpha_bar[i] = i < S ? 0.f : abar[i];
  std::vector<float> tbar(S);
  float R = 0.f;  // R_{i+1}
  for (int i = S - 1; i >= 0; --i) {
    tbar[i] = P[i] * R;
    float Pb = (P[i] >= 0.f && P[i] <= 1.f) ? abar[i] * (ap[i] - am[i]) : 0.f;
    R = Pb + t[i] * R;
  }
  float s_bar = 0.f, beta_bar = 0.f, gamma_bar = 0.f;
  for (int i = 0; i < S; ++i) {
    float vis = clampf_(P[i], 0.f, 1.f);
    float ap_bar = abar[i] * vis, am_bar = abar[i] * (1.f - vis);
    float ic = iter_cos_forward(tc[i], c->has_r, c->r);
    float sb1, ib1, s1, sb2, ib2, s2;
    neus_alpha_backward(udf[i], ic, dists[i], c->inv_s, ap_bar, &sb1, &ib1, &s1);
    neus_alpha_backward(-udf[i], ic, dists[i], c->inv_s, am_bar, &sb2, &ib2, &s2);
    float u_bar = sb1 - sb2;
    s_bar += s1 + s2;
    float tc_bar = (ib1 + ib2) * iter_cos_dtc(tc[i], c->has_r, c->r);
    float q_bar = (q[i] >= 0.f && q[i] <= 1.f) ? tbar[i] : 0.f;
    float ub2, bb, gb;
    occ_backward(udf[i], dists[i], c->beta, c->gamma, -q_bar, &ub2, &bb, &gb);
    u_bar += ub2; beta_bar += bb; gamma_bar += gb;
    float pn = sqrtf(pts[3 * i] * pts[3 * i] + pts[3 * i + 1] * pts[3 * i + 1] + pts[3 * i + 2] * pts[3 * i + 2]);
    float eik = (pn < 1.2f ? coef[0] : 0.f) + (udf[i] < 0.05f ? coef[1] : 0.f);
    u_bar += coef[2] * (-c->ssf) * expf(-c->ssf * udf[i]);
    grad_quantities_backward(g + 3 * i, d, c->use_norm, tc_bar, eik, g_bar + 3 * i);
    udf_bar[i] = u_bar;
  }
  scalar_bar[0] = s_bar; scalar_bar[1] = beta_bar; scalar_bar[2] = gamma_bar;
}

}  // extern "C"
