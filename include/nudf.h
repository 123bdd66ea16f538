/* nudf.h -- C ABI of libnudf.so, the B200 (sm_100a) implementation of NeuralUDF's volume-rendering hot path.
 *
 * The reference (xxlong0/NeuralUDF) is pure PyTorch and has NO plugin / FFI layer;
 * its "boundary" for this path is the Python surface `exp_runner_blending.py:15-19` imports.  This header is the
 * C-ABI a drop-in replacement binds instead (INTEGRATION.md shows the ctypes stub): every entry point below cites
 * the reference interface it replaces.  Conventions:
 *   - all pointers are DEVICE pointers into caller-owned (torch) storages unless marked "host";
 *   - fp32, row-major, explicit leading dimensions (ld, in floats);
 *   - the library never allocates device memory: callers size workspaces with the *_floats() queries;
 *   - every call enqueues on `stream` (a cudaStream_t passed as void*) and returns immediately;
 *   - return value: 0 ok, -1 invalid argument, -2 CUDA error; nudf_last_error() has the text (thread-local).
 */
#ifndef NUDF_H_
#define NUDF_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NUDF_MAX_LAYERS 16
#define NUDF_ABI_VERSION 3
/* bits of the device-side status word (nudf_render_out.status, `status` of the sampling entry points): set by the kernels,
 * never cleared by the library; the caller reads it at a host synchronisation point of its choice */
#define NUDF_STATUS_NONFINITE_SAMPLES 1   /* sample_pdf / up_sample produced a non-finite sample position (:97-101, 265-269) */
#define NUDF_STATUS_NONFINITE_RENDER 2    /* compositing produced a non-finite per-ray result (:543-544) */

int nudf_abi_version(void);
const char* nudf_last_error(void);
/* GEMM engine for the wide (K,N in {128,256}) layers: 0 = exact-fp32 FFMA, 1 = tcgen05 3xBF16 split (default 1
 * when built with the tensor path).  Small / odd-shaped contractions always use the FFMA engine. */
int nudf_set_engine(int engine);
int nudf_get_engine(void);
/* Which contraction chains may use the tensor engine (bit mask; default 126 = everything but the forward value chains:
 * the udf head feeds exp(-25000 u) and needs fp32-grade accuracy, and the ReLU gates of the colour / NeRF++ networks
 * must not flip more often than under fp32 rounding): 1 UDF forward value, 2 reverse sweep (grad_x udf), 4 tangent,
 * 8 backward, 16 weight gradients, 32 colour-network backward, 64 NeRF++ backward, 128 colour / NeRF++ forward. */
int nudf_set_tc_mask(int mask);
int nudf_get_tc_mask(void);
int nudf_default_tc_mask(void);   /* the mask the library ships (what bench.py times and the parity suite pins) */
/* Plane-fed reverse-sweep / tangent chains (engine 1 only): intermediate tensors of those chains are kept as split-bf16
 * plane tensors (see nudf_pack_planes) and fetched with cp.async.bulk.  Default 0 (env NUDF_PLANES).  Like the engine
 * and the mask it must not change between a forward call and its backward (the ctx layout depends on it). */
int nudf_set_chain_planes(int on);
int nudf_get_chain_planes(void);
/* --- tensor engine building blocks (unit-tested on their own) ---
 * weight image: bf16 hi/lo split of B(n,k) in UMMA shared-memory order; `transposed` selects B(n,k) = W[k*ldw+n]. */
/* planes: 2 = hi/lo (3 products, ~4e-6 vs fp64), 3 = hi/mid/lo (6 products, fp32-grade; used by the value chain) */
int64_t nudf_tc_image_elems(int32_t N, int32_t K, int32_t planes);
int nudf_tc_prepare_weights(const float* W, int64_t ldw, int32_t N, int32_t K, int32_t transposed, int32_t planes,
                            uint16_t* img, void* stream);
int nudf_dense_forward_tc(const float* X, int64_t ldx, const uint16_t* img, int32_t planes, const float* bias, float* Y,
                          int64_t ldy, int64_t M, int32_t N, int32_t K, int32_t act, void* stream);
/* dW[n_out, n_in] += dZ[P, n_out]^T X[P, n_in]  (engine 0: fp32 FFMA, 1: tcgen05) */
int nudf_wgrad(const float* dZ, int64_t ldz, const float* X, int64_t ldx, int32_t n_out, int32_t n_in, int64_t P,
               float* dW, int64_t ldw, int32_t engine, void* stream);

/* Split-bf16 plane tensors (csrc/gemm_pl.cuh): a fp32 matrix [rows x cols] stored as hi = bf16(x), lo = bf16(x - hi) in
 * 64 x 64 blocks, [row block][col block][plane][64 rows x 128 B SWIZZLE_128B]; 1024-byte aligned; pad rows are zero.
 * The tcgen05 kernels fetch these blocks with cp.async.bulk and use them as K-major (layer chains) or MN-major
 * (weight gradients) operands without conversion.  nudf_planes_elems: uint16 elements to allocate. */
int64_t nudf_planes_elems(int64_t rows, int32_t cols);
int nudf_pack_planes(const float* X, int64_t ldx, int64_t rows, int32_t cols, uint16_t* planes, void* stream);
int nudf_unpack_planes(const uint16_t* planes, int64_t rows, int32_t cols, float* X, int64_t ldx, void* stream);
/* Y = act(X W^T + b), X given as a plane tensor allocated for round_up(M, 128) rows, W as a 2-plane weight image
 * (nudf_tc_prepare_weights, transposed = 0), K <= 256: the plane-fed weights-resident chain kernel on one layer. */
int nudf_dense_forward_planes(const uint16_t* X_planes, const uint16_t* img, const float* bias, float* Y, int64_t ldy, int64_t M,
                              int32_t N, int32_t K, int32_t act, void* stream);
/* dW[n_out, n_in] += dZ[P, n_out]^T X[P, n_in], both operands plane tensors (replaces the autograd weight gradient of one
 * nn.Linear, models/fields.py:185; tensor engine only). */
int nudf_wgrad_planes(const uint16_t* dZ_planes, const uint16_t* X_planes, int32_t n_out, int32_t n_in, int64_t P, float* dW,
                      int64_t ldw, void* stream);
/* profiling aid: with NUDF_TC_DEBUG & 16 the persistent tensor kernel records clock64() stamps of CTA 0 (4 roles x 256) */
int nudf_tc_read_trace(long long* host_buf);
/* number of CUDA kernels this library has launched in this process (bench.py reports it as gpu_launches) */
int64_t nudf_launch_count(void);
/* Per-kernel-family device times for the benchmark's roofline table: while enabled, the library brackets its launches with
 * cudaEvent pairs on the launching stream.  nudf_read_launch_timing synchronises, writes the summed milliseconds and the
 * launch counts per family (host arrays of nudf_launch_family_count() entries; order: fused UDF value chain, tcgen05
 * reverse-sweep / tangent / backward / other layers, tcgen05 weight gradients, fp32 FFMA GEMMs, ray kernels, element-wise,
 * fused tangent + backward chains)
 * and clears the record. */
int nudf_launch_family_count(void);
int nudf_set_launch_timing(int on);
int nudf_read_launch_timing(float* ms_per_family, int32_t* launches_per_family);
/* One fused dense layer Y[M,N] = act(X[M,K] W[N,K]^T + bias), act: 0 none, 1 relu, 2 softplus(beta=100), 3 sigmoid.
 * The building block of every network below (an nn.Linear + activation of the reference, e.g. fields.py:205-208). */
int nudf_dense_forward(const float* X, int64_t ldx, const float* W, int64_t ldw, const float* bias, float* Y, int64_t ldy,
                       int64_t M, int32_t N, int32_t K, int32_t act, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * UDFNetwork  (reference: models/fields.py:115-231; forward :192-211, gradient :219-231)
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct nudf_udf_desc {
  int32_t n_lin;      /* number of linear layers = n_layers + 1                                   */
  int32_t d_in;       /* 3                                                                        */
  int32_t multires;   /* positional-encoding octaves L (models/embedder.py:39-51)                 */
  int32_t d_out;      /* 257 = 1 (udf) + feature width                                            */
  int32_t skip_layer; /* layer whose input is cat(h, PE)/sqrt(2) (fields.py:202-203), -1 if none  */
  float scale;        /* fields.py:193                                                            */
  int32_t in_dim[NUDF_MAX_LAYERS];
  int32_t out_dim[NUDF_MAX_LAYERS];
  const float* weight_g[NUDF_MAX_LAYERS]; /* [out,1]  legacy weight_norm g  (lin{l}.weight_g)     */
  const float* weight_v[NUDF_MAX_LAYERS]; /* [out,in] legacy weight_norm v  (lin{l}.weight_v)     */
  const float* bias[NUDF_MAX_LAYERS];     /* [out]                          (lin{l}.bias)         */
} nudf_udf_desc;

/* floats needed for the folded weights W_l = g_l v_l/||v_l|| (rows padded to a multiple of 4 floats) */
int64_t nudf_udf_folded_floats(const nudf_udf_desc* d);
/* folds weight-norm once per optimiser step (replaces torch._weight_norm inside every nn.Linear call) */
int nudf_udf_fold_weights(const nudf_udf_desc* d, float* wfold, void* stream);
/* floats of the activation context saved by forward for `P` points (with_grad: also the reverse-sweep tensors) */
int64_t nudf_udf_ctx_floats(const nudf_udf_desc* d, int64_t P, int with_grad);
/* floats of the scratch needed by backward */
int64_t nudf_udf_scratch_floats(const nudf_udf_desc* d, int64_t P);
/* out[P, ld_out] <- cat(|y0|/scale, y[1:])   (UDFNetwork.forward, fields.py:192-211)
 * grad[P,3]     <- d udf / d x exactly, by a reverse sweep (UDFNetwork.gradient, fields.py:219-231); may be NULL.
 * ctx           <- saved activations (required; sized by nudf_udf_ctx_floats(d, P, grad != NULL)). */
int nudf_udf_forward(const nudf_udf_desc* d, const float* wfold, const float* pts, int64_t P, float* out,
                     int64_t ld_out, float* grad, float* ctx, void* stream);
/* Value only (no context kept): udf[P] <- |y0|/scale.  Used by importance sampling and grid queries
 * (udf_renderer_blending.py:731-733, 282-284; extract_mesh.py:60-73). `work` needs nudf_udf_ctx_floats(d,P,0). */
int nudf_udf_value(const nudf_udf_desc* d, const float* wfold, const float* pts, int64_t P, float* udf, float* work,
                   void* stream);
/* Parameter gradients for  L = <out_bar, out> + <grad_bar, grad>  (first- and second-order terms; the latter is
 * what autograd's double backward through create_graph=True computes in the reference).  out_bar [P, ld_ob] and
 * grad_bar [P,3] may each be NULL (= zero).  dwfold (same layout as wfold) and dbias (sum of out_dim floats) are
 * OVERWRITTEN.  ctx must come from nudf_udf_forward with grad != NULL on the same points. */
int nudf_udf_backward(const nudf_udf_desc* d, const float* wfold, const float* pts, int64_t P, const float* out_bar,
                      int64_t ld_ob, const float* grad_bar, const float* ctx, float* scratch, float* dwfold,
                      float* dbias, void* stream);
/* The same pair with the value and the feature part as SEPARATE tensors (udf [P], feat [P, ld_feat]; udf_bar [P], feat_bar
 * [P, ld_fb], each may be NULL = zero): what render_core consumes (udf_renderer_blending.py:364-366 slices udf_nn_output[:, :1]
 * and [:, 1:]) without an odd-width [P, 257] tensor in between and without re-assembling its gradient. */
int nudf_udf_forward_split(const nudf_udf_desc* d, const float* wfold, const float* pts, int64_t P, float* udf, float* feat,
                           int64_t ld_feat, float* grad, float* ctx, void* stream);
int nudf_udf_backward_split(const nudf_udf_desc* d, const float* wfold, const float* pts, int64_t P, const float* udf_bar,
                            const float* feat_bar, int64_t ld_fb, const float* grad_bar, const float* ctx, float* scratch,
                            float* dwfold, float* dbias, void* stream);
/* weight-norm backward: dwfold -> (dg[l] [out,1], dv[l] [out,in]) for every layer (overwrites) */
int nudf_udf_unfold_grads(const nudf_udf_desc* d, const float* dwfold, float* const* dg, float* const* dv, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * ResidualRenderingNetwork, mode 'no_normal'  (reference: models/fields.py:400-495)
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct nudf_color_desc {
  int32_t n_lin;        /* n_layers + 1 (5) */
  int32_t d_feature;    /* 256 */
  int32_t d_hidden;     /* 128 */
  int32_t d_out;        /* 3   */
  int32_t n_blend;      /* blending_cand_views (10) */
  int32_t multires_view;/* 4   */
  const float* base_g[NUDF_MAX_LAYERS]; const float* base_v[NUDF_MAX_LAYERS]; const float* base_b[NUDF_MAX_LAYERS];
  const float* main_g[NUDF_MAX_LAYERS]; const float* main_v[NUDF_MAX_LAYERS]; const float* main_b[NUDF_MAX_LAYERS];
} nudf_color_desc;

int64_t nudf_color_folded_floats(const nudf_color_desc* d);
int nudf_color_fold_weights(const nudf_color_desc* d, float* wfold, void* stream);
int64_t nudf_color_ctx_floats(const nudf_color_desc* d, int64_t P);
int64_t nudf_color_scratch_floats(const nudf_color_desc* d, int64_t P);
/* color_base[P,3], color[P,3], blend[P,n_blend] <- forward(points, view_dirs, feature_vectors) (fields.py:452-495).
 * dirs is [P,3] when samples_per_ray <= 1; otherwise it is [P / samples_per_ray, 3] and row p uses
 * dirs[p / samples_per_ray] (the reference materialises that expansion of rays_d, udf_renderer_blending.py:359-362). */
int nudf_color_forward(const nudf_color_desc* d, const float* wfold, const float* pts, const float* dirs,
                       int32_t samples_per_ray, const float* feat, int64_t ld_feat, int64_t P, float* color_base,
                       float* color, float* blend, float* ctx, void* stream);
/* grads wrt parameters (dwfold/dbias overwritten) and wrt feature_vectors (dfeat [P, ld_df] overwritten).
 * cb_bar/c_bar [P,3], blend_bar [P,n_blend]; any may be NULL (= zero). */
int nudf_color_backward(const nudf_color_desc* d, const float* wfold, int64_t P, const float* cb_bar,
                        const float* c_bar, const float* blend_bar, const float* ctx, float* scratch, float* dfeat,
                        int64_t ld_df, float* dwfold, float* dbias, void* stream);
int nudf_color_unfold_grads(const nudf_color_desc* d, const float* dwfold, float* const* dg_base, float* const* dv_base,
                            float* const* dg_main, float* const* dv_main, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * NeRF++ background network  (reference: models/fields.py:541-628, use_viewdirs=True)
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct nudf_nerf_desc {
  int32_t D;             /* 8   */
  int32_t W;             /* 256 */
  int32_t d_in;          /* 4   */
  int32_t multires;      /* 10  */
  int32_t multires_view; /* 4   */
  int32_t skip;          /* skips[0] = 4: output of layer `skip` is concatenated as cat(PE, h) */
  const float* pts_w[NUDF_MAX_LAYERS]; const float* pts_b[NUDF_MAX_LAYERS];
  const float* views_w; const float* views_b;
  const float* feature_w; const float* feature_b;
  const float* alpha_w; const float* alpha_b;
  const float* rgb_w; const float* rgb_b;
} nudf_nerf_desc;

/* tensor-engine weight images of the NeRF layers (floats); build with nudf_nerf_prepare once per optimiser step and pass
 * to forward/backward (NULL = exact-fp32 engine) */
int64_t nudf_nerf_image_floats(const nudf_nerf_desc* d);
int nudf_nerf_prepare(const nudf_nerf_desc* d, float* wimg, void* stream);
int64_t nudf_nerf_ctx_floats(const nudf_nerf_desc* d, int64_t P);
int64_t nudf_nerf_scratch_floats(const nudf_nerf_desc* d, int64_t P);
/* sigma[P], rgb[P,3] <- NeRF.forward(pts4, view_dirs) (fields.py:599-628; no sigmoid on rgb) */
int nudf_nerf_forward(const nudf_nerf_desc* d, const float* wimg, const float* pts, const float* dirs,
                      int32_t samples_per_ray, int64_t P, float* sigma, float* rgb, float* ctx, void* stream);
/* parameter gradients; dparams[] order: pts_w[0], pts_b[0], ..., views_w, views_b, feature_w, feature_b, alpha_w,
 * alpha_b, rgb_w, rgb_b  (each overwritten, same shape as the parameter). */
int nudf_nerf_backward(const nudf_nerf_desc* d, const float* wimg, int64_t P, const float* sigma_bar,
                       const float* rgb_bar, const float* ctx, float* scratch, float* const* dparams, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * render_core ray kernels  (reference: models/udf_renderer_blending.py:327-584)
 * ------------------------------------------------------------------------------------------------------------ */
typedef struct nudf_render_cfg {
  int32_t n_rays;           /* N */
  int32_t n_samples;        /* S  (columns of z_vals entering render_core) */
  int32_t n_outside;        /* O  (0 when no NeRF++ background is composited) */
  float sample_dist;        /* last-interval length, :353 */
  float cos_anneal_ratio;   /* :296-297; used only when has_cos_anneal != 0 */
  int32_t has_cos_anneal;
  float flip_saturation;    /* :409 */
  float sparse_scale_factor;/* :553 */
  int32_t use_norm_grad_for_cosine; /* :380-383 */
  int32_t has_background_rgb; float background_rgb[3]; /* :527-528 */
} nudf_render_cfg;

/* pts[N*S,3], mid_z[N,S], dists[N,S] <- rays and z_vals (:352-362) */
int nudf_ray_points(const float* rays_o, const float* rays_d, const float* z_vals, int32_t n_rays, int32_t n_samples,
                    float sample_dist, float* pts, float* mid_z, float* dists, void* stream);

/* per-sample / per-ray output pointers of the compositing pass; any per-sample pointer may be NULL */
typedef struct nudf_render_out {
  float* color_base; float* color; float* depth; float* normals;  /* [N,3] [N,3] [N,1] [N,3] */
  float* weights;                                                 /* [N,S+O] */
  float* weight_sum; float* weight_sum_fg_bg;                     /* [N,1] */
  float* ray_sums;  /* [N,5]: sum relax*(|g|-1)^2, sum relax, sum near*(|g|-1)^2, sum near, sum exp(-k udf) */
  float* gradient_mag; float* true_cos; float* vis_prob; float* alpha; float* alpha_plus; float* alpha_minus;
  float* alpha_occ; float* raw_occ; float* inside_sphere;        /* [N,S] each */
  float* gradients_flip;                                         /* [N,S,3] */
  int32_t* status;  /* DEVICE int or NULL: NUDF_STATUS_NONFINITE_RENDER is OR-ed in when a ray's colour / depth / weight sum /
                     * regulariser sums are not finite (the reference traps this with pdb, :543-544) */
} nudf_render_out;

/* alpha compositing with the visibility-weighted UDF density (:364-553 minus the networks).
 * heads: DEVICE float[3] = (inv_s, beta, gamma) after the clips of :373-377 (kept on the device so that a training
 * step needs no host synchronisation).
 * udf [N*S] (stride ld_udf floats between samples), grads [N*S,3], sampled colours [N*S,3] x2,
 * bg_alpha [N,S+O] / bg_color [N,S+O,3] from render_core_outside (only columns >= S are read), may be NULL. */
int nudf_render_composite_forward(const nudf_render_cfg* cfg, const float* heads, const float* rays_d, const float* pts,
                                  const float* mid_z, const float* dists, const float* udf, int64_t ld_udf,
                                  const float* grads, const float* sampled_color_base, const float* sampled_color,
                                  const float* bg_alpha, const float* bg_color, const nudf_render_out* out,
                                  void* stream);

typedef struct nudf_render_bar {   /* upstream gradients, any may be NULL */
  const float* color_base; const float* color; const float* depth;       /* [N,3] [N,3] [N,1] */
  const float* weight_sum; const float* weight_sum_fg_bg;                 /* [N,1] */
  const float* weights;  /* [N,S+O] d loss / d weights (pixel / patch blending composites are formed from `weights`) */
  const float* ray_sums; /* [N,5] d loss / d ray_sums (columns 1 and 3, the detached mask counts, are ignored) */
} nudf_render_bar;

/* Backward of the compositing pass.  Outputs (overwritten): udf_bar [N*S], grads_bar [N*S,3], scb_bar/sc_bar
 * [N*S,3], bg_alpha_bar [N,S+O] / bg_color_bar [N,S+O,3] (may be NULL), scalar_bar[N,3] = per-ray partial
 * d/d(inv_s, beta, gamma) (caller sums over rays). */
int nudf_render_composite_backward(const nudf_render_cfg* cfg, const float* heads, const float* rays_d, const float* pts,
                                   const float* mid_z, const float* dists, const float* udf, int64_t ld_udf,
                                   const float* grads, const float* sampled_color_base, const float* sampled_color,
                                   const float* bg_alpha, const float* bg_color, const nudf_render_bar* bar,
                                   float* udf_bar, float* grads_bar, float* scb_bar,
                                   float* sc_bar, float* bg_alpha_bar, float* bg_color_bar, float* scalar_bar,
                                   void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * hierarchical sampling  (reference: models/udf_renderer_blending.py:66-104, 197-290, 723-755, 834-866)
 * ------------------------------------------------------------------------------------------------------------ */
/* One up-sampling round: new_z[N,m] (and optionally the searchsorted indices inds[N,m], int64) from z[N,n], udf[N,n].
 * mode 0 = up_sample_unbias (:197-272), 1 = up_sample_no_occ_aware (:834-866).  Scans are accumulated in fp64 and
 * rounded to fp32 per element, like torch's CPU cumsum/cumprod, so that indices are reproducible.
 * u_lin: DEVICE float[m] = torch.linspace(0.5/m, 1-0.5/m, m) (:76), supplied by the caller so that its rounding is
 * exactly torch's.  status: DEVICE int or NULL, see NUDF_STATUS_NONFINITE_SAMPLES. */
int nudf_up_sample(int32_t mode, const float* rays_o, const float* rays_d, const float* z, const float* udf,
                   int32_t n_rays, int32_t n, int32_t m, float sample_dist, float inv_s, float beta, float gamma,
                   const float* u_lin, float* new_z, int64_t* inds, int32_t* status, void* stream);
/* sample_pdf(det=True) alone (:66-104): bins [N,n], weights [N,n-1] -> samples [N,m], inds [N,m] */
int nudf_sample_pdf(const float* bins, const float* weights, int32_t n_rays, int32_t n, int32_t m, const float* u_lin,
                    float* samples, int64_t* inds, int32_t* status, void* stream);
/* cat_z_vals merge (:274-290): z_out[N,n+m] sorted union, udf_out gathered likewise (udf/new_udf/udf_out may be
 * NULL for the `last` round).  new_pts[N*m,3] <- o + d*new_z (points to evaluate before the merge), optional. */
int nudf_merge_z(const float* z, const float* new_z, const float* udf, const float* new_udf, int32_t n_rays, int32_t n,
                 int32_t m, float* z_out, float* udf_out, void* stream);
int nudf_points_on_rays(const float* rays_o, const float* rays_d, const float* z, int32_t n_rays, int32_t n,
                        float* pts, void* stream);
/* NeRF++ inverted-sphere inputs and alpha (:161-184): pts4[N*n,4], then alpha = 1-exp(-relu(sigma) dists) */
int nudf_outside_points(const float* rays_o, const float* rays_d, const float* z, int32_t n_rays, int32_t n,
                        int32_t col0, float sample_dist, float* pts4, float* dists, void* stream);

/* --- pixel / patch blending of the fine-tuning stage (replaces patch_projector.pixel_warp / patch_warp +
 * fields.color_blend, models/patch_projector.py:21-166, models/projector_utils.py:8-85, models/fields.py:498-537, as used by
 * render_core, models/udf_renderer_blending.py:431-480).  One call fuses, for every sample point, the projection into each
 * source view, the bilinear gathers (pixel colour and homography-warped patch) and the masked-softmax fusion over views.
 *   pts    [P,3]  sample points (P = n_rays * n_samples, ray-major)
 *   proj   [V,12] row-major 3x4 matrices K[:3,:3] @ w2c[:3,:] of the source views
 *   hom    [V,P,9] row-major plane-induced homographies query pixel -> source pixel, or NULL (pixel blending only)
 *   px     [n_rays,2] pixel coordinates of each ray in the query image (patch centre)
 *   imgs   [V,3,H,W] source images;  logits [P, ld_logits]: blending logits, the first V columns are used
 *   c_pix  [P,3] blended pixel colour;  c_pat [P,(2h+1)^2,3] blended patch colours;  m_pat [P] 1 if any view sees the
 *   whole patch.  Backward: gradients w.r.t. the logits only ([P,V]); everything else is a constant of the graph. */
typedef struct {
  int32_t n_rays, n_samples, n_views, height, width, h_patch;
} nudf_blend_cfg;
int nudf_blend_forward(const nudf_blend_cfg* cfg, const float* pts, const float* proj, const float* hom, const float* px,
                       const float* imgs, const float* logits, int64_t ld_logits, float* c_pix, float* c_pat, float* m_pat,
                       void* stream);
int nudf_blend_backward(const nudf_blend_cfg* cfg, const float* pts, const float* proj, const float* hom, const float* px,
                        const float* imgs, const float* logits, int64_t ld_logits, const float* g_pix, const float* g_pat,
                        float* g_logits, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * ray generation on the device (replaces the per-step tensor algebra of the reference's data loader)
 * ------------------------------------------------------------------------------------------------------------ */
/* gen_random_rays_patches_at (dataset/dataset.py:228-294) without the patch crop: for n pixels (px, py: DEVICE int64, drawn by
 * the caller with torch.randint like the reference) of one image:
 *   rays[n,10] = (origin, unit direction in world space, rgb gathered from image[H,W,3], mask[H,W,3] > 0)
 *   ndc_uv[n,2] = 2 px/(W-1) - 1, 2 py/(H-1) - 1 (may be NULL);  near/far[n] = near_far_from_sphere (:329-335; may be NULL)
 * intrinsics_inv: DEVICE row-major 3x3 (top-left block of intrinsics_all_inv[idx], compact), pose: DEVICE row-major 4x4 c2w. */
int nudf_gen_rays(const float* intrinsics_inv, const float* pose, const int64_t* px, const int64_t* py, int32_t n,
                  const float* image, const float* mask, int32_t H, int32_t W, float* rays, float* ndc_uv, float* near,
                  float* far, void* stream);
/* gen_rays_at (dataset/dataset.py:151-164): rays of the [Hl, Wl] = [H // level, W // level] grid of pixel centres
 * linspace(0, W-1, Wl) x linspace(0, H-1, Hl); rays_o / rays_d [Hl, Wl, 3] (already transposed to image order). */
int nudf_gen_rays_grid(const float* intrinsics_inv, const float* pose, int32_t W, int32_t H, int32_t Wl, int32_t Hl, float* rays_o,
                       float* rays_d, float* near, float* far, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NUDF_H_ */
