"""Pins oracle/oracle_torch.py against outputs of the UNMODIFIED reference (tests/golden/*.npz).

The fp64 oracle must agree with the fp64 reference run to round-off (1e-9), and the fp32 oracle with the
fp32 reference run to fp32 noise: both execute the same torch ops, only organised differently.
"""
import pytest
import torch

from oracle import oracle_torch as O
from tests.golden_util import rel_err

TOL = {torch.float64: 1e-8, torch.float32: 2e-5}
TAGS = [(torch.float32, "f32"), (torch.float64, "f64")]


@pytest.mark.parametrize("dtype,tag", TAGS)
def test_udf_value_and_gradient(golden, dtype, tag):
    g = golden
    x = g.t("udf_x", dtype)
    for name, cfg in (("udf", g.udf_c), ("udf_small", g.udf_small_c)):
        p = O.to_dtype(g.params[name], dtype)
        out = O.udf_mlp(p, cfg, x)
        grad = O.udf_gradient_autograd(p, cfg, x, create_graph=False)
        assert rel_err(out, g.t(name + "_out_" + tag)) < TOL[dtype]
        assert rel_err(grad, g.t(name + "_grad_" + tag)) < TOL[dtype]
        out2, grad2 = O.udf_value_and_gradient_analytic(p, cfg, x)
        assert rel_err(out2, g.t(name + "_out_" + tag)) < TOL[dtype]
        assert rel_err(grad2, g.t(name + "_grad_" + tag)) < 5 * TOL[dtype]


@pytest.mark.parametrize("dtype,tag", TAGS)
def test_color_and_nerf(golden, dtype, tag):
    g = golden
    cb, c, bl = O.color_mlp(O.to_dtype(g.params["color"], dtype), g.col_c, g.t("col_pts", dtype),
                            g.t("col_dirs", dtype), g.t("col_feat", dtype))
    assert rel_err(cb, g.t("col_base_" + tag)) < TOL[dtype]
    assert rel_err(c, g.t("col_color_" + tag)) < TOL[dtype]
    assert rel_err(bl, g.t("col_blend_" + tag)) < TOL[dtype]
    a, rgb = O.nerf_mlp(O.to_dtype(g.params["nerf"], dtype), g.nerf_c, g.t("nerf_pts", dtype),
                        g.t("nerf_dirs", dtype))
    assert rel_err(a, g.t("nerf_alpha_" + tag)) < TOL[dtype]
    assert rel_err(rgb, g.t("nerf_rgb_" + tag)) < TOL[dtype]


@pytest.mark.parametrize("dtype,tag", TAGS)
def test_sample_pdf(golden, dtype, tag):
    g = golden
    s = O.sample_pdf_det(g.t("pdf_bins", dtype), g.t("pdf_weights", dtype), 16)
    assert torch.equal(s, g.t("pdf_samples_" + tag))


@pytest.mark.parametrize("dtype,tag", TAGS)
def test_up_sampling_rounds(golden, dtype, tag):
    g = golden
    o, d = g.t("rays_o", dtype), g.t("rays_d", dtype)
    near, far = g.t("near", dtype), g.t("far", dtype)
    z, udf = g.t("up_z_" + tag), g.t("up_udf_" + tag)
    sd = ((far - near) / 64).mean().item()
    for i in range(5):
        gamma = float(min(max(20 * 2 ** (5 - i), 20), 320))
        nz = O.up_sample_unbias(o, d, z, udf, sd, 10, 64 * 2 ** i, 64 * 2 ** (i + 1), gamma)
        assert torch.equal(nz, g.t("up_newz_r%d_%s" % (i, tag))), i
    nz = O.up_sample_no_occ_aware(o, d, z, udf, sd, 13, 128, float(torch.exp(torch.tensor(3.0, dtype=torch.float64))))
    assert rel_err(nz, g.t("up_noocc_newz_" + tag)) < 1e-6


@pytest.mark.parametrize("dtype,tag", TAGS)
def test_importance_sampling(golden, dtype, tag):
    g = golden
    p = O.to_dtype(g.params["udf"], dtype)
    o, d = g.t("rays_o", dtype), g.t("rays_d", dtype)
    near, far = g.t("near", dtype), g.t("far", dtype)
    z0, _, sd = O.coarse_z(near, far, 64, 0)
    udf_fn = lambda x: O.udf_mlp(p, g.udf_c, x)[:, 0]
    with torch.no_grad():
        z = O.importance_sample(udf_fn, o, d, z0, sd, 50, 5)
        assert rel_err(z, g.t("imp_z_" + tag)) < 10 * TOL[dtype]
        _, beta, gamma = O.scalar_heads(O.to_dtype(g.params["sc"], dtype))
        zm = O.importance_sample_mix(udf_fn, o, d, z0, sd, 78, 5, beta, gamma)
        assert rel_err(zm, g.t("impmix_z_" + tag)) < 10 * TOL[dtype]


def _rc_loss(ret, S, dtype):
    tgt = torch.full((ret["color"].shape[0], 3), 0.4, dtype=dtype)
    return ((ret["color"] - tgt).abs().mean() + 0.01 * (ret["color_base"] - tgt).abs().mean()
            + 0.1 * ret["gradient_error"] + 1e-3 * ret["sparse_error"]
            + 0.05 * ret["gradient_error_near_surface"]
            + 0.1 * ((ret["weights"][:, :S].sum(-1) - 0.5) ** 2).mean())


RC_KEYS = ["color_base", "color", "weights", "depth", "gradient_error", "gradient_error_near_surface", "normals",
           "gradients", "gradients_flip", "inside_sphere", "udf", "gradient_mag", "true_cos", "vis_prob", "alpha",
           "alpha_plus", "alpha_minus", "mid_z_vals", "dists", "sparse_error", "alpha_occ", "raw_occ", "s_val",
           "beta", "gamma"]


@pytest.mark.parametrize("dtype,tag", TAGS)
@pytest.mark.parametrize("case", ["rc", "rc_na"])
def test_render_core_and_grads(golden, dtype, tag, case):
    g = golden
    kw = dict(cos_anneal_ratio=0.5, flip_saturation=0.3) if case == "rc" else dict(cos_anneal_ratio=None,
                                                                                     flip_saturation=0.0)
    up = {k: v.clone().requires_grad_(True) for k, v in O.to_dtype(g.params["udf"], dtype).items()}
    cp = {k: v.clone().requires_grad_(True) for k, v in O.to_dtype(g.params["color"], dtype).items()}
    sc = {k: v.clone().requires_grad_(True) for k, v in O.to_dtype(g.params["sc"], dtype).items()}
    o, d = g.t("rays_o", dtype), g.t("rays_d", dtype)
    near, far = g.t("near", dtype), g.t("far", dtype)
    S = 128
    z = near + (far - near) * torch.linspace(0.0, 1.0, S, dtype=dtype)[None, :]
    sd = ((far - near) / S).mean().item()
    ret = O.render_core(up, g.udf_c, cp, g.col_c, sc, o, d, z, sd, **kw)
    # composites amplify fp32 rounding through exp(-25000 udf), sigmoid(400 x): loosen fp32 (reference noise floor)
    tol = TOL[dtype] if dtype == torch.float64 else 2e-3
    for k in RC_KEYS:
        assert rel_err(ret[k], g.t("%s_%s_%s" % (case, k, tag))) < tol, k
    loss = _rc_loss(ret, S, dtype)
    assert rel_err(loss, g.t("%s_loss_%s" % (case, tag))) < tol
    if dtype != torch.float64:
        return
    loss.backward()
    from oracle.make_golden import GRAD_STRIDE
    n_checked = 0
    for mn, pd in (("udf", up), ("color", cp)):
        for pn, p in pd.items():
            key = "%s_grad.%s.%s_f64" % (case, mn, pn)
            if g.has(key):
                assert rel_err(p.grad, g.t(key)) < 1e-7, key
                n_checked += 1
            elif g.has(key + "_sub"):
                assert rel_err(p.grad.reshape(-1)[::GRAD_STRIDE], g.t(key + "_sub")) < 1e-7, key
                assert rel_err(p.grad.norm(), g.t(key + "_norm")) < 1e-7, key
                n_checked += 1
    assert n_checked >= 50
    assert rel_err(sc["variance"].grad, g.t("%s_grad.var.variance_f64" % case)) < 1e-7
    assert rel_err(sc["beta"].grad, g.t("%s_grad.beta.beta_f64" % case)) < 1e-7


@pytest.mark.parametrize("dtype,tag", [(torch.float64, "f64")])
def test_whole_render_dtu(golden, dtype, tag):
    g = golden
    up = {k: v.clone().requires_grad_(True) for k, v in O.to_dtype(g.params["udf"], dtype).items()}
    cp = {k: v.clone().requires_grad_(True) for k, v in O.to_dtype(g.params["color"], dtype).items()}
    npar = {k: v.clone().requires_grad_(True) for k, v in O.to_dtype(g.params["nerf"], dtype).items()}
    sc = {k: v.clone().requires_grad_(True) for k, v in O.to_dtype(g.params["sc"], dtype).items()}
    o, d = g.t("rays_o", dtype)[:32], g.t("rays_d", dtype)[:32]
    near, far = g.t("near", dtype)[:32], g.t("far", dtype)[:32]
    ret = O.render(up, g.udf_c, cp, g.col_c, npar, g.nerf_c, sc, o, d, near, far, 64, 50, 32, 5,
                   cos_anneal_ratio=0.7, flip_saturation=0.2)
    for k in ["z_vals", "color", "color_base", "weights", "depth", "weight_sum", "weight_sum_fg_bg", "udf",
              "gradients", "gradient_error", "sparse_error", "normals"]:
        assert rel_err(ret[k], g.t("render_%s_%s" % (k, tag))) < 1e-8, k
    tgt = torch.full((32, 3), 0.4, dtype=dtype)
    loss = ((ret["color"] - tgt).abs().mean() + 0.01 * (ret["color_base"] - tgt).abs().mean()
            + 0.1 * ret["gradient_error"])
    loss.backward()
    from oracle.make_golden import GRAD_STRIDE
    for pn in ("pts_linears.0.weight", "pts_linears.5.weight", "rgb_linear.weight", "alpha_linear.weight"):
        key = "render_grad.nerf.%s_f64" % pn
        if g.has(key):
            assert rel_err(npar[pn].grad, g.t(key)) < 1e-7, key
        else:
            assert rel_err(npar[pn].grad.reshape(-1)[::GRAD_STRIDE], g.t(key + "_sub")) < 1e-7, key
    key = "render_grad.udf.lin4.weight_v_f64"
    assert rel_err(up["lin4.weight_v"].grad.reshape(-1)[::GRAD_STRIDE], g.t(key + "_sub")) < 1e-7


# ---------------------------------------------------------------------------------------------------------------
# a16: pixel / patch blending (fixtures of oracle/make_golden_blend.py)
# ---------------------------------------------------------------------------------------------------------------
def _blend_fx():
    import os
    import numpy as np
    from tests.golden_util import GOLDEN
    return np.load(os.path.join(GOLDEN, "blend_outputs.npz"))


def test_oracle_patch_projector_matches_reference():
    fx = _blend_fx()
    v = O.make_blend_views(16, n_views=6, seed=0)
    pts, nrm = torch.from_numpy(fx["proj_pts"]), torch.from_numpy(fx["proj_normals"])
    c, m = O.pixel_warp(pts, v["color_maps"], v["intrinsics"], v["w2cs"])
    assert torch.equal(m, torch.from_numpy(fx["proj_pixel_mask"]))
    assert (c - torch.from_numpy(fx["proj_pixel_color"])).abs().max() < 2e-6
    c, m = O.patch_warp(pts, v["rays_uv"], nrm, v["color_maps"], v["intrinsics"][0], v["intrinsics"], v["query_c2w"],
                        torch.inverse(v["w2cs"]))
    assert (m != torch.from_numpy(fx["proj_patch_mask"])).float().mean() < 1e-4
    assert (c - torch.from_numpy(fx["proj_patch_color"])).abs().max() < 2e-4      # fp32, pixel coordinates up to 64


def test_oracle_render_core_with_blending_matches_reference_fp64(golden):
    fx = _blend_fx()
    g = golden
    dt = torch.float64
    v = {k: t.to(dt) for k, t in O.make_blend_views(16, n_views=6, seed=0).items()}
    udf_p, col_p, nerf_p = (O.to_dtype(g.params[k], dt) for k in ("udf", "color", "nerf"))
    sc = O.to_dtype(g.params["sc"], dt)
    z = torch.from_numpy(fx["blend_z"]).to(dt)
    z_feed = torch.from_numpy(fx["blend_z_feed"]).to(dt)
    # the reference takes sample_dist from its fp64 run; the fixture stores the fp32 run's value
    sd = ((v["far"] - v["near"]) / 32).mean().item()
    z64 = v["near"] + (v["far"] - v["near"]) * torch.linspace(0.0, 1.0, 32, dtype=dt)[None, :]
    zo = torch.linspace(1e-3, 1.0 - 1.0 / 9.0, 8, dtype=dt)
    zo = v["far"] / torch.flip(zo, dims=[-1]) + 1.0 / 32
    zf64, _ = torch.sort(torch.cat([z64, zo], dim=-1), dim=-1)
    assert rel_err(z64, z) < 1e-6 and rel_err(zf64, z_feed) < 1e-6
    bg = O.render_core_outside(lambda p, d: O.nerf_mlp(nerf_p, g.nerf_c, p, d), v["rays_o"], v["rays_d"], zf64, sd, 8)
    ret = O.render_core(udf_p, g.udf_c, col_p, g.col_c, sc, v["rays_o"], v["rays_d"], z64, sd, cos_anneal_ratio=0.8,
                        background_alpha=bg["alpha"], background_sampled_color=bg["sampled_color"], flip_saturation=0.1,
                        blending=dict(color_maps=v["color_maps"], w2cs=v["w2cs"], intrinsics=v["intrinsics"],
                                      query_c2w=v["query_c2w"], rays_uv=v["rays_uv"]))
    for k in ("color_base", "color", "color_pixel", "patch_colors", "patch_mask", "weights", "depth"):
        ref = torch.from_numpy(fx["blend_%s_f64" % k])
        assert rel_err(ret[k].reshape(ref.shape), ref) < 2e-7, k        # the warps run through fp64 grid_sample
