"""GPU parity of the pixel / patch blending stage (config C3): render_core with colour maps, uv and
a NeRF++ background against fixtures of the UNMODIFIED reference (oracle/make_golden_blend.py), fp64 reference = arbiter.
"""
import os

import numpy as np
import pytest
import torch

from neuraludf_b200.synthetic import make_blend_views
from tests.golden_util import GOLDEN
from tests.gpu_util import build_modules, err_inf, parity, report, scale_inf

pytestmark = pytest.mark.gpu
DEV = "cuda"
N_RAYS, S, N_OUT, N_VIEWS = 16, 32, 8, 6


@pytest.fixture(scope="module")
def fx():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return np.load(os.path.join(GOLDEN, "blend_outputs.npz"))


def _loss(ret):
    n = ret["color"].shape[0]
    tgt = torch.full((n, 3), 0.4, device=DEV)
    loss = (ret["color"] - tgt).abs().mean() + 0.5 * (ret["color_pixel"] - tgt).abs().mean()
    loss = loss + 0.01 * (ret["color_base"] - tgt).abs().mean() + 0.1 * ret["gradient_error"]
    pm = ret["patch_mask"].detach()
    return loss + 0.5 * ((ret["patch_colors"] - 0.4).abs().mean(dim=(1, 2)) * pm).sum() / (pm.sum() + 1e-5)


@pytest.mark.parametrize("engine", [0, 1, 2])
def test_render_core_blending_vs_reference(golden, fx, engine):
    """engine 0: exact fp32; 1: tensor engine (default chains); 2: tensor engine with the plane-fed reverse-sweep / tangent
    chains and plane-fed weight gradients (nudf_set_chain_planes)"""
    from neuraludf_b200 import _lib
    from neuraludf_b200.models.udf_renderer_blending import UDFRendererBlending
    from oracle.make_golden import GRAD_STRIDE
    L = _lib.lib()
    old, old_pl = L.nudf_get_engine(), L.nudf_get_chain_planes()
    L.nudf_set_engine(min(engine, 1))
    L.nudf_set_chain_planes(1 if engine == 2 else 0)
    try:
        udf, col, nerf, var, beta = build_modules(golden, DEV)
        ren = UDFRendererBlending(nerf, udf, var, col, beta, n_samples=S, n_importance=0, n_outside=N_OUT,
                                  up_sample_steps=0, perturb=0.0)
        v = {k: t.to(DEV) for k, t in make_blend_views(N_RAYS, n_views=N_VIEWS, seed=0).items()}
        o, d = v["rays_o"], v["rays_d"]
        z = torch.from_numpy(fx["blend_z"]).to(DEV).contiguous()
        z_feed = torch.from_numpy(fx["blend_z_feed"]).to(DEV).contiguous()
        sd = float(fx["blend_sample_dist"])
        bg = ren.render_core_outside(o, d, z_feed, sd, nerf)
        uv0 = v["rays_uv"].clone()
        ret = ren.render_core(o, d, z, sd, udf, var, col, beta_network=beta, cos_anneal_ratio=0.8,
                              background_alpha=bg["alpha"], background_sampled_color=bg["sampled_color"],
                              flip_saturation=0.1, color_maps=v["color_maps"], w2cs=v["w2cs"],
                              intrinsics=v["intrinsics"], query_c2w=v["query_c2w"], img_index=None,
                              rays_uv=v["rays_uv"])
        assert torch.equal(uv0, v["rays_uv"])
        tag = "blend.e%d." % engine
        tol = 2e-4 if engine == 0 else 5e-4
        for k in ("color_base", "color", "color_pixel", "patch_colors", "patch_mask", "weights", "depth"):
            r64 = torch.from_numpy(fx["blend_%s_f64" % k])
            r32 = torch.from_numpy(fx["blend_%s_f32" % k])
            parity(tag + k, ret[k].reshape(r64.shape), r64, r32, tol=tol)
        loss = _loss(ret)
        parity(tag + "loss", loss, torch.from_numpy(fx["blend_loss_f64"]), torch.from_numpy(fx["blend_loss_f32"]),
               tol=tol)
        loss.backward()
        worst, n = 0.0, 0
        for mn, m in (("udf", udf), ("color", col), ("nerf", nerf)):
            for pn, p in m.named_parameters():
                key = "blend_grad.%s.%s_f64" % (mn, pn)
                if key in fx.files:
                    ref, new = torch.from_numpy(fx[key]), p.grad.cpu()
                elif key + "_sub" in fx.files:
                    ref, new = torch.from_numpy(fx[key + "_sub"]), p.grad.reshape(-1)[::GRAD_STRIDE].cpu()
                else:
                    assert p.grad is None or float(p.grad.abs().max()) == 0.0, key
                    continue
                e = err_inf(new, ref) / scale_inf(ref)
                worst = max(worst, e)
                n += 1
                report(tag + "dparam.%s.%s" % (mn, pn), rel=e)
                assert e < (5e-3 if engine == 0 else 2e-2), (key, e)
        assert n >= 60
        report(tag + "dparam.worst_rel", rel=worst)
        # the blending logits (10 output rows of the colour head) must receive a gradient
        assert float(col.lin4.weight_v.grad[3:].abs().max()) > 0
    finally:
        L.nudf_set_engine(old)
        L.nudf_set_chain_planes(old_pl)


def test_whole_render_with_blending_runs_and_trains_all_networks(golden):
    from neuraludf_b200.models.udf_renderer_blending import UDFRendererBlending
    udf, col, nerf, var, beta = build_modules(golden, DEV)
    ren = UDFRendererBlending(nerf, udf, var, col, beta, n_samples=32, n_importance=24, n_outside=8, up_sample_steps=2,
                              perturb=1.0)
    v = {k: t.to(DEV) for k, t in make_blend_views(64, n_views=N_VIEWS, seed=3).items()}
    ret = ren.render(v["rays_o"], v["rays_d"], v["near"], v["far"], cos_anneal_ratio=1.0, flip_saturation=0.0,
                     color_maps=v["color_maps"], w2cs=v["w2cs"], intrinsics=v["intrinsics"], query_c2w=v["query_c2w"],
                     img_index=None, rays_uv=v["rays_uv"])
    assert ret["color_pixel"].shape == (64, 3) and ret["patch_colors"].shape == (64, 49, 3)
    assert ret["patch_mask"].shape == (64,)
    for k in ("color", "color_pixel", "patch_colors", "patch_mask"):
        assert torch.isfinite(ret[k]).all(), k
    _loss(ret).backward()
    for m in (udf, col, nerf):
        gs = [p.grad for p in m.parameters() if p.grad is not None]
        assert gs and all(torch.isfinite(g_).all() for g_ in gs) and sum(float(g_.abs().sum()) for g_ in gs) > 0
    # pixel-only blending (colour maps without uv) is a valid configuration of the trainer as well
    ret2 = ren.render(v["rays_o"], v["rays_d"], v["near"], v["far"], cos_anneal_ratio=1.0, perturb_overwrite=0,
                      color_maps=v["color_maps"], w2cs=v["w2cs"], intrinsics=v["intrinsics"], query_c2w=v["query_c2w"])
    assert ret2["color_pixel"] is not None and ret2["patch_colors"] is None and ret2["patch_mask"] is None
