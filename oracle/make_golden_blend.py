"""Golden fixtures for the pixel / patch blending stage (config C3), from the UNMODIFIED reference (dev container only).

    python oracle/make_golden_blend.py      # writes tests/golden/blend_outputs.npz

Same scene parameters as make_golden.py (same seeds -> tests/golden/scene_params.npz); inputs come from
neuraludf_b200.synthetic.make_blend_views.  Two groups:
  * proj_*: PatchProjector.pixel_warp / patch_warp of the reference on fixed points and normals (fp32);
  * blend_*: render_core with colour maps, uv and a NeRF++ background produced by the reference's render_core_outside,
    in fp32 and fp64, with the gradients of a trainer-like loss.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_torch as O  # noqa: E402
from oracle import refshim  # noqa: E402
from oracle.make_golden import GRAD_STRIDE, OUT, build_ref_nets, np_  # noqa: E402

N_RAYS, S, N_OUT, N_VIEWS = 16, 32, 8, 6


def blend_loss(ret, dtype):
    """Shape of the fine-tuning loss (exp_runner_blending.py:318-371) with plain L1 terms."""
    n = ret["color"].shape[0]
    tgt = torch.full((n, 3), 0.4, dtype=dtype)
    loss = (ret["color"] - tgt).abs().mean() + 0.5 * (ret["color_pixel"] - tgt).abs().mean()
    loss = loss + 0.01 * (ret["color_base"] - tgt).abs().mean() + 0.1 * ret["gradient_error"]
    pm = ret["patch_mask"].detach()
    loss = loss + 0.5 * ((ret["patch_colors"] - 0.4).abs().mean(dim=(1, 2)) * pm).sum() / (pm.sum() + 1e-5)
    return loss


def main():
    F, R = refshim.load()
    udf_c, col_c, nerf_c = O.udf_cfg(), O.color_cfg(), O.nerf_cfg()
    udf_p, col_p = O.make_udf_params(udf_c, seed=0), O.make_color_params(col_c, seed=1)
    nerf_p, sc = O.make_nerf_params(nerf_c, seed=2), O.make_scalars()
    views = O.make_blend_views(N_RAYS, n_views=N_VIEWS, seed=0)
    fx = {}

    # ---- projector alone (fp32, CPU) ----
    torch.set_default_dtype(torch.float32)
    from models.patch_projector import PatchProjector
    pp = PatchProjector(3)
    z = views["near"] + (views["far"] - views["near"]) * torch.linspace(0.0, 1.0, 24)[None, :]
    pts = views["rays_o"][:, None, :] + views["rays_d"][:, None, :] * z[..., None]
    gg = torch.Generator().manual_seed(21)
    nrm = -views["rays_d"][:, None, :] + 0.5 * torch.randn(N_RAYS, 24, 3, generator=gg)
    nrm = nrm / nrm.norm(dim=-1, keepdim=True)
    fx["proj_pts"], fx["proj_normals"] = np_(pts), np_(nrm)
    c, m = pp.pixel_warp(pts, views["color_maps"], views["intrinsics"], views["w2cs"])
    fx["proj_pixel_color"], fx["proj_pixel_mask"] = np_(c), np_(m)
    c, m = pp.patch_warp(pts, views["rays_uv"].clone(), nrm, views["color_maps"], views["intrinsics"][0],
                         views["intrinsics"], views["query_c2w"], torch.inverse(views["w2cs"]))
    fx["proj_patch_color"], fx["proj_patch_mask"] = np_(c), np_(m)

    # ---- render_core with blending ----
    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        udf, col, nerf, var, beta = build_ref_nets(F, udf_c, col_c, nerf_c, udf_p, col_p, nerf_p, sc, dtype)
        ren = R.UDFRendererBlending(nerf, udf, var, col, beta, n_samples=S, n_importance=0, n_outside=N_OUT,
                                    up_sample_steps=0, perturb=0.0)
        ren.patch_projector.z_axis = ren.patch_projector.z_axis.to(dtype)
        v = {k: t.to(dtype) for k, t in views.items()}
        o, d, near, far = v["rays_o"], v["rays_d"], v["near"], v["far"]
        z = near + (far - near) * torch.linspace(0.0, 1.0, S)[None, :]
        sd = ((far - near) / S).mean().item()
        z_out = torch.linspace(1e-3, 1.0 - 1.0 / (N_OUT + 1.0), N_OUT)
        z_out = far / torch.flip(z_out, dims=[-1]) + 1.0 / S
        z_feed, _ = torch.sort(torch.cat([z, z_out], dim=-1), dim=-1)
        if tag == "f32":
            fx["blend_z"], fx["blend_z_feed"], fx["blend_sample_dist"] = np_(z), np_(z_feed), np.array(sd)
        for m_ in (udf, col, var, beta, nerf):
            m_.zero_grad(set_to_none=True)
        bg = ren.render_core_outside(o, d, z_feed, sd, nerf)
        ret = ren.render_core(o, d, z, sd, udf, var, col, beta_network=beta, cos_anneal_ratio=0.8,
                              background_alpha=bg["alpha"], background_sampled_color=bg["sampled_color"],
                              flip_saturation=0.1, color_maps=v["color_maps"], w2cs=v["w2cs"],
                              intrinsics=v["intrinsics"], query_c2w=v["query_c2w"], img_index=None,
                              rays_uv=v["rays_uv"].clone())
        loss = blend_loss(ret, dtype)
        loss.backward()
        for k in ("color_base", "color", "color_pixel", "patch_colors", "patch_mask", "weights", "depth"):
            fx["blend_%s_%s" % (k, tag)] = np_(ret[k])
        fx["blend_loss_" + tag] = np_(loss)
        for mn, m_ in (("udf", udf), ("color", col), ("var", var), ("beta", beta), ("nerf", nerf)):
            for pn, p in m_.named_parameters():
                if p.grad is not None:
                    fx["blend_grad.%s.%s_%s" % (mn, pn, tag)] = np_(p.grad)

    torch.set_default_dtype(torch.float32)
    for k in list(fx):
        if "_grad." in k and fx[k].size > 4096:
            if k.endswith("_f32"):
                del fx[k]
                continue
            full = fx.pop(k).astype(np.float64).reshape(-1)
            fx[k + "_sub"] = full[::GRAD_STRIDE].copy()
            fx[k + "_norm"] = np.array(np.sqrt((full ** 2).sum()))
    np.savez_compressed(os.path.join(OUT, "blend_outputs.npz"), **fx)
    print("wrote", len(fx), "arrays;", sum(a.nbytes for a in fx.values()) / 1e6, "MB raw")


if __name__ == "__main__":
    main()
