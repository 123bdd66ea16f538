"""Generate golden fixtures by running the UNMODIFIED reference (dev container only).

    python oracle/make_golden.py            # writes tests/golden/*.npz

The reference has no tests / golden vectors of its own; these fixtures are outputs of
its real code (imported through oracle/refshim.py) on seeded synthetic inputs, in fp32 and in fp64
(the fp64 run of the reference is the arbiter of the parity protocol).  They pin oracle/oracle_torch.py
(tests/test_oracle_pinned.py) and are compared directly with the CUDA path (tests/test_gpu_*.py).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle_torch as O  # noqa: E402
from oracle import refshim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
GRAD_STRIDE = 7


def np_(t):
    return t.detach().cpu().numpy()


def build_ref_nets(F, udf_c, col_c, nerf_c, udf_p, col_p, nerf_p, sc, dtype):
    torch.set_default_dtype(dtype)
    udf = F.UDFNetwork(d_in=3, d_out=udf_c["d_out"], d_hidden=udf_c["d_hidden"], n_layers=udf_c["n_layers"],
                       skip_in=udf_c["skip_in"], multires=udf_c["multires"], scale=udf_c["scale"],
                       bias=udf_c["bias"], geometric_init=False, weight_norm=True, udf_type="abs")
    udf.load_state_dict(O.to_dtype(udf_p, dtype))
    col = F.ResidualRenderingNetwork(d_feature=col_c["d_feature"], mode="no_normal", d_in=6, d_out=3,
                                     d_hidden=col_c["d_hidden"], n_layers=col_c["n_layers"], weight_norm=True,
                                     multires_view=col_c["multires_view"], squeeze_out=True,
                                     blending_cand_views=col_c["blending_cand_views"])
    col.load_state_dict(O.to_dtype(col_p, dtype))
    nerf = F.NeRF(D=nerf_c["D"], W=nerf_c["W"], d_in=4, d_in_view=3, multires=nerf_c["multires"],
                  multires_view=nerf_c["multires_view"], output_ch=4, skips=list(nerf_c["skips"]),
                  use_viewdirs=True)
    nerf.load_state_dict(O.to_dtype(nerf_p, dtype))
    var = F.SingleVarianceNetwork(init_val=float(sc["variance"]))
    beta = F.BetaNetwork(init_var_beta=float(sc["beta"]), init_var_gamma=float(sc["gamma"]),
                         init_var_zeta=float(sc["zeta"]), beta_min=5e-5, requires_grad_beta=True,
                         requires_grad_gamma=False, requires_grad_zeta=False)
    var.variance.data = sc["variance"].to(dtype).clone()
    beta.beta.data = sc["beta"].to(dtype).clone()
    beta.gamma.data = sc["gamma"].to(dtype).clone()
    return udf, col, nerf, var, beta


def main():
    os.makedirs(OUT, exist_ok=True)
    F, R = refshim.load()
    torch.manual_seed(0)

    udf_c = O.udf_cfg()
    udf_s_c = O.udf_cfg(d_hidden=128, n_layers=4)
    col_c = O.color_cfg()
    nerf_c = O.nerf_cfg()
    udf_p = O.make_udf_params(udf_c, seed=0)
    udf_s_p = O.make_udf_params(udf_s_c, seed=3)
    col_p = O.make_color_params(col_c, seed=1)
    nerf_p = O.make_nerf_params(nerf_c, seed=2)
    sc = O.make_scalars()

    scene = {}
    for pre, p in (("udf.", udf_p), ("udf_small.", udf_s_p), ("color.", col_p), ("nerf.", nerf_p), ("sc.", sc)):
        for k, v in p.items():
            scene[pre + k] = np_(v)
    np.savez(os.path.join(OUT, "scene_params.npz"), **scene)

    g = torch.Generator().manual_seed(7)
    fx = {}

    for dtype, tag in ((torch.float32, "f32"), (torch.float64, "f64")):
        udf, col, nerf, var, beta = build_ref_nets(F, udf_c, col_c, nerf_c, udf_p, col_p, nerf_p, sc, dtype)
        udf_s = F.UDFNetwork(d_in=3, d_out=257, d_hidden=128, n_layers=4, skip_in=(4,), multires=6, scale=1.0,
                             bias=0.5, geometric_init=False, weight_norm=True, udf_type="abs")
        udf_s.load_state_dict(O.to_dtype(udf_s_p, dtype))

        # ---- stage: UDF value + gradient (a2, a3) ----
        gg = torch.Generator().manual_seed(11)
        x = (torch.rand(256, 3, generator=gg, dtype=torch.float64) * 2 - 1) * 0.9
        fx["udf_x"] = np_(x.float())
        xx = x.float().to(dtype)
        fx["udf_out_" + tag] = np_(udf(xx))
        fx["udf_grad_" + tag] = np_(udf.gradient(xx.clone())[:, 0])
        fx["udf_small_out_" + tag] = np_(udf_s(xx))
        fx["udf_small_grad_" + tag] = np_(udf_s.gradient(xx.clone())[:, 0])

        # ---- stage: colour net (a4) ----
        gg = torch.Generator().manual_seed(12)
        cp = (torch.rand(512, 3, generator=gg, dtype=torch.float64) * 2 - 1).float()
        cd = torch.randn(512, 3, generator=gg, dtype=torch.float64)
        cd = (cd / cd.norm(dim=1, keepdim=True)).float()
        cf = (0.3 * torch.randn(512, 256, generator=gg, dtype=torch.float64)).float()
        fx["col_pts"], fx["col_dirs"], fx["col_feat"] = np_(cp), np_(cd), np_(cf)
        cb, c, bl = col(cp.to(dtype), cd.to(dtype), cd.to(dtype), cf.to(dtype))
        fx["col_base_" + tag], fx["col_color_" + tag], fx["col_blend_" + tag] = np_(cb), np_(c), np_(bl)

        # ---- stage: NeRF (a5) ----
        gg = torch.Generator().manual_seed(13)
        npnt = torch.randn(256, 4, generator=gg, dtype=torch.float64)
        npnt = (npnt / npnt[:, :3].norm(dim=1, keepdim=True)).float()
        npnt[:, 3] = torch.rand(256, generator=gg, dtype=torch.float64).float()
        nd = torch.randn(256, 3, generator=gg, dtype=torch.float64)
        nd = (nd / nd.norm(dim=1, keepdim=True)).float()
        fx["nerf_pts"], fx["nerf_dirs"] = np_(npnt), np_(nd)
        na, nrgb = nerf(npnt.to(dtype), nd.to(dtype))
        fx["nerf_alpha_" + tag], fx["nerf_rgb_" + tag] = np_(na), np_(nrgb)

        # ---- renderer ----
        ren = R.UDFRendererBlending(nerf, udf, var, col, beta, n_samples=64, n_importance=50, n_outside=32,
                                    up_sample_steps=5, perturb=0.0)
        o, d, near, far = O.make_rays(64, seed=0)
        o, d, near, far = o.to(dtype), d.to(dtype), near.to(dtype), far.to(dtype)
        if tag == "f32":
            fx["rays_o"], fx["rays_d"], fx["near"], fx["far"] = np_(o), np_(d), np_(near), np_(far)

        # ---- stage: sample_pdf (a9) on synthetic peaky weights ----
        gg = torch.Generator().manual_seed(14)
        bins = torch.sort(torch.rand(64, 65, generator=gg, dtype=torch.float64) * 2 + 1.5, dim=-1)[0].float()
        w = torch.rand(64, 64, generator=gg, dtype=torch.float64) ** 8
        w[::3] *= 1e-6
        w = w.float()
        fx["pdf_bins"], fx["pdf_weights"] = np_(bins), np_(w)
        sm = R.sample_pdf(bins.to(dtype), w.to(dtype), 16, det=True)
        fx["pdf_samples_" + tag] = np_(sm)

        # ---- stage: importance sampling, DTU schedule (a8, a10) ----
        sample_dist = ((far - near) / 64).mean().item()
        z0 = near + (far - near) * torch.linspace(0.0, 1.0, 64)[None, :]
        with torch.no_grad():
            pts = o[:, None, :] + d[:, None, :] * z0[..., :, None]
            u0 = udf(pts.reshape(-1, 3))[:, 0].reshape(64, 64)
            fx["up_z_" + tag], fx["up_udf_" + tag] = np_(z0), np_(u0)
            for i in range(5):
                nz = ren.up_sample_unbias(o, d, z0, u0, sample_dist, 10, 64 * 2 ** i, 64 * 2 ** (i + 1),
                                          gamma=float(np.clip(20 * 2 ** (5 - i), 20, 320)))
                fx["up_newz_r%d_%s" % (i, tag)] = np_(nz)
            nz2 = ren.up_sample_no_occ_aware(o, d, z0, u0, sample_dist, 13, 64, 128, float(np.exp(3.0)))
            fx["up_noocc_newz_" + tag] = np_(nz2)
            zf = ren.importance_sample(o, d, z0, sample_dist)
            fx["imp_z_" + tag] = np_(zf)
            ren.upsampling_type = "mix"
            ren.n_importance, ren.up_sample_steps = 78, 5
            zm = ren.importance_sample_mix(o, d, z0, sample_dist)
            fx["impmix_z_" + tag] = np_(zm)
            ren.upsampling_type = "classical"
            ren.n_importance, ren.up_sample_steps = 50, 5

        # ---- render_core on uniform z, 64 rays x 128 samples (C2 shape, fewer rays) + grads ----
        S = 128
        z = near + (far - near) * torch.linspace(0.0, 1.0, S)[None, :]
        sd = ((far - near) / S).mean().item()
        for name, kw in (("rc", dict(cos_anneal_ratio=0.5, flip_saturation=0.3)),
                         ("rc_na", dict(cos_anneal_ratio=None, flip_saturation=0.0))):
            for m in (udf, col, var, beta):
                m.zero_grad(set_to_none=True)
            ret = ren.render_core(o, d, z, sd, udf, var, col, beta_network=beta, **kw)
            tgt = torch.full((64, 3), 0.4, dtype=dtype)
            loss = ((ret["color"] - tgt).abs().mean() + 0.01 * (ret["color_base"] - tgt).abs().mean()
                    + 0.1 * ret["gradient_error"] + 1e-3 * ret["sparse_error"]
                    + 0.05 * ret["gradient_error_near_surface"]
                    + 0.1 * ((ret["weights"][:, :S].sum(-1) - 0.5) ** 2).mean())
            loss.backward()
            for k, v in ret.items():
                if isinstance(v, torch.Tensor):
                    fx["%s_%s_%s" % (name, k, tag)] = np_(v)
            fx["%s_loss_%s" % (name, tag)] = np_(loss)
            for mn, m in (("udf", udf), ("color", col), ("var", var), ("beta", beta)):
                for pn, p in m.named_parameters():
                    if p.grad is not None:
                        fx["%s_grad.%s.%s_%s" % (name, mn, pn, tag)] = np_(p.grad)

        # ---- whole render(), DTU conf, perturb 0, 32 rays ----
        for m in (udf, col, var, beta, nerf):
            m.zero_grad(set_to_none=True)
        o2, d2, n2, f2 = o[:32], d[:32], near[:32], far[:32]
        # render() draws `torch.rand([1024,3]).float()` for sparse_random_error (:683), which breaks an fp64
        # run; cast inside .udf() only (that output is not part of any comparison).
        udf.udf = (lambda x, _m=udf, _dt=dtype: F.UDFNetwork.udf(_m, x.to(_dt)))
        ret = ren.render(o2, d2, n2, f2, cos_anneal_ratio=0.7, perturb_overwrite=0, flip_saturation=0.2)
        tgt = torch.full((32, 3), 0.4, dtype=dtype)
        loss = ((ret["color"] - tgt).abs().mean() + 0.01 * (ret["color_base"] - tgt).abs().mean()
                + 0.1 * ret["gradient_error"])
        loss.backward()
        for k, v in ret.items():
            if isinstance(v, torch.Tensor):
                fx["render_%s_%s" % (k, tag)] = np_(v)
        fx["render_loss_" + tag] = np_(loss)
        for mn, m in (("udf", udf), ("color", col), ("var", var), ("beta", beta), ("nerf", nerf)):
            for pn, p in m.named_parameters():
                if p.grad is not None:
                    fx["render_grad.%s.%s_%s" % (mn, pn, tag)] = np_(p.grad)

    torch.set_default_dtype(torch.float32)
    # Large per-parameter gradients: keep the fp64 arbiter only, as a strided subsample (flat[::GRAD_STRIDE])
    # plus its L2 norm -- enough to catch any indexing / scaling error while keeping the fixture small.
    for k in list(fx):
        if "_grad." in k and fx[k].size > 4096:
            if k.endswith("_f32"):
                del fx[k]
                continue
            full = fx.pop(k).astype(np.float64).reshape(-1)
            fx[k + "_sub"] = full[::GRAD_STRIDE].copy()
            fx[k + "_norm"] = np.array(np.sqrt((full ** 2).sum()))
    np.savez_compressed(os.path.join(OUT, "reference_outputs.npz"), **fx)
    print("wrote", len(fx), "arrays;", sum(v.nbytes for v in fx.values()) / 1e6, "MB raw")


if __name__ == "__main__":
    main()
