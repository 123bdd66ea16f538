"""Helpers for the `-m gpu` parity tests: module construction from the golden scene, parity protocol, report file."""
import json
import os

import torch

from oracle import oracle_torch as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REPORT = os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")


def report(name, **kv):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    with open(REPORT, "a") as f:
        f.write(json.dumps(dict(name=name, **{k: (float(v) if hasattr(v, "__float__") else v) for k, v in kv.items()})) + "\n")


def err_inf(a, b):
    return float((a.double().cpu() - b.double().cpu()).abs().max())


def scale_inf(b):
    return float(b.double().abs().max()) + 1e-300


def parity(name, new, ref64, ref32=None, tol=1e-4, noise_mult=2.0):
    """max|new-ref64| <= max(tol * max|ref64|, noise_mult * max|ref32-ref64|)."""
    e = err_inf(new, ref64)
    s = scale_inf(ref64)
    noise = err_inf(ref32, ref64) if ref32 is not None else 0.0
    bound = max(tol * s, noise_mult * noise)
    report(name, err=e, rel=e / s, ref_noise_rel=noise / s, bound_rel=bound / s, ok=bool(e <= bound))
    assert e <= bound, "%s: err %.3e (rel %.3e) > bound %.3e (rel %.3e; ref noise rel %.3e)" % (
        name, e, e / s, bound, bound / s, noise / s)


def build_modules(g, device="cuda", udf_name="udf"):
    """Our modules loaded with the golden scene parameters."""
    from neuraludf_b200.models import fields as F
    cfg = g.udf_c if udf_name == "udf" else g.udf_small_c
    udf = F.UDFNetwork(d_in=3, d_out=cfg["d_out"], d_hidden=cfg["d_hidden"], n_layers=cfg["n_layers"],
                       skip_in=cfg["skip_in"], multires=cfg["multires"], scale=cfg["scale"], bias=cfg["bias"],
                       geometric_init=True, weight_norm=True, udf_type="abs")
    udf.load_state_dict(g.params[udf_name])
    cc = g.col_c
    col = F.ResidualRenderingNetwork(d_feature=cc["d_feature"], mode="no_normal", d_in=6, d_out=3,
                                     d_hidden=cc["d_hidden"], n_layers=cc["n_layers"], weight_norm=True,
                                     multires_view=cc["multires_view"], squeeze_out=True,
                                     blending_cand_views=cc["blending_cand_views"])
    col.load_state_dict(g.params["color"])
    nc = g.nerf_c
    nerf = F.NeRF(D=nc["D"], W=nc["W"], d_in=4, d_in_view=3, multires=nc["multires"], multires_view=nc["multires_view"],
                  output_ch=4, skips=list(nc["skips"]), use_viewdirs=True)
    nerf.load_state_dict(g.params["nerf"])
    sc = g.params["sc"]
    var = F.SingleVarianceNetwork(init_val=float(sc["variance"]))
    beta = F.BetaNetwork(init_var_beta=float(sc["beta"]), init_var_gamma=float(sc["gamma"]),
                         init_var_zeta=float(sc["zeta"]), beta_min=5e-5, requires_grad_beta=True,
                         requires_grad_gamma=False, requires_grad_zeta=False)
    with torch.no_grad():
        var.variance.copy_(sc["variance"])
        beta.beta.copy_(sc["beta"])
        beta.gamma.copy_(sc["gamma"])
    mods = [m.to(device) for m in (udf, col, nerf, var, beta)]
    return mods


def oracle_params(g, name, dtype, requires_grad=False):
    p = O.to_dtype(g.params[name], dtype)
    if requires_grad:
        p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    return p
