"""Per-sample formulas + hand-derived derivatives of the compositing pass (neuraludf_b200/csrc/raymath.cuh),
compiled for the host and compared with autograd of the pinned oracle.  CPU only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import oracle_torch as O

HERE = os.path.dirname(os.path.abspath(__file__))


class Cfg(ctypes.Structure):
    _fields_ = [("S", ctypes.c_int), ("O", ctypes.c_int), ("inv_s", ctypes.c_float), ("beta", ctypes.c_float),
                ("gamma", ctypes.c_float), ("r", ctypes.c_float), ("has_r", ctypes.c_int), ("fs", ctypes.c_float),
                ("ssf", ctypes.c_float), ("use_norm", ctypes.c_int), ("has_bg_rgb", ctypes.c_int),
                ("bg_rgb", ctypes.c_float * 3)]


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(HERE, "host", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libraymath_host.so")
    src = os.path.join(HERE, "host", "raymath_host.cpp")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, src])
    return ctypes.CDLL(so)


def fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def make_case(seed, S, O, near_surface):
    g = torch.Generator().manual_seed(seed)
    N = 6
    o, d, near, far = O_rays(N, seed)
    z = near + (far - near) * torch.linspace(0, 1, S)[None, :]
    sd = float(((far - near) / S).mean())
    dists = torch.cat([z[:, 1:] - z[:, :-1], torch.full((N, 1), sd)], -1)
    mid = z + dists * 0.5
    pts = o[:, None, :] + d[:, None, :] * mid[..., None]
    # a synthetic sphere-like UDF along the ray plus noise, and noisy gradients
    udf = ((pts.norm(dim=-1) - 0.5).abs() * (0.05 if near_surface else 1.0)
           + 1e-3 * torch.rand(N, S, generator=g)).float()
    grads = (pts / pts.norm(dim=-1, keepdim=True) * torch.sign(pts.norm(dim=-1, keepdim=True) - 0.5)
             + 0.2 * torch.randn(N, S, 3, generator=g)).float()
    scb = torch.rand(N, S, 3, generator=g)
    sc = torch.rand(N, S, 3, generator=g)
    bga = torch.rand(N, S + O, generator=g) * 0.3
    bgc = torch.rand(N, S + O, 3, generator=g)
    return dict(o=o, d=d, z=z, dists=dists, mid=mid, pts=pts, udf=udf, grads=grads, scb=scb, sc=sc, bga=bga, bgc=bgc)


def O_rays(n, seed):
    return O.make_rays(n, seed)


@pytest.mark.parametrize("S,Oo,has_r,use_norm,bg_rgb,near", [(40, 0, 1, 0, 0, False), (70, 9, 1, 0, 1, True),
                                                               (33, 5, 0, 1, 0, True), (64, 0, 0, 0, 0, True)])
def test_composite_forward_backward(lib, S, Oo, has_r, use_norm, bg_rgb, near):
    c = make_case(3 + S, S, Oo, near)
    N = c["udf"].shape[0]
    inv_s, beta, gamma, r, fs, ssf = 403.4, 148.4, 20.1, 0.35, 0.4, 300.0
    bgv = torch.tensor([0.2, 0.5, 0.9])
    # ---- oracle in fp64 with autograd ----
    dt = torch.float64
    leaves = {k: c[k].to(dt).clone().requires_grad_(True) for k in ("udf", "grads", "scb", "sc", "bga", "bgc")}
    heads = [torch.tensor(v, dtype=dt, requires_grad=True) for v in (inv_s, beta, gamma)]
    ret = O.composite(c["d"].to(dt), c["pts"].to(dt), c["mid"].to(dt), c["dists"].to(dt), leaves["udf"],
                      leaves["grads"], leaves["scb"], leaves["sc"], heads[0], heads[1], heads[2],
                      cos_anneal_ratio=r if has_r else None, flip_saturation=fs,
                      background_rgb=bgv.to(dt) if bg_rgb else None,
                      background_alpha=leaves["bga"] if Oo else None,
                      background_sampled_color=leaves["bgc"] if Oo else None, sparse_scale_factor=ssf,
                      use_norm_grad_for_cosine=bool(use_norm))
    gen = torch.Generator().manual_seed(99)
    bars = {k: torch.randn(ret[k].shape, generator=gen, dtype=dt) for k in
            ("color_base", "color", "depth", "weight_sum", "weight_sum_fg_bg")}
    sb = torch.randn(3, generator=gen, dtype=dt)
    loss = sum((ret[k] * bars[k]).sum() for k in bars) + sb[0] * ret["gradient_error"] \
        + sb[1] * ret["gradient_error_near_surface"] + sb[2] * ret["sparse_error"]
    wanted = [leaves["udf"], leaves["grads"], leaves["scb"], leaves["sc"]] + heads + \
        ([leaves["bga"], leaves["bgc"]] if Oo else [])
    gr = torch.autograd.grad(loss, wanted)
    # ---- host harness, ray by ray ----
    cfg = Cfg(S, Oo, inv_s, beta, gamma, r, has_r, fs, ssf, use_norm, bg_rgb, (ctypes.c_float * 3)(*bgv.tolist()))
    f32 = lambda t: np.ascontiguousarray(t.detach().float().numpy())
    sums = np.zeros((N, 5), np.float32)
    outs = np.zeros((N, 14), np.float32)
    W = np.zeros((N, S + Oo), np.float32)
    arr = {k: f32(c[k]) for k in ("d", "pts", "mid", "dists", "udf", "grads", "scb", "sc", "bga", "bgc")}
    for i in range(N):
        lib.ray_forward_host(ctypes.byref(cfg), fp(arr["d"][i]), fp(arr["pts"][i]), fp(arr["mid"][i]), fp(arr["dists"][i]),
                             fp(arr["udf"][i]), fp(arr["grads"][i]), fp(arr["scb"][i]), fp(arr["sc"][i]), fp(arr["bga"][i]),
                             fp(arr["bgc"][i]), fp(outs[i]), fp(W[i]))
    def close(a, b, tol, name):
        a = np.asarray(a, np.float64); b = np.asarray(b.detach().numpy(), np.float64)
        err = np.abs(a - b).max() / (np.abs(b).max() + 1e-30)
        assert err < tol, (name, err)
    close(outs[:, 0:3], ret["color_base"], 2e-4, "color_base")
    close(outs[:, 3:6], ret["color"], 2e-4, "color")
    close(outs[:, 6:7], ret["depth"], 2e-4, "depth")
    close(outs[:, 7:8], ret["weight_sum"], 2e-4, "ws")
    close(outs[:, 8:9], ret["weight_sum_fg_bg"], 2e-4, "ws_all")
    close(W, ret["weights"], 2e-4, "weights")
    relax_sum, near_sum = outs[:, 10].sum(), outs[:, 12].sum()
    close(outs[:, 9].sum() / (relax_sum + 1e-5), ret["gradient_error"], 1e-4, "ge")
    close(outs[:, 13].sum() / N, ret["sparse_error"], 1e-4, "sparse")
    coef = np.array([sb[0] / (relax_sum + 1e-5), sb[1] / (near_sum + 1e-5), sb[2] / N], np.float32)
    ub = np.zeros((N, S), np.float32); gb = np.zeros((N, S, 3), np.float32)
    scbb = np.zeros((N, S, 3), np.float32); scb_ = np.zeros((N, S, 3), np.float32)
    bab = np.zeros((N, S + Oo), np.float32); bcb = np.zeros((N, S + Oo, 3), np.float32)
    scal = np.zeros((N, 3), np.float32)
    for i in range(N):
        bar = np.concatenate([f32(bars["color_base"][i]), f32(bars["color"][i]), f32(bars["depth"][i]),
                              f32(bars["weight_sum"][i]), f32(bars["weight_sum_fg_bg"][i])]).astype(np.float32)
        lib.ray_backward_host(ctypes.byref(cfg), fp(arr["d"][i]), fp(arr["pts"][i]), fp(arr["mid"][i]), fp(arr["dists"][i]),
                              fp(arr["udf"][i]), fp(arr["grads"][i]), fp(arr["scb"][i]), fp(arr["sc"][i]),
                              fp(arr["bga"][i]), fp(arr["bgc"][i]), fp(bar), fp(coef), fp(ub[i]), fp(gb[i]), fp(scbb[i]),
                              fp(scb_[i]), fp(bab[i]), fp(bcb[i]), fp(scal[i]))
    tol = 2e-3   # fp32 harness vs fp64 autograd through sigmoid(400 x) chains
    close(ub, gr[0], tol, "udf_bar")
    close(gb, gr[1], tol, "grads_bar")
    close(scbb, gr[2], tol, "scb_bar")
    close(scb_, gr[3], tol, "sc_bar")
    close(scal[:, 0].sum(), gr[4], tol, "inv_s_bar")
    close(scal[:, 1].sum(), gr[5], tol, "beta_bar")
    close(scal[:, 2].sum(), gr[6], tol, "gamma_bar")
    if Oo:
        close(bab[:, S:], gr[7][:, S:], tol, "bg_alpha_bar")
        close(bcb[:, S:], gr[8][:, S:], tol, "bg_color_bar")
