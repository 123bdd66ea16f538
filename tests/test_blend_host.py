"""Per-view / per-pixel math of the fused blending kernel (neuraludf_b200/csrc/blendmath.cuh), compiled for the host and
compared with the op-by-op torch form (PatchProjector.pixel_warp / patch_warp + fields.color_blend, themselves pinned to the
reference's outputs in tests/test_patch_projector.py), forward and logits gradient.  CPU only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from neuraludf_b200.models.fields import color_blend
from neuraludf_b200.models.patch_projector import PatchProjector
from neuraludf_b200.synthetic import make_blend_views
from tests.golden_util import GOLDEN

HERE = os.path.dirname(os.path.abspath(__file__))


class Cfg(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("n_rays", "n_samples", "n_views", "height", "width", "h_patch")]


@pytest.fixture(scope="module")
def lib():
    out = os.path.join(HERE, "host", "_build")
    os.makedirs(out, exist_ok=True)
    so = os.path.join(out, "libblend_host.so")
    subprocess.check_call(["g++", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "host", "blend_host.cpp")])
    return ctypes.CDLL(so)


def fp(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


@pytest.mark.parametrize("with_patch", [True, False])
def test_blend_math_matches_op_by_op_form(lib, with_patch):
    fx = np.load(os.path.join(GOLDEN, "blend_outputs.npz"))
    v = make_blend_views(16, n_views=6, seed=0)
    pts, nrm = torch.from_numpy(fx["proj_pts"]), torch.from_numpy(fx["proj_normals"])
    N, S, V, h = 16, 24, 6, 3
    g = torch.Generator().manual_seed(5)
    logits = (torch.randn(N, S, 10, generator=g) * 1.5).requires_grad_(True)
    pp = PatchProjector(h)
    pix_col, pix_mask = pp.pixel_warp(pts, v["color_maps"], v["intrinsics"], v["w2cs"])
    pat_col = pat_mask = None
    if with_patch:
        pat_col, pat_mask = pp.patch_warp(pts, v["rays_uv"], nrm, v["color_maps"], v["intrinsics"][0], v["intrinsics"],
                                          v["query_c2w"], torch.inverse(v["w2cs"]))
    c_pix, _, c_pat, m_pat = color_blend(logits, None, pix_col, pix_mask, pat_col, pat_mask)
    g_pix = torch.randn(N, S, 3, generator=g)
    loss = (c_pix * g_pix).sum()
    g_pat = None
    if with_patch:
        g_pat = torch.randn(N, S, 49, 3, generator=g)
        loss = loss + (c_pat * g_pat).sum()
    loss.backward()

    proj = (v["intrinsics"][:, :3, :3] @ v["w2cs"][:, :3, :]).reshape(V, 12).contiguous().numpy()
    hom = px = None
    if with_patch:
        hm, pxt = pp.homographies(pts, v["rays_uv"], nrm, (48, 64), v["intrinsics"][0], v["intrinsics"], v["query_c2w"],
                                  torch.inverse(v["w2cs"]))
        hom, px = hm.reshape(V, -1, 9).contiguous().numpy(), pxt.contiguous().numpy()
    cfg = Cfg(N, S, V, 48, 64, h)
    P = N * S
    o_pix = np.zeros((P, 3), np.float32)
    o_pat = np.zeros((P, 49, 3), np.float32) if with_patch else None
    o_m = np.zeros(P, np.float32) if with_patch else None
    lg = logits.detach().reshape(P, 10).contiguous().numpy()
    imgs = v["color_maps"].contiguous().numpy()
    p_np = pts.reshape(P, 3).contiguous().numpy()
    lib.blend_host_forward(ctypes.byref(cfg), fp(p_np), fp(proj), fp(hom), fp(px), fp(imgs), fp(lg), ctypes.c_int64(10), fp(o_pix),
                           fp(o_pat), fp(o_m))
    assert np.abs(o_pix - c_pix.detach().reshape(P, 3).numpy()).max() < 3e-6
    if with_patch:
        ref_m = m_pat.reshape(P).float().numpy()
        assert (o_m != ref_m).mean() < 2e-3                        # a pixel exactly on the border may flip a view
        same = o_m == ref_m
        d = np.abs(o_pat - c_pat.detach().reshape(P, 49, 3).numpy()).reshape(P, -1).max(-1)
        assert d[same].max() < 5e-6 and 0.3 < ref_m.mean() < 1.0
    gl = np.zeros((P, V), np.float32)
    gp = g_pix.reshape(P, 3).contiguous().numpy()
    gq = g_pat.reshape(P, 49, 3).contiguous().numpy() if with_patch else None
    lib.blend_host_backward(ctypes.byref(cfg), fp(p_np), fp(proj), fp(hom), fp(px), fp(imgs), fp(lg), ctypes.c_int64(10), fp(gp),
                            fp(gq), fp(gl))
    ref_g = logits.grad.reshape(P, 10).numpy()
    assert np.abs(ref_g[:, V:]).max() == 0.0
    err = np.abs(gl - ref_g[:, :V]).max(-1)
    assert err.max() < 3e-5 * max(1.0, np.abs(ref_g).max()), err.max()
