"""Seeded synthetic scene for benchmarks, smoke tests and parity tests.

Network shape records and parameter generators for a geometric-init sphere UDF (radius 0.5) with small weight noise,
default-init colour / NeRF++ networks, and camera rays on a radius-2.5 shell.  Pure torch-CPU data generation: no
kernels, no dependency on the CUDA library or on the oracle.  Parameter dicts use the reference's state_dict names.
"""
import math

import torch

# ----------------------------------------------------------------------------------------------
# network configuration records (plain dicts)
# ----------------------------------------------------------------------------------------------

def udf_cfg(d_in=3, d_out=257, d_hidden=256, n_layers=8, skip_in=(4,), multires=6, scale=1.0,
            bias=0.5, udf_type="abs"):
    """Shapes of UDFNetwork, models/fields.py:116-178."""
    d_pe = d_in * (1 + 2 * multires) if multires > 0 else d_in
    dims = [d_pe] + [d_hidden] * n_layers + [d_out]
    layers = []
    for l in range(len(dims) - 1):
        out = dims[l + 1] - dims[0] if (l + 1) in skip_in else dims[l + 1]
        layers.append((dims[l], out))
    return dict(d_in=d_in, d_out=d_out, d_hidden=d_hidden, n_layers=n_layers, skip_in=tuple(skip_in),
                multires=multires, scale=float(scale), bias=float(bias), udf_type=udf_type,
                d_pe=d_pe, layers=layers)


def color_cfg(d_feature=256, d_in=6, d_out=3, d_hidden=128, n_layers=4, multires_view=4,
              blending_cand_views=10, mode="no_normal"):
    """Shapes of ResidualRenderingNetwork, models/fields.py:401-450 (mode no_normal only)."""
    assert mode == "no_normal"
    d_view = 3 * (1 + 2 * multires_view) if multires_view > 0 else 3
    dims_base = [d_in - 3 + d_feature] + [d_hidden] * n_layers + [d_out]
    dims = [d_hidden + d_out + d_view] + [d_hidden] * n_layers + [d_out + blending_cand_views]
    return dict(d_feature=d_feature, d_in=d_in, d_out=d_out, d_hidden=d_hidden, n_layers=n_layers,
                multires_view=multires_view, blending_cand_views=blending_cand_views, mode=mode,
                d_view=d_view, dims_base=dims_base, dims=dims)


def nerf_cfg(D=8, W=256, d_in=4, d_in_view=3, multires=10, multires_view=4, skips=(4,)):
    """Shapes of NeRF (NeRF++ background), models/fields.py:542-594, use_viewdirs=True."""
    ch = d_in * (1 + 2 * multires) if multires > 0 else d_in
    chv = d_in_view * (1 + 2 * multires_view) if multires_view > 0 else d_in_view
    return dict(D=D, W=W, d_in=d_in, d_in_view=d_in_view, multires=multires,
                multires_view=multires_view, skips=tuple(skips), input_ch=ch, input_ch_view=chv)


# ----------------------------------------------------------------------------------------------
# seeded synthetic scene: geometric-init sphere + small noise ----------------------
# ----------------------------------------------------------------------------------------------

def _randn(gen, *shape):
    return torch.randn(*shape, generator=gen, dtype=torch.float64)


def _rand(gen, *shape):
    return torch.rand(*shape, generator=gen, dtype=torch.float64)


def make_udf_params(cfg, seed=0, noise=1e-3):
    """Geometric initialisation, models/fields.py:156-173, then weight-norm split (g = row norms)
    and a small perturbation so that the scene is not exactly a sphere."""
    gen = torch.Generator().manual_seed(seed)
    p = {}
    n_lin = len(cfg["layers"])
    d_pe = cfg["d_pe"]
    for l, (din, dout) in enumerate(cfg["layers"]):
        if l == n_lin - 1:
            w = math.sqrt(math.pi) / math.sqrt(din) + 1e-4 * _randn(gen, dout, din)
            b = torch.full((dout,), -cfg["bias"], dtype=torch.float64)
        elif cfg["multires"] > 0 and l == 0:
            w = torch.zeros(dout, din, dtype=torch.float64)
            w[:, :3] = _randn(gen, dout, 3) * math.sqrt(2) / math.sqrt(dout)
            b = torch.zeros(dout, dtype=torch.float64)
        elif cfg["multires"] > 0 and l in cfg["skip_in"]:
            w = _randn(gen, dout, din) * math.sqrt(2) / math.sqrt(dout)
            w[:, -(d_pe - 3):] = 0.0
            b = torch.zeros(dout, dtype=torch.float64)
        else:
            w = _randn(gen, dout, din) * math.sqrt(2) / math.sqrt(dout)
            b = torch.zeros(dout, dtype=torch.float64)
        w = w + noise * _randn(gen, dout, din)
        b = b + noise * _randn(gen, dout)
        # legacy nn.utils.weight_norm(dim=0): g = ||w||_row, v = w
        p["lin%d.weight_g" % l] = w.norm(dim=1, keepdim=True).float()
        p["lin%d.weight_v" % l] = w.float()
        p["lin%d.bias" % l] = b.float()
    return p


def _default_linear(gen, dout, din):
    bound = 1.0 / math.sqrt(din)
    w = (torch.rand(dout, din, generator=gen, dtype=torch.float64) * 2 - 1) * bound
    b = (torch.rand(dout, generator=gen, dtype=torch.float64) * 2 - 1) * bound
    return w, b


def make_color_params(cfg, seed=1, noise=2e-2):
    gen = torch.Generator().manual_seed(seed)
    p = {}
    for prefix, dims in (("lin", cfg["dims"]), ("lin_base", cfg["dims_base"])):
        for l in range(len(dims) - 1):
            w, b = _default_linear(gen, dims[l + 1], dims[l])
            w = w + noise * _randn(gen, dims[l + 1], dims[l])
            b = b + noise * _randn(gen, dims[l + 1])
            p["%s%d.weight_g" % (prefix, l)] = w.norm(dim=1, keepdim=True).float()
            p["%s%d.weight_v" % (prefix, l)] = w.float()
            p["%s%d.bias" % (prefix, l)] = b.float()
    return p


def make_nerf_params(cfg, seed=2, noise=2e-2):
    gen = torch.Generator().manual_seed(seed)
    W, ch, chv = cfg["W"], cfg["input_ch"], cfg["input_ch_view"]
    shapes = {}
    shapes["pts_linears.0"] = (W, ch)
    for i in range(cfg["D"] - 1):
        shapes["pts_linears.%d" % (i + 1)] = (W, W + ch) if i in cfg["skips"] else (W, W)
    shapes["views_linears.0"] = (W // 2, chv + W)
    shapes["feature_linear"] = (W, W)
    shapes["alpha_linear"] = (1, W)
    shapes["rgb_linear"] = (3, W // 2)
    p = {}
    for name, (dout, din) in shapes.items():
        w, b = _default_linear(gen, dout, din)
        w = w + noise * _randn(gen, dout, din) / math.sqrt(din)
        p[name + ".weight"] = w.float()
        p[name + ".bias"] = b.float()
    return p


def make_scalars(variance=0.6, beta=0.5, gamma=0.3, zeta=0.3):
    """deviation_network.variance and beta_network.{beta,gamma,zeta}; confs/udf_dtu_blending.conf:83-106."""
    return {"variance": torch.tensor([variance]), "beta": torch.tensor([beta]),
            "gamma": torch.tensor([gamma]), "zeta": torch.tensor([zeta])}


def make_rays(n_rays, seed=0):
    """Cameras on a radius-2.5 shell looking at the origin; near/far = mid -/+ 1 (dataset/dataset.py:329-335)."""
    gen = torch.Generator().manual_seed(1000 + seed)
    o = _randn(gen, n_rays, 3)
    o = 2.5 * o / o.norm(dim=1, keepdim=True)
    d = -o / o.norm(dim=1, keepdim=True) + 0.15 * _randn(gen, n_rays, 3)
    d = d / d.norm(dim=1, keepdim=True)
    a = (d * d).sum(-1, keepdim=True)
    b = 2.0 * (o * d).sum(-1, keepdim=True)
    mid = 0.5 * (-b) / a
    return o.float(), d.float(), (mid - 1.0).float(), (mid + 1.0).float()


def _look_at(cam_pos):
    """OpenCV-convention camera-to-world matrix (x right, y down, z forward) of a camera at cam_pos looking at 0."""
    z = -cam_pos / cam_pos.norm()
    up = torch.tensor([0.0, 0.0, 1.0], dtype=cam_pos.dtype)
    if abs(float((z * up).sum())) > 0.95:
        up = torch.tensor([0.0, 1.0, 0.0], dtype=cam_pos.dtype)
    x = torch.linalg.cross(z, up)
    x = x / x.norm()
    y = torch.linalg.cross(z, x)
    c2w = torch.eye(4, dtype=cam_pos.dtype)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = x, y, z, cam_pos
    return c2w


def make_blend_views(n_rays, n_views=6, height=48, width=64, seed=0):
    """A query camera plus n_views source views around it for the pixel / patch blending stage (config C3): the inputs
    exp_runner_blending.py:270-306 takes from dataset/dataset.py -- rays of random query pixels, their normalised uv,
    source images (smooth synthetic textures), world-to-camera matrices and a shared 4x4 intrinsic matrix."""
    gen = torch.Generator().manual_seed(2000 + seed)
    q = _randn(gen, 3)
    q = 2.5 * q / q.norm()
    focal = 1.1 * width
    K = torch.eye(4, dtype=torch.float64)
    K[0, 0], K[1, 1], K[0, 2], K[1, 2] = focal, focal, (width - 1) / 2.0, (height - 1) / 2.0
    query_c2w = _look_at(q.double())
    c2ws = []
    for _ in range(n_views):
        p = q.double() + 0.9 * _randn(gen, 3).double()
        c2ws.append(_look_at(2.5 * p / p.norm()))
    c2ws = torch.stack(c2ws)
    w2cs = torch.inverse(c2ws)
    # rays through random (non-integer) pixels of the query image, kept 6 px away from the border
    u = 6 + _rand(gen, n_rays).double() * (width - 13)
    v = 6 + _rand(gen, n_rays).double() * (height - 13)
    pix = torch.stack([u, v, torch.ones_like(u)], dim=-1)
    d_cam = pix @ torch.inverse(K[:3, :3]).T
    d = d_cam @ query_c2w[:3, :3].T
    d = d / d.norm(dim=1, keepdim=True)
    o = query_c2w[:3, 3].expand(n_rays, 3).clone()
    b = 2.0 * (o * d).sum(-1, keepdim=True)
    mid = 0.5 * (-b)
    uv = torch.stack([2 * u / (width - 1) - 1, 2 * v / (height - 1) - 1], dim=-1)
    # smooth textures: a few random plane waves per channel, in (0,1)
    yy, xx = torch.meshgrid(torch.arange(height, dtype=torch.float64), torch.arange(width, dtype=torch.float64),
                            indexing="ij")
    imgs = torch.zeros(n_views, 3, height, width, dtype=torch.float64)
    for k in range(4):
        fx = (_rand(gen, n_views, 3, 1, 1).double() - 0.5) * 0.6
        fy = (_rand(gen, n_views, 3, 1, 1).double() - 0.5) * 0.6
        ph = _rand(gen, n_views, 3, 1, 1).double() * 6.28318
        imgs = imgs + torch.sin(fx * xx + fy * yy + ph) / 8.0
    imgs = imgs + 0.5
    return {"rays_o": o.float(), "rays_d": d.float(), "near": (mid - 1.0).float(), "far": (mid + 1.0).float(),
            "rays_uv": uv.float(), "color_maps": imgs.float(), "w2cs": w2cs.float(),
            "intrinsics": K.float().expand(n_views, 4, 4).contiguous(), "query_c2w": query_c2w.float()}
