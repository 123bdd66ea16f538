"""NeRF++ parameter gradients: tensor engine vs exact-fp32 engine on identical inputs (locates engine-specific errors)."""
import sys

import torch

sys.path.insert(0, ".")
from neuraludf_b200 import _lib as L
from tests.golden_util import load_golden
from tests.gpu_util import build_modules

lib = L.lib()
g = load_golden()
gen = torch.Generator().manual_seed(21)
P = 700
pts4 = torch.randn(P, 4, generator=gen, dtype=torch.float64)
pts4 = (pts4 / pts4[:, :3].norm(dim=1, keepdim=True)).float().cuda()
dirs = torch.randn(P, 3, generator=gen, dtype=torch.float64)
dirs = (dirs / dirs.norm(dim=1, keepdim=True)).float().cuda()
ab = torch.randn(P, 1, generator=gen, dtype=torch.float64).float().cuda()
rb = torch.randn(P, 3, generator=gen, dtype=torch.float64).float().cuda()
res = {}
for engine in (0, 1):
    lib.nudf_set_engine(engine)
    _, _, nerf, _, _ = build_modules(g, "cuda")
    a, rgb = nerf(pts4, dirs)
    ((a * ab).sum() + (rgb * rb).sum()).backward()
    res[engine] = {k: v.grad.clone() for k, v in nerf.named_parameters()}
for k in res[0]:
    d = (res[0][k] - res[1][k]).abs()
    s = res[0][k].abs().max()
    line = "%-28s rel %.3e" % (k, float(d.max() / s))
    if d.dim() == 2 and d.shape[1] > 256:
        line += "   cols<256: %.3e   cols>=256: %.3e" % (float(d[:, :256].max() / s), float(d[:, 256:].max() / s))
        idx = (d == d.max()).nonzero()[0].tolist()
        line += "   argmax %s" % idx
    print(line)
