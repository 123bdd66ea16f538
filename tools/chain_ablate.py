"""Profiling aid: times the fused F+R and T+B launches (and the weight gradients) of 65 536 points through the library's own
per-family event pairs.  `python tools/chain_ablate.py 0` is the plain A/B timer used to compare library builds on one box.
The masks 1..63 were the NUDF_CHAIN_DEBUG ablation switches of an instrumented intermediate build of udf_chain.cuh (no global
stores / no MUFU / no TMEM parking / no accumulator loads / no auxiliary loads / no operand slicing): the shipped kernel carries
no instrumentation and ignores them; the recorded ablation is profiles/r02_chain_ablation.txt."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from neuraludf_b200 import _lib  # noqa: E402

dev = torch.device("cuda", 0)
udf, col, var, beta = bench.scene(dev)
x = (torch.rand(65536, 3, device=dev) * 2 - 1) * 0.9
ob = torch.randn(65536, 257, device=dev)
gb = torch.randn(65536, 3, device=dev)
L = _lib.lib()
names = _lib.LAUNCH_FAMILIES


def run(mask, n=5):
    import ctypes
    os.environ["NUDF_CHAIN_DEBUG"] = str(mask)
    for it in range(n + 2):
        if it == 2:
            torch.cuda.synchronize()
            _lib.check(L.nudf_set_launch_timing(1), "timing on")
        for p in udf.parameters():
            p.grad = None
        out, grad = udf.value_and_gradient(x)
        ((out * ob).sum() + (grad * gb).sum()).backward()
    torch.cuda.synchronize()
    nf = L.nudf_launch_family_count()
    ms = (ctypes.c_float * nf)()
    cnt = (ctypes.c_int32 * nf)()
    _lib.check(L.nudf_read_launch_timing(ms, cnt), "read")
    L.nudf_set_launch_timing(0)
    return {names[i]: (ms[i] * 1e3 / max(cnt[i], 1), cnt[i]) for i in range(nf)}


masks = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 4, 8, 16, 32, 1 | 16, 2 | 1 | 16, 63]
for mask in masks:
    fam = run(mask)
    f = fam.get("udf_fwd_chain_fused", (0, 0))
    b = fam.get("udf_bwd_chain_fused", (0, 0))
    w = fam.get("tc_weight_gradient", (0, 0))
    print("mask %2d   F+R %8.1f us (%d)   T+B %8.1f us (%d)   wgrad %7.1f us/launch (%d)" % (mask, f[0], f[1], b[0], b[1], w[0], w[1]), flush=True)
