"""The hand-derived chains the CUDA path implements (reverse sweep for grad_x udf, tangent + backward chains
for the second-order parameter gradients, weight-norm backward, compositing backward) agree with autograd of
the pinned oracle in fp64."""
import pytest
import torch

from oracle import oracle_torch as O
from tests.golden_util import rel_err
from tests.proto import udf_pipeline as UP


@pytest.mark.parametrize("name", ["udf", "udf_small"])
def test_udf_chains_match_autograd(golden, name):
    g = golden
    cfg = g.udf_c if name == "udf" else g.udf_small_c
    dt = torch.float64
    p = {k: v.clone().requires_grad_(True) for k, v in O.to_dtype(g.params[name], dt).items()}
    gen = torch.Generator().manual_seed(5)
    x = (torch.rand(96, 3, generator=gen, dtype=dt) * 2 - 1) * 0.8
    out_bar = torch.randn(96, cfg["d_out"], generator=gen, dtype=dt)
    grad_bar = torch.randn(96, 3, generator=gen, dtype=dt)

    xg = x.clone().requires_grad_(True)
    out = O.udf_mlp(p, cfg, xg)
    grad = torch.autograd.grad(out[:, :1], xg, torch.ones_like(out[:, :1]), create_graph=True)[0]
    loss = (out * out_bar).sum() + (grad * grad_bar).sum()
    ref = torch.autograd.grad(loss, list(p.values()))
    ref = dict(zip(p.keys(), ref))

    with torch.no_grad():
        pd = {k: v.detach() for k, v in p.items()}
        sv = UP.forward(pd, cfg, x)
        assert rel_err(sv["out"], out) < 1e-10
        assert rel_err(sv["grad"], grad) < 1e-10
        mine = UP.backward(pd, cfg, sv, out_bar, grad_bar)
    for k in ref:
        assert rel_err(mine[k], ref[k]) < 1e-8, k
