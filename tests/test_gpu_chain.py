"""GPU tests of the fused UDF value chain (csrc/udf_chain.cuh): one tcgen05 kernel walks all layers of UDFNetwork.forward
(reference models/fields.py:192-211) with the activations resident on chip and an exact-main fp16 slice scheme.  Checked
against the pinned oracle in fp64 (arbiter) and fp32 (the reference's own rounding noise), and against the exact-fp32 FFMA
engine, on both network shapes, ragged point counts and the value-only / full-output / saved-context variants."""
import pytest
import torch

from oracle import oracle_torch as O
from tests.gpu_util import build_modules, err_inf, oracle_params, parity, report, scale_inf

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from neuraludf_b200 import _lib
    L = _lib.lib()
    old, old_mask = L.nudf_get_engine(), L.nudf_get_tc_mask()
    L.nudf_set_engine(1)
    L.nudf_set_tc_mask(L.nudf_default_tc_mask())
    assert L.nudf_default_tc_mask() & 1, "the shipped mask must route the UDF value chain to the fused kernel"
    yield
    L.nudf_set_engine(old)
    L.nudf_set_tc_mask(old_mask)


def _points(P, seed):
    gen = torch.Generator().manual_seed(seed)
    x = (torch.rand(P, 3, generator=gen, dtype=torch.float64) * 2 - 1) * 0.9
    # a third of the points close to the zero level set (radius 0.5 sphere), where udf is small and exp(-25000 udf) matters
    n = P // 3
    if n:
        x[:n] = x[:n] / x[:n].norm(dim=1, keepdim=True) * (0.5 + 2e-4 * torch.randn(n, 1, generator=gen, dtype=torch.float64))
    return x


@pytest.mark.parametrize("name", ["udf", "udf_small"])
@pytest.mark.parametrize("P", [1, 127, 129, 1000, 20000])
def test_chain_outputs_vs_oracle(golden, name, P):
    g = golden
    cfg = g.udf_c if name == "udf" else g.udf_small_c
    udf = build_modules(g, DEV, name)[0]
    x = _points(P, 17 + P)
    ref64 = O.udf_mlp(oracle_params(g, name, torch.float64), cfg, x)
    ref32 = O.udf_mlp(oracle_params(g, name, torch.float32), cfg, x.float())
    xd = x.float().to(DEV)
    out = udf(xd)                                     # full output (udf | 256 features), context saved
    parity("chain.%s.P%d.out" % (name, P), out, ref64, ref32, tol=2e-5)
    vals = udf.udf_values(xd)                         # value-only variant (last layer restricted to the udf head)
    e = err_inf(vals, ref64[:, 0])
    noise = err_inf(ref32[:, 0], ref64[:, 0])
    report("chain.%s.P%d.udf_head" % (name, P), abs_err=e, ref32_abs_noise=noise)
    # the udf head feeds exp(-25000 udf): its ABSOLUTE error must stay at the level of the reference's own fp32 rounding
    assert e <= max(2.5 * noise, 1.5e-6), (e, noise)
    assert err_inf(vals, out[:, 0]) <= 1e-6


def test_chain_matches_ffma_engine_and_saves_the_same_context(golden):
    """forward + the exact input-gradient and every parameter gradient computed from the context the fused kernel saved must
    agree with the exact-fp32 engine (whose forward kernels write the same tensors layer by layer)."""
    from neuraludf_b200 import _lib
    L = _lib.lib()
    g = golden
    x = _points(3000, 5).float().to(DEV)
    gen = torch.Generator().manual_seed(8)
    ob = torch.randn(3000, 257, generator=gen).to(DEV)
    gb = torch.randn(3000, 3, generator=gen).to(DEV)
    res = {}
    for engine in (0, 1):
        L.nudf_set_engine(engine)
        udf = build_modules(g, DEV, "udf")[0]
        out, grad = udf.value_and_gradient(x)
        ((out * ob).sum() + (grad * gb).sum()).backward()
        res[engine] = (out.detach(), grad.detach(), {k: v.grad.clone() for k, v in udf.named_parameters()})
    L.nudf_set_engine(1)
    s = scale_inf(res[0][0])
    report("chain.vs_ffma", out=err_inf(res[1][0], res[0][0]) / s, grad=err_inf(res[1][1], res[0][1]) / scale_inf(res[0][1]))
    assert err_inf(res[1][0], res[0][0]) <= 4e-6 * s
    assert err_inf(res[1][1], res[0][1]) <= 1e-4 * scale_inf(res[0][1])
    for k in res[0][2]:
        a, b = res[1][2][k], res[0][2][k]
        assert err_inf(a, b) <= 2e-3 * scale_inf(b) + 1e-12, k


def test_chain_large_batch_throughput_report(golden):
    """65 536 points (the C2 step's point count): timing of the fused chain, value-only and with the saved context; reported."""
    g = golden
    udf = build_modules(g, DEV, "udf")[0]
    x = _points(65536, 3).float().to(DEV)
    for _ in range(3):
        udf.udf_values(x)
    rep = {}
    for tag, fn in (("value_only", lambda: udf.udf_values(x)), ("with_context", lambda: udf(x))):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            with torch.no_grad():
                fn()
        e1.record()
        torch.cuda.synchronize()
        rep[tag + "_us"] = e0.elapsed_time(e1) / 5 * 1e3
    rep["value_only_algorithmic_tflops"] = 65536 * 918016 / (rep["value_only_us"] * 1e-6) / 1e12
    report("chain.throughput_65536", **rep)
    assert rep["value_only_us"] > 0


def test_split_outputs_match_the_packed_form(golden):
    """value_feature_gradient (separate udf / feature tensors, what render_core consumes) == value_and_gradient ([P, 1+F])
    in the outputs and in every parameter gradient, with upstream gradients on all three outputs or on a subset."""
    g = golden
    x = _points(2500, 11).float().to(DEV)
    gen = torch.Generator().manual_seed(3)
    ob = torch.randn(2500, 257, generator=gen).to(DEV)
    gb = torch.randn(2500, 3, generator=gen).to(DEV)
    for use in ((1, 1, 1), (0, 1, 0), (1, 0, 1)):
        res = []
        for split in (False, True):
            udf = build_modules(g, DEV, "udf")[0]
            if split:
                u, f, grad = udf.value_feature_gradient(x)
                out = torch.cat([u, f], dim=1)
            else:
                out, grad = udf.value_and_gradient(x)
                u, f = out[:, :1], out[:, 1:]
            loss = 0.0
            if use[0]:
                loss = loss + (u * ob[:, :1]).sum()
            if use[1]:
                loss = loss + (f * ob[:, 1:]).sum()
            if use[2]:
                loss = loss + (grad * gb).sum()
            loss.backward()
            res.append((out.detach(), grad.detach(), {k: v.grad.clone() for k, v in udf.named_parameters()}))
        assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
        for k in res[0][2]:
            a, b = res[1][2][k], res[0][2][k]
            assert err_inf(a, b) <= 1e-5 * scale_inf(b) + 1e-12, (use, k)
