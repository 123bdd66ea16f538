"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/nudf.h
declares; the product path refuses to run without CUDA (no silent fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    txt = open(os.path.join(ROOT, "include", "nudf.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(nudf_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    from neuraludf_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "libnudf.so missing: run `python -m neuraludf_b200.build`"
    L = _lib.lib()
    declared = _header_functions()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(L, name), "libnudf.so does not export %s" % name
    assert sorted(_lib.exported_symbols()) == declared, "python binding and include/nudf.h disagree"
    assert L.nudf_abi_version() == 3


def test_descriptor_validation_runs_without_gpu():
    from neuraludf_b200 import _lib
    import ctypes
    L = _lib.lib()
    d = _lib.UdfDesc()
    d.n_lin = 1          # invalid: needs >= 2
    assert L.nudf_udf_folded_floats(ctypes.byref(d)) == -1
    assert b"n_lin" in L.nudf_last_error()
    # a valid DTU-shaped descriptor: sizes are computed on the host
    d = _lib.UdfDesc()
    d.n_lin, d.d_in, d.multires, d.d_out, d.skip_layer, d.scale = 9, 3, 6, 257, 4, 1.0
    dims = [(39, 256), (256, 256), (256, 256), (256, 217), (256, 256), (256, 256), (256, 256), (256, 256), (256, 257)]
    for l, (i, o) in enumerate(dims):
        d.in_dim[l], d.out_dim[l] = i, o
    n = L.nudf_udf_folded_floats(ctypes.byref(d))
    # fp32 folded weights (>= 524 544 floats) followed by the bf16 hi/lo tensor-engine images of every layer
    assert 524544 <= n < 8 * 1024 * 1024
    assert L.nudf_udf_ctx_floats(ctypes.byref(d), 1024, 1) > L.nudf_udf_ctx_floats(ctypes.byref(d), 1024, 0) > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from neuraludf_b200.models.fields import UDFNetwork
    net = UDFNetwork(d_in=3, d_out=257, d_hidden=64, n_layers=4, skip_in=(2,), multires=6)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(torch.zeros(4, 3))


def test_state_dict_layout_matches_reference_names():
    from neuraludf_b200.models.fields import UDFNetwork, ResidualRenderingNetwork, NeRF, SingleVarianceNetwork, BetaNetwork
    udf = UDFNetwork(d_in=3, d_out=257, d_hidden=256, n_layers=8, skip_in=(4,), multires=6, bias=0.5, scale=1.0,
                     geometric_init=True, weight_norm=True, udf_type='abs')
    sd = udf.state_dict()
    assert sum(v.numel() for v in sd.values()) == 529076
    assert tuple(sd["lin3.weight_v"].shape) == (217, 256) and tuple(sd["lin3.weight_g"].shape) == (217, 1)
    assert tuple(sd["lin8.weight_v"].shape) == (257, 256) and tuple(sd["lin0.weight_v"].shape) == (256, 39)
    col = ResidualRenderingNetwork(d_feature=256, mode='no_normal', d_in=6, d_out=3, d_hidden=128, n_layers=4,
                                   weight_norm=True, multires_view=4, squeeze_out=True, blending_cand_views=10)
    sdc = col.state_dict()
    assert sum(v.numel() for v in sdc.values()) == 155808
    assert tuple(sdc["lin0.weight_v"].shape) == (128, 158) and tuple(sdc["lin_base0.weight_v"].shape) == (128, 259)
    nerf = NeRF(D=8, d_in=4, d_in_view=3, W=256, multires=10, multires_view=4, output_ch=4, skips=[4], use_viewdirs=True)
    assert sum(v.numel() for v in nerf.state_dict().values()) == 606596
    assert tuple(nerf.state_dict()["pts_linears.5.weight"].shape) == (256, 340)
    assert list(SingleVarianceNetwork(0.3).state_dict()) == ["variance"]
    assert sorted(BetaNetwork().state_dict()) == ["beta", "gamma", "zeta"]


def test_golden_scene_loads_into_modules(golden):
    from tests.gpu_util import build_modules
    build_modules(golden, device="cpu")
    build_modules(golden, device="cpu", udf_name="udf_small")


def test_launcher_shadows_the_reference_module_names():
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r); from neuraludf_b200.launch import install_shadow_modules; "
            "install_shadow_modules(); from models.fields import UDFNetwork, ResidualRenderingNetwork, NeRF, "
            "SingleVarianceNetwork, BetaNetwork, SDFNetwork; from models.udf_renderer_blending import "
            "UDFRendererBlending, extract_fields, extract_gradient_fields, sample_pdf; from models.embedder import "
            "get_embedder; import models.fields as f; assert 'neuraludf_b200' in f.__file__; print('ok')" % ROOT)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]
