"""Device-side grid query / near-surface cell emission (neuraludf_b200/grid.py) against the reference's own
get_udf_normals_grid_slow (extract_mesh.py:18-105, staged copy) driven with the same network callbacks."""
import importlib
import os
import sys

import pytest
import torch

from oracle import refshim
from tests.gpu_util import build_modules

pytestmark = pytest.mark.gpu


def _ref_extract_mesh():
    from tests import runner_env
    runner_env.install_stubs()
    if refshim.REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, refshim.REFERENCE_ROOT)
    for k in [k for k in sys.modules if k == "extract_mesh"]:
        del sys.modules[k]
    return importlib.import_module("extract_mesh")


@pytest.mark.skipif(not os.path.isfile(os.path.join(refshim.REFERENCE_ROOT, "extract_mesh.py")), reason="no staged reference copy")
def test_grid_and_near_surface_cells_match_reference(golden):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from neuraludf_b200 import grid
    udf = build_modules(golden, "cuda")[0]
    N = 48
    func = udf.udf

    def func_grad(xyz):                                    # exp_runner_blending.py:767-771
        g = udf.gradient(xyz)
        return g / (torch.linalg.norm(g, ord=2, dim=-1, keepdim=True) + 1e-5)

    em = _ref_extract_mesh()
    # the reference mixes device-less tensors with .cuda() results and relies on the runner's global default
    # (exp_runner_blending.py:872)
    torch.set_default_tensor_type("torch.cuda.FloatTensor")
    try:
        df_ref, vec_ref, samples_ref = em.get_udf_normals_grid_slow(func, func_grad, N=N, max_batch=1 << 14)
    finally:
        torch.set_default_tensor_type("torch.FloatTensor")
    df_ref, vec_ref, samples_ref = df_ref.cpu(), vec_ref.cpu(), samples_ref.cpu()
    df, vec, samples = grid.get_udf_normals_grid_slow(udf, N=N, max_batch=1 << 15)
    assert df.shape == df_ref.shape and vec.shape == vec_ref.shape and samples.shape == samples_ref.shape
    assert float((samples[:, :3] - samples_ref[:, :3]).abs().max()) < 1e-6          # lattice coordinates
    assert float((df - df_ref).abs().max()) < 1e-6                                   # same kernels, different batching
    mask_ref = (vec_ref.abs().sum(-1) > 0)
    mask = (vec.abs().sum(-1) > 0)
    assert torch.equal(mask, mask_ref) and int(mask.sum()) > 100                     # the same near-surface cells
    assert float((vec - vec_ref).abs().max()) < 1e-4
    idx, normals = grid.near_surface_cells(udf, N)
    assert idx.numel() == int(mask.sum())
    assert float((normals.norm(dim=1) - 1).abs().max()) < 1e-4
    # slab partition (multi-GPU sweeps): two halves reproduce the whole grid
    half = N ** 3 // 2
    a = grid.udf_grid(udf, N, lo=0, hi=half)
    b = grid.udf_grid(udf, N, lo=half, hi=N ** 3)
    assert torch.equal(torch.cat([a, b]).cpu(), df.reshape(-1))
