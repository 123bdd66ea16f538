"""Drop-in proof: the reference's UNMODIFIED `exp_runner_blending.py` (byte-for-byte staged
copy under oracle/_ref, see oracle/make_ref.py) trains on top of the nudf modules through `neuraludf_b200.launch`:

    python -m neuraludf_b200.launch <ref>/exp_runner_blending.py --mode train --conf <conf> --case synth

on a synthetic DTU-layout dataset, with the reference's own dataset loader, losses, Adam groups, LR schedules, TensorBoard
writer, checkpointing (`save_checkpoint` / `--is_continue` -> `load_checkpoint`) and `validate()` image rendering.  Only
third-party modules this image lacks are stubbed (tests/runner_env.py)."""
import glob
import os
import subprocess
import sys

import pytest
import torch

from oracle import refshim

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

DRIVER = """
import sys
sys.path.insert(0, {root!r})
from tests import runner_env
runner_env.install_stubs()
import torch
_load = torch.load
# the reference targets torch 2.0: its checkpoints hold numpy scalars (`iter_step`, learning rates), which torch >= 2.6 refuses
# under the new `weights_only=True` default of torch.load -- restore the old default for the runner's own load_checkpoint
torch.load = lambda *a, **k: _load(*a, **dict(dict(weights_only=False), **k))
from neuraludf_b200 import launch
try:
    rc = launch.main({argv!r})
except NotImplementedError as e:
    # `--mode train` ends with runner.extract_udf_mesh(resolution=512) (exp_runner_blending.py:900-902): the dense 512^3 UDF +
    # gradient grid query runs on the nudf kernels, then the Cython MeshUDF marching cubes (custom_mc, SURVEY.md 2 row 10, out
    # of scope, absent from this image) is called -- the stub raises there, after the training loop has finished
    if "custom_mc" not in str(e):
        raise
    print("REACHED_CUSTOM_MC_STUB_AFTER_TRAINING")
    rc = 0
sys.exit(rc)
"""


def _run(tmp, argv, timeout=900):
    drv = os.path.join(tmp, "drive.py")
    with open(drv, "w") as f:
        f.write(DRIVER.format(root=ROOT, argv=argv))
    env = dict(os.environ, PYTHONUNBUFFERED="1")
    r = subprocess.run([sys.executable, drv], cwd=tmp, env=env, capture_output=True, text=True, timeout=timeout)
    return r


@pytest.mark.skipif(not refshim.available(), reason="no staged reference copy (oracle/make_ref.py)")
def test_unmodified_runner_trains_checkpoints_and_validates(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from tests import runner_env
    ref = refshim.REFERENCE_ROOT
    runner = os.path.join(ref, "exp_runner_blending.py")
    tmp = str(tmp_path)
    data = runner_env.write_synthetic_dtu(os.path.join(tmp, "data", "synth"), n_images=12, width=96, height=72)
    exp = os.path.join(tmp, "exp", "CASE_NAME") + "/"
    conf = runner_env.write_conf(ref, os.path.join(tmp, "synth.conf"), os.path.join(tmp, "data", "CASE_NAME") + "/", exp,
                                 end_iter=4, batch_size=256, save_freq=2, val_freq=3)
    r = _run(tmp, [runner, "--mode", "train", "--conf", conf, "--case", "synth", "--gpu", "0"])
    tail = (r.stdout[-3000:] + "\n---- stderr ----\n" + r.stderr[-3000:])
    assert r.returncode == 0, tail
    exp_dir = os.path.join(tmp, "exp", "synth", "udf_dtu")
    ck = sorted(glob.glob(os.path.join(exp_dir, "checkpoints", "ckpt_*.pth")))
    assert [os.path.basename(c) for c in ck] == ["ckpt_000002.pth", "ckpt_000004.pth"], tail
    sd = torch.load(ck[-1], map_location="cpu", weights_only=False)
    assert sd["iter_step"] == 4
    assert set(sd) == {"nerf", "udf_network_fine", "variance_network_fine", "color_network_fine", "beta_network", "optimizer",
                       "iter_step"}
    assert "lin8.weight_v" in sd["udf_network_fine"] and sd["udf_network_fine"]["lin8.weight_v"].shape == (257, 256)
    assert "lin_base0.weight_g" in sd["color_network_fine"] and "pts_linears.5.weight" in sd["nerf"]
    assert all(torch.isfinite(v).all() for v in sd["udf_network_fine"].values())
    # the steps changed the parameters (Adam ran on our gradients).  The UDF network itself is frozen during the first
    # `fix_geo_end` = 500 iterations (exp_runner_blending.py:80, 186-191), so look at the colour network
    sd2 = torch.load(ck[0], map_location="cpu", weights_only=False)
    # (and the DTU conf renders with n_outside = 0: the NeRF++ background receives no gradient either)
    assert not torch.equal(sd["color_network_fine"]["lin0.weight_v"], sd2["color_network_fine"]["lin0.weight_v"])
    assert not torch.equal(sd["color_network_fine"]["lin_base0.weight_v"], sd2["color_network_fine"]["lin_base0.weight_v"])
    assert torch.equal(sd["udf_network_fine"]["lin4.weight_v"], sd2["udf_network_fine"]["lin4.weight_v"])
    # validate() at iteration 3 wrote its images (exp_runner_blending.py:604-719)
    imgs = glob.glob(os.path.join(exp_dir, "**", "*.png"), recursive=True)
    assert len(imgs) >= 1, tail
    assert "iter:" in r.stdout and "psnr" in r.stdout
    assert "REACHED_CUSTOM_MC_STUB_AFTER_TRAINING" in r.stdout       # the post-training 512^3 grid query ran on our kernels
    # resume: --is_continue loads the last checkpoint (exp_runner_blending.py:150-162, 467-482) and trains on to iteration 6
    conf2 = runner_env.write_conf(ref, os.path.join(tmp, "synth2.conf"), os.path.join(tmp, "data", "CASE_NAME") + "/", exp,
                                  end_iter=6, batch_size=256, save_freq=2, val_freq=100)
    r2 = _run(tmp, [runner, "--mode", "train", "--conf", conf2, "--case", "synth", "--gpu", "0", "--is_continue"])
    tail2 = (r2.stdout[-3000:] + "\n---- stderr ----\n" + r2.stderr[-3000:])
    assert r2.returncode == 0, tail2
    assert os.path.exists(os.path.join(exp_dir, "checkpoints", "ckpt_000006.pth")), tail2
