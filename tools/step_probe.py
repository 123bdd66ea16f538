"""One C2 training step (render_core fwd + bwd) with the fused-chain pipeline trace switched on:
    NUDF_CHAIN_TRACE=2 python tools/step_probe.py        # traces the F+R launch and the T+B launch of the first step
then times 10 steps."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from neuraludf_b200.models.udf_renderer_blending import UDFRendererBlending  # noqa: E402

dev = torch.device("cuda", 0)
udf, col, var, beta = bench.scene(dev)
ren = UDFRendererBlending(None, udf, var, col, beta, n_samples=128, n_importance=0, n_outside=0, up_sample_steps=1, perturb=0.0)
ren.want_diagnostics = False
o, d, z, sd = bench.rays(seed=0, device=dev)
tgt = torch.full((512, 3), 0.4, device=dev)
params = [p for m in (udf, col, var, beta) for p in m.parameters() if p.requires_grad]


def step():
    for p in params:
        p.grad = None
    ret = ren.render_core(o, d, z, sd, udf, var, col, beta_network=beta, cos_anneal_ratio=0.5)
    bench.loss_fn(ret, tgt).backward()


for _ in range(4):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record()
torch.cuda.synchronize()
print("C2 step (no optimiser): %.3f ms" % (e0.elapsed_time(e1) / 10))
