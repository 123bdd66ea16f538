"""Timing / tracing harness of the fused UDF value chain (csrc/udf_chain.cuh) on the C2 point count.

    python tools/chain_probe.py [P] [value|full]            # CUDA-event timing, 10 launches
    NUDF_CHAIN_TRACE=1 python tools/chain_probe.py          # + CTA 0's pipeline stamps (stderr)
    ncu --set full -k regex:udf_chain -c 2 ... python tools/chain_probe.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuraludf_b200 import synthetic as S          # noqa: E402
from neuraludf_b200.models import fields as F      # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
mode = sys.argv[2] if len(sys.argv) > 2 else "value"
dev = torch.device("cuda", 0)
udf = F.UDFNetwork(d_in=3, d_out=257, d_hidden=256, n_layers=8, skip_in=(4,), multires=6, bias=0.5, scale=1.0,
                   geometric_init=True, weight_norm=True, udf_type="abs")
udf.load_state_dict(S.make_udf_params(S.udf_cfg(), seed=0))
udf = udf.to(dev)
x = (torch.rand(P, 3, device=dev) * 1.8 - 0.9).contiguous()
fn = (lambda: udf.udf_values(x)) if mode == "value" else (lambda: udf(x))
with torch.no_grad():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 10 * 1e3
print("chain %s P=%d: %.1f us per launch, %.1f algorithmic TFLOP/s (value chain 918016 FLOP/pt)" % (mode, P, us, P * 918016 / us / 1e6))
