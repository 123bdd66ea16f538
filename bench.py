#!/usr/bin/env python
"""Benchmark of the NeuralUDF volume-rendering hot path (BASELINE.json metric: ray-samples/s, render_core fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], "C2" in SURVEY.md 8(d)): 512 rays x 128 uniform samples per GPU, UDF 8x256 +
ResidualRenderingNetwork 2x(4x128) (reference conf dims), synthetic seeded sphere scene, one step =
render_core forward + backward of  L1 colour + 0.01 L1 base colour + 0.1 eikonal  (exp_runner_blending.py:330-371).
For N > 1 (torchrun, one rank per GPU) rays are sharded (weak scaling: 512 rays per GPU) and the step ends with one
NCCL all-reduce of the flat gradient bucket.

Prints ONE JSON line (rank 0).  `value` = device-resident throughput (CUDA events, max over ranks); `e2e` = the same
step driven from pinned HOST buffers through the public module API with the H2D / D2H copies inside the timed region;
`roofline` = the dominant kernel (the 256x256 dense layer) timed alone; `cpu_baseline` = the oracle port on host cores.
`--impl reference` times the oracle port (the reference is pure Python/PyTorch and cannot travel to the GPU box) on
the host cores; under torchrun only rank 0 works.
"""
import argparse
import datetime
import json
import os
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RAYS, N_SAMPLES = 512, 128
FLOP_FWD = 2274560           # per ray-sample, SURVEY 8(d): UDF value 1 049 088 + input-gradient sweep 918 016 + colour 307 456
FLOP_FWD_BWD = 6823680       # forward + backward (data-grad + weight-grad of every contraction)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    hbm=d["hbm_gbs"], source="measured (MEASURED_PEAKS.json)")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback (B200_PROFILING.md)")


def ncu_traffic(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch) of the dominant kernel from the committed
    `ncu --set full` summary under profiles/ (None if not captured)."""
    import glob
    best = None
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_gemm_kernels.txt"))):
        cur, rd, wr = None, None, None
        for line in open(f):
            if line.startswith("== "):
                cur, rd, wr = line, None, None
            elif cur and kernel_substr.replace("nudf::", "").replace("tc::", "") in cur.replace("nudf::", "").replace("tc::", ""):
                parts = line.split()
                if parts and parts[0] == "dram__bytes_read.sum":
                    rd = float(parts[1]) * (1e6 if parts[2].startswith("Mbyte") else 1e3 if parts[2].startswith("Kbyte") else 1e9 if parts[2].startswith("Gbyte") else 1)
                if parts and parts[0] == "dram__bytes_write.sum":
                    wr = float(parts[1]) * (1e6 if parts[2].startswith("Mbyte") else 1e3 if parts[2].startswith("Kbyte") else 1e9 if parts[2].startswith("Gbyte") else 1)
                if rd is not None and wr is not None:
                    best = rd + wr
                    cur = None
    return best


def nerf_module(device):
    from neuraludf_b200 import synthetic as O
    from neuraludf_b200.models import fields as F
    nerf = F.NeRF(D=8, d_in=4, d_in_view=3, W=256, multires=10, multires_view=4, output_ch=4, skips=[4], use_viewdirs=True)
    nerf.load_state_dict(O.make_nerf_params(O.nerf_cfg(), seed=2))
    return nerf.to(device)


def scene(device):
    from neuraludf_b200 import synthetic as O
    from neuraludf_b200.models import fields as F
    udf_c, col_c = O.udf_cfg(), O.color_cfg()
    udf = F.UDFNetwork(d_in=3, d_out=257, d_hidden=256, n_layers=8, skip_in=(4,), multires=6, bias=0.5, scale=1.0,
                       geometric_init=True, weight_norm=True, udf_type="abs")
    udf.load_state_dict(O.make_udf_params(udf_c, seed=0))
    col = F.ResidualRenderingNetwork(d_feature=256, mode="no_normal", d_in=6, d_out=3, d_hidden=128, n_layers=4,
                                     weight_norm=True, multires_view=4, squeeze_out=True, blending_cand_views=10)
    col.load_state_dict(O.make_color_params(col_c, seed=1))
    var = F.SingleVarianceNetwork(0.6)
    beta = F.BetaNetwork(init_var_beta=0.5, init_var_gamma=0.3, init_var_zeta=0.3, beta_min=5e-5,
                         requires_grad_beta=True, requires_grad_gamma=False, requires_grad_zeta=False)
    return [m.to(device) for m in (udf, col, var, beta)]


def rays(seed, device=None):
    from neuraludf_b200 import synthetic as O
    o, d, near, far = O.make_rays(N_RAYS, seed=seed)
    z = near + (far - near) * torch.linspace(0.0, 1.0, N_SAMPLES)[None, :]
    sd = float(((far - near) / N_SAMPLES).mean())
    if device is not None:
        o, d, z = o.to(device), d.to(device), z.to(device).contiguous()
    return o, d, z, sd


class Clocks:
    """nvidia-smi sampler (B200_PROFILING.md clocks line).  Started BEFORE the warm-up (nvidia-smi needs ~0.2 s to come
    up), every sample is time-stamped, and only the samples inside the [begin(), end()] wall-clock window -- the timed
    region -- are reported."""

    def __init__(self, idx):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.t0 = self.t1 = None
        q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(idx), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                       "-lms", "10"], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def begin(self):
        deadline = time.time() + 2.0            # the first sample must exist before the timed region starts
        while self.p is not None and os.path.getsize(self.f.name) == 0 and time.time() < deadline:
            time.sleep(0.01)
        self.t0 = datetime.datetime.now()

    def end(self):
        self.t1 = datetime.datetime.now()

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"], "samples": 0}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [l.strip().split(", ") for l in open(self.f.name) if l.strip()]
        os.unlink(self.f.name)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        parsed = []
        for r in rows:
            try:
                ts = datetime.datetime.strptime(r[0].strip(), "%Y/%m/%d %H:%M:%S.%f")
                parsed.append((ts, float(r[1]), float(r[2]), [n for n, v in zip(names, r[3:7])
                                                               if v.strip().lower() == "active"]))
            except Exception:
                pass
        slack = datetime.timedelta(milliseconds=15)
        inside = [x for x in parsed if self.t0 is not None and self.t0 - slack <= x[0] <= self.t1 + slack]
        window = "timed region"
        if not inside and parsed and self.t0 is not None:
            # region shorter than the sampling period: the sample closest to it
            inside = [min(parsed, key=lambda x: abs((x[0] - self.t0).total_seconds()))]
            window = "nearest sample (timed region shorter than the sampling period)"
        sm = sorted(x[1] for x in inside)
        reasons = sorted({n for x in inside for n in x[3]})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(x[2] for x in inside) if inside else None,
                "reasons": reasons, "samples": len(sm), "window": window}


def loss_fn(ret, tgt):
    return ((ret["color"] - tgt).abs().mean() + 0.01 * (ret["color_base"] - tgt).abs().mean()
            + 0.1 * ret["gradient_error"])


def run_ours(args):
    import torch.distributed as dist
    from neuraludf_b200 import _lib
    from neuraludf_b200.models.udf_renderer_blending import UDFRendererBlending
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl ours) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.lib()
    udf, col, var, beta = scene(dev)
    params = [p for m in (udf, col, var, beta) for p in m.parameters() if p.requires_grad]
    ren = UDFRendererBlending(None, udf, var, col, beta, n_samples=N_SAMPLES, n_importance=0, n_outside=0,
                              up_sample_steps=1, perturb=0.0)
    if args.workload == "c4":
        return run_c4(args, dev, lib, udf, col, var, beta, rank, world)
    if args.workload == "c3":
        return run_c3(args, dev, lib, udf, col, var, beta, rank, world)
    o, d, z, sd = rays(seed=rank, device=dev)
    tgt = torch.full((N_RAYS, 3), 0.4, device=dev)
    from neuraludf_b200.dp import GradBucket
    bucket = GradBucket(params)

    def step(o_, d_, z_):
        for p in params:
            p.grad = None
        ret = ren.render_core(o_, d_, z_, sd, udf, var, col, beta_network=beta, cos_anneal_ratio=0.5)
        loss = loss_fn(ret, tgt)
        loss.backward()
        if world > 1:
            bucket.allreduce_mean()                    # ONE NCCL all-reduce of the flat bucket (0.69 M floats) over NVLink
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = Clocks(local) if rank == 0 else None
    for _ in range(max(args.warmup, 3)):
        step(o, d, z)
    barrier()
    l0 = lib.nudf_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if clocks:
        clocks.begin()
    barrier()
    e0.record()
    for _ in range(args.steps):
        step(o, d, z)
    e1.record()
    barrier()
    if clocks:
        clocks.end()
    ms = e0.elapsed_time(e1)
    launches = lib.nudf_launch_count() - l0
    clk = clocks.stop() if clocks else None

    if args.quick:
        if rank == 0:
            print(json.dumps({"quick": True, "ms_per_step": ms / args.steps, "gpu_launches": int(launches),
                              "value": world * N_RAYS * N_SAMPLES * args.steps / (ms * 1e-3)}), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return None

    # ---- end-to-end: host (pinned) inputs, H2D + D2H inside the timed region, through the public module API ----
    ho, hd, hz, _ = rays(seed=rank)
    ho, hd, hz = ho.pin_memory(), hd.pin_memory(), hz.contiguous().pin_memory()
    hres = torch.empty(N_RAYS, 3).pin_memory()
    for _ in range(2):
        l = step(ho.to(dev, non_blocking=True), hd.to(dev, non_blocking=True), hz.to(dev, non_blocking=True))
    barrier()
    t0 = time.perf_counter()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        l = step(ho.to(dev, non_blocking=True), hd.to(dev, non_blocking=True), hz.to(dev, non_blocking=True))
        float(l.item())                                # D2H read of the step's result (loss)
    f1.record()
    barrier()
    ms_e2e = f0.elapsed_time(f1)
    wall_e2e = (time.perf_counter() - t0) * 1e3
    ms_e2e = max(ms_e2e, wall_e2e)                     # host-driven loop: wall clock bounds it from above

    times = torch.tensor([ms, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(times[0]), float(times[1])

    out = None
    if rank == 0:
        pk = peaks()
        samples = world * N_RAYS * N_SAMPLES * args.steps
        value = samples / (ms * 1e-3)
        # ---- dominant kernels alone: one 256x256 dense layer (softplus epilogue) over the step's 65 536 points, on each
        #      engine, and the matching weight-gradient contraction; L2 flushed between timed launches ----
        import ctypes
        P = N_RAYS * N_SAMPLES
        X = torch.randn(P, 256, device=dev) * 0.1
        W = torch.randn(256, 256, device=dev) * 0.06
        b = torch.zeros(256, device=dev)
        Y = torch.empty(P, 256, device=dev)
        dW = torch.zeros(256, 256, device=dev)
        flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
        st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        imgs = {}
        for npl in (2, 3):
            im = torch.zeros(lib.nudf_tc_image_elems(256, 256, npl), dtype=torch.int16, device=dev)
            lib.nudf_tc_prepare_weights(_lib.ptr(W), 256, 256, 256, 0, npl, _lib.ptr(im), st)
            imgs[npl] = im
        calls = {
            "dense_fp32_ffma": lambda: lib.nudf_dense_forward(_lib.ptr(X), 256, _lib.ptr(W), 256, _lib.ptr(b), _lib.ptr(Y), 256, P, 256, 256, 2, st),
            "dense_tcgen05_3xbf16": lambda: lib.nudf_dense_forward_tc(_lib.ptr(X), 256, _lib.ptr(imgs[2]), 2, _lib.ptr(b), _lib.ptr(Y), 256, P, 256, 256, 2, st),
            "dense_tcgen05_6xbf16": lambda: lib.nudf_dense_forward_tc(_lib.ptr(X), 256, _lib.ptr(imgs[3]), 3, _lib.ptr(b), _lib.ptr(Y), 256, P, 256, 256, 2, st),
            "wgrad_tcgen05_3xbf16": lambda: lib.nudf_wgrad(_lib.ptr(Y), 256, _lib.ptr(X), 256, 256, 256, P, _lib.ptr(dW), 256, 1, st),
            "wgrad_fp32_ffma": lambda: lib.nudf_wgrad(_lib.ptr(Y), 256, _lib.ptr(X), 256, 256, 256, P, _lib.ptr(dW), 256, 0, st),
        }
        kflops = 2.0 * P * 256 * 256
        ktimes = {}
        for name, call in calls.items():
            for _ in range(3):
                call()
            tk = 0.0
            reps = 10
            for _ in range(reps):
                flush.zero_()                          # evict L2 between timed launches
                a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); call(); bb.record()
                torch.cuda.synchronize()
                tk += a.elapsed_time(bb)
            ktimes[name] = {"us": tk / reps * 1e3, "algorithmic_tflops": kflops / (tk / reps * 1e-3) / 1e12}
        dom = "dense_tcgen05_3xbf16" if lib.nudf_get_engine() == 1 else "dense_fp32_ffma"
        ach = ktimes[dom]["algorithmic_tflops"]
        engine = lib.nudf_get_engine()
        cpu = cpu_baseline(steps=2, warmup=1, n_rays=256) if world == 1 else None      # host baseline: N = 1 only
        out = {
            "metric": "ray-samples/sec (render_core fwd+bwd)", "value": value, "unit": "ray-samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic (seeded geometric-init sphere scene, random-perturbed weights)",
            "config": {"workload": "C2: 512 rays x 128 uniform samples per GPU, UDF 8x256 + colour 2x(4x128), "
                                   "render_core forward+backward", "rays_per_gpu": N_RAYS, "samples_per_ray": N_SAMPLES,
                       "parallelism": "dp%d (rays sharded, NCCL all-reduce of the flat gradient bucket)" % world,
                       "l2": "per-step working set (~2.7 GB of saved activations) exceeds the 126 MB L2",
                       "engine": ("tcgen05 3xBF16 on chains mask %d + fp32 FFMA elsewhere" % lib.nudf_get_tc_mask())
                       if engine == 1 else "fp32 FFMA",
                       "algorithmic_flop_per_sample": FLOP_FWD_BWD,
                       "step_algorithmic_tflops": value * FLOP_FWD_BWD / 1e12 / world},
            "clocks": clk,
            "e2e": {"value": world * N_RAYS * N_SAMPLES * args.steps / (ms_e2e * 1e-3), "unit": "ray-samples/s",
                    "h2d_bytes_per_step": int((ho.numel() + hd.numel() + hz.numel()) * 4), "d2h_bytes_per_step": 4},
            "gpu_launches": int(launches),
            "kernels": ktimes,
            "roofline": {"bound": "tensor", "kernel": dom + ": dense 65536x256x256 + bias + softplus epilogue (UDF hidden layer)",
                         "achieved": ach, "peak": pk["bf16_burst"], "unit": "TFLOP/s", "frac": ach / pk["bf16_burst"],
                         "traffic": ncu_traffic("gemm_wr_kernel<nudf::EpiAct>" if dom.startswith("dense_tc") else "gemm_simt_kernel<1, 1, nudf::EpiAct>"),
                         "traffic_unit": "bytes per launch (dram read+write, ncu --set full, profiles/)", "peak_source": pk["source"] + ", bf16 burst",
                         "note": "algorithmic FLOPs (2MNK); the tensor engine executes 3x that (3xBF16 split)",
                         # the same launch seen from the memory side: it reads A and writes Y once (fp32, 128 MiB); at the
                         # measured peaks the tensor bound (3 x 8.6 GFLOP) and the HBM bound are both ~16-20 us
                         "hbm_view": {"algorithmic_bytes": 2 * 65536 * 256 * 4,
                                      "achieved_gbs": 2 * 65536 * 256 * 4 / (ktimes[dom]["us"] * 1e-6) / 1e9,
                                      "peak_gbs": pk["hbm"],
                                      "frac": 2 * 65536 * 256 * 4 / (ktimes[dom]["us"] * 1e-6) / 1e9 / pk["hbm"]}},
            "cpu_baseline": cpu,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


def run_c4(args, dev, lib, udf, col, var, beta, rank, world):
    """Whole render() of the DTU conf (one rank's 512 rays): sampling + NeRF++ background + fine pass,
    forward + backward.  Secondary number (not the BASELINE.json headline); printed as a reduced JSON line."""
    import torch.distributed as dist
    from neuraludf_b200 import synthetic as O
    from neuraludf_b200.dp import GradBucket
    from neuraludf_b200.models.udf_renderer_blending import UDFRendererBlending
    nerf = nerf_module(dev)
    ren = UDFRendererBlending(nerf, udf, var, col, beta, n_samples=64, n_importance=50, n_outside=32, up_sample_steps=5,
                              perturb=1.0)
    params = [p for m in (udf, col, var, beta, nerf) for p in m.parameters() if p.requires_grad]
    bucket = GradBucket(params)
    o, d, near, far = [t.to(dev) for t in O.make_rays(N_RAYS, seed=rank)]
    tgt = torch.full((N_RAYS, 3), 0.4, device=dev)

    def step():
        for p in params:
            p.grad = None
        ret = ren.render(o, d, near, far, cos_anneal_ratio=0.5, flip_saturation=0.1)
        loss = loss_fn(ret, tgt)
        loss.backward()
        if world > 1:
            bucket.allreduce_mean()
        return loss

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    l0 = lib.nudf_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if rank == 0:
        print(json.dumps({"workload": "c4: render() fwd+bwd, confs/udf_dtu_blending.conf shapes, 512 rays per GPU",
                          "ms_per_step": ms, "rays_per_s": world * N_RAYS / (ms * 1e-3),
                          "fine_ray_samples_per_s": world * N_RAYS * 114 / (ms * 1e-3), "n_gpus": world,
                          "gpu_launches_per_step": (lib.nudf_launch_count() - l0) / args.steps}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return None


def run_c3(args, dev, lib, udf, col, var, beta, rank, world):
    """Whole render() of the fine-tuning stage on open-surface shapes: 1024 rays x (64 + 64 importance)
    samples, no outside samples, pixel + patch blending on (8 source views of 1024 x 1024, 7 x 7 patches), forward +
    backward.  Secondary number; printed as a reduced JSON line."""
    import torch.distributed as dist
    from neuraludf_b200 import synthetic as O
    from neuraludf_b200.dp import GradBucket
    from neuraludf_b200.models.udf_renderer_blending import UDFRendererBlending
    n_rays = 1024
    ren = UDFRendererBlending(None, udf, var, col, beta, n_samples=64, n_importance=64, n_outside=0, up_sample_steps=4,
                              perturb=1.0, upsampling_type="classical", h_patch_size=3, use_norm_grad_for_cosine=True)
    params = [p for m in (udf, col, var, beta) for p in m.parameters() if p.requires_grad]
    bucket = GradBucket(params)
    v = {k: t.to(dev) for k, t in O.make_blend_views(n_rays, n_views=8, height=1024, width=1024, seed=rank).items()}
    tgt = torch.full((n_rays, 3), 0.4, device=dev)

    def step():
        for p in params:
            p.grad = None
        ret = ren.render(v["rays_o"], v["rays_d"], v["near"], v["far"], cos_anneal_ratio=1.0, flip_saturation=0.0,
                         color_maps=v["color_maps"], w2cs=v["w2cs"], intrinsics=v["intrinsics"], query_c2w=v["query_c2w"],
                         rays_uv=v["rays_uv"])
        pm = ret["patch_mask"].detach()
        loss = (loss_fn(ret, tgt) + 0.5 * (ret["color_pixel"] - tgt).abs().mean()
                + 0.5 * ((ret["patch_colors"] - 0.4).abs().mean(dim=(1, 2)) * pm).sum() / (pm.sum() + 1e-5))
        loss.backward()
        if world > 1:
            bucket.allreduce_mean()
        return loss

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    l0 = lib.nudf_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if rank == 0:
        print(json.dumps({"workload": "c3: render() fwd+bwd with pixel + patch blending, 1024 rays x (64+64) samples, "
                                      "8 views 1024x1024, 49-pixel patches", "ms_per_step": ms,
                          "rays_per_s": world * n_rays / (ms * 1e-3),
                          "fine_ray_samples_per_s": world * n_rays * 128 / (ms * 1e-3), "n_gpus": world,
                          "gpu_launches_per_step": (lib.nudf_launch_count() - l0) / args.steps}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return None


def cpu_baseline(steps, warmup, n_rays):
    """The oracle port (pinned restatement of the reference's PyTorch code) on the host cores: a bounded sample of the
    same workload (n_rays of the 512 rays x 128 samples, forward + backward)."""
    from oracle import oracle_torch as O
    udf_c, col_c = O.udf_cfg(), O.color_cfg()
    up = {k: v.clone().requires_grad_(True) for k, v in O.make_udf_params(udf_c, seed=0).items()}
    cp = {k: v.clone().requires_grad_(True) for k, v in O.make_color_params(col_c, seed=1).items()}
    sc = {k: v.clone().requires_grad_(True) for k, v in O.make_scalars().items()}
    o, d, z, sd = rays(seed=0)
    o, d, z = o[:n_rays], d[:n_rays], z[:n_rays]
    tgt = torch.full((n_rays, 3), 0.4)

    def one(nr):
        t0 = time.perf_counter()
        ret = O.render_core(up, udf_c, cp, col_c, sc, o[:nr], d[:nr], z[:nr], sd, cos_anneal_ratio=0.5)
        loss = loss_fn(ret, tgt[:nr])
        torch.autograd.grad(loss, list(up.values()) + list(cp.values()) + [sc["variance"], sc["beta"]])
        return time.perf_counter() - t0

    # give the CPU path its best thread count (oversubscribing a 128-thread host makes torch 10x slower)
    ncpu = os.cpu_count() or 1
    cands = sorted(set(c for c in (4, 8, 16, 32, 64, ncpu) if c <= ncpu))
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        one(16)
        t = one(16)
        if t < best_t:
            best, best_t = c, t
    torch.set_num_threads(best)
    ts = []
    for i in range(warmup + steps):
        dt = one(n_rays)
        if i >= warmup:
            ts.append(dt)
    per = sum(ts) / len(ts)
    return {"value": n_rays * N_SAMPLES / per, "unit": "ray-samples/s", "cores": torch.get_num_threads(),
            "kind": "port", "sample": "%d of 512 rays x 128 samples, render_core fwd+bwd, %d timed steps after %d warm-up; "
                                      "thread count chosen as the fastest of %s on this host" % (n_rays, steps, warmup, cands),
            "s_per_step": per}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    n_rays = 256          # the same bounded sample as the cpu_baseline leg of the CUDA arm
    cpu = cpu_baseline(steps=args.steps, warmup=max(1, min(args.warmup, 2)), n_rays=n_rays)
    return {"impl": "reference", "metric": "ray-samples/sec (render_core fwd+bwd)", "value": cpu["value"],
            "unit": "ray-samples/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": cpu["s_per_step"] * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (same seeded scene as the CUDA arm)",
            "config": {"workload": "C2 sample: %d of 512 rays x 128 uniform samples, UDF 8x256 + colour 2x(4x128), "
                                   "render_core forward+backward on the host CPU (oracle port of the reference's "
                                   "PyTorch code; the reference itself is not installable: no setup.py, imports absent "
                                   "modules)" % n_rays},
            "cpu_baseline": cpu,
            "e2e": {"value": cpu["value"], "unit": "ray-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=["c2", "c3", "c4"],
                    help="c2 (default, the BASELINE.json headline): render_core fwd+bwd on 512x128 uniform samples; "
                         "c4: whole render() of confs/udf_dtu_blending.conf (64+50 samples, 32 outside, perturb) fwd+bwd; "
                         "c3: whole render() with pixel + patch blending on, 1024 rays x (64+64) samples, 8 views")
    ap.add_argument("--quick", action="store_true", help="main timed loop only (for profiler runs): no e2e leg, no "
                    "single-kernel probes, no CPU baseline; prints a reduced JSON line")
    args = ap.parse_args()
    out = run_reference(args) if args.impl == "reference" else run_ours(args)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
