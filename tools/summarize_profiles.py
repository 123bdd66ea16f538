"""Turns the ncu outputs brought back in gpurun_out/ into the text summaries committed under profiles/.

    python tools/summarize_profiles.py r01
"""
import collections
import csv
import re
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
out_l = "profiles/%s_launches.txt" % tag
out_k = "profiles/%s_gemm_kernels.txt" % tag

# ---- launch list ----
rows = [r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r) > 5]
for i, r in enumerate(rows):
    if "Kernel Name" in r:
        hdr, start = r, i + 1
        break
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
ls = []
for r in rows[start:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    ls.append((r[ki], v))
per = len(ls) // 4                     # bench.py --steps 1 --warmup 3 --quick  => 4 identical steps
step = ls[-per:]
agg = collections.defaultdict(lambda: [0, 0.0])
for k, v in step:
    k = re.sub(r"\(.*", "", k)
    k = re.sub(r"^void ", "", k)
    agg[k[:110]][0] += 1
    agg[k[:110]][1] += v
tot = sum(v for _, v in step)
with open(out_l, "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 1 --warmup 3 --quick\n")
    f.write("# last of 4 identical render_core fwd+bwd steps (C2: 512 rays x 128 samples); per-launch times are cold-cache and\n")
    f.write("# serialised by the profiler: compare SHARES, not absolutes.  %d launches, %.1f us total.\n" % (per, tot))
    f.write("%10s %6s %6s %9s  %s\n" % ("total_us", "share", "count", "each_us", "kernel"))
    for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write("%10.1f %5.1f%% %6d %9.1f  %s\n" % (t, 100 * t / tot, c, t / c, k))
print(open(out_l).read()[:2500])

# ---- full-set metrics of the GEMM kernels ----
raw = subprocess.run(["ncu", "-i", "gpurun_out/prof_%s_gemm.ncu-rep" % tag, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "l1tex__t_bytes.sum"]
with open(out_k, "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on  python tools/kernel_probe.py\n")
    f.write("# one dense layer / weight-gradient contraction at the C2 layer size (M = 65 536 points, N = K = 256), fp32 in/out;\n")
    f.write("# algorithmic work 8.59 GFLOP per launch, algorithmic HBM bytes 128 MiB (dense: read A, write Y) / 128 MiB (wgrad).\n")
    last = {}
    for r in rows[2:]:                       # our kernels only, the last (warm) launch of each
        n = r[idx["Kernel Name"]]
        if "gemm" in n or "pack_planes" in n:
            last[n] = r
    for r in last.values():
        f.write("\n== %s\n" % r[idx["Kernel Name"]][:150])
        for w in want:
            if w in idx:
                f.write("  %-72s %18s %s\n" % (w, r[idx[w]], units[idx[w]]))
print(open(out_k).read()[:6000])
