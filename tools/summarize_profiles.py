"""Turns the ncu outputs brought back in gpurun_out/ into the text summaries committed under profiles/.

    python tools/summarize_profiles.py --tag r02 --launches gpurun_out/launches22.csv \
        --report gpurun_out/prof_chain22.ncu-rep:chain --report gpurun_out/prof_tn22.ncu-rep:wgrad

* launch list (`ncu --metrics gpu__time_duration.sum --clock-control none ... python tools/step_probe.py`): one C2 step
  (from one launch of the fused F + R chain to the next), aggregated by kernel -> profiles/<tag>_launches.txt
* `ncu --set full` reports -> profiles/<tag>_<name>_kernel.txt: duration, DRAM bytes (the `traffic` of bench.py's roofline),
  pipe activity, issue activity, the largest stall reasons
"""
import argparse
import collections
import csv
import re
import subprocess


def launches(path, out):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi = hdr.index("Grid Size") if "Grid Size" in hdr else None
    ls = []
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
        ls.append((r[ki], v, r[gi] if gi is not None else ""))
    idx = [i for i, (k, _, _) in enumerate(ls) if "udf_chain_kernel" in k]
    assert len(idx) >= 3, "need at least one whole step (F+R chain ... next F+R chain) in the launch list"
    step = ls[idx[0]:idx[2]]
    agg = collections.OrderedDict()
    for k, v, _ in step:
        k = re.sub(r"\(.*", "", k)
        k = re.sub(r"^void ", "", k)[:110]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v, _ in step)
    with open(out, "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv  python tools/step_probe.py\n")
        f.write("# one C2 step (512 rays x 128 samples: render_core forward + backward, no optimiser), from the launch of the fused\n")
        f.write("# F + R chain to the next one.  Per-launch times under the profiler are cold-cache and serialised: compare SHARES.\n")
        f.write("# %d launches, %.1f us in total.\n" % (len(step), tot))
        f.write("%10s %6s %6s %9s  %s\n" % ("total_us", "share", "count", "each_us", "kernel"))
        for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%10.1f %5.1f%% %6d %9.1f  %s\n" % (t, 100 * t / tot, c, t / c, k))
        f.write("\n# in launch order (GEMM-class kernels only)\n")
        for k, v, g in step:
            if any(s in k for s in ("gemm", "udf_chain", "composite", "blend")):
                f.write("%9.1f us  %-14s %s\n" % (v, g, re.sub(r"\(.*", "", k)[:100]))
    print(open(out).read()[:1800])


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_tensor.sum",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"]


def report(path, name, out, header):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on -k regex:<kernel>  python tools/step_probe.py  (%s)\n" % path)
        f.write(header)
        for n, r in enumerate(rows[2:]):
            f.write("\n== launch %d: %s  grid %s block %s\n" % (n, r[hdr.index("Kernel Name")][:90], r[hdr.index("Grid Size")], r[hdr.index("Block Size")]))
            for w in WANT:
                if w in hdr:
                    i = hdr.index(w)
                    f.write("  %-82s %16s %s\n" % (w, r[i], units[i]))
            st = []
            for i, h in enumerate(hdr):
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                    try:
                        st.append((float(r[i].replace(",", "")), h))
                    except ValueError:
                        pass
            f.write("  warp stall reasons (cycles stalled per issue-active cycle, top 8):\n")
            for v, h in sorted(st, reverse=True)[:8]:
                f.write("    %6.2f  %s\n" % (v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    print(open(out).read()[:2500])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r02")
    ap.add_argument("--launches", default=None)
    ap.add_argument("--report", action="append", default=[], help="path.ncu-rep:name")
    a = ap.parse_args()
    if a.launches:
        launches(a.launches, "profiles/%s_launches.txt" % a.tag)
    for spec in a.report:
        path, name = spec.rsplit(":", 1)
        hdr = {"chain": "# udf_chain_kernel: launch 0 = F + R (value chain + reverse sweep), launch 1 = T + B (tangent + backward chains) of a C2 step,\n"
                        "# 65 536 points = 512 tiles of 128 points on 148 persistent CTAs.  Algorithmic MACs: F+R 2 x 524 544 per point, T+B 2 x 524 544.\n",
               "wgrad": "# gemm_tn2_kernel<.., 0> (T128 operands): dW_l += D_l^T Adot_l + Zbar_l^T A_l (two operand pairs, 4 x [65 536 x 256] fp32 = 268 MB\n"
                        "# algorithmic read).  Captured before the last round of address hoisting (73 us per launch in the final launch list).\n",
               "wgrad_rm": "# gemm_tn2_kernel<.., 1> (row-major operands, colour network): dW += dZ^T X, [65 536 x 128] x [65 536 x 128] fp32 = 67 MB read\n"}.get(name, "")
        report(path, name, "profiles/%s_%s_kernel.txt" % (a.tag, name), hdr)
