#!/usr/bin/env python
"""Summaries of ncu output for profiles/:
   summarize.py launches <launch.csv> <out.txt> <title>     kernel | launches | total ms | share, plus the share of
                                                            gmr_retarget_kernel inside the timed steps
   summarize.py raw <x.ncu-rep> <out.txt> <title>            selected raw metrics of the first kernel in a --set full capture"""
import csv, io, subprocess, sys, collections

def launches(path, out, title):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if not l.startswith("=="))]
    hdr = rows[0]; c = {h: i for i, h in enumerate(hdr)}
    tot = collections.OrderedDict(); seq = []
    for r in rows[1:]:
        if len(r) < len(hdr) or r[c["Metric Name"]] != "gpu__time_duration.sum": continue
        v = float(r[c["Metric Value"]].replace(",", "")); u = r[c["Metric Unit"]]
        ms = v * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[u]
        k = r[c["Kernel Name"]]
        tot.setdefault(k, [0, 0.0]); tot[k][0] += 1; tot[k][1] += ms; seq.append((k, ms))
    allms = sum(v[1] for v in tot.values())
    with open(out, "w") as f:
        f.write(title + "\n")
        f.write("kernel | launches | total ms | share of all GPU time in the process\n")
        for k, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:12]:
            f.write(f"{k[:90]} | {n} | {ms:.2f} | {100 * ms / allms:.1f}%\n")
        g = [i for i, (k, _) in enumerate(seq) if "gmr_retarget_kernel" in k]
        if g:
            # the timed steps: from the first to the last solve launch, everything in between
            inside = seq[g[0]: g[-1] + 1]
            gm = sum(ms for k, ms in inside if "gmr_retarget_kernel" in k); al = sum(ms for _, ms in inside)
            f.write(f"\nbetween the first and the last solve launch: {len(inside)} launches, {al:.2f} ms, "
                    f"gmr_retarget_kernel = {len(g)} launches, {gm:.2f} ms = {100 * gm / al:.2f}% of the GPU time\n")
            f.write("per-launch ms of gmr_retarget_kernel (cold cache, serialised by ncu): " + " ".join(f"{seq[i][1]:.1f}" for i in g) + "\n")

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed"]

def raw(rep, out, title):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    with open(out, "w") as f:
        f.write(title + "\n\n")
        for h, u, v in zip(hdr, units, vals):
            if h == "Kernel Name" or h in KEEP or "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                f.write(f"{h} [{u}] = {v}\n")

if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](*sys.argv[2:5])
