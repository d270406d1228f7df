#!/usr/bin/env python
"""Summarise ncu output brought back in gpurun_out/ into tracked files under profiles/.

    python tools/ncu_summary.py <tag>       e.g. r01b  ->  profiles/<tag>_launches.csv, profiles/<tag>_kernels.md

Launch list: the `--metrics gpu__time_duration.sum --clock-control none` pass (cold-cache, serialised: shares, not
absolutes).  Per-kernel detail: one `--set full` capture per kernel (<tag>_prof_*.ncu-rep), read with
`ncu -i ... --page raw --csv`.
"""
import csv
import glob
import json
import io
import os
import re
import subprocess
import sys
from collections import OrderedDict

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ncu_hot  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__inst_executed.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__cycles_elapsed.max",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
]


def short(name):
    m = re.search(r"(\w+)(<[^>]*>)?\(", name)
    return (m.group(1) + (m.group(2) or "")) if m else name


def launches(tag, out):
    src = os.path.join(ROOT, "gpurun_out", f"{tag}_launches.csv")
    if not os.path.exists(src):
        return None
    text = "".join(l for l in open(src) if l.startswith('"'))
    rows = list(csv.DictReader(io.StringIO(text)))
    agg = OrderedDict()
    for r in rows:
        k = short(r["Kernel Name"])
        a = agg.setdefault(k, [0, 0.0, r["Grid Size"], r["Block Size"]])
        a[0] += 1
        a[1] += float(r["Metric Value"])
    total = sum(a[1] for a in agg.values())
    with open(os.path.join(out, f"{tag}_launches.csv"), "w") as f:
        f.write("kernel,launches,total_ns,avg_ns,share_of_listed_time,grid,block\n")
        for k, a in agg.items():
            f.write(f"\"{k}\",{a[0]},{a[1]:.0f},{a[1] / a[0]:.0f},{a[1] / total:.4f},\"{a[2]}\",\"{a[3]}\"\n")
    return agg, total


def kernel_detail(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    if len(rows) < 3:
        return None
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    return d


def main():
    tag = sys.argv[1]
    # optional second argument: where to write (on the GPU box only gpurun_out/ travels back, so the summaries are written
    # there next to the captures, the multi-megabyte .ncu-rep files are deleted, and the summaries are copied to profiles/ here)
    out = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles")
    os.makedirs(out, exist_ok=True)
    md = [f"# ncu summary `{tag}`", "",
          "Source: `tools/gpu_round.sh` on one B200 (`ncu --clock-control none`); numbers printed under ncu are never bench values.", ""]
    traffic = {}
    la = launches(tag, out)
    if la:
        agg, total = la
        md += ["## Launch list (gpu__time_duration.sum; cold-cache, serialised -- compare shares)", "",
               "| kernel | launches | avg us | share |", "|---|---:|---:|---:|"]
        for k, a in agg.items():
            md.append(f"| `{k}` | {a[0]} | {a[1] / a[0] / 1e3:.1f} | {100 * a[1] / total:.1f}% |")
        md.append("")
        # the timed step of bench.py is rho + tau-correlation (one pair per batch) + one bin fold: the share inside THAT
        # is what bench.py's roofline.share_of_step claims (the list above also holds the secondary legs' kernels)
        step = {k: a for k, a in agg.items() if k.startswith(("rho_lattice", "isf_corr_mma_kernel<2, 1>", "bins_fold"))}
        rho = [a for k, a in step.items() if k.startswith("rho_lattice")]
        cor = [a for k, a in step.items() if k.startswith("isf_corr")]
        if rho and cor:
            r_us, c_us = rho[0][1] / rho[0][0] / 1e3, cor[0][1] / cor[0][0] / 1e3
            md += [f"Inside the timed step (one rho + one tau-correlation launch per batch): rho {r_us:.1f} us of {r_us + c_us:.1f} us = "
                   f"**{100 * r_us / (r_us + c_us):.1f} %** -- compare `roofline.share_of_step` of the bench line.", ""]
    for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"{tag}_prof_*.ncu-rep"))):
        d = kernel_detail(rep)
        if not d:
            continue
        name = d.get("Kernel Name", ("?", ""))[0]
        try:
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
            traffic[re.sub(r"<.*", "", short(name)) + "_dram_bytes_per_launch"] = float(rd[0]) * scale[rd[1]] + float(wr[0]) * scale[wr[1]]
        except (KeyError, ValueError):
            pass
        md += [f"## `{short(name)}`  ({os.path.basename(rep)}, --set full)", "", "| metric | value | unit |", "|---|---:|---|"]
        for k in KEYS:
            if k in d:
                md.append(f"| {k} | {d[k][0]} | {d[k][1]} |")
        md.append("")
        hot = ncu_hot.digest(rep)
        if hot:
            md += ["Hot spots (source page, SASS; `tools/ncu_hot.py`):", "", "```"] + hot + ["```", ""]
    if traffic:
        # kernels captured in this visit replace their entries; entries of kernels captured earlier stay
        path = os.path.join(out, "traffic.json")
        merged = {}
        seed = os.path.join(ROOT, "profiles", "traffic.json")
        if not os.path.exists(path) and os.path.exists(seed):
            try:
                merged = json.load(open(seed))
            except ValueError:
                merged = {}
        if os.path.exists(path):
            try:
                merged = json.load(open(path))
            except ValueError:
                merged = {}
        src = merged.get("sources", {}) if isinstance(merged.get("sources"), dict) else {}
        per = merged.get("configurations_per_launch", {}) if isinstance(merged.get("configurations_per_launch"), dict) else {}
        batch = int(os.environ.get("PIMCB_PROFILE_BATCH", "512"))      # configurations per launch of the profiled bench run
        for k, v in traffic.items():
            merged[k] = v
            src[k] = tag
            per[k] = batch
        merged["sources"] = src
        merged["configurations_per_launch"] = per
        merged["source"] = ("ncu --set full --clock-control none, one launch each; per-kernel visit tag in `sources`, C2 configurations in that "
                            "launch in `configurations_per_launch` (bench.py scales the figure to its own batch)")
        with open(path, "w") as f:
            json.dump(merged, f, indent=1)
    with open(os.path.join(out, f"{tag}_kernels.md"), "w") as f:
        f.write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main()
