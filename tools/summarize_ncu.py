#!/usr/bin/env python
"""Turn the scratch ncu outputs in gpurun_out/ into the small tracked summaries under profiles/.

  python tools/summarize_ncu.py <tag>      e.g. r01b

  gpurun_out/launches.csv            -> profiles/<tag>_launches.csv (per-kernel totals + share of the step)
  gpurun_out/prof_r02g_build.ncu-rep   -> profiles/<tag>_ncu_full.txt (selected --set full metrics per kernel)
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SCR = os.path.join(ROOT, "gpurun_out")
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def launches(tag):
    src = os.path.join(SCR, "launches.csv")
    if not os.path.exists(src):
        return
    rows = [r for r in csv.reader(open(src, errors="ignore")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot = collections.OrderedDict()
    for r in rows:
        if r is hdr or len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0].replace("void ", "").replace("xp::", "")
        t = float(r[iv].replace(",", ""))
        n, s = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, s + t)
    unit = next((r[hdr.index("Metric Unit")] for r in rows if r is not hdr and r[im] == "gpu__time_duration.sum"), "ns")
    allt = sum(s for _, s in tot.values())
    with open(os.path.join(OUT, f"{tag}_launches.csv"), "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none over `bench.py --steps 2 --warmup 3 --no-cpu`\n")
        f.write(f"# per-launch times are cold-cache and serialised: compare SHARES, not absolutes. unit={unit}\n")
        f.write("kernel,launches,total,mean,share\n")
        for k, (n, s) in sorted(tot.items(), key=lambda x: -x[1][1]):
            f.write(f"{k},{n},{s:.1f},{s / n:.1f},{s / allt:.4f}\n")
    print("wrote", f"{tag}_launches.csv")


def full(tag, rep="prof_r02g_build.ncu-rep"):
    src = os.path.join(SCR, rep)
    if not os.path.exists(src):
        return
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(OUT, f"{tag}_ncu_full.txt"), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on, selected metrics per captured launch\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write("\n== " + d["Kernel Name"].split("(")[0] + "\n")
            for k in KEYS:
                if k in d:
                    f.write(f"{k:95s} {d[k]:>18s} {units[hdr.index(k)]}\n")
            tr = float(d.get("dram__bytes_read.sum", "0").replace(",", "")) + float(d.get("dram__bytes_write.sum", "0").replace(",", ""))
            f.write(f"{'traffic = dram read + write':95s} {tr:18.3f} {units[hdr.index('dram__bytes_read.sum')]}\n")
    print("wrote", f"{tag}_ncu_full.txt")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    full(tag, sys.argv[2] if len(sys.argv) > 2 else "prof_r02g_build.ncu-rep")
