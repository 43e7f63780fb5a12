#!/usr/bin/env python
"""Per-source-line instruction and stall-sample totals of one kernel from an ncu report.

  python tools/ncu_lines.py gpurun_out/prof.ncu-rep k_broadphase_grid [top]

Uses `ncu --page source --print-source cuda,sass --csv` (needs -lineinfo and --import-source on).
"""
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name",
                          "regex:" + kern, "--launch-skip", "0", "--launch-count", "1"], capture_output=True, text=True).stdout
    cur_file, hdr, out = None, None, []
    for r in csv.reader(raw.splitlines()):
        if len(r) == 2 and r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
        elif len(r) > 5 and r[0] == "Line No":
            hdr = r
            ii, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        elif hdr and len(r) > isamp and r[0].isdigit() and r[2] == "-":
            out.append((cur_file, int(r[0]), r[1], int(r[ii] or 0), int(r[it] or 0), int(r[isamp] or 0)))
    tot_i = sum(o[3] for o in out) or 1
    tot_s = sum(o[5] for o in out) or 1
    print(f"kernel {kern}: {tot_i} warp instructions, {sum(o[4] for o in out)} thread instructions, {tot_s} samples")
    print(f"{'file:line':28s} {'warp inst':>11s} {'%':>5s} {'lanes':>5s} {'smp%':>5s}  source")
    for o in sorted(out, key=lambda o: -o[3])[:top]:
        print(f"{o[0] + ':' + str(o[1]):28s} {o[3]:11d} {100.0 * o[3] / tot_i:5.1f} {o[4] / max(o[3], 1):5.1f} {100.0 * o[5] / tot_s:5.1f}  {o[2].strip()[:110]}")


if __name__ == "__main__":
    main()
