#!/usr/bin/env python
"""Turns an ncu report (gpurun_out/*.ncu-rep, captured on the B200 box with
`ncu --set full --clock-control none --import-source on`) into the text summary committed
under profiles/:  python profiles/summarize.py gpurun_out/prof.ncu-rep > profiles/rNN_name.txt"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def ncu(rep, *args):
    out = subprocess.run(["ncu", "-i", rep, "--csv", *args], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main(rep, top=40):
    rows = ncu(rep, "--page", "raw")
    hdr, units = rows[0], rows[1]
    print(f"# {rep}")
    for r in rows[2:]:
        print(f"\n## kernel: {r[hdr.index('Kernel Name')]}")
        for k in KEYS:
            if k in hdr:
                print(f"{k:72s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
        stalls = []
        for i, h in enumerate(hdr):
            if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
                try:
                    stalls.append((float(r[i]), h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
                except ValueError:
                    pass
        tot = sum(v for v, _ in stalls) or 1.0
        print("warp stall samples: " + ", ".join(f"{h} {100 * v / tot:.0f}%" for v, h in sorted(stalls, reverse=True)[:7]))
    src = ncu(rep, "--page", "source", "--print-source", "cuda,sass")
    seg, agg, hdr2 = 0, {}, None
    first_kernel = None
    for r in src:
        if len(r) == 2 and r[0] == "Function Name":
            first_kernel = first_kernel or r[1]
            cur_kernel = r[1]
            continue
        if len(r) > 2 and r[0] == "Line No":
            hdr2, seg = r, seg + 1
            ie = hdr2.index("Instructions Executed")
            continue
        if hdr2 is None or len(r) <= 8 or not r[0].isdigit():
            continue
        key = (r[1].strip()[:110])
        try:
            agg[key] = agg.get(key, 0) + int(r[ie])
        except ValueError:
            pass
    tot = sum(agg.values()) or 1
    print(f"\n## warp-level instructions executed per source line (all captured launches, top {top})")
    for src_line, v in sorted(agg.items(), key=lambda kv: -kv[1])[:top]:
        print(f"{100 * v / tot:5.1f}%  {src_line}")


if __name__ == "__main__":
    main(sys.argv[1])
