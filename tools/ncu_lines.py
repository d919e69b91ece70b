#!/usr/bin/env python3
"""Summarise an .ncu-rep here (no GPU needed): headline metrics + executed instructions per source line.

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [units_per_launch] [top]

`units_per_launch` (entries or pixels the profiled launch processed) turns instruction counts into
thread-instructions per unit.
"""
import collections
import csv
import subprocess
import sys


def raw_metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


WANT = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "smsp__warps_eligible.avg.per_cycle_active",
]


def main():
    rep = sys.argv[1]
    units = float(sys.argv[2]) if len(sys.argv) > 2 else None
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    m = raw_metrics(rep)
    for k in WANT:
        if k in m:
            print(f"{k:75s} {m[k][0]:>18s} {m[k][1]}")
    for k, (v, u) in m.items():
        if "warp_issue_stalled" in k and k.endswith("per_warp_active.pct"):
            try:
                if float(v.replace(",", "")) >= 3.0:
                    print(f"{k:75s} {v:>18s} {u}")
            except ValueError:
                pass
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    per_line = collections.Counter()
    cur = None
    for r in csv.reader(out.splitlines()):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if len(r) < 8 or r[0] in ("Line No", "Function Name", ""):
            continue
        try:
            per_line[(cur, int(r[0]), r[1].strip()[:100])] += int(r[7])
        except ValueError:
            continue
    tot = sum(per_line.values())
    print(f"\nwarp instructions executed (source-attributed): {tot}")
    if units:
        print(f"thread instructions per unit: {tot * 32 / units:.1f}")
    for (f, l, s), v in per_line.most_common(top):
        per = f"{v * 32 / units:8.1f}/unit" if units else ""
        print(f"{v / tot * 100:5.1f}% {per}  {f}:{l}  {s}")


if __name__ == "__main__":
    main()
