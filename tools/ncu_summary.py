#!/usr/bin/env python
"""Per-kernel summary of an `ncu --set full` report (runs anywhere ncu is installed; no GPU needed):

  python tools/ncu_summary.py gpurun_out/x.ncu-rep [kernel-regex] > profiles/rNN_ncu_<what>.txt

Prints, per captured launch: duration, DRAM bytes, pipe utilisation (ALU / FMA / XU / LSU / tensor),
issue activity, occupancy, registers and the leading warp-stall reasons.
"""
import csv
import re
import subprocess
import sys

KEEP = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 % of peak"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/tex % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__issue_active.avg.pct_of_peak_sustained_elapsed", "issue active %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "pipe ALU %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "pipe FMA %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "pipe XU (popc) %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "pipe LSU %"),
    ("sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "pipe uniform %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active %"),
    ("sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active", "tensor IMMA subpipe %"),
    ("sm__inst_executed_pipe_tensor_subpipe_imma.avg.pct_of_peak_sustained_active", "tensor IMMA inst %"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "pipe TMEM %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem / block"),
    ("launch__shared_mem_per_block_static", "static smem / block"),
    ("smsp__inst_executed.sum", "warp instructions"),
]
STALL = re.compile(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active\.ratio")


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# {rep}: {len(rows) - 2} captured launches (ncu --set full --clock-control none)")
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        if pat and not pat.search(name):
            continue
        print(f"\n== {name[:140]}")
        for key, label in KEEP:
            if key in col and r[col[key]] not in ("", "n/a"):
                print(f"  {label:28s} {r[col[key]]} {units[col[key]]}")
        stalls = []
        for h, i in col.items():
            m = STALL.match(h)
            if m and m.group(1) != "selected":
                try:
                    stalls.append((float(r[i]), m.group(1)))
                except ValueError:
                    pass
        stalls.sort(reverse=True)
        print("  top stalls (warps per issue): " + ", ".join(f"{n} {v:.2f}" for v, n in stalls[:5]))


if __name__ == "__main__":
    main()
