#!/usr/bin/env python
"""Stall-reason breakdown and hottest SASS instructions of one kernel from an `ncu --set full --import-source on`
report (runs anywhere `ncu` is installed; no GPU needed):

  python tools/ncu_stalls.py gpurun_out/v16_top.ncu-rep describe_kernel [launch_index]

Reads `ncu -i REPORT --page source --csv --print-source sass` and prints, for the chosen launch of the kernel:
the warp-stall sampling totals by reason, and the instructions with the most samples together with their two
leading stall reasons.  A stall attributed to an instruction is the reason that instruction could not issue --
`stall_long_sb` on an arithmetic instruction means it waits for the global / local load that feeds it.
"""
import csv
import subprocess
import sys


def main():
    if len(sys.argv) < 3:
        print(__doc__)
        return 2
    report, kernel = sys.argv[1], sys.argv[2]
    which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    out = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}", "--print-source", "sass"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    if not starts:
        print("no such kernel in the report")
        return 1
    s = starts[min(which, len(starts) - 1)]
    e = starts[starts.index(s) + 1] if starts.index(s) + 1 < len(starts) else len(rows)
    hdr = rows[s + 1]
    si, src = hdr.index("# Samples"), hdr.index("Source")
    ex = hdr.index("Instructions Executed")
    stall_cols = [i for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c]
    data = [r for r in rows[s + 2:e] if len(r) >= len(hdr) and r[0].startswith("0x")]
    total = sum(int(r[si]) for r in data)
    print(f"# {rows[s][1][:110]}")
    print(f"# launch {which} of {len(starts)} in {report}: {len(data)} SASS instructions, {total} warp-stall samples, "
          f"{sum(int(r[ex]) for r in data)} warp instructions executed")
    agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stall_cols}
    tot = max(1, sum(agg.values()))
    print("stall reason (all samples)          samples   share")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:10]:
        print(f"  {k:32s} {v:9d}  {100 * v / tot:5.1f} %")
    print("hottest instructions: samples, share, instruction, leading stall reasons")
    for r in sorted(data, key=lambda r: -int(r[si]))[:20]:
        st = sorted(((hdr[i], int(r[i] or 0)) for i in stall_cols), key=lambda kv: -kv[1])[:2]
        print(f"  {int(r[si]):7d} {100 * int(r[si]) / max(1, total):5.1f} %  {r[src].strip()[:64]:64s} " + ", ".join(f"{k[6:]} {v}" for k, v in st if v))
    return 0


if __name__ == "__main__":
    sys.exit(main())
