#!/usr/bin/env python
"""DRAM traffic per frame of every stage from an `ncu --set full` capture of one bench step (no GPU needed):

  python tools/ncu_traffic.py gpurun_out/x.ncu-rep --frames 256 --source "<command that was profiled>" \
      > profiles/r02_dram_traffic.json

Sums dram__bytes_read.sum + dram__bytes_write.sum over the captured launches of each kernel (one step: every kernel of
the step must be captured exactly once per launch it makes in a step), maps kernels to the stages bench.py reports and
divides by the frames one launch processed.  bench.py reads the result for `roofline.traffic`.
"""
import argparse
import csv
import json
import re
import subprocess

STAGE_OF = [("pyramid_kernel", "pyramid"), ("agast_detect", "detect"), ("row_scan|corner_fill|corner_list|layer_start|tile_corner", "lists"),
            ("nms_|refine_kernel|compact_kernel|score_tile", "nms"), ("integral_", "integral"), ("describe_", "describe"), ("harris_", "harris"),
            ("hamming_|knn_", "knn")]


def to_bytes(v, unit):
    f = float(v)
    u = unit.lower()
    return f * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--frames", type=int, required=True, help="frames processed by one launch of the per-chunk kernels")
    ap.add_argument("--source", default="")
    a = ap.parse_args()
    out = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    kernels, stages = {}, {}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        short = re.sub(r"\(.*", "", name).replace("void ", "").replace("briskb200::", "").strip()
        try:
            b = to_bytes(r[col["dram__bytes_read.sum"]], units[col["dram__bytes_read.sum"]]) + \
                to_bytes(r[col["dram__bytes_write.sum"]], units[col["dram__bytes_write.sum"]])
            ms = float(r[col["gpu__time_duration.sum"]]) * {"ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(units[col["gpu__time_duration.sum"]], 1)
        except ValueError:
            continue
        k = kernels.setdefault(short, {"launches": 0, "dram_bytes": 0.0, "ms": 0.0})
        k["launches"] += 1; k["dram_bytes"] += b; k["ms"] += ms
        for pat, st in STAGE_OF:
            if re.search(pat, short):
                k["stage"] = st
                stages[st] = stages.get(st, 0.0) + b / a.frames
                break
    print(json.dumps({"source": a.source or a.report, "frames_per_launch": a.frames, "stages": stages, "kernels": kernels}, indent=1))


if __name__ == "__main__":
    main()
