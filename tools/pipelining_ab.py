#!/usr/bin/env python
"""A/B: resident 1080p detect+describe step time with the two-slot chunk pipeline on and off.
usage: pipelining_ab.py [frames] [workspace_gb]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import ethzasl_brisk_b200 as bb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
gb = int(sys.argv[2]) if len(sys.argv) > 2 else 32
uniq = torch.from_numpy(bb.synthetic_batch(8, 1920, 1080, 2000)).cuda()
frames = torch.stack([torch.roll(uniq[j % 8], (j // 8) * 5, 1) for j in range(n)])
ctx = bb.Context(0, workspace_limit=gb << 30)
det, ext = bb.BriskFeatureDetector(60, 4, ctx=ctx), bb.BriskDescriptorExtractor(ctx=ctx)
out = (torch.empty((n, 12288, 7), device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda"),
       torch.empty((n, 12288, 48), dtype=torch.uint8, device="cuda"))
for mode in (True, False, True, False):
    ctx.set_pipelining(mode)
    for _ in range(2):
        bb.detect_and_compute_batch(det, ext, frames, cap=12288, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        bb.detect_and_compute_batch(det, ext, frames, cap=12288, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"pipelining={mode}: {ms:.2f} ms per {n} frames = {n / ms * 1e3:.0f} frames/s")
