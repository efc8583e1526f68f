#!/usr/bin/env python
"""One resident detect+describe step on N synthetic 1080p frames (for `ncu`): warm-up calls first,
then exactly one more call.  usage: profile_step.py [frames] [warmup_calls]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import ethzasl_brisk_b200 as bb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 2
uniq = torch.from_numpy(bb.synthetic_batch(8, 1920, 1080, 2000)).cuda()
frames = torch.stack([torch.roll(uniq[j % 8], (j // 8) * 5, 1) for j in range(n)])
ctx = bb.Context(0)
ctx.set_pipelining(False)
det, ext = bb.BriskFeatureDetector(60, 4, ctx=ctx), bb.BriskDescriptorExtractor(ctx=ctx)
out = (torch.empty((n, 12288, 7), device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda"),
       torch.empty((n, 12288, 48), dtype=torch.uint8, device="cuda"))
for _ in range(warm + 1):
    bb.detect_and_compute_batch(det, ext, frames, cap=12288, out=out)
torch.cuda.synchronize()
print("kps/frame", float(out[1].float().mean()))
