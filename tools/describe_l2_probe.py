"""Experiment: how much faster is the descriptor kernel when the integral blocks of its frames are still in L2?
brisk_describe on n device-resident frames builds the block images and describes right after: for n = 2 or 3 the
blocks (33 MB per 1080p frame) fit the 126 MB L2, for n = 64 they do not."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
import ethzasl_brisk_b200 as bb  # noqa: E402
from ethzasl_brisk_b200.api import _ptr  # noqa: E402


def main(cap=12288):
    cfg = bench.FRAME_CONFIGS["C3"]
    uniq = [torch.from_numpy(np.ascontiguousarray(f)).cuda() for f in bench.unique_frames(cfg)]
    ctx = bb.Context(0, timing=True)
    det = bb.BriskFeatureDetector(cfg["thresh"], cfg["octaves"], ctx=ctx)
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    for n in (1, 2, 3, 4, 8, 64):
        frames = torch.stack([torch.roll(uniq[j % len(uniq)], shifts=(j // len(uniq)) * 5, dims=1) for j in range(n)])
        kps, counts = det.detect_batch(frames.cpu().numpy(), cap=cap)
        dk0 = torch.from_numpy(kps.view(np.float32).reshape(n, cap, 7)).cuda()
        dc0 = torch.from_numpy(counts.copy()).cuda()
        desc = torch.empty((n, cap, 48), dtype=torch.uint8, device="cuda")
        best = None
        for rep in range(5):
            kk, cc = dk0.clone(), dc0.clone()
            torch.cuda.synchronize()
            ctx._check(ctx._lib.brisk_describe(ctx._h, ext._h, _ptr(frames), n, 1920, 1080, C.c_size_t(1920), C.c_size_t(1920 * 1080),
                                               _ptr(kk), _ptr(cc), int(cap), _ptr(desc)))
            t = ctx.last_timing()[0]
            if best is None or t["describe"] < best["describe"]:
                best = t
        print(f"n = {n:3d}: describe {1e3 * best['describe'] / n:7.1f} us per frame, integral {1e3 * best['integral'] / n:7.1f} us per frame "
              f"({int(cc.sum().item()) / n:.0f} key points per frame)")


if __name__ == "__main__":
    main()
