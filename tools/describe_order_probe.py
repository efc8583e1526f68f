"""Experiment: does the ORDER in which a frame's key points are described matter (L1 / L2 locality of the integral-block
gathers)?  Describes the same key points in raster order (as detected), in 64x64-tile order and in random order."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
import ethzasl_brisk_b200 as bb  # noqa: E402


def main(n=256, cap=12288):
    cfg = bench.FRAME_CONFIGS["C3"]
    uniq = [torch.from_numpy(np.ascontiguousarray(f)).cuda() for f in bench.unique_frames(cfg)]
    frames = torch.stack([torch.roll(uniq[j % len(uniq)], shifts=(j // len(uniq)) * 5, dims=1) for j in range(n)])
    ctx = bb.Context(0, timing=True)
    det = bb.BriskFeatureDetector(cfg["thresh"], cfg["octaves"], ctx=ctx)
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    host_frames = frames.cpu().numpy()
    kps, counts = det.detect_batch(host_frames, cap=cap)
    rng = np.random.default_rng(0)
    for name in ("raster", "tile64", "tile128", "scale_then_tile", "random"):
        k2 = kps.copy()
        for f in range(n):
            c = int(counts[f]); k = kps[f, :c]
            tx, ty = (k["x"] // 64).astype(np.int64), (k["y"] // 64).astype(np.int64)
            if name == "raster": order = np.arange(c)
            elif name == "tile64": order = np.lexsort((k["x"], tx, ty))
            elif name == "tile128": order = np.lexsort((k["x"], (k["x"] // 128).astype(np.int64), (k["y"] // 128).astype(np.int64)))
            elif name == "scale_then_tile": order = np.lexsort((tx, ty, np.log2(np.maximum(k["size"], 1.0)).astype(np.int64)))
            else: order = rng.permutation(c)
            k2[f, :c] = k[order]
        import ctypes as C
        from ethzasl_brisk_b200.api import _ptr
        dk0 = torch.from_numpy(k2.view(np.float32).reshape(n, cap, 7)).cuda()
        dc0 = torch.from_numpy(counts.copy()).cuda()
        desc = torch.empty((n, cap, 48), dtype=torch.uint8, device="cuda")
        times = []
        for rep in range(3):
            kk, cc = dk0.clone(), dc0.clone()
            torch.cuda.synchronize()
            rc = ctx._lib.brisk_describe(ctx._h, ext._h, _ptr(frames), n, 1920, 1080, C.c_size_t(1920), C.c_size_t(1920 * 1080),
                                         _ptr(kk), _ptr(cc), int(cap), _ptr(desc))
            ctx._check(rc)
            times.append(ctx.last_timing()[0])
        kept = int(cc.sum().item())
        t = times[-1]
        print(f"{name:16s} describe {t.get('describe', 0):7.3f} ms  integral {t.get('integral', 0):6.3f} ms   ({n} frames, {int(counts.sum())} key points in, {kept} described)")


if __name__ == "__main__":
    main()
