#!/usr/bin/env python
"""Throughput of the other BASELINE.json configurations (bench.py itself measures C3).

  python tools/bench_configs.py --config C2      # 256 x 752x480, Harris scale space + BRISK2, 1 GPU
  python tools/bench_configs.py --config C4      # 3840x2160, 6 octaves, describe stress
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port 29531 \
      tools/bench_configs.py --config C5          # kNN k=2, Q x T 512-bit, train set sharded over N ranks

Each run prints one JSON line (device time from CUDA events, max over ranks).  Frames are synthetic
(ethzasl_brisk_b200.synthetic); inputs are resident in HBM for `value` and in pinned host memory for `e2e`.
"""
import argparse
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ethzasl_brisk_b200 as bb  # noqa: E402
from ethzasl_brisk_b200.distributed import shard_range, sharded_knn  # noqa: E402
from ethzasl_brisk_b200.synthetic import synthetic_frame  # noqa: E402


def timed(fn, steps, stream, world, dev):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        fn()
    e1.record(stream)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def frames_on_device(w, h, n, seed0, unique, dev, **kw):
    uniq = torch.from_numpy(np.stack([synthetic_frame(w, h, seed0 + i, **kw) for i in range(unique)])).to(dev)
    d = torch.empty((n, h, w), dtype=torch.uint8, device=dev)
    for j in range(n):
        d[j] = torch.roll(uniq[j % unique], shifts=(j // unique) * 5, dims=1)
    return d


def run_frames(args, ctx, dev, stream, world, rank):
    if args.config == "C2":
        w, h, n, cap, nbytes = 752, 480, args.frames or 256, 4096, 48
        det = bb.ScaleSpaceFeatureDetector(4, 30.0, 20.0, ctx=ctx)
        name = "C2: Harris scale space (octaves=4, radius=30, absThr=20) + BRISK2, 752x480 synthetic frames"
        d_frames = frames_on_device(w, h, n, 1000 + 64 * rank, 32, dev)
    else:
        w, h, n, cap, nbytes = 3840, 2160, args.frames or 128, 49152, 48
        det = bb.BriskFeatureDetector(60, 6, ctx=ctx)
        name = "C4: AGAST(60, 6 octaves) + BRISK2, 3840x2160 synthetic frames (describe stress)"
        d_frames = frames_on_device(w, h, n, 3000 + 8 * rank, 4, dev, n_shapes=args.c4_shapes)
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    h_frames = torch.empty((n, h, w), dtype=torch.uint8).pin_memory()
    h_frames.copy_(d_frames)
    d_out = (torch.empty((n, cap, 7), dtype=torch.float32, device=dev), torch.empty(n, dtype=torch.int32, device=dev),
             torch.empty((n, cap, nbytes), dtype=torch.uint8, device=dev))
    h_out = (torch.empty((n, cap, 7), dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory(),
             torch.empty((n, cap, nbytes), dtype=torch.uint8).pin_memory())
    resident = lambda: bb.detect_and_compute_batch(det, ext, d_frames, cap=cap, out=d_out)
    e2e = lambda: bb.detect_and_compute_batch(det, ext, h_frames, cap=cap, out=h_out)
    for _ in range(args.warmup):
        resident()
    ms = timed(resident, args.steps, stream, world, dev)
    e2e()
    ms_e2e = timed(e2e, args.steps, stream, world, dev)
    counts = d_out[1].cpu().numpy()
    ctx.set_pipelining(False)
    resident()
    resident()
    stage_ms, launches = ctx.last_timing()
    ctx.set_pipelining(True)
    kp = float(counts.mean())
    return {"metric": f"{w}x{h}_frames_per_s_detect_describe", "value": world * n * args.steps / (ms * 1e-3), "unit": "frames/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "config": {"workload": name, "frames_per_gpu_per_step": n, "keypoints_per_frame": kp, "kp_capacity": cap,
                       "max_keypoints": int(counts.max())},
            "keypoints_per_s": world * kp * n * args.steps / (ms * 1e-3),
            "e2e": {"value": world * n * args.steps / (ms_e2e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": n * h * w,
                    "d2h_bytes_per_step": int(4 * n + counts.sum() * (28 + nbytes))},
            "stages_ms_serial": stage_ms, "gpu_launches_per_step": launches, "data": "synthetic", "dtype": "u8"}


def run_knn(args, ctx, dev, stream, world, rank):
    nq, nt, nbytes = args.knn_q, args.knn_t, args.knn_bytes
    begin, end = shard_range(nt, rank, world)
    q = torch.from_numpy(bb.random_descriptors(nq, nbytes, 5)).to(dev)
    # every rank draws its shard from its own seed; the global train set is their concatenation
    t = torch.from_numpy(bb.random_descriptors(end - begin, nbytes, 6 + rank)).to(dev)
    m = bb.BruteForceMatcher(ctx=ctx)
    out = {}
    for name, variant in (("tensor_core", 1), ("popc", 0)):
        if variant == 0 and not args.knn_popc:
            continue
        ctx.set_knn_variant(variant)
        if world > 1:
            fn = lambda: sharded_knn(m, q, t, 2, begin)
        else:
            fn = lambda: m.knn(q, t, 2)
        fn()
        ms = timed(fn, args.steps, stream, world, dev)
        out[name] = {"ms_per_step": ms / args.steps, "Gcmp/s": nq * nt * args.steps / (ms * 1e-3) / 1e9}
    res = fn()
    chk = int(res[0][:, 0].to(torch.int64).sum().item())
    best = out["tensor_core"]
    return {"metric": f"hamming_knn_k2_{8 * nbytes}bit", "value": best["Gcmp/s"], "unit": "Gcmp/s", "n_gpus": world, "steps": args.steps,
            "ms_per_step": best["ms_per_step"], "variants": out,
            "config": {"workload": f"C5: brute-force Hamming kNN k=2, {nq} queries x {nt} train rows, {8 * nbytes}-bit, "
                                   f"train set sharded over {world} rank(s)" + (", NCCL all-gather + top-k merge" if world > 1 else ""),
                       "queries": nq, "train": nt},
            "int8_TOPS": best["Gcmp/s"] * 2 * 8 * nbytes / 1e3, "index_checksum": chk, "data": "synthetic uniform random bits", "dtype": "u8"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", choices=["C2", "C4", "C5"], required=True)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=0)
    ap.add_argument("--c4-shapes", type=int, default=4200)
    ap.add_argument("--knn-q", type=int, default=1000000)
    ap.add_argument("--knn-t", type=int, default=10000000)
    ap.add_argument("--knn-bytes", type=int, default=64)
    ap.add_argument("--knn-popc", action="store_true")
    ap.add_argument("--workspace-gb", type=int, default=32)
    args = ap.parse_args()
    os.environ["NCCL_DEBUG"] = os.environ.get("BENCH_NCCL_DEBUG", "WARN")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.current_stream()
    ctx = bb.Context(local, stream=stream.cuda_stream, timing=True, workspace_limit=args.workspace_gb << 30)
    line = run_knn(args, ctx, dev, stream, world, rank) if args.config == "C5" else run_frames(args, ctx, dev, stream, world, rank)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
