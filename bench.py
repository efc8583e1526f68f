#!/usr/bin/env python
"""Benchmark of the BRISK hot path on B200 (contract: see the task description).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[2], "C3"): AGAST/OAST9-16 scale-space detect
(BriskFeatureDetector(60, 4)) + BRISK2 describe on synthetic 1920x1080 frames.
One step = one batch of --frames frames per GPU (default 1024, sharded by
frame across ranks without any collective: weak scaling).  `value` is frames/s
with the batch resident in HBM; `e2e` is the same call with pinned HOST
buffers in and out (H2D + D2H inside the timed region).  The reference arm
(--impl reference) times the unmodified reference (oracle/_ref) on the host
cores on a bounded sample of the same frames.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "1080p_frames_per_s_detect_describe"
UNIT = "frames/s"
W, H = 1920, 1080
THRESH, OCTAVES = 60, 4
UNIQUE = 16  # distinct procedurally generated frames; the batch cycles them with per-frame shifts


def unique_frames(seed0, n=UNIQUE):
    from ethzasl_brisk_b200.synthetic import synthetic_frame
    return np.stack([synthetic_frame(W, H, seed0 + i) for i in range(n)])


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.keep, self.on = [], False

    def mark(self):
        """Start of a timed region: lines read from now on count."""
        self.on = True

    def pause(self):
        self.on = False

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            row = [c.strip() for c in line.split(",")]
            self.rows.append(row)
            if self.on:
                self.keep.append(row)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in (self.keep or self.rows[-1:]):
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pin_to_gpu_numa_node(index):
    """One process per GPU: run on the CPUs NVML reports as local to the GPU, so that the pinned host buffers
    of the end-to-end path are allocated on the GPU's NUMA node (first touch).  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, wd in enumerate(words) for b in range(64) if (wd >> b) & 1}
        full = os.sched_getaffinity(0)
        cpus &= full
        if cpus:
            os.sched_setaffinity(0, cpus)
            pin_to_gpu_numa_node.full = full
            return len(cpus)
    except Exception:
        pass
    return None


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return json.loads(p.read_text()).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def stage_bytes(n_layers_dims, n_corners, n_kps, desc_bytes=48):
    """Algorithmic bytes per frame of each stage."""
    px = [w * h for w, h in n_layers_dims]
    return {
        "pyramid": px[0] + sum(px),                 # read L0 once, write every layer (L0 copy included)
        "detect": sum(px) + 2 * sum(px),            # read u8 layers, write u16 corner maps
        "lists": 2 * sum(px) + 4 * n_corners,       # read corner maps, write packed corners
        "nms": 3 * sum(px) + 100 * n_corners,       # touch-map clear + corner-map traffic + per-corner records
        "integral": px[0] + 4 * (W + 1) * (H + 1),  # u8 in, i32 out.  Real traffic is 2 x px[0] in + 16 B per pixel
                                                    # out (one 2x2 block of the integral image per pixel, for the sampler)
        "describe": n_kps * (28 * 2 + desc_bytes),  # compulsory HBM only; the gathers hit L2
    }


def layer_dims(w, h, octaves):
    dims = [(w, h), (2 * (w // 3), 2 * (h // 3))]
    for i in range(2, 2 * octaves):
        dims.append((dims[i - 2][0] // 2, dims[i - 2][1] // 2))
    return dims[:max(1, 2 * octaves)]


def run_reference(args):
    """CPU arm: the unmodified reference (oracle/_ref), frame-parallel over the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import ref
    cores = host_cores()
    frames = unique_frames(2000)
    per_step = min(256, max(UNIQUE, 2 * cores))
    frames = frames[np.arange(per_step) % UNIQUE]
    sample = f"{per_step} synthetic 1080p frames per step ({UNIQUE} distinct), {cores} threads (one detector per thread)"
    for _ in range(args.warmup):
        ref.bench_detect_describe(frames[:cores], False, THRESH, OCTAVES, nthreads=cores)
    t = kp = 0
    for _ in range(args.steps):
        s, k = ref.bench_detect_describe(frames, False, THRESH, OCTAVES, nthreads=cores)
        t += s; kp += k
    value = args.steps * len(frames) / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "C3: AGAST(60,4) detect + BRISK2 describe, 1920x1080 synthetic frames", "frames_per_step": len(frames),
                       "keypoints_per_frame": kp / (args.steps * len(frames))},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def run_gpu(args):
    import torch
    import torch.distributed as dist

    import ethzasl_brisk_b200 as bb

    # keep stdout to the one JSON line: NCCL prints its version banner to stdout from level WARN upwards
    if "BENCH_NCCL_DEBUG" in os.environ:
        os.environ["NCCL_DEBUG"] = os.environ["BENCH_NCCL_DEBUG"]
    else:
        os.environ.pop("NCCL_DEBUG", None)
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    numa = pin_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    n = args.frames
    cap = args.cap

    # frames: UNIQUE generated frames per rank, cycled with a per-copy horizontal roll so that all n differ
    uniq = torch.from_numpy(unique_frames(2000 + rank * UNIQUE)).to(dev)
    d_frames = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
    for j in range(n):
        d_frames[j] = torch.roll(uniq[j % UNIQUE], shifts=(j // UNIQUE) * 5, dims=1)
    h_frames = torch.empty((n, H, W), dtype=torch.uint8).pin_memory()
    h_frames.copy_(d_frames)
    torch.cuda.synchronize()

    stream = torch.cuda.current_stream()
    ctx = bb.Context(local, stream=stream.cuda_stream, timing=True, workspace_limit=args.workspace_gb << 30)
    det = bb.BriskFeatureDetector(THRESH, OCTAVES, ctx=ctx)
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    d_out = (torch.empty((n, cap, 7), dtype=torch.float32, device=dev), torch.empty(n, dtype=torch.int32, device=dev),
             torch.empty((n, cap, 48), dtype=torch.uint8, device=dev))
    h_out = (torch.empty((n, cap, 7), dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory(),
             torch.empty((n, cap, 48), dtype=torch.uint8).pin_memory())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        stages, launches = {}, 0
        for _ in range(steps):
            fn()
            ms, l = ctx.last_timing()
            launches += l
            for k, v in ms.items():
                stages[k] = stages.get(k, 0.0) + v
        e1.record(stream)
        barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), stages, launches

    resident = lambda: bb.detect_and_compute_batch(det, ext, d_frames, cap=cap, out=d_out)
    e2e = lambda: bb.detect_and_compute_batch(det, ext, h_frames, cap=cap, out=h_out)

    # nvidia-smi takes a few hundred ms to deliver its first line: start it before the warm-up, count only the
    # lines of the two timed regions (resident and host end-to-end)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        resident()
    if rank == 0:
        sampler.mark()
    ms_res, stages, launches = timed(resident, args.steps)
    if rank == 0:
        sampler.pause()
    counts = d_out[1].cpu().numpy()
    # per-stage device times without cross-stream overlap (same kernels, one stream, stages back to
    # back): these feed the per-kernel roofline numbers; the headline numbers above keep pipelining on
    ctx.set_pipelining(False)
    resident()
    _, stages_serial, _ = timed(resident, 1)
    ctx.set_pipelining(True)
    for _ in range(max(1, args.warmup // 2)):
        e2e()
    if rank == 0:
        sampler.mark()
    ms_e2e, _, _ = timed(e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    if getattr(pin_to_gpu_numa_node, "full", None):
        os.sched_setaffinity(0, pin_to_gpu_numa_node.full)  # the CPU baseline below uses every host core
    hc = h_out[1].numpy()
    assert np.array_equal(hc, counts), "host and device paths disagree"
    kp_total = int(np.minimum(counts, cap).sum())

    # secondary metric: brute-force Hamming kNN (k=2), 512-bit descriptors, reduced C5 shape per GPU
    q = torch.from_numpy(bb.random_descriptors(args.knn_q, 64, 5)).to(dev)
    t = torch.from_numpy(bb.random_descriptors(args.knn_t, 64, 6 + rank)).to(dev)
    m = bb.BruteForceMatcher(ctx=ctx)
    knn_variants = {}
    for name, variant in (("popc", 0), ("tensor_core", 1)):
        ctx.set_knn_variant(variant)
        m.knn(q, t, 2)
        v_ms, _, _ = timed(lambda: m.knn(q, t, 2), 3)
        knn_variants[name] = {"Gcmp/s": world * args.knn_q * args.knn_t * 3 / (v_ms * 1e-3) / 1e9, "ms": v_ms / 3}
    knn_ms = 3 * knn_variants["tensor_core"]["ms"]   # the default path
    gcmp = knn_variants["tensor_core"]["Gcmp/s"]

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    value = world * n * args.steps / (ms_res * 1e-3)
    e2e_value = world * n * args.steps / (ms_e2e * 1e-3)
    h2d = n * H * W
    d2h = 4 * n + kp_total * (28 + 48)
    # roofline of the dominant stage (device time from CUDA events on the launching stream)
    peak, peak_src = measured_peaks()
    dims = layer_dims(W, H, OCTAVES)
    compute_stages = {k: v for k, v in stages_serial.items() if k in ("pyramid", "detect", "lists", "nms", "integral", "describe")}
    total_ms = sum(compute_stages.values())
    top = max(compute_stages, key=compute_stages.get)
    # raw corners per frame are not returned by the API; the measured mean on these frames is ~2.1x the key points
    kps_per_frame = kp_total / n
    bytes_per_frame = stage_bytes(dims, int(2.1 * kps_per_frame), int(kps_per_frame))
    stage_report = {}
    for k, v in compute_stages.items():
        gbs = bytes_per_frame[k] * n / (v * 1e-3) / 1e9 if v > 0 else 0.0
        stage_report[k] = {"ms_per_step": v, "share": v / total_ms if total_ms else 0.0, "algorithmic_GBps": gbs,
                           "frac_of_hbm_peak": gbs / peak, "ms_per_step_pipelined": stages.get(k, 0.0) / args.steps}
    roof = {"bound": "hbm", "kernel": top, "achieved": stage_report[top]["algorithmic_GBps"], "peak": peak, "unit": "GB/s",
            "frac": stage_report[top]["frac_of_hbm_peak"], "traffic": None, "peak_source": peak_src,
            "note": "dominant stage by CUDA-event time of a non-pipelined step (stages back to back on one stream); see stages"}
    # DRAM bytes per frame of the stages' kernels from the committed `ncu --set full` captures (dram__bytes_read.sum +
    # dram__bytes_write.sum of one launch / frames of that launch); reported for one frame of the dominant stage,
    # like `achieved`, which is per frame too (algorithmic bytes of n frames / time of n frames)
    NCU_DRAM_BYTES_PER_FRAME = {"describe": (2.195719e9 + 87.774e6) / 256,        # profiles/r01_ncu_full_top_kernels_v14.txt
                                "pyramid": (0.354613e9 + 0.626522e9) / 256,       # (inputs partly L2 resident in that capture)
                                # nms_prefix + nms_checks (v14 capture, 256 frames per launch) + nms_chain (v15 capture, 171 frames
                                # per launch); refine / compact (a few per cent of the stage's time) were not captured
                                "nms": (1.080348e9 + 0.191219e9 + 1.312184e9 + 0.260362e9) / 256 + (961.691392e6 + 79.735040e6) / 171}
    LIMITER = {"describe": "L1/L2 sector rate of scattered 4-byte gathers (ncu: l1tex 77 %, lts 55 % of peak, DRAM 30 %); the "
                           "integral images of a chunk do not fit L2, so ~8.9 MB per frame come from DRAM although only "
                           "104 B per key point are compulsory",
               "nms": "ALU pipe / instruction issue of the per-corner kernels and the tie chain (ncu: ALU 73-79 %)",
               "detect": "ALU pipe (packed min/max at half rate; ncu: ALU 73 %, DRAM 10 %)"}
    roof["algorithmic_bytes"] = bytes_per_frame[top] * n
    if "integral" in stage_report:  # the block layout trades 3.6x the output bytes for 3.5x fewer gather instructions in describe
        real = 2 * W * H + 16 * W * H
        stage_report["integral"]["hbm_traffic_GBps"] = real * n / (compute_stages["integral"] * 1e-3) / 1e9
        stage_report["integral"]["hbm_traffic_frac_of_peak"] = stage_report["integral"]["hbm_traffic_GBps"] / peak
    if stage_report.get("describe", {}).get("ms_per_step", 0) > 0:
        # describe is gather bound, not HBM bound -- its meaningful rate is key points per second (132 box-filter
        # samples each); ncu of the round: L1 71 %, L2 51 % of peak, 46 % of the stall samples on dependent global loads
        stage_report["describe"]["keypoints_per_s"] = kp_total / (stage_report["describe"]["ms_per_step"] * 1e-3)
    if top in NCU_DRAM_BYTES_PER_FRAME:
        roof["traffic"] = NCU_DRAM_BYTES_PER_FRAME[top] * n
        roof["traffic_note"] = ("ncu dram bytes per frame x frames of the step (captures: profiles/r01_ncu_full_top_kernels_v14.txt, "
                                "profiles/r01_ncu_full_top_kernels_v15.txt)")
    if top in LIMITER:
        roof["limiter"] = LIMITER[top]

    # CPU baseline beside it: the unmodified reference on a bounded sample of the same frames
    cpu = None
    try:
        from oracle import ref
        if ref.available():
            cores = host_cores()
            sample_frames = uniq.cpu().numpy()[np.arange(min(256, max(UNIQUE, 2 * cores))) % UNIQUE]
            ref.bench_detect_describe(sample_frames[:2], False, THRESH, OCTAVES, nthreads=min(2, cores))
            s, _ = ref.bench_detect_describe(sample_frames, False, THRESH, OCTAVES, nthreads=cores)
            cpu = {"value": len(sample_frames) / s, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"{len(sample_frames)} frames drawn from the same {UNIQUE} distinct 1080p frames, unmodified reference (oracle/_ref), {cores} threads"}
    except Exception as e:  # the baseline is informative; never fail the bench on it
        cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
    # ... and of the matcher: the reference's brisk::Hamming primitive in its k-successive-arg-min loop (brute-force-matcher.cc:
    # 80-162), one std::thread per core over the queries, on a bounded slice of the same descriptors
    knn_cpu = None
    try:
        from oracle import ref
        if ref.available() and world == 1:
            cores = host_cores()
            cq, ct = q[:min(4096, args.knn_q)].cpu().numpy(), t[:min(1000000, args.knn_t)].cpu().numpy()
            t0 = time.perf_counter()
            ref.knn(cq, ct, 2, nthreads=cores)
            knn_cpu = {"value": len(cq) * len(ct) / (time.perf_counter() - t0) / 1e9, "unit": "Gcmp/s", "cores": cores, "kind": "reference",
                       "sample": f"{len(cq)} queries x {len(ct)} train rows of the same descriptors, k = 2, {cores} threads"}
    except Exception as e:
        knn_cpu = {"value": None, "unit": "Gcmp/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": "C3: AGAST(60,4) detect + BRISK2 describe, 1920x1080 synthetic frames", "frames_per_gpu_per_step": n,
                       "global_frames_per_step": world * n, "keypoints_per_frame": kps_per_frame, "parallelism": f"frame-sharded x{world}, no collective", "cpus_per_rank": numa,
                       "l2": f"inputs ({n * H * W / 1e6:.0f} MB per step) exceed the 126 MB L2; no explicit flush", "kp_capacity": cap},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "roofline": roof, "stages": stage_report, "cpu_baseline": cpu, "clocks": clocks,
            "secondary": {"metric": "hamming_knn_k2_512bit", "value": gcmp, "unit": "Gcmp/s", "queries": args.knn_q, "train": args.knn_t,
                          "ms": knn_ms / 3, "variants": knn_variants, "cpu_baseline": knn_cpu,
                          "int8_TOPS": gcmp * 1024 / 1e3,
                          "note": "default = s8 x u8 IMMA (mma.sync m16n8k32) on byte-expanded bits, 512 MACs per comparison; "
                                  "tensor-pipe activity from ncu is in profiles/"}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames", type=int, default=1024, help="frames per GPU per step")
    ap.add_argument("--cap", type=int, default=12288, help="key-point capacity per frame")
    ap.add_argument("--workspace-gb", type=int, default=32)
    ap.add_argument("--knn-q", type=int, default=100000)
    ap.add_argument("--knn-t", type=int, default=1000000)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
