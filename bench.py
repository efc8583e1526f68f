#!/usr/bin/env python
"""Benchmark of the BRISK hot path on B200 (contract: see the task description).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C3|C2|C4|C5]

Default workload (BASELINE.json configs[2], "C3"): AGAST/OAST9-16 scale-space detect
(BriskFeatureDetector(60, 4)) + BRISK2 describe on synthetic 1920x1080 frames.
One step = one batch of --frames frames per GPU (default 1024, sharded by
frame across ranks without any collective: weak scaling).  `value` is frames/s
with the batch resident in HBM; `e2e` is the same call with pinned HOST
buffers in and out (H2D + D2H inside the timed region).  The reference arm
(--impl reference) times the unmodified reference (oracle/_ref) on the host
cores on a bounded sample of the same frames.

The other BASELINE.json configurations print the same schema with --config:
  C2  256 x 752x480, Harris scale space (octaves 4, radius 30, absThr 20) + BRISK2
  C4  3840x2160, AGAST(60, 6 octaves) + BRISK2 (descriptor-extraction stress)
  C5  brute-force Hamming kNN (k = 2), 512-bit rows, train set sharded over the ranks (NCCL all-gather + merge)
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

# BASELINE.json configs[1], [2], [3]: frame size, detector, frames per GPU per step, key-point capacity,
# distinct generated frames (the batch cycles them with per-copy horizontal shifts), generator arguments
FRAME_CONFIGS = {
    "C2": dict(w=752, h=480, frames=256, cap=4096, unique=32, seed=1000, gen={}, harris=True, octaves=4, radius=30.0, abs_thr=20.0,
               metric="752x480_frames_per_s_harris_detect_describe",
               workload="C2: HarrisScaleSpaceFeatureDetector(octaves=4, uniformityRadius=30, absThr=20) + BRISK2 describe, 752x480 synthetic frames"),
    "C3": dict(w=1920, h=1080, frames=1024, cap=12288, unique=16, seed=2000, gen={}, harris=False, thresh=60, octaves=4, metric=METRIC,
               workload="C3: AGAST(60,4) detect + BRISK2 describe, 1920x1080 synthetic frames"),
    "C4": dict(w=3840, h=2160, frames=128, cap=49152, unique=4, seed=3000, gen=dict(n_shapes=4200), harris=False, thresh=60, octaves=6,
               metric="2160p_frames_per_s_detect_describe",
               workload="C4: AGAST(60, 6 octaves) detect + BRISK2 describe, 3840x2160 synthetic frames (describe stress)"),
}


def unique_frames(cfg, rank=0):
    from ethzasl_brisk_b200.synthetic import synthetic_frame
    s0 = cfg["seed"] + rank * cfg["unique"]
    return np.stack([synthetic_frame(cfg["w"], cfg["h"], s0 + i, **cfg["gen"]) for i in range(cfg["unique"])])


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
    """(HBM GB/s, dense int8 TOP/s, source).  The tensor figure is 2 x the measured cuBLAS bf16 burst rate (int8 runs at twice
    the bf16 rate on the same pipe; no int8 library GEMM was measured)."""
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get("hbm_gbs", 6650.0), 2.0 * d.get("bf16_tflops", 1590.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 2.0 * 1590.0, "fallback (B200_PROFILING.md)"


def layer_dims(w, h, octaves):
    dims = [(w, h), (2 * (w // 3), 2 * (h // 3))]
    for i in range(2, 2 * octaves):
        dims.append((dims[i - 2][0] // 2, dims[i - 2][1] // 2))
    return dims[:max(1, 2 * octaves)]


def stage_bytes(cfg, n_corners, n_kps, desc_bytes=48):
    """ALGORITHMIC bytes per frame of each stage: unique bytes in + bytes out (SURVEY.md 8d's per-unit figures; DESIGN.md
    section 5).  `nms` has no figure in SURVEY (its input is the scoring stage's output): the per-corner records it must read
    and the key points it writes."""
    w, h = cfg["w"], cfg["h"]
    px = [a * b for a, b in layer_dims(w, h, cfg["octaves"])]
    b = {
        "pyramid": px[0] + sum(px[1:]),                       # L0 in once (one fused kernel) + every derived layer out
        "integral": px[0] + 4 * (w + 1) * (h + 1),            # u8 in, (h+1) x (w+1) i32 out
        "describe": n_kps * (28 * 2 + desc_bytes),            # compulsory HBM only: key point in / out, descriptor out
    }
    if cfg["harris"]:
        b["harris"] = 5 * sum(px)                             # 1 B/px in + 4 B/px scores out, all layers
    else:
        b["detect"] = sum(px) + 12 * n_corners                # 1 B/px over all layers + 12 B per emitted corner
        b["lists"] = 12 * n_corners                           # (the corner lists are part of SURVEY's scoring figure; kept apart: own kernels)
        b["nms"] = 12 * n_corners + 28 * n_kps                # corner records in, key points out
    return b


def traffic_table():
    """DRAM bytes per frame of every stage's kernels, from the committed ncu --set full capture of the CURRENT kernels
    (profiles/r02_dram_traffic.json, written by tools/ncu_traffic.py from the .ncu-rep; dram__bytes_read.sum +
    dram__bytes_write.sum per launch / frames per launch).  None when the file is absent."""
    p = ROOT / "profiles" / "r02_dram_traffic.json"
    if not p.exists():
        return None
    return json.loads(p.read_text())


LIMITER = {
    "describe": "L1/L2 sector rate of scattered 16-byte gathers from the integral blocks (not HBM); see profiles/",
    "nms": "ALU pipe / instruction issue of the per-corner score evaluation and the tie chain (not HBM); see profiles/",
    "detect": "ALU pipe: packed min/max of the threshold map and the arc test (not HBM); see profiles/",
    "harris": "the std::sort replay and the sequential uniformity stamping per (frame, layer); see DESIGN.md section 7",
}


def ref_detect_describe(ref, cfg, frame):
    if cfg["harris"]:
        k = ref.harris_detect(frame, cfg["octaves"], cfg["radius"], cfg["abs_thr"], -1)
    else:
        k = ref.agast_detect(frame, cfg["thresh"], cfg["octaves"], cap=1 << 19)
    return ref.describe(frame, k)


def ref_bench(ref, cfg, frames, nthreads):
    return ref.bench_detect_describe(frames, cfg["harris"], cfg.get("thresh", 60), cfg["octaves"], cfg.get("radius", 30.0),
                                     cfg.get("abs_thr", 20.0), nthreads=nthreads)


def run_reference(args):
    """CPU arm: the unmodified reference (oracle/_ref) on the host cores, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import ref
    cores = host_cores()
    if args.config == "C5":
        import ethzasl_brisk_b200 as bb
        nq, nt = min(4096, args.knn_q), min(1000000, args.knn_t)
        q, t = bb.random_descriptors(nq, 64, 5), bb.random_descriptors(nt, 64, 6)
        for _ in range(min(args.warmup, 1)):
            ref.knn(q[:64], t, 2, nthreads=cores)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ref.knn(q, t, 2, nthreads=cores)
        dt = time.perf_counter() - t0
        value = args.steps * nq * nt / dt / 1e9
        sample = f"{nq} queries x {nt} train rows per step (uniform random 512-bit rows), k = 2, {cores} threads"
        line = {"impl": "reference", "metric": "hamming_knn_k2_512bit", "value": value, "unit": "Gcmp/s", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u8", "data": "synthetic",
                "config": {"workload": "C5: brute-force Hamming kNN k=2, 512-bit descriptors", "queries": nq, "train": nt},
                "cpu_baseline": {"value": value, "unit": "Gcmp/s", "cores": cores, "kind": "reference", "sample": sample},
                "e2e": {"value": value, "unit": "Gcmp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        emit(line)
        return 0
    cfg = FRAME_CONFIGS[args.config]
    frames = unique_frames(cfg)
    px_scale = (1920 * 1080) / (cfg["w"] * cfg["h"])
    per_step = int(min(256 * max(px_scale, 1), max(cfg["unique"], 2 * cores * max(int(px_scale), 1))))
    per_step = max(per_step, cores)
    frames = frames[np.arange(per_step) % cfg["unique"]]
    sample = f"{per_step} synthetic {cfg['w']}x{cfg['h']} frames per step ({cfg['unique']} distinct), {cores} threads (one detector per thread)"
    for _ in range(args.warmup):
        ref_bench(ref, cfg, frames[:cores], cores)
    t = kp = 0
    for _ in range(args.steps):
        s, k = ref_bench(ref, cfg, frames, cores)
        t += s; kp += k
    value = args.steps * len(frames) / t
    line = {"impl": "reference", "metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": cfg["workload"], "frames_per_step": len(frames),
                       "keypoints_per_frame": kp / (args.steps * len(frames))},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)
    return 0


class Rig:
    """Process group, device, clocks and timing helpers shared by the GPU arms."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
        # NCCL's own output (version banner, INFO lines with the communicator's nranks) ends up on stderr: main() pointed
        # file descriptor 1 there.  At N > 1 INFO is on unless the caller chose a level.
        if self.world > 1:
            # INFO (communicator setup lines incl. nranks) unless BENCH_NCCL_DEBUG says otherwise; a preset lower level
            # (images often export NCCL_DEBUG=WARN / VERSION) would hide them
            os.environ["NCCL_DEBUG"] = os.environ.get("BENCH_NCCL_DEBUG", "INFO")
        torch.cuda.set_device(self.local)
        self.numa = pin_to_gpu_numa_node(self.local) if self.world > 1 else None
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.current_stream()
        self.sampler = ClockSampler(self.local)
        if self.rank == 0:
            self.sampler.start()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, ctx=None):
        """K calls bracketed by barrier + synchronize on both sides, CUDA events on the launching stream, max over ranks."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        stages, launches = {}, 0
        for _ in range(steps):
            fn()
            if ctx is not None:
                ms, l = ctx.last_timing()
                launches += l
                for k, v in ms.items():
                    stages[k] = stages.get(k, 0.0) + v
        e1.record(self.stream)
        self.barrier()
        t = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item()), stages, launches

    def finish(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def run_frames(args):
    import ethzasl_brisk_b200 as bb
    cfg = FRAME_CONFIGS[args.config]
    rig = Rig(args)
    torch, dev, world, rank = rig.torch, rig.dev, rig.world, rig.rank
    W, H = cfg["w"], cfg["h"]
    n = args.frames or cfg["frames"]
    cap = args.cap or cfg["cap"]

    # frames: `unique` generated frames per rank, cycled with a per-copy horizontal roll so that all n differ
    uniq_host = unique_frames(cfg, rank)
    uniq = torch.from_numpy(uniq_host).to(dev)
    U = cfg["unique"]
    d_frames = torch.empty((n, H, W), dtype=torch.uint8, device=dev)
    for j in range(n):
        d_frames[j] = torch.roll(uniq[j % U], shifts=(j // U) * 5, dims=1)
    h_frames = torch.empty((n, H, W), dtype=torch.uint8).pin_memory()
    h_frames.copy_(d_frames)
    torch.cuda.synchronize()

    ctx = bb.Context(rig.local, stream=rig.stream.cuda_stream, timing=True, workspace_limit=args.workspace_gb << 30)
    if cfg["harris"]:
        det = bb.ScaleSpaceFeatureDetector(cfg["octaves"], cfg["radius"], cfg["abs_thr"], ctx=ctx)
    else:
        det = bb.BriskFeatureDetector(cfg["thresh"], cfg["octaves"], ctx=ctx)
    ext = bb.BriskDescriptorExtractor(ctx=ctx)
    d_out = (torch.empty((n, cap, 7), dtype=torch.float32, device=dev), torch.empty(n, dtype=torch.int32, device=dev),
             torch.empty((n, cap, 48), dtype=torch.uint8, device=dev))
    h_out = (torch.empty((n, cap, 7), dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory(),
             torch.empty((n, cap, 48), dtype=torch.uint8).pin_memory())
    h_out2 = tuple(torch.empty_like(t).pin_memory() for t in h_out)
    resident = lambda: bb.detect_and_compute_batch(det, ext, d_frames, cap=cap, out=d_out)
    e2e = lambda: bb.detect_and_compute_batch(det, ext, h_frames, cap=cap, out=h_out)

    def e2e_streamed(steps):
        # the streaming form of the same call (brisk_detect_describe_async + brisk_sync): the caller alternates two sets
        # of host output buffers, so the first upload / last download of a step run under the neighbouring steps' kernels
        for s_ in range(steps):
            bb.detect_and_compute_batch(det, ext, h_frames, cap=cap, out=(h_out, h_out2)[s_ % 2], async_=True)
        ctx.sync()

    # nvidia-smi takes a few hundred ms to deliver its first line: it was started before the warm-up; only the
    # lines of the two timed regions (resident and host end-to-end) count
    # The headline regions run the library the way a caller does: no per-stage events.  Stage times come from a separate pass.
    ctx.enable_timing(False)
    for _ in range(args.warmup):
        resident()
    if rank == 0:
        rig.sampler.mark()
    ms_res, _, launches = rig.timed(resident, args.steps, ctx)
    if rank == 0:
        rig.sampler.pause()
    counts = d_out[1].cpu().numpy()
    # per-stage device times without cross-stream overlap (same kernels, one stream, stages back to
    # back): these feed the per-kernel roofline numbers
    ctx.enable_timing(True)
    ctx.set_pipelining(False)
    resident()
    _, stages_serial, _ = rig.timed(resident, 1, ctx)
    raw_corners = ctx.last_raw_corners()
    stages = {k: v * args.steps for k, v in stages_serial.items()}
    ctx.set_pipelining(True)
    ctx.enable_timing(False)
    for _ in range(max(1, args.warmup // 2)):
        e2e()
    if rank == 0:
        rig.sampler.mark()
    ms_e2e_blocking, _, _ = rig.timed(e2e, args.steps, ctx)
    e2e_streamed(2)
    ms_e2e, _, _ = rig.timed(lambda: e2e_streamed(args.steps), 1)
    clocks = rig.sampler.stop() if rank == 0 else None
    if getattr(pin_to_gpu_numa_node, "full", None):
        os.sched_setaffinity(0, pin_to_gpu_numa_node.full)  # the CPU legs below use every host core
    hc = h_out[1].numpy()
    assert np.array_equal(hc, counts) and np.array_equal(h_out2[1].numpy(), counts), "host and device paths disagree"
    assert torch.equal(h_out[2][-1, :int(counts[-1])], h_out2[2][-1, :int(counts[-1])]), "the two host buffer sets disagree"
    assert counts.max() <= cap, "key-point capacity exceeded"
    kp_total = int(counts.sum())

    # secondary metric (C3 only): brute-force Hamming kNN (k = 2), 512-bit descriptors.  One rank: reduced C5 shape.
    # N ranks: the TRAIN set is sharded over the ranks (each holds --knn-t rows), every rank searches its shard,
    # NCCL all-gather of the per-shard top-2 keys + merge (distributed.sharded_knn): weak scaling in the train size.
    secondary = None
    if args.config == "C3" and not args.no_knn:
        secondary = knn_measure(args, rig, ctx, bb, args.knn_q, args.knn_t * world, steps=3)

    if rank != 0:
        rig.finish()
        return 0

    value = world * n * args.steps / (ms_res * 1e-3)
    # end to end through the C ABI with pinned host buffers, both ways a caller can make the call; `value` is the faster
    # one on this box (streaming wins where the copies hide behind kernels, the blocking call where N ranks saturate the
    # host's memory system and overlapping steps only adds contention)
    streamed_value = world * n * args.steps / (ms_e2e * 1e-3)
    blocking_value = world * n * args.steps / (ms_e2e_blocking * 1e-3)
    streamed = streamed_value >= blocking_value
    e2e_value = max(streamed_value, blocking_value)
    e2e_obj = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * H * W, "d2h_bytes_per_step": 4 * n + kp_total * (28 + 48),
               "ms_per_step": (ms_e2e if streamed else ms_e2e_blocking) / args.steps,
               "call": ("brisk_detect_describe_async x steps + brisk_sync (outputs double-buffered by the caller)" if streamed
                        else "brisk_detect_describe (blocking)") + ", pinned host buffers",
               "streaming_call_value": streamed_value, "streaming_call_ms_per_step": ms_e2e / args.steps,
               "blocking_call_value": blocking_value, "blocking_call_ms_per_step": ms_e2e_blocking / args.steps}
    h2d = n * H * W
    d2h = 4 * n + kp_total * (28 + 48)
    # roofline of the dominant stage (device time from CUDA events on the launching stream)
    hbm_peak, _, peak_src = measured_peaks()
    kps_per_frame = kp_total / n
    corners_per_frame = raw_corners / n
    bytes_per_frame = stage_bytes(cfg, corners_per_frame, kps_per_frame)
    names = ("pyramid", "harris", "integral", "describe") if cfg["harris"] else ("pyramid", "detect", "lists", "nms", "integral", "describe")
    serial = dict(stages_serial)
    if cfg["harris"]:
        serial["harris"] = serial.get("nms", 0.0)  # the Harris kernels run between the NMS stage marks
        stages["harris"] = stages.get("nms", 0.0)
    compute_stages = {k: serial.get(k, 0.0) for k in names}
    total_ms = sum(compute_stages.values())
    top = max(compute_stages, key=compute_stages.get)
    traffic = traffic_table() if args.config == "C3" else None
    stage_report = {}
    for k, v in compute_stages.items():
        gbs = bytes_per_frame[k] * n / (v * 1e-3) / 1e9 if v > 0 else 0.0
        stage_report[k] = {"ms_per_step": v, "share": v / total_ms if total_ms else 0.0, "algorithmic_bytes_per_frame": bytes_per_frame[k],
                           "algorithmic_GBps": gbs, "frac_of_hbm_peak": gbs / hbm_peak}
        if traffic and k in traffic.get("stages", {}):
            stage_report[k]["ncu_dram_bytes_per_frame"] = traffic["stages"][k]
    roof = {"bound": "hbm", "kernel": top, "achieved": stage_report[top]["algorithmic_GBps"], "peak": hbm_peak, "unit": "GB/s",
            "frac": stage_report[top]["frac_of_hbm_peak"], "traffic": None, "peak_source": peak_src,
            "algorithmic_bytes": bytes_per_frame[top] * n,
            "note": "dominant stage by CUDA-event time of a non-pipelined step (stages back to back on one stream); algorithmic bytes = "
                    "SURVEY.md 8d per-frame figure x frames per step; see `stages` for every stage"}
    if traffic and top in traffic.get("stages", {}):
        roof["traffic"] = traffic["stages"][top] * n
        roof["traffic_note"] = f"ncu dram__bytes_read.sum + dram__bytes_write.sum per frame x frames of the step (capture: {traffic.get('source')})"
    if top in LIMITER:
        roof["limiter"] = LIMITER[top]
    if "integral" in stage_report:  # the block layout trades 3.6x the output bytes for 3.5x fewer gather instructions in describe
        real = 2 * W * H + 16 * W * H
        stage_report["integral"]["hbm_traffic_GBps"] = real * n / (compute_stages["integral"] * 1e-3) / 1e9
        stage_report["integral"]["hbm_traffic_frac_of_peak"] = stage_report["integral"]["hbm_traffic_GBps"] / hbm_peak
    if stage_report.get("describe", {}).get("ms_per_step", 0) > 0:
        # describe is gather bound, not HBM bound -- its meaningful rate is key points per second
        stage_report["describe"]["keypoints_per_s"] = kp_total / (stage_report["describe"]["ms_per_step"] * 1e-3)

    # parity of the benched workload itself: randomly chosen frames of the batch (key points + descriptors of the end-to-end
    # run) against the unmodified reference, and the CPU baseline beside the GPU numbers on a bounded sample of the same frames
    parity, cpu = None, None
    try:
        from oracle import ref
        if ref.available():
            rng = np.random.default_rng(12345)
            pick = sorted(rng.choice(n, size=min(args.parity_frames, n), replace=False).tolist())
            hk = h_out[0].numpy().view(np.uint8).reshape(n, cap, 28)
            ok, kp_checked, bad = True, 0, []
            for f in pick:
                frame = h_frames[f].numpy()
                k2, d2 = ref_detect_describe(ref, cfg, frame)
                m = int(hc[f])
                k1 = np.frombuffer(hk[f, :m].tobytes(), bb.KP_DTYPE)
                good = m == len(k2) and all(np.array_equal(k1[fld], k2[fld]) for fld in ("x", "y", "size", "response", "octave", "class_id")) \
                    and (m == 0 or float(np.abs(k1["angle"] - k2["angle"]).max()) <= 1e-4) and np.array_equal(h_out[2][f, :m].numpy(), d2)
                ok &= bool(good)
                kp_checked += m
                if not good:
                    bad.append(f)
            parity = {"parity_checked_frames": len(pick), "parity_ok": ok, "keypoints_checked": kp_checked, "frames": pick, "mismatching_frames": bad,
                      "against": "oracle/_ref (the unmodified reference), bit-exact key points and descriptors, angle within 1e-4 deg"}
            cores = host_cores()
            px_scale = max(1, int((1920 * 1080) / (W * H)))
            ns = min(256 * px_scale, max(U, 2 * cores * px_scale))
            sample_frames = uniq_host[np.arange(ns) % U]
            ref_bench(ref, cfg, sample_frames[:2], min(2, cores))
            s, _ = ref_bench(ref, cfg, sample_frames, cores)
            cpu = {"value": len(sample_frames) / s, "unit": UNIT, "cores": cores, "kind": "reference",
                   "sample": f"{len(sample_frames)} frames drawn from the same {U} distinct {W}x{H} frames, unmodified reference (oracle/_ref), {cores} threads"}
    except Exception as e:  # the baseline is informative; never fail the bench on it
        cpu = cpu or {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}

    line = {"metric": cfg["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_res / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": cfg["workload"], "frames_per_gpu_per_step": n,
                       "global_frames_per_step": world * n, "keypoints_per_frame": kps_per_frame, "raw_corners_per_frame": corners_per_frame,
                       "parallelism": f"frame-sharded x{world}, no collective", "cpus_per_rank": rig.numa,
                       "l2": f"inputs ({n * H * W / 1e6:.0f} MB per step) exceed the 126 MB L2; no explicit flush", "kp_capacity": cap},
            "e2e": e2e_obj,
            "gpu_launches": launches, "roofline": roof, "stages": stage_report, "cpu_baseline": cpu, "clocks": clocks}
    if parity:
        line.update({"parity_checked_frames": parity["parity_checked_frames"], "parity_ok": parity["parity_ok"], "parity": parity})
    if secondary:
        line["secondary"] = secondary
    emit(line)
    rig.finish()
    return 0


def knn_measure(args, rig, ctx, bb, nq, nt_total, steps, full_report=False):
    """Brute-force Hamming 2-NN of nq queries against nt_total 512-bit train rows sharded contiguously over the ranks
    (rank r draws its shard from seed 6 + r; the global train set is their concatenation).  Collective on all ranks;
    returns the report on rank 0."""
    from ethzasl_brisk_b200.distributed import shard_range, sharded_knn
    torch, dev, world, rank = rig.torch, rig.dev, rig.world, rig.rank
    begin, end = shard_range(nt_total, rank, world)
    q = torch.from_numpy(bb.random_descriptors(nq, 64, 5)).to(dev)
    t = torch.from_numpy(bb.random_descriptors(end - begin, 64, 6 + rank)).to(dev)
    m = bb.BruteForceMatcher(ctx=ctx)
    variants, res = {}, None
    for name, variant in (("popc", 0), ("mma_sync_imma", 1), ("tcgen05_i8", 2), ("tcgen05_fp4", 3)):
        if variant < 2 and full_report and not args.knn_popc:
            continue
        ctx.set_knn_variant(variant)
        fn = (lambda: sharded_knn(m, q, t, 2, begin)) if world > 1 else (lambda: m.knn(q, t, 2))
        res = fn()
        v_ms, _, _ = rig.timed(fn, steps)
        variants[name] = {"Gcmp/s": nq * nt_total * steps / (v_ms * 1e-3) / 1e9, "ms": v_ms / steps}
    best = variants["tcgen05_fp4"]   # the default path: tcgen05.mma kind::mxf4 on +-1.0 E2M1 operands, TMEM accumulators, TMA operands (hamming_tc5.cu)
    # end to end: queries and the train shard start in pinned host memory, the result is read back
    hq, ht = q.cpu().pin_memory(), t.cpu().pin_memory()

    def e2e():
        qq, tt = hq.to(dev, non_blocking=True), ht.to(dev, non_blocking=True)
        r = sharded_knn(m, qq, tt, 2, begin) if world > 1 else m.knn(qq, tt, 2)
        return r[0].cpu(), r[1].cpu()
    e2e()
    e_ms, _, _ = rig.timed(e2e, steps)
    idx, dist_ = (res[0].cpu().numpy(), res[1].cpu().numpy())
    if rank != 0:
        return None
    _, int8_peak, peak_src = measured_peaks()
    fp4_peak = 2.0 * int8_peak   # dense FP4 runs at twice the int8 / fp8 rate (nominal 9 against 4.5 PFLOP/s)
    gcmp = best["Gcmp/s"]
    tops = gcmp * 2 * 512 / 1e3  # 512 MACs = 1024 operations per 512-bit comparison on the tensor pipe
    rep = {"metric": "hamming_knn_k2_512bit", "value": gcmp, "unit": "Gcmp/s", "queries": nq, "train": nt_total, "train_rows_per_rank": end - begin,
           "ms": best["ms"], "variants": variants, "sharding": (f"train set sharded over {world} ranks, NCCL all-gather of per-shard top-2 keys + merge"
                                                                 if world > 1 else "one rank, no collective"),
           "e2e": {"value": nq * nt_total * steps / (e_ms * 1e-3) / 1e9, "unit": "Gcmp/s", "h2d_bytes_per_step": int((nq + end - begin) * 64),
                   "d2h_bytes_per_step": int(nq * 2 * 8), "ms_per_step": e_ms / steps},
           "roofline": {"bound": "tensor", "achieved": tops / world, "peak": fp4_peak, "unit": "TFLOP/s", "frac": tops / world / fp4_peak, "traffic": None,
                        "peak_source": peak_src + ": 4 x the measured bf16 cuBLAS burst rate (dense FP4 runs at four times the bf16 rate; no FP4 library GEMM was measured)",
                        "frac_of_int8_peak": tops / world / int8_peak,
                        "note": "per GPU; multiply-adds of the +-1.0 E2M1 operands counted as flops (1024 per 512-bit comparison); the kind::i8 kernel's "
                                "figure is in `variants`"}}
    # parity of the benched result: a random sample of the queries against the reference's own matcher loop on the full train set
    try:
        from oracle import ref
        if ref.available():
            cores = host_cores()
            train = np.concatenate([bb.random_descriptors(shard_range(nt_total, r, world)[1] - shard_range(nt_total, r, world)[0], 64, 6 + r)
                                    for r in range(world)]) if world > 1 else t.cpu().numpy()
            rng = np.random.default_rng(7)
            pick = np.sort(rng.choice(nq, size=min(256, nq), replace=False))
            qs = q.cpu().numpy()[pick]
            cap_t = min(len(train), max(1, int(4.0e9 // len(pick))))  # bounds the CPU work to a few seconds
            if cap_t == len(train):
                i2, d2 = ref.knn(qs, train, 2, nthreads=cores)
                rep["parity_ok"] = bool(np.array_equal(i2, idx[pick]) and np.array_equal(d2, dist_[pick]))
                rep["parity_checked_queries"] = int(len(pick))
            tq = min(4096, nq)
            t0 = time.perf_counter()
            ref.knn(q.cpu().numpy()[:tq], train[:min(1000000, len(train))], 2, nthreads=cores)
            rep["cpu_baseline"] = {"value": tq * min(1000000, len(train)) / (time.perf_counter() - t0) / 1e9, "unit": "Gcmp/s", "cores": cores, "kind": "reference",
                                   "sample": f"{tq} queries x {min(1000000, len(train))} train rows of the same descriptors, k = 2, the reference's Hamming primitive in "
                                             f"its k-successive-arg-min loop, {cores} threads"}
    except Exception as e:
        rep["cpu_baseline"] = {"value": None, "unit": "Gcmp/s", "cores": 0, "kind": "reference", "sample": f"unavailable: {e}"}
    return rep


def run_knn(args):
    """--config C5: the matcher as the headline line (same schema)."""
    import ethzasl_brisk_b200 as bb
    rig = Rig(args)
    ctx = bb.Context(rig.local, stream=rig.stream.cuda_stream, timing=True, workspace_limit=args.workspace_gb << 30)
    if rig.rank == 0:
        rig.sampler.mark()
    rep = knn_measure(args, rig, ctx, bb, args.knn_q, args.knn_t, steps=args.steps, full_report=True)
    clocks = rig.sampler.stop() if rig.rank == 0 else None
    if rig.rank == 0:
        line = {"metric": rep["metric"], "value": rep["value"], "unit": "Gcmp/s", "n_gpus": rig.world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": rep["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": f"C5: brute-force Hamming kNN k=2, {args.knn_q} queries x {args.knn_t} train rows, 512-bit, {rep['sharding']}",
                           "queries": args.knn_q, "train": args.knn_t, "l2": "train set exceeds the 126 MB L2"},
                "e2e": rep["e2e"], "gpu_launches": 2 * args.steps, "roofline": rep["roofline"], "cpu_baseline": rep.get("cpu_baseline"),
                "clocks": clocks, "variants": rep["variants"]}
        for k in ("parity_ok", "parity_checked_queries"):
            if k in rep:
                line[k] = rep[k]
        emit(line)
    rig.finish()
    return 0


_JSON_FD = None


def emit(line):
    """The one JSON line, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    # stdout carries exactly one JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its
    # version banner there whatever NCCL_DEBUG_FILE says): point fd 1 at stderr for the whole run and keep the original
    # stdout for emit().
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C3", choices=["C2", "C3", "C4", "C5"])
    ap.add_argument("--frames", type=int, default=0, help="frames per GPU per step (default: the configuration's)")
    ap.add_argument("--cap", type=int, default=0, help="key-point capacity per frame (default: the configuration's)")
    ap.add_argument("--workspace-gb", type=int, default=32)
    ap.add_argument("--knn-q", type=int, default=None)
    ap.add_argument("--knn-t", type=int, default=None, help="train rows (C3's secondary metric: per rank; C5: in total)")
    ap.add_argument("--knn-popc", action="store_true", help="C5: also time the POPC and mma.sync kernels")
    ap.add_argument("--no-knn", action="store_true", help="C3: skip the secondary matcher metric")
    ap.add_argument("--parity-frames", type=int, default=8)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.knn_q is None:
        args.knn_q = 1000000 if args.config == "C5" else 100000
    if args.knn_t is None:
        args.knn_t = 10000000 if args.config == "C5" else 1000000
    if args.impl == "reference":
        return run_reference(args)
    if args.config == "C5":
        return run_knn(args)
    return run_frames(args)


if __name__ == "__main__":
    sys.exit(main())
