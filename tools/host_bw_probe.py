#!/usr/bin/env python
"""Host <-> device copy bandwidth of the box with all ranks copying at once (the ceiling of the end-to-end path).

  python tools/host_bw_probe.py                                   # one GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/host_bw_probe.py

Every rank moves what one bench step moves per GPU (2.12 GB up, 0.52 GB down by default) between pinned host memory and
its GPU, H2D and D2H on two streams at the same time, `--reps` times, bracketed by barriers.  Rank 0 prints one JSON line:
per-rank and aggregate GB/s for H2D alone, D2H alone and both together, and the step time the copies alone would take.
"""
import argparse
import json
import os

import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--h2d-mb", type=int, default=2123)
    ap.add_argument("--d2h-mb", type=int, default=520)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    out_fd = os.dup(1)
    os.dup2(2, 1)   # NCCL's banner goes to fd 1: keep stdout for the JSON line
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    up_h = torch.empty(a.h2d_mb << 20, dtype=torch.uint8).pin_memory()
    up_d = torch.empty(a.h2d_mb << 20, dtype=torch.uint8, device=dev)
    dn_h = torch.empty(a.d2h_mb << 20, dtype=torch.uint8).pin_memory()
    dn_d = torch.empty(a.d2h_mb << 20, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(do_up, do_dn):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s_up.wait_stream(torch.cuda.current_stream()); s_dn.wait_stream(torch.cuda.current_stream())
        for _ in range(a.reps):
            if do_up:
                with torch.cuda.stream(s_up):
                    up_d.copy_(up_h, non_blocking=True)
            if do_dn:
                with torch.cuda.stream(s_dn):
                    dn_h.copy_(dn_d, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s_up); torch.cuda.current_stream().wait_stream(s_dn)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / a.reps

    timed(True, True)
    ms_up, ms_dn, ms_both = timed(True, False), timed(False, True), timed(True, True)
    if rank == 0:
        gb_up, gb_dn = a.h2d_mb * (1 << 20) / 1e9, a.d2h_mb * (1 << 20) / 1e9
        os.write(out_fd, (json.dumps({"n_gpus": world, "h2d_GB_per_rank": gb_up, "d2h_GB_per_rank": gb_dn,
                          "h2d_alone": {"ms": ms_up, "GBps_per_rank": gb_up / ms_up * 1e3, "GBps_aggregate": world * gb_up / ms_up * 1e3},
                          "d2h_alone": {"ms": ms_dn, "GBps_per_rank": gb_dn / ms_dn * 1e3, "GBps_aggregate": world * gb_dn / ms_dn * 1e3},
                          "both": {"ms": ms_both, "GBps_per_rank": (gb_up + gb_dn) / ms_both * 1e3, "GBps_aggregate": world * (gb_up + gb_dn) / ms_both * 1e3},
                          "note": "ms = time of one bench step's host traffic with every rank copying at once (max over ranks); the end-to-end "
                                  "step cannot be shorter than `both.ms` on this box"}) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
