"""SM clock, power draw and throttle reasons while a matcher kernel runs back to back (GPU box only):
usage knn_power_probe.py [variant ...]"""
import subprocess, sys, threading, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ethzasl_brisk_b200 as bb

variants = [int(v) for v in sys.argv[1:]] or [2, 3]
ctx = bb.Context(0, timing=True)
m = bb.BruteForceMatcher(ctx=ctx)
nq, nt = 100000, 1000000
q = torch.from_numpy(bb.random_descriptors(nq, 64, 5)).cuda(); t = torch.from_numpy(bb.random_descriptors(nt, 64, 6)).cuda()
for v in variants:
    ctx.set_knn_variant(v)
    m.knn(q, t, 2)
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_throttle_reasons.sw_power_cap,clocks_throttle_reasons.hw_slowdown",
                          "--format=csv,noheader,nounits", "-lms", "100", "-i", "0"], stdout=subprocess.PIPE, text=True)
    t0 = time.time(); ms = []
    while time.time() - t0 < 4.0:
        m.knn(q, t, 2); ms.append(ctx.last_timing()[0]["knn"])
    p.terminate()
    rows = [l.strip().split(", ") for l in p.stdout.read().splitlines() if l.strip()]
    rows = rows[len(rows) // 3:]   # steady state
    clk = sorted(float(r[0]) for r in rows); pw = sorted(float(r[1]) for r in rows)
    cap = sum(1 for r in rows if r[2].lower().startswith("active")) / max(len(rows), 1)
    print(f"variant {v}: {nq * nt / min(ms) / 1e9:.2f} Tcmp/s best, {nq * nt / (sum(ms[-10:]) / 10) / 1e9:.2f} sustained; SM clock median {clk[len(clk) // 2]:.0f} MHz, "
          f"power median {pw[len(pw) // 2]:.0f} W, sw_power_cap active in {100 * cap:.0f} % of the samples")
