"""FP4 (kind::mxf4) matcher against the POPC kernel on a few shapes, then timing at 100 k x 1 M.  Run under `timeout`."""
import sys, time
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import ethzasl_brisk_b200 as bb

ctx = bb.Context(0, timing=True)
m = bb.BruteForceMatcher(ctx=ctx)
ok = True
for nb, nq, nt in ((64, 300, 9000), (48, 300, 9000), (48, 1000, 50000), (48, 257, 97), (64, 5000, 200001), (48, 5000, 200001)):
    q, t = bb.random_descriptors(nq, nb, 1), bb.random_descriptors(nt, nb, 2)
    t[min(100, nt - 1)] = q[3]; t[min(200, nt - 1)] = q[3]
    ctx.set_knn_variant(0); i0, d0 = m.knn(q, t, 2)
    ctx.set_knn_variant(3); i3, d3 = m.knn(q, t, 2)
    same = np.array_equal(i0, i3) and np.array_equal(d0, d3)
    print(f"{nb} B, {nq} x {nt}: identical to POPC: {same}", flush=True)
    if not same:
        bad = np.argwhere((i0 != i3) | (d0 != d3))[:5]
        print("  first mismatches (query, k):", bad.tolist(), i0[bad[:, 0]].tolist(), i3[bad[:, 0]].tolist(), d0[bad[:, 0]].tolist(), d3[bad[:, 0]].tolist())
        ok = False
if ok and len(sys.argv) > 1:
  for nb in (64, 48):
    nq, nt = 100000, 1000000
    q = torch.from_numpy(bb.random_descriptors(nq, nb, 5)).cuda(); t = torch.from_numpy(bb.random_descriptors(nt, nb, 6)).cuda()
    for variant in (2, 3):
        ctx.set_knn_variant(variant)
        best = 1e9
        for rep in range(4):
            m.knn(q, t, 2)
            best = min(best, ctx.last_timing()[0].get("knn", 1e9))
        print(f"{nb} B variant {variant}: {best:.2f} ms, {nq * nt / best / 1e9:.2f} Tcmp/s", flush=True)
