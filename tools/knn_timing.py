"""Wall / device time of the Hamming 2-NN variants (GPU box only): usage knn_timing.py [nq nt [variants]]"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
import ethzasl_brisk_b200 as bb

nq, nt = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (100000, 1000000)
variants = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [2]
ctx = bb.Context(0, timing=True)
m = bb.BruteForceMatcher(ctx=ctx)
out = {}
for nbytes in (64, 48, 64, 48):
    q = torch.from_numpy(bb.random_descriptors(nq, nbytes, 5)).cuda()
    t = torch.from_numpy(bb.random_descriptors(nt, nbytes, 6)).cuda()
    for v in variants:
        ctx.set_knn_variant(v)
        m.knn(q, t, 2)
        walls, devs = [], []
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            m.knn(q, t, 2)
            walls.append(1e3 * (time.perf_counter() - t0))
            devs.append(ctx.last_timing()[0]["knn"])
        out.setdefault(f"{nbytes}B_v{v}", []).append({"wall_ms": [round(w, 2) for w in walls], "dev_ms": [round(d, 2) for d in devs],
                                                      "Gcmp/s_dev": round(nq * nt / min(devs) / 1e6, 1)})
print(json.dumps(out))
