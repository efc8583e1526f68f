"""Time the POPC and tensor-core Hamming 2-NN kernels on the same inputs (GPU box only)."""
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import ethzasl_brisk_b200 as bb


def main():
    nq, nt = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (100000, 1000000)
    ctx = bb.Context(0)
    m = bb.BruteForceMatcher(ctx=ctx)
    out = {}
    for nbytes in (48, 64):
        q = torch.from_numpy(bb.random_descriptors(nq, nbytes, 5)).cuda()
        t = torch.from_numpy(bb.random_descriptors(nt, nbytes, 6)).cuda()
        res = []
        for variant in (0, 1, 2, 3):
            ctx.set_knn_variant(variant)
            m.knn(q, t, 2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(3):
                r = m.knn(q, t, 2)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            res.append(r)
            out[f"{nbytes}B_variant{variant}"] = {"ms": ms, "Gcmp/s": nq * nt / ms / 1e6}
        same = all(np.array_equal(np.asarray(a.cpu() if hasattr(a, "cpu") else a), np.asarray(b.cpu() if hasattr(b, "cpu") else b))
                   for other in res[1:] for a, b in zip(res[0], other))
        out[f"{nbytes}B_identical"] = bool(same)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
