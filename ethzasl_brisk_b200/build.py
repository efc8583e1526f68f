"""Build libbrisk_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

The library is linked against the static CUDA runtime and has no Python or
torch dependency; it travels to the GPU box with the repo snapshot.
"""
import os
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libbrisk_b200.so"
SOURCES = ["pyramid.cu", "detect.cu", "nms.cu", "describe.cu", "hamming.cu", "hamming_mma.cu", "hamming_tc5.cu", "harris.cu", "capi.cu", "pattern.cc"]
# -fmad=false: the reference is built without FMA contraction and float results
# are compared bit for bit.  The system g++ is pinned as host
# compiler (the image's CXX links libstdc++ statically).
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
              "-ccbin", "/usr/bin/g++", "-Xcompiler", "-fPIC,-ffp-contract=off,-O2", "-cudart", "static"]


def _stale():
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*")) + [PKG.parent / "include" / "brisk_b200.h", Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a and link the shared library."""
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    obj_dir = PKG / "build"
    obj_dir.mkdir(exist_ok=True)
    procs = []
    for src in SOURCES:
        if not (CSRC / src).exists():
            continue
        obj = obj_dir / (src + ".o")
        # BRISK_B200_NVCC_EXTRA: extra flags for tuning experiments (e.g. -DBRISK_DESC_MIN_BLOCKS=5)
        extra = os.environ.get("BRISK_B200_NVCC_EXTRA", "").split()
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-Xptxas", "-v", "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, obj, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            print(f"--- {src}\n{out}")
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
        objs.append(str(obj))
    cmd = [nvcc, "-shared", "-ccbin", "/usr/bin/g++", "-cudart", "static", "-o", str(LIB), *objs]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
