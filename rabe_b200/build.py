"""In-tree build of librabe_b200.so (nvcc, sm_100a only) -- no JIT cache, the .so travels with the tree."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "librabe_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CUDA_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]
HOST_SRCS = ["host_policy.cpp", "host_capi.cpp"]


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None):
    """defines/out: kernel-variant experiments (tools/sweep_variants.py); the product build uses neither."""
    if out is not None:
        objs = [os.path.join(CSRC, src.replace(".cpp", ".o")) for src in HOST_SRCS]
        cmd = [NVCC] + CUDA_FLAGS + ["-D" + d for d in defines] + ["-shared", "-o", out, os.path.join(CSRC, "engine.cu")] + objs
        subprocess.check_call(cmd)
        return out
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "rabe_b200.h")]
    if not force and not _newer(OUT, deps):
        return OUT
    objs = []
    for src in HOST_SRCS:
        obj = os.path.join(CSRC, src.replace(".cpp", ".o"))
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-Wall", "-c", "-o", obj, os.path.join(CSRC, src)])
        objs.append(obj)
    cmd = [NVCC] + CUDA_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", OUT, os.path.join(CSRC, "engine.cu")] + objs
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
