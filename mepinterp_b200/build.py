"""Build libmepb.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.  No torch, no JIT cache."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["mepb_small.cu", "mepb_sub.cu", "mepb_mep.cu", "mepb_eig.cu", "mepb_sweep.cu", "mepb_sweep_mma.cu", "mepb_large.cu", "mepb_tri.cu", "mepb_posop.cu", "mepb_api.cu", "mepb_comm.cu"]
LIB = os.path.join(HERE, "libmepb.so")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O3"]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    hdrs = [os.path.join(CSRC, "mepb_internal.cuh"), os.path.join(CSRC, "mepb_ctx.cuh"), os.path.join(HERE, "..", "include", "mepb.h")]
    objs = []
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(CSRC, s.replace(".cu", ".o"))
        if force or _stale(obj, [src] + hdrs):
            cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            subprocess.check_call(cmd)
        objs.append(obj)
    if force or _stale(LIB, objs):
        subprocess.check_call([nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs + ["-lcudart", "-ldl"])
    return LIB


DRIVER = os.path.join(HERE, "mepInterp")


def build_driver(force: bool = False) -> str:
    """g++ host driver (no CUDA code in it): links libmepb.so, rpath = its own directory."""
    src = os.path.join(CSRC, "driver", "mepinterp_main.cpp")
    build()
    if force or _stale(DRIVER, [src, LIB, os.path.join(HERE, "..", "include", "mepb.h")]):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-pthread", src, "-o", DRIVER, "-L" + HERE, "-lmepb",
                               "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + "/usr/local/cuda/lib64"])
    return DRIVER


TOOLS = os.path.join(HERE, "..", "tools")


def build_tools(force: bool = False) -> list:
    """Stand-alone measurement tools (not part of the product): the cuSOLVER heevjBatched yardstick for the eigensolvers."""
    out = []
    for name, libs in (("heevj_compare", ["-lcusolver"]),):
        src, exe = os.path.join(TOOLS, name + ".cu"), os.path.join(TOOLS, name)
        if os.path.exists(src) and (force or _stale(exe, [src])):
            subprocess.check_call([nvcc(), "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, src] + libs)
        out.append(exe)
    return out


if __name__ == "__main__":
    build_driver(force="--force" in sys.argv)
    build_tools(force="--force" in sys.argv)
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
