"""Builds the in-tree native libraries of voxeltoy_b200 for sm_100a.

    python -m voxeltoy_b200.build

libvoxeltoy_b200.so  CUDA kernels + the C ABI (include/voxeltoy_b200.h)
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    # arithmetic contract (csrc/vt_math.cuh): no FMA contraction, IEEE div/sqrt, no flush-to-zero
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off", "-shared",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _sources(d, exts):
    out = []
    for base, _, files in os.walk(d):
        for f in files:
            if f.endswith(exts):
                out.append(os.path.join(base, f))
    return sorted(out)


def build(force=False, verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    lib = os.path.join(HERE, "libvoxeltoy_b200.so")
    deps = _sources(CSRC, (".cu", ".cuh", ".h", ".inl")) + _sources(HOST, (".cpp", ".h")) + [os.path.join(ROOT, "include", "voxeltoy_b200.h")]
    if force or _newer(lib, deps):
        cus = _sources(CSRC, (".cu",))
        cpps = _sources(HOST, (".cpp",))
        cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-I", os.path.join(ROOT, "include"), "-I", HERE,
                                                                            "-o", lib] + cus + cpps + ["-lz"]
        subprocess.check_call(cmd)
    # the C++ multi-GPU demo (tools/vt_group_demo.cpp): a host program without Python over the same library
    demo_src = os.path.join(ROOT, "tools", "vt_group_demo.cpp")
    demo = os.path.join(HERE, "vt_group_demo")
    if os.path.exists(demo_src) and (force or _newer(demo, [demo_src, lib])):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", os.path.join(ROOT, "include"), "-I", "/usr/local/cuda/include", "-o", demo, demo_src,
                               "-L", HERE, "-lvoxeltoy_b200", "-Wl,-rpath,$ORIGIN"])
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
