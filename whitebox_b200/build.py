"""Builds whitebox_b200/libwbx.so (CUDA kernels + C ABI + host engine) in-tree for sm_100a with nvcc."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libwbx.so")
SOURCES = ["wbx_kernels.cu", "wbx_fir_tc.cu", "wbx_fir_fft.cu", "wbx_api.cu", "wbx_host.cpp"]
HEADERS = [os.path.join(CSRC, "wbx_device.cuh")] + [
    os.path.join(ROOT, "include", h) for h in ("wbx.h", "wbx_host.h", "wbx_engine.hpp")]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                      # no FMA contraction: parity with the reference's separately rounded ops
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2,-Wall",
    "-cudart", "static",
]


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES] + HEADERS + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    # translation units are compiled in parallel, and only those older than their source / the shared headers
    hdr_t = max(os.path.getmtime(d) for d in HEADERS + [os.path.abspath(__file__)])
    objs, jobs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        o = os.path.join(CSRC, os.path.splitext(s)[0] + ".o")
        objs.append(o)
        if not force and os.path.exists(o) and os.path.getmtime(o) > max(hdr_t, os.path.getmtime(src)):
            continue
        cmd = [NVCC] + FLAGS + (["-x", "cu"] if s.endswith(".cpp") else []) + ["-c", src, "-o", o]
        if verbose:
            print(" ".join(cmd))
        jobs.append((cmd, subprocess.Popen(cmd)))
    for cmd, j in jobs:
        if j.wait() != 0:
            raise subprocess.CalledProcessError(j.returncode, cmd)
    cmd = [NVCC, "-shared", "-cudart", "static", "-Wno-deprecated-gpu-targets", "-o", OUT] + objs
    if verbose:
        print(" ".join(cmd))
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(OUT)
