"""Build recipes for the native libraries (explicit nvcc / g++ command lines, sm_100a only, in-tree outputs)."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB_DEVICE = os.path.join(HERE, "libmccba.so")
LIB_HOST = os.path.join(HERE, "libmccba_host.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
              "-shared", "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _glob(d, exts):
    out = []
    for base, _, files in os.walk(d):
        for f in files:
            if f.endswith(exts):
                out.append(os.path.join(base, f))
    return out


def build_device(force=False, verbose=False):
    """libmccba.so: CUDA kernels + C-ABI shim (include/mccba.h)."""
    srcs = _glob(CSRC, (".cu", ".cuh")) + [os.path.join(ROOT, "include", "mccba.h")]
    srcs = [s for s in srcs if os.sep + "host" + os.sep not in s]
    if force or _stale(LIB_DEVICE, srcs):
        cmd = ["nvcc"] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
              ["-o", LIB_DEVICE, os.path.join(CSRC, "mccba_capi.cu"), "-ldl"]
        subprocess.check_call(cmd)
    return LIB_DEVICE


def build_host(force=False):
    """libmccba_host.so: C++17 host class mirroring MultiCameraCalibration, XML writer, synthetic rig generator."""
    hdir = os.path.join(CSRC, "host")
    srcs = _glob(hdir, (".cpp", ".hpp")) + _glob(os.path.join(ROOT, "include"), (".h", ".hpp"))
    cpps = sorted(s for s in srcs if s.endswith(".cpp"))
    if not cpps:
        return None
    if force or _stale(LIB_HOST, srcs + [LIB_DEVICE]):
        cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-o", LIB_HOST] + cpps + \
              ["-L", HERE, "-lmccba", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return LIB_HOST


def build_all(force=False):
    build_device(force)
    build_host(force)


if __name__ == "__main__":
    build_all(force=True)
    print("built", LIB_DEVICE, LIB_HOST if os.path.exists(LIB_HOST) else "")
