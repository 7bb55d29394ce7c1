"""Builds kontiki_b200/lib/libkontiki_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libkontiki_b200.so")
SOURCES = ["ktk.cu"]
HEADERS = ["spline_math.cuh", "split_math.cuh", "sensor_jac.cuh", "newton_math.cuh", "lie_math.cuh", "dualnum.cuh", "gn_device.cuh"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
              "-cudart", "shared", "-Xlinker", "-rpath=/usr/local/cuda/lib64"]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(os.path.dirname(_HERE), "include", "kontiki_b200.h")]
    return any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB_PATH
    os.makedirs(LIB_DIR, exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + [os.path.join(CSRC, s) for s in SOURCES]
    subprocess.check_call(cmd)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force=True, verbose=True))
