"""In-tree build of the CUDA library (sm_100a only).

    python -m ipcr_b200.build [--force]

Produces ipcr_b200/libipcr_b200.so next to this file.  nvcc cross-compiles
without a GPU; the .so is git-ignored but travels to the GPU box with gpurun.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libipcr_b200.so")
CLI = os.path.join(HERE, "ipcr_gpu_cli")
CUDA_HOME = os.environ.get("CUDA_HOME", "/usr/local/cuda")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
    "-shared",
]


def sources():
    out = []
    for root, _dirs, files in os.walk(CSRC):
        for f in files:
            if f.endswith((".cu", ".cuh", ".hpp", ".h", ".cpp")):
                out.append(os.path.join(root, f))
    out.append(os.path.join(os.path.dirname(HERE), "include", "ipcr_gpu.h"))
    return out


def is_stale() -> bool:
    if not os.path.exists(LIB) or not os.path.exists(CLI):
        return True
    t = min(os.path.getmtime(LIB), os.path.getmtime(CLI))
    return any(os.path.getmtime(s) > t for s in sources() if os.path.exists(s))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    nvcc = shutil.which("nvcc") or os.path.join(CUDA_HOME, "bin", "nvcc")
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libipcr_b200.so")
    lib64 = os.path.join(CUDA_HOME, "lib64")
    cmd = [nvcc] + NVCC_FLAGS + [
        "-o", LIB, os.path.join(CSRC, "engine.cu"),
        "-L" + lib64, "-lnvrtc", "-lz", "-Xlinker", "-rpath," + lib64,
    ]
    if verbose:
        cmd += ["-Xptxas", "-v"]
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    # the C++ host (flags, FASTA, visitors, --sort, TSV) above the C ABI
    gxx = shutil.which("g++") or "g++"
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-Wall", "-o", CLI,
                           os.path.join(CSRC, "host", "ipcr_gpu_cli.cpp"),
                           "-L" + HERE, "-lipcr_b200", "-Wl,-rpath," + HERE, "-Wl,-rpath," + lib64])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
