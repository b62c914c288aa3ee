"""Build libabnn_b200.so in-tree with nvcc for sm_100a (no JIT cache: the .so travels with the repo).

    python -m abnn_b200.build [--force]

-fmad=false keeps the plasticity arithmetic an unfused IEEE sequence (bit-parity with the oracle);
-lineinfo lets ncu's source page map to these files.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.environ.get("ABNN_B200_OUT") or os.path.join(HERE, "libabnn_b200.so")   # override: tuning variants only
OBJ = os.path.join(HERE, "_obj" + os.environ.get("ABNN_B200_OBJ_SUFFIX", ""))
SOURCES = ["traversal.cu", "exact.cu", "io_kernels.cu", "structural.cu", "init.cu", "exchange.cu", "capi.cu"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-fmad=false",
              "-gencode", "arch=compute_100a,code=sm_100a",
              "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libabnn_b200.so cannot be built (and there is no CPU fallback)")


def _deps():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    files.append(os.path.join(HERE, "..", "include", "abnn.h"))
    files.append(os.path.abspath(__file__))
    return files


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(f) > t for f in _deps() if os.path.isfile(f))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    objs = []
    procs = []
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(OBJ, src + ".o")
        cmd = [nvcc] + NVCC_FLAGS + os.environ.get("ABNN_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-c", path, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose and out:
            print(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    link = [nvcc, "-shared", "-o", OUT] + objs + ["-lnccl", "-Xcompiler", "-fPIC"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return OUT


def build_host_tests() -> str:
    """Compile tests/host_cpp/engine_parity.cpp (host C++17 over include/abnn_brain.hpp) against the library."""
    root = os.path.abspath(os.path.join(HERE, ".."))
    src = os.path.join(root, "tests", "host_cpp", "engine_parity.cpp")
    exe = os.path.join(root, "tests", "host_cpp", "engine_parity")
    lib = build()
    deps = [src, os.path.join(root, "include", "abnn_brain.hpp"), os.path.join(root, "include", "abnn.h"), lib]
    if os.path.exists(exe) and all(os.path.getmtime(exe) >= os.path.getmtime(d) for d in deps):
        return exe
    cmd = ["g++", "-std=c++17", "-O2", "-Wall", "-I", os.path.join(root, "include"), src, "-o", exe,
           "-L", HERE, "-labnn_b200", "-Wl,-rpath," + HERE, "-pthread"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed on engine_parity.cpp:\n{r.stdout}")
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
