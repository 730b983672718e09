"""Build recipe for the sm_100a shared library (and the oracle): explicit nvcc / g++ commands.

The library is built IN-TREE (tdnn-f_nas_b200/lib/libtdnnf_nas_b200.so) so that it travels to the
GPU box with the repository snapshot.  nvcc cross-compiles for sm_100a without a GPU.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libtdnnf_nas_b200.so")

CUDA_SOURCES = ["context.cu", "splice_gemm.cu", "mixing.cu", "den.cu", "den_slices.cu", "num.cu", "neighbours.cu", "ng.cu", "orthonormal.cu", "chain_step.cu"]
CXX_SOURCES = [
    "nnet3/shim.cc",
    "nnet3/indexes.cc",
    "nnet3/components.cc",
    "nnet3/natural_gradient.cc",
    "nnet3/handle_api.cc",
    "chain_io.cc",
    "egs_io.cc",
]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-I", os.path.join(ROOT, "include"),
]
CXX_FLAGS = ["-O2", "-std=c++17", "-fPIC", "-Wall", "-I", os.path.join(ROOT, "include"),
             "-I", "/usr/local/cuda/include"]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: the CUDA extension cannot be built")
    return exe


def _digest(paths, extra) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode())
            h.update(f.read())
    h.update(repr(extra).encode())
    return h.hexdigest()


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build command failed: " + " ".join(cmd[:3]) + " ...")
    return r.stdout + r.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA / C++ source into libtdnnf_nas_b200.so; returns its path."""
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    cu = [os.path.join(CSRC, s) for s in CUDA_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    cc = [os.path.join(CSRC, s) for s in CXX_SOURCES if os.path.exists(os.path.join(CSRC, s))]
    headers = []
    for d, _, files in os.walk(CSRC):
        headers += [os.path.join(d, f) for f in files if f.endswith((".h", ".cuh"))]
    headers.append(os.path.join(ROOT, "include", "tdnnf_nas_b200.h"))
    stamp = os.path.join(LIBDIR, ".stamp")
    dig = _digest(cu + cc + headers, (NVCC_FLAGS, CXX_FLAGS))
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = _nvcc()
    jobs = []
    objs = []
    for src in cu:
        obj = os.path.join(OBJDIR, os.path.basename(src) + ".o")
        objs.append(obj)
        jobs.append([nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])
    for src in cc:
        obj = os.path.join(OBJDIR, os.path.basename(src) + ".o")
        objs.append(obj)
        jobs.append(["g++"] + CXX_FLAGS + ["-c", src, "-o", obj])
    with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
        outs = list(ex.map(_run, jobs))
    if verbose:
        print("\n".join(outs))
    _run([nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static", "-ldl"])
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
