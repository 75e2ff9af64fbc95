"""Build libgeneclust_b200.so in-tree with nvcc for sm_100a (no torch / pybind dependency: plain C ABI).

    python -m scgeneclust_b200.csrc.build          # or __graft_entry__.build()
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(os.path.dirname(HERE), "libgeneclust_b200.so")

# (source, extra flags).  -fmad=false where float results must round like the numpy/sklearn expression.
SOURCES = [
    ("lib.cu", []),
    ("mi_kernels.cu", ["-fmad=false"]),
    ("kmeans_kernels.cu", ["-fmad=false"]),
    ("synth_kernels.cu", ["-fmad=false"]),
    ("embed_kernels.cu", []),
    ("small_dense.cu", []),
    ("rsvd_tc.cu", []),
    ("rng_kernels.cu", ["-fmad=false"]),
    ("deviance_kernels.cu", ["-fmad=false"]),
    ("mst_kernels.cu", []),
    ("confidence_kernels.cu", []),
    ("host_tail.cu", []),
]
COMMON = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
          "-I", os.path.join(ROOT, "include"), "-I", HERE]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libgeneclust_b200.so")
    return nvcc


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    obj_dir = os.path.join(HERE, "build")
    os.makedirs(obj_dir, exist_ok=True)
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "geneclust_b200.h"))
    newest_header = max(os.path.getmtime(h) for h in headers)
    objs, rebuilt = [], False
    procs = []
    for src, extra in SOURCES:
        path = os.path.join(HERE, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        objs.append(obj)
        if (not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(path), newest_header)):
            continue
        cmd = [nvcc, *COMMON, *extra, "-c", path, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        rebuilt = True
    for src, proc in procs:
        out, _ = proc.communicate()
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
        if verbose and out:
            print(out)
    if rebuilt or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
